#!/bin/bash
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernels']
print('$1', 'step_ms %.4f' % d['ms_per_step'], {n.replace('sss_tc_','').replace('_kernel',''):round(v['avg_ms']*1000*v['launches']/d['steps'],1) for n,v in k.items()})
"; }
for m in 0 1; do
  SNB200_SSS_TC_CHAIN=$m python bench.py --steps 30 --quick --no-cpu-baseline --global-batch 65536 2>/dev/null | show "B=65536 chain=$m"
done
SNB200_SSS_TC_CHAIN=0 python bench.py --steps 30 --quick --no-cpu-baseline --global-batch 131072 2>/dev/null | show "B=131072 chain=0"
SNB200_SSS_TC_CHAIN=1 python bench.py --steps 30 --quick --no-cpu-baseline --global-batch 131072 2>/dev/null | show "B=131072 chain=1"
