#!/bin/bash
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernels']
print('$1', 'step_ms %.4f' % d['ms_per_step'], {n.replace('sss_tc_','').replace('_kernel',''):round(v['avg_ms']*1000*v['launches']/d['steps'],1) for n,v in k.items()})
"; }
python -m pytest tests/test_sss_tc_gpu.py tests/test_config_size_gpu.py tests/test_sss_gpu.py -x -q -m gpu -k "sss or SSS" 2>&1 | tail -2
for gb in 8192 12288 16384; do
  SNB200_SSS_TC_CHAIN=0 python bench.py --steps 20 --quick --no-cpu-baseline --global-batch $gb 2>/dev/null | show "B=$gb mma-scans"
done
python bench.py --steps 20 --quick --no-cpu-baseline --global-batch 4096 2>/dev/null | show "B=4096"
