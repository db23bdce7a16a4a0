#!/bin/bash
# Ablations of the persistent gradient GEMM (measurements only): 0 = as shipped, 1 = no reductions into dM, 2 = items range-fastest.
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernels']
print('$1', 'step_ms %.4f' % d['ms_per_step'], {n.replace('sss_tc_','').replace('_kernel',''):round(v['avg_ms']*1000*v['launches']/d['steps'],1) for n,v in k.items() if 'grad' in n})
"; }
for abl in 1 2 0; do
  SNB200_NVCC_EXTRA="-DSN_G2_ABL=$abl" python -m structurednets_b200.build --force > /dev/null 2>&1 || { echo "build failed abl=$abl"; continue; }
  for gb in 8192 65536; do
    python bench.py --steps 20 --quick --no-cpu-baseline --global-batch $gb 2>/dev/null | show "abl=$abl B=$gb"
  done
  if [ $abl = 1 ]; then SNB200_SSS_G2_PER=4 python bench.py --steps 20 --quick --no-cpu-baseline --global-batch 8192 2>/dev/null | show "abl=$abl B=8192 per=4"; fi
done
