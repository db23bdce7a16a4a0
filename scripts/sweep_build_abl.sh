#!/bin/bash
# Ablations of the tensor-core build kernel (run on the GPU box): 0 = full kernel, 1 = parameter staging only.
for abl in 1 0; do
  SNB200_NVCC_EXTRA="-DSN_BUILD_ABL=$abl" python -m structurednets_b200.build --force > /dev/null 2>&1 || { echo "build failed abl=$abl"; continue; }
  python bench.py --steps 20 --quick --no-cpu-baseline --global-batch 8192 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernels']
print('abl=$abl', {n.replace('sss_tc_','').replace('_kernel',''):round(v['avg_ms']*1000,1) for n,v in k.items() if 'build' in n or 'pack' in n})
"
done
