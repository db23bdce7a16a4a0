#!/bin/bash
# Lanes-per-column sweep of the SSS build kernels (run on the GPU box): rebuilds libsnb200.so per setting, checks the kernel-level
# parity test, prints the build kernels' durations from a short eager bench at 8 192 samples.
for lpc in 4 8 16; do
  SNB200_NVCC_EXTRA="-DSN_B4_LPC=$lpc" python -m structurednets_b200.build --force > /dev/null 2>&1 || { echo "build failed lpc=$lpc"; continue; }
  python -m pytest tests/test_sss_tc_gpu.py -m gpu -q -x -k "emulator or alexnet" 2>&1 | tail -1
  python bench.py --steps 20 --quick --no-cpu-baseline --global-batch 8192 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernels']
print('lpc=$lpc', 'step_ms', round(d['ms_per_step'],4), {n.replace('sss_tc_','').replace('_kernel',''):round(v['avg_ms']*1000,1) for n,v in k.items() if 'build' in n or 'pack' in n})
"
done
python -m structurednets_b200.build --force > /dev/null 2>&1
