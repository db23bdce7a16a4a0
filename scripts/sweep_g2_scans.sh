#!/bin/bash
# Round-2 measurements: persistent gradient GEMM at the per-GPU batches of 1/2/4/8 GPUs, and the chain kernels against the
# warp-level tensor-core scans above the current crossover.
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernels']
print('$1', 'step_ms %.4f' % d['ms_per_step'], {n.replace('sss_tc_','').replace('_kernel',''):round(v['avg_ms']*1000*v['launches']/d['steps'],1) for n,v in k.items()})
"; }
python -m pytest tests/test_sss_tc_gpu.py -x -q -m gpu 2>&1 | tail -2
for gb in 8192 16384 32768 65536; do
  python bench.py --steps 20 --quick --no-cpu-baseline --global-batch $gb 2>/dev/null | show "B=$gb default"
done
for gb in 16384 32768 65536; do
  SNB200_SSS_TC_CHAIN=0 python bench.py --steps 20 --quick --no-cpu-baseline --global-batch $gb 2>/dev/null | show "B=$gb mma-scans"
done
for per in 4 16 32; do
  SNB200_SSS_G2_PER=$per python bench.py --steps 20 --quick --no-cpu-baseline --global-batch 8192 2>/dev/null | show "B=8192 per=$per"
done
for per in 16 28; do
  SNB200_SSS_G2_PER=$per python bench.py --steps 20 --quick --no-cpu-baseline --global-batch 65536 2>/dev/null | show "B=65536 per=$per"
done
