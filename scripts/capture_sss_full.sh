#!/bin/bash
# One ncu --set full capture of the four dominant SSS kernels of the C5 bench step (after the same command exits 0 without ncu).
set -e
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --no-graph"
$CMD > gpurun_out/r2_full_pre.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"sss_tc_(grad_gemm|chain_fwd|local_gemm|chain_bwd)_kernel" --launch-skip 12 -c 8 \
    -o gpurun_out/r2_sss_top -f $CMD > gpurun_out/r2_full_ncu.log 2>&1
ncu -i gpurun_out/r2_sss_top.ncu-rep --page raw --csv > gpurun_out/r2_sss_top_raw.csv
ls -la gpurun_out/r2_sss_top*
