#!/bin/bash
# ncu --set full (with source) of the batch-independent and small-batch SSS kernels at the per-GPU batch of 8 GPUs.
set -e
CMD="python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --no-graph --global-batch 8192"
$CMD > gpurun_out/r2v_pre.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"sss_tc_(scan_.*_m|buildm|build_bwdm)_kernel" --launch-skip 15 -c 5 \
    -o gpurun_out/${1:-r2v_small} -f $CMD > gpurun_out/r2v_ncu.log 2>&1
ls -la gpurun_out/${1:-r2v_small}*
