#!/bin/bash
# ncu launch list (durations only) of the SSS step at the per-GPU batch of 8 GPUs
CMD="python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --no-graph --global-batch 8192"
$CMD > gpurun_out/small_pre.log 2>&1 || { echo plain run failed; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c5_8192.csv $CMD > gpurun_out/small_ncu.log 2>&1
python profiles/summarize.py launches gpurun_out/r2_launches_c5_8192.csv | grep -E "sss_tc|colsum|total"
