#!/bin/bash
set -e
CMD="python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --no-graph"
$CMD > gpurun_out/r3g_pre.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sss_tc_scan_bwd_m_kernel" --launch-skip 3 -c 1 -o gpurun_out/r3g_scanbwd65k -f $CMD > gpurun_out/r3g_ncu.log 2>&1
ls -la gpurun_out/r3g_scanbwd65k*
