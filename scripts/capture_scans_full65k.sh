#!/bin/bash
# ncu --set full of the scan kernels at the full C5 batch
set -e
CMD="python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --no-graph"
$CMD > gpurun_out/r2zi_pre.log 2>&1
python scripts/show_bench.py gpurun_out/r2zi_pre.log | head -3
ncu --set full --clock-control none --import-source on -k regex:"sss_tc_scan_.*_m_kernel" --launch-skip 9 -c 3 -o gpurun_out/r2zi_scans65k -f $CMD > gpurun_out/r2zi_ncu.log 2>&1
ls -la gpurun_out/r2zi_scans65k*
