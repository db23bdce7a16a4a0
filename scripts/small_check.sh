#!/bin/bash
# tests of the SSS paths, graph-replay step at small per-GPU batches, and the ncu launch list (durations) at 8192 samples
python -m pytest tests/test_sss_tc_gpu.py tests/test_config_size_gpu.py tests/test_sss_gpu.py -x -q -m gpu -k "sss or SSS" 2>&1 | tail -2
for gb in 4096 8192 12288; do
  python bench.py --steps 30 --quick --no-cpu-baseline --global-batch $gb 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=$gb step_ms %.4f' % d['ms_per_step'])"
done
CMD="python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --no-graph --global-batch 8192"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c5_8192.csv $CMD > gpurun_out/small_ncu.log 2>&1
python profiles/summarize.py launches gpurun_out/r2_launches_c5_8192.csv | grep -E "sss_tc|colsum|total"
