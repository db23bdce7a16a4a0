#!/bin/bash
# graph-replay step: tcgen05 chain kernels vs warp-level tensor-core scans around the crossover
for gb in 16384 24576 32768 49152; do
  for m in 0 1; do
    SNB200_SSS_TC_CHAIN=$m python bench.py --steps 30 --quick --no-cpu-baseline --global-batch $gb 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=$gb chain=$m step_ms %.4f' % d['ms_per_step'])"
  done
done
