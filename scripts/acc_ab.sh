#!/bin/bash
# accuracy and speed with the lo part of the warp-level tensor-core operands left as the exact remainder (shipped) or rounded to tf32
for r in 0 1; do
  SNB200_NVCC_EXTRA="-DSN_MMA_LO_ROUND=$r" python -m structurednets_b200.build --force > /dev/null 2>&1 || echo build failed
  echo "== SN_MMA_LO_ROUND=$r"
  python scripts/acc_check.py 1000 2>&1 | tail -2
  for gb in 8192 65536; do
    python bench.py --steps 30 --quick --no-cpu-baseline --global-batch $gb 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   B=$gb step_ms %.4f' % d['ms_per_step'])"
  done
done
