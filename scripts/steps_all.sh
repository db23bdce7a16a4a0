#!/bin/bash
# SSS tests + graph-replay steps at the per-GPU batches of 8 / 4 / 2 / 1 GPUs with the per-kernel table
python -m pytest tests/test_sss_tc_gpu.py tests/test_config_size_gpu.py tests/test_sss_gpu.py -x -q -m gpu -k "sss or SSS" 2>&1 | tail -1
for gb in 8192 16384 32768 65536; do
  python bench.py --steps 30 --quick --no-cpu-baseline --global-batch $gb 2>/dev/null | python scripts/show_bench.py /dev/stdin | head -2
done
