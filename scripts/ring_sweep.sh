#!/bin/bash
for gb in 32768 65536; do for r in 3 2; do
  SNB200_SSS_SCAN_RING=$r python bench.py --steps 30 --quick --no-cpu-baseline --global-batch $gb 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels']['sss_tc_scan_bwd_m_kernel']
print('B=$gb ring=$r step_ms %.4f scan_bwd %.1f us' % (d['ms_per_step'], k['avg_ms']*1000))"
done; done
