#!/bin/bash
# cycle counters of the warp-level tensor-core scans (measurements only)
SNB200_NVCC_EXTRA="-DSN_SCAN_PROF" python -m structurednets_b200.build --force > /dev/null 2>&1 || echo build failed
python bench.py --steps 2 --warmup 1 --quick --no-cpu-baseline --no-graph --global-batch 8192 > gpurun_out/scanprof_raw.log 2>&1
grep CHTOT gpurun_out/scanprof_raw.log | tail -16
grep "CHPROF scan bwd blk 0" gpurun_out/scanprof_raw.log | tail -5
tail -1 gpurun_out/scanprof_raw.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernels']
print('step_ms %.4f' % d['ms_per_step'], {n.replace('sss_tc_','').replace('_kernel',''):round(v['avg_ms']*1000*v['launches']/d['steps'],1) for n,v in k.items()})
"
