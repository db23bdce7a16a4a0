#!/bin/bash
# round-end validation on one GPU: the whole GPU test suite, smoke(), the default bench line and the reference arm
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/r3a_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3a_ref.json 2> gpurun_out/r3a_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r3a_ref.json
