#!/bin/bash
# sanity at HEAD: GPU tests, smoke, default bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_tests.log
tail -5 gpurun_out/r2j_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2j_bench.json
