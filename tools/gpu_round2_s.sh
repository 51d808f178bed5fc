#!/bin/bash
# full GPU test run after the feed fix (marker tree unmerged)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_tests.log
tail -4 gpurun_out/r2s_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2s_smoke.log; tail -3 gpurun_out/r2s_smoke.log
