#!/bin/bash
# product library (TMA issue spread over warps in the early-out variant): full GPU tests; then A/B of the prefetch-a variant
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2ad_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ad_tests.log
tail -2 gpurun_out/r2ad_tests.log
SKIP_VARIANT_TESTS=1 bash tools/gpu_round2_ac.sh gpurun_in/lib_prefa.so r2adv
