#!/bin/bash
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_step.py -m gpu -x -q -k "golden_fixtures or committed_oracle or specialized_kernels_match or leaf_update or reference_source" > gpurun_out/r2ap_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ap_tests.log
tail -3 gpurun_out/r2ap_tests.log
