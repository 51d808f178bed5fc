#!/bin/bash
# second GPU pass of round 2
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2b_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2b_tests.log
timeout 600 python tools/parity_study.py --out gpurun_out/r2b_parity_newton.json > gpurun_out/r2b_parity_newton.log 2>&1
RMP2_BUILD_OUT=/tmp/librmp2_nonewton.so RMP2_NVCC_EXTRA=-DRMP2_SQRT_NEWTON=0 python riemannian_motion_policies_b200/build.py --force > gpurun_out/r2b_build_nonewton.log 2>&1
RMP2_B200_LIB=/tmp/librmp2_nonewton.so timeout 600 python tools/parity_study.py --out gpurun_out/r2b_parity_nonewton.json > gpurun_out/r2b_parity_nonewton.log 2>&1
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench exit $?" >> gpurun_out/r2b_bench.err
RMP2_B200_LIB=/tmp/librmp2_nonewton.so timeout 300 python bench.py --steps 20 --skip-e2e --skip-checks > gpurun_out/r2b_bench_nonewton.json 2> gpurun_out/r2b_bench_nonewton.err
grep -E "passed|failed|error" gpurun_out/r2b_tests.log | tail -5
tail -2 gpurun_out/r2b_bench.err
