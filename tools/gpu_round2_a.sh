#!/bin/bash
# first GPU pass of round 2: tests, parity study (with / without the Newton-refined distance), quick bench in both resolve modes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2a_tests.log
timeout 600 python tools/parity_study.py --out gpurun_out/r2a_parity_newton.json > gpurun_out/r2a_parity_newton.log 2>&1
RMP2_BUILD_OUT=/tmp/librmp2_nonewton.so RMP2_NVCC_EXTRA=-DRMP2_SQRT_NEWTON=0 python riemannian_motion_policies_b200/build.py --force > gpurun_out/r2a_build_nonewton.log 2>&1
RMP2_B200_LIB=/tmp/librmp2_nonewton.so timeout 600 python tools/parity_study.py --out gpurun_out/r2a_parity_nonewton.json > gpurun_out/r2a_parity_nonewton.log 2>&1
for split in 1 0; do
  RMP2_SPLIT_RESOLVE=$split timeout 600 python bench.py --steps 20 --warmup 3 --skip-e2e --skip-checks > gpurun_out/r2a_bench_split$split.json 2> gpurun_out/r2a_bench_split$split.err
done
RMP2_B200_LIB=/tmp/librmp2_nonewton.so RMP2_SPLIT_RESOLVE=1 timeout 600 python bench.py --steps 20 --warmup 3 --skip-e2e --skip-checks > gpurun_out/r2a_bench_nonewton.json 2> gpurun_out/r2a_bench_nonewton.err
for c in 2 3 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 --skip-e2e --skip-checks > gpurun_out/r2a_bench_c$c.json 2> gpurun_out/r2a_bench_c$c.err
done
tail -3 gpurun_out/r2a_tests.log
