#!/bin/bash
# third GPU pass of round 2: tests with the two-clause criterion, parity study, bench, ncu of the pair kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2c_tests.log
timeout 600 python tools/parity_study.py --out gpurun_out/r2c_parity.json > gpurun_out/r2c_parity.log 2>&1
timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench exit $?" >> gpurun_out/r2c_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv \
  python bench.py --steps 2 --warmup 1 --skip-e2e --skip-checks > gpurun_out/r2c_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rmp2_spheres -c 6 -o gpurun_out/r2c_spheres \
  python bench.py --steps 2 --warmup 1 --skip-e2e --skip-checks > gpurun_out/r2c_ncu_spheres.log 2>&1
grep -E "passed|failed|error" gpurun_out/r2c_tests.log | tail -5
tail -2 gpurun_out/r2c_bench.err
