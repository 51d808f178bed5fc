#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2f_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2f_tests.log
grep -E "passed|failed|error" gpurun_out/r2f_tests.log | tail -3
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit $?" >> gpurun_out/r2f_bench.err
RMP2_BUILD_OUT=/tmp/lib_skipall.so RMP2_NVCC_EXTRA="-DRMP2_SKIP_MIN_SPHERES=0" python riemannian_motion_policies_b200/build.py --force > /dev/null 2>&1
for lib in default skipall; do
  for c in 3; do
    if [ $lib = default ]; then unset RMP2_B200_LIB; else export RMP2_B200_LIB=/tmp/lib_skipall.so; fi
    python bench.py --config $c --steps 20 --warmup 3 --skip-e2e --skip-checks 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib config$c', 'all-pairs ms', round(d['ms_per_step'],4), '| early-out ms', round(d['early_out']['ms_per_step'],4), 'speedup', round(d['early_out']['speedup_over_all_pairs'],3))"
  done
done > gpurun_out/r2f_config3.txt 2>&1
unset RMP2_B200_LIB
cat gpurun_out/r2f_config3.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -3 gpurun_out/r2f_smoke.log
