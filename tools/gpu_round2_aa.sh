#!/bin/bash
# transposed reach test of the sorted early-out (variant library built ahead): exactness tests, then A/B timing
mkdir -p gpurun_out
export T=$PWD/gpurun_in/lib_transposed.so
RMP2_B200_LIB=$T timeout 900 python -m pytest tests -m gpu -x -q -k "early_out or merged or full_size or determin or independent or chunk" > gpurun_out/r2aa_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aa_tests.log
tail -3 gpurun_out/r2aa_tests.log
run() {
  RMP2_B200_LIB=$2 python bench.py --steps 50 --warmup 5 --skip-e2e --skip-checks 2>gpurun_out/r2aa_err_$1.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel_ms']; eo=d['early_out']; ld=d['library_default']
print('$1 | all pairs ms %.4f |' % d['ms_per_step'], ' '.join('%s %.4f' % (n, k[n]['ms_per_step']) for n in k), '| early_out ms %.4f | default ms %.4f' % (eo['ms_per_step'], ld['ms_per_step']), {a: round(b,4) for a,b in ld['kernel_ms'].items()})
"
}
{
run base ""
run transposed $T
run base_again ""
run transposed_again $T
} > gpurun_out/r2aa_timing.txt 2>&1
cat gpurun_out/r2aa_timing.txt
