#!/bin/bash
# third session of round 2: GPU tests on the current tree, then A/B of the frame-record load position in the pair kernel
# (RMP2_REC_EARLY 0 / 1; variant libraries built ahead into gpurun_in/)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_tests.log
tail -3 gpurun_out/r2q_tests.log
run() {  # name, library
  RMP2_B200_LIB=$2 python bench.py --steps 30 --warmup 5 --skip-e2e --skip-checks 2>gpurun_out/r2q_err_$1.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel_ms']; eo=d.get('early_out') or {}
print('$1 |', 'ms %.4f |' % d['ms_per_step'], ' '.join('%s %.4f' % (n, k[n]['ms_per_step']) for n in k), '| early_out ms %.4f' % eo.get('ms_per_step', 0))
"
}
{
run rec0 $PWD/gpurun_in/lib_rec0.so
run rec1 $PWD/gpurun_in/lib_rec1.so
run rec0_again $PWD/gpurun_in/lib_rec0.so
run rec1_again $PWD/gpurun_in/lib_rec1.so
} > gpurun_out/r2q_rec.txt 2>&1
cat gpurun_out/r2q_rec.txt
