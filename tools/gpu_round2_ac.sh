#!/bin/bash
# A/B of a variant library (built ahead into gpurun_in/) against the in-tree product library; bench only
# usage: bash tools/gpu_round2_ac.sh <variant.so> <tag>
mkdir -p gpurun_out
V=$PWD/$1; TAG=$2
run() {
  RMP2_B200_LIB=$2 python bench.py --steps 50 --warmup 5 --skip-e2e --skip-checks 2>gpurun_out/${TAG}_err_$1.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel_ms']; eo=d['early_out']; ld=d['library_default']
print('$1 | all pairs ms %.4f |' % d['ms_per_step'], ' '.join('%s %.4f' % (n, k[n]['ms_per_step']) for n in k), '| early_out ms %.4f | default ms %.4f' % (eo['ms_per_step'], ld['ms_per_step']), {a: round(b,4) for a,b in ld['kernel_ms'].items()})
"
}
{
run base ""
run variant $V
run base_again ""
run variant_again $V
} > gpurun_out/${TAG}_timing.txt 2>&1
cat gpurun_out/${TAG}_timing.txt
if [ -z "$SKIP_VARIANT_TESTS" ]; then RMP2_B200_LIB=$V timeout 600 python -m pytest tests -m gpu -x -q -k "early_out or tma or full_size or edge" > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log; fi
[ -z "$SKIP_VARIANT_TESTS" ] && tail -2 gpurun_out/${TAG}_tests.log
