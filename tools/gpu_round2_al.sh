#!/bin/bash
# slot-fastest thread mapping of the pair kernel (variant libraries built ahead): A/B timing, then tests on the variant
mkdir -p gpurun_out
run() {
  RMP2_B200_LIB=$2 python bench.py --steps 50 --warmup 5 --skip-e2e --skip-checks 2>gpurun_out/r2al_err_$1.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel_ms']; eo=d['early_out']; ld=d['library_default']
print('$1 | all pairs ms %.4f |' % d['ms_per_step'], ' '.join('%s %.4f' % (n, k[n]['ms_per_step']) for n in k), '| early_out ms %.4f | default ms %.4f' % (eo['ms_per_step'], ld['ms_per_step']), {a: round(b,4) for a,b in ld['kernel_ms'].items()})
"
}
{
run base ""
run sf2 $PWD/gpurun_in/lib_sf2.so
run sf3 $PWD/gpurun_in/lib_sf3.so
run base_again ""
run sf2_again $PWD/gpurun_in/lib_sf2.so
} > gpurun_out/r2al_timing.txt 2>&1
cat gpurun_out/r2al_timing.txt
RMP2_B200_LIB=$PWD/gpurun_in/lib_sf3.so timeout 600 python -m pytest tests -m gpu -x -q -k "early_out or merged or full_size or edge or tma or golden" > gpurun_out/r2al_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2al_tests.log
tail -2 gpurun_out/r2al_tests.log
