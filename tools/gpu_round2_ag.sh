#!/bin/bash
# all-pairs pair kernel: compile-time variants (built ahead into gpurun_in/) against the product library, all-pairs phase only
mkdir -p gpurun_out
run() {
  RMP2_B200_LIB=$2 python bench.py --steps 50 --warmup 5 --skip-e2e --skip-checks --skip-early-out 2>gpurun_out/r2ag_err_$1.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel_ms']
print('$1 | all pairs ms %.4f |' % d['ms_per_step'], ' '.join('%s %.4f' % (n, k[n]['ms_per_step']) for n in k))
"
}
{
run base ""
for v in spread3 trip2 trip8 mb4 mb6; do run $v $PWD/gpurun_in/lib_$v.so; done
run base_again ""
} > gpurun_out/r2ag_timing.txt 2>&1
cat gpurun_out/r2ag_timing.txt
