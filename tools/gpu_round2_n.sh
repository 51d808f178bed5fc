#!/bin/bash
# step kernel: instruction-fetch experiment -- larger blocks and block barriers keep the warps of an SM sub-partition at
# the same place of the ~110 KB straight-line specialised code
mkdir -p gpurun_out
run() {  # name, step block, NVRTC flags
  RMP2_SPEC_STEP_BLOCK=$2 RMP2_JIT_EXTRA="$3" python bench.py --steps 20 --warmup 3 --skip-e2e --skip-checks --skip-early-out 2>gpurun_out/r2n_err_$1.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel_ms']
print('$1 |', 'ms %.4f |' % d['ms_per_step'], ' '.join('%s %.4f' % (n, k[n]['ms_per_step']) for n in k))
"
}
{
run b256 256 "-DRMP2_BLOCK_THREADS=256 -DRMP2_STEP_MIN_BLOCKS(N)=2"
run b512 512 "-DRMP2_BLOCK_THREADS=512 -DRMP2_STEP_MIN_BLOCKS(N)=1"
run b512_again 512 "-DRMP2_BLOCK_THREADS=512 -DRMP2_STEP_MIN_BLOCKS(N)=1"
PROBE_CONFIG=5 true
} > gpurun_out/r2n_lockstep2.txt 2>&1
cat gpurun_out/r2n_lockstep2.txt; tail -3 gpurun_out/r2n_err_b512.txt
