#!/bin/bash
# Rebuild the library with different spheres-kernel knobs on the GPU box and time config 4.
# usage (under gpurun): bash tools/tune_spheres.sh "-DRMP2_SPHERES_MIN_BLOCKS=8" "-DRMP2_SPHERES_STEPS_PER_TRIP=4" ...
for flags in "$@"; do
  RMP2_NVCC_EXTRA="$flags" python riemannian_motion_policies_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  regs=$(grep -A2 "rmp2_spheres_kernelILb1ELb0" riemannian_motion_policies_b200/csrc/build.log | grep -o "Used [0-9]* registers" | head -1)
  python bench.py --steps 20 --warmup 3 --skip-e2e --skip-checks ${BENCH_EXTRA} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel_ms']; eo=d.get('early_out') or {}
print('$flags | $regs |', 'value %.4g ms %.4f |' % (d['value'], d['ms_per_step']), ' '.join('%s %.4f' % (n, k[n]['ms_per_step']) for n in k), '| early_out %.4g' % eo.get('value', 0))
"
done
