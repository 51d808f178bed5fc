#!/bin/bash
# Time config 4 with different NVRTC options for the tree-specialised kernels (under gpurun):
#   bash tools/tune_jit.sh "" "-DRMP2_SPLIT_MIN_BLOCKS(N)=5" ...
for flags in "$@"; do
  RMP2_JIT_EXTRA="$flags" python bench.py --steps 20 --warmup 3 --skip-e2e --skip-checks --skip-early-out 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel_ms']
print('$flags |', 'value %.4g ms %.4f |' % (d['value'], d['ms_per_step']), ' '.join('%s %.4f' % (n, k[n]['ms_per_step']) for n in k), '| step regs', d['kernel']['step']['registers'], 'frames regs', d['kernel']['frames']['registers'])
"
done
