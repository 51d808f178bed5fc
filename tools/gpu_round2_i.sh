#!/bin/bash
mkdir -p gpurun_out
for c in 4 5 3 2; do for split in 0 1; do
RMP2_SPLIT_RESOLVE=$split python bench.py --config $c --steps 30 --warmup 3 --skip-e2e --skip-checks 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('config$c split$split', 'ms', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['kernel_ms'].items()}, 'early ms', d['early_out'] and round(d['early_out']['ms_per_step'],4))"
done; done > gpurun_out/r2i_split.txt 2>&1
cat gpurun_out/r2i_split.txt
