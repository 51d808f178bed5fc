#!/bin/bash
# inert obstacle leaves pruned: the merged-vs-unmerged test and one timing run
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_step.py -m gpu -x -q -s -k "merged or early_out_is_exact" > gpurun_out/r2ao_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ao_tests.log
grep -E "merged vs|passed|failed|rc=" gpurun_out/r2ao_tests.log | tail -6
python bench.py --steps 30 --warmup 3 --skip-e2e --skip-checks 2>gpurun_out/r2ao_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel_ms']; eo=d['early_out']; ld=d['library_default']
print('all pairs ms %.4f |' % d['ms_per_step'], ' '.join('%s %.4f' % (n, k[n]['ms_per_step']) for n in k), '| early_out ms %.4f | default ms %.4f' % (eo['ms_per_step'], ld['ms_per_step']), {a: round(b,4) for a,b in ld['kernel_ms'].items()}, ld['pair_loops'], 'merged all pairs', ld['all_pairs_merged']['ms_per_step'])
" > gpurun_out/r2ao_timing.txt 2>&1; cat gpurun_out/r2ao_timing.txt
