#!/bin/bash
# full GPU tests, then two timing runs of the bench (all pairs / early-out / library default with per-kernel times)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2ae_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ae_tests.log
tail -2 gpurun_out/r2ae_tests.log
for i in 1 2; do
python bench.py --steps 50 --warmup 5 --skip-e2e --skip-checks 2>gpurun_out/r2ae_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernel_ms']; eo=d['early_out']; ld=d['library_default']
print('all pairs ms %.4f |' % d['ms_per_step'], ' '.join('%s %.4f' % (n, k[n]['ms_per_step']) for n in k), '| early_out ms %.4f | default ms %.4f' % (eo['ms_per_step'], ld['ms_per_step']), {a: round(b,4) for a,b in ld['kernel_ms'].items()})
"
done > gpurun_out/r2ae_timing.txt 2>&1
cat gpurun_out/r2ae_timing.txt
