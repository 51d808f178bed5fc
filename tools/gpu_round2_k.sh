#!/bin/bash
# full GPU test suite + config 4/5 bench lines (kernel times, early-out)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2k_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_tests.log
tail -15 gpurun_out/r2k_tests.log
for c in 4 5; do
python bench.py --config $c --steps 30 --warmup 3 --skip-e2e --skip-checks 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('config$c', 'ms', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['kernel_ms'].items()}, 'early ms', d['early_out'] and round(d['early_out']['ms_per_step'],4), 'roofline', d.get('roofline_fp32'))"
done > gpurun_out/r2k_bench.txt 2>&1
cat gpurun_out/r2k_bench.txt
