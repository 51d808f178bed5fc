#!/bin/bash
# the driver's own commands on the final code: reference arm, then the bench line
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2aj_ref.json 2> gpurun_out/r2aj_ref.err; echo "ref exit $? in $(( $(date +%s) - t0 )) s" >> gpurun_out/r2aj_ref.err
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err; echo "bench exit $? in $(( $(date +%s) - t0 )) s" >> gpurun_out/r2aj_bench.err
tail -1 gpurun_out/r2aj_ref.err; tail -1 gpurun_out/r2aj_bench.err
python - <<'PY'
import json
r=json.loads(open('gpurun_out/r2aj_ref.json').read().strip().splitlines()[-1]); print('reference', r['value'], r['cpu_baseline']['cores'], r['ms_per_step'])
d=json.loads(open('gpurun_out/r2aj_bench.json').read().strip().splitlines()[-1])
print('value %.4g ms %.4f' % (d['value'], d['ms_per_step']), 'default', d['library_default']['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks'])
PY
