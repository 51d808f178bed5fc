#!/bin/bash
# final code of the third session: full GPU tests, smoke, the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2am_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2am_tests.log
tail -2 gpurun_out/r2am_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2am_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2am_smoke.log; tail -2 gpurun_out/r2am_smoke.log
timeout 600 python bench.py > gpurun_out/r2am_bench.json 2> gpurun_out/r2am_bench.err; echo "bench exit $?" >> gpurun_out/r2am_bench.err; tail -1 gpurun_out/r2am_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2am_bench.json').read().strip().splitlines()[-1])
ld=d['library_default']
print('value %.4g ms %.4f' % (d['value'], d['ms_per_step']), {k:round(v['ms_per_step'],4) for k,v in d['kernel_ms'].items()})
print('early', d['early_out']['ms_per_step'], d['early_out']['value'], 'default', ld['ms_per_step'], ld['value'], ld['speedup_over_all_pairs'], ld['kernel_ms'])
for kk,v in d['other_configs'].items(): print(kk, {a:b for a,b in v.items() if a in ('ms_per_step','value','value_early_out','value_library_default')})
print('rollout', {kk:(v['ms_per_rollout'], v['control_steps_per_s']) for kk,v in d['rollout'].items()}, 'latency', d['latency_b1_us'])
print('parity', d['parity']['pass_a_strict_1e-5_vs_f32'], d['parity']['kept'], 'e2e', d['e2e']['value'])
PY
