#!/bin/bash
# merged coincident obstacle leaves (RMP2_OPT_MERGE_COINCIDENT): GPU tests, then the full bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s -k "merged or early_out or specialized" > gpurun_out/r2r_tests_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_tests_new.log
grep -E "config . n=|passed|failed|rc=" gpurun_out/r2r_tests_new.log | tail -8
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_tests.log
tail -3 gpurun_out/r2r_tests.log
timeout 900 python bench.py --steps 50 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench exit $?" >> gpurun_out/r2r_bench.err
tail -2 gpurun_out/r2r_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2r_bench.json').read().strip().splitlines()[-1])
print('value %.4g ms %.4f' % (d['value'], d['ms_per_step']), {k:round(v['ms_per_step'],4) for k,v in d['kernel_ms'].items()})
print('early_out', d['early_out']['ms_per_step'], 'library_default', json.dumps(d['library_default'])[:700])
print('parity', json.dumps(d['parity'])[:400])
for k,v in d['other_configs'].items(): print(k, {kk:vv for kk,vv in v.items() if kk!='parity'})
PY
