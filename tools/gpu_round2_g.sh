#!/bin/bash
# N-GPU run: bench under torchrun (collect + e2e at N ranks), plus the reference arm on rank 0
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2ai_bench_n$N.json 2> gpurun_out/r2ai_bench_n$N.err
echo "exit $?" >> gpurun_out/r2ai_bench_n$N.err
tail -3 gpurun_out/r2ai_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ai_bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'])
print("early", d["early_out"]["value"], "default", d["library_default"]["value"])
print('e2e', d['e2e']['value'], d['e2e']['frac_of_h2d_peak'], d['e2e']['h2d_peak_gbs_per_rank'], d['e2e']['h2d_gbs_per_rank'])
print('collect', d['collect'])
PY
