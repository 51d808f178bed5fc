#!/bin/bash
# final artefacts of round 2 (second half): bench line, parity study, ncu launch list and full captures
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench exit $?" >> gpurun_out/r2p_bench.err
timeout 600 python tools/parity_study.py --out gpurun_out/r2p_parity.json > gpurun_out/r2p_parity.log 2>&1
CMD="python bench.py --steps 2 --warmup 1 --skip-e2e --skip-checks"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:rmp2_' -c 400 --csv --log-file gpurun_out/r2p_launches.csv $CMD > gpurun_out/r2p_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:rmp2_(spec|spheres|resolve)' -s 12 -c 4 -f -o gpurun_out/r2p_step $CMD > gpurun_out/r2p_ncu_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rmp2_spheres -s 5 -c 1 -f -o gpurun_out/r2p_spheres_skip $CMD > gpurun_out/r2p_ncu_skip.log 2>&1
tail -2 gpurun_out/r2p_bench.err
ls -la gpurun_out/r2p_*
