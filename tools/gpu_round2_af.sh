#!/bin/bash
# third session of round 2, final artefacts: bench line (driver's command), ncu launch list, full captures of one
# all-pairs step (value / roofline setting) and one library-default step
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err; echo "bench exit $? in $(( $(date +%s) - t0 )) s" >> gpurun_out/r2af_bench.err
tail -1 gpurun_out/r2af_bench.err
CMD="python bench.py --steps 2 --warmup 1 --skip-e2e --skip-checks"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:rmp2_' -c 400 --csv --log-file gpurun_out/r2af_launches.csv $CMD > gpurun_out/r2af_ncu_launches.log 2>&1
# 5 steps x 4 kernels per phase: launches 0..19 all pairs / every leaf, 20..39 early-out, 40..59 library default
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:rmp2_(spec|spheres|resolve)' -s 12 -c 4 -f -o gpurun_out/r2af_step $CMD > gpurun_out/r2af_ncu_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:rmp2_(spec|spheres|resolve)' -s 44 -c 4 -f -o gpurun_out/r2af_default $CMD > gpurun_out/r2af_ncu_default.log 2>&1
tail -1 gpurun_out/r2af_ncu_step.log; tail -1 gpurun_out/r2af_ncu_default.log
ls -la gpurun_out/r2af_*
