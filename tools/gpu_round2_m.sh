#!/bin/bash
# ncu --set full capture of one launch of every kernel of the step (fused resolve, all pairs) + the early-out variant
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --skip-e2e --skip-checks"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:rmp2_(spec|spheres|resolve)' -s 12 -c 4 -f -o gpurun_out/r2m_step $CMD > gpurun_out/r2m_ncu_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rmp2_spheres -s 5 -c 1 -f -o gpurun_out/r2m_spheres_skip $CMD > gpurun_out/r2m_ncu_skip.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:rmp2_' -c 400 --csv --log-file gpurun_out/r2m_launches.csv $CMD > gpurun_out/r2m_ncu_launches.log 2>&1
ls -la gpurun_out/r2m_*
