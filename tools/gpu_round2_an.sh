#!/bin/bash
# one full ncu capture of the library-default pair kernel on the final code (the 12th launch of rmp2_spheres_kernel in the
# bench process: phases of 5 steps -- all pairs / early-out every leaf / library default)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rmp2_spheres -s 11 -c 1 -f -o gpurun_out/r2an_spheres_default python bench.py --steps 2 --warmup 1 --skip-e2e --skip-checks > gpurun_out/r2an_ncu.log 2>&1
tail -2 gpurun_out/r2an_ncu.log
