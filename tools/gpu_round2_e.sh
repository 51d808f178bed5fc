#!/bin/bash
mkdir -p gpurun_out
run() {  # name, nvcc flags
  RMP2_BUILD_OUT=/tmp/lib_$1.so RMP2_NVCC_EXTRA="$2" python riemannian_motion_policies_b200/build.py --force > /dev/null 2>&1 || { echo "$1 build failed"; return; }
  RMP2_B200_LIB=/tmp/lib_$1.so python bench.py --steps 20 --warmup 3 --skip-e2e --skip-checks 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'all-pairs ms', round(d['ms_per_step'],4), 'spheres', round(d['kernel_ms']['spheres']['ms_per_step'],4), '| early-out ms', round(d['early_out']['ms_per_step'],4), 'speedup', round(d['early_out']['speedup_over_all_pairs'],3))"
}
{
run base ""
run nosort "-DRMP2_SKIP_SORT=0"
run skipmb5 "-DRMP2_SPHERES_SKIP_MIN_BLOCKS=5"
run skipmb7 "-DRMP2_SPHERES_SKIP_MIN_BLOCKS=7"
run skipmb8 "-DRMP2_SPHERES_SKIP_MIN_BLOCKS=8"
run nosort_mb8 "-DRMP2_SKIP_SORT=0 -DRMP2_SPHERES_SKIP_MIN_BLOCKS=8"
} > gpurun_out/r2e_variants.txt 2>&1
cat gpurun_out/r2e_variants.txt
python -m pytest tests/test_gpu_step.py -m gpu -q -k "early_out or edge or tma or specialized_kernels_match" 2>&1 | tail -3
python -m pytest tests/test_gpu_properties.py -m gpu -q -k "specialized_kernels_full_size or independent" 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rmp2_spheres -s 5 -c 1 -o gpurun_out/r2e_spheres_skip \
  python bench.py --steps 2 --warmup 1 --skip-e2e --skip-checks > gpurun_out/r2e_ncu.log 2>&1
