#!/bin/bash
# early-out variants: time the spheres kernel with the library default (early-out on), then one ncu capture
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
run unroll2 "-DRMP2_SKIP_UNROLL=2"
run nosort_unroll2 "-DRMP2_SKIP_SORT=0 -DRMP2_SKIP_UNROLL=2"
run minblk4 "-DRMP2_SPHERES_MIN_BLOCKS=4"
run minblk6 "-DRMP2_SPHERES_MIN_BLOCKS=6"
} > gpurun_out/r2d_variants.txt 2>&1
cat gpurun_out/r2d_variants.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rmp2_spheres -s 5 -c 2 -o gpurun_out/r2d_spheres_skip \
  python bench.py --steps 2 --warmup 1 --skip-e2e --skip-checks > gpurun_out/r2d_ncu.log 2>&1
