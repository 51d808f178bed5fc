#!/bin/bash
# early-out variants: expanded reach test on/off, ncu capture of each
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
run scalar "-DRMP2_REACH_SCALAR=1"
run scalar_mb8 "-DRMP2_REACH_SCALAR=1 -DRMP2_SPHERES_SKIP_MIN_BLOCKS=8"
run mb8 "-DRMP2_SPHERES_SKIP_MIN_BLOCKS=8"
run mb6 "-DRMP2_SPHERES_SKIP_MIN_BLOCKS=6"
} > gpurun_out/r2l_variants.txt 2>&1
cat gpurun_out/r2l_variants.txt
[ -n "$NO_NCU" ] && exit 0
for v in x1 x0; do
RMP2_B200_LIB=/tmp/lib_$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:rmp2_spheres -s 5 -c 1 -f -o gpurun_out/r2l_skip_$v \
  python bench.py --steps 2 --warmup 1 --skip-e2e --skip-checks > gpurun_out/r2l_ncu_$v.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
