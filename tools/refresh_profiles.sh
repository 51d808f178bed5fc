#!/bin/bash
# gpurun_out/<tag>_{bench.json,launches.csv,step.ncu-rep,default.ncu-rep} (tools/gpu_round2_z.sh and its copies) ->
# profiles/r2_* (bench line, launch list, ncu summaries of one all-pairs step and one library-default step, traffic)
# usage (build container, no GPU): bash tools/refresh_profiles.sh r2af
TAG=$1
CMD="python bench.py --steps 2 --warmup 1 --skip-e2e --skip-checks"
python tools/summarize_ncu.py gpurun_out/${TAG}_step.ncu-rep gpurun_out/${TAG}_launches.csv profiles/r2_ncu_summary.md "$CMD  (ncu --set full --clock-control none --import-source on -k 'regex:rmp2_(spec|spheres|resolve)' -s 12 -c 4: one step of the all-pairs / every-leaf phase = the value and roofline setting; launch list: ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rmp2_ -c 400 --csv over the whole process, which runs 5 steps each of: all pairs every leaf / early-out every leaf / library default / all pairs merged)" profiles/r2_traffic.json 4 1048576 > /dev/null
python tools/summarize_ncu.py gpurun_out/${TAG}_default.ncu-rep gpurun_out/${TAG}_launches.csv profiles/r2_ncu_summary_library_default.md "$CMD  (ncu --set full ... -s 44 -c 4: one step of the library-default phase -- RMP2_OPT_EARLY_OUT and RMP2_OPT_MERGE_COINCIDENT on; launch list as in r2_ncu_summary.md)" profiles/r2_traffic_library_default.json 4 1048576 > /dev/null
cp gpurun_out/${TAG}_bench.json profiles/r2_bench_config4.json; cp gpurun_out/${TAG}_launches.csv profiles/r2_launches_config4_B1048576.csv
TAG=$TAG python - <<'PY'
import csv, collections, json, os
tag=os.environ['TAG']
rows=[]
with open(f'gpurun_out/{tag}_launches.csv') as fh:
    rd=csv.reader(l for l in fh if not l.startswith('=='))
    h=next(rd); ik,iv,im=h.index('Kernel Name'),h.index('Metric Value'),h.index('Metric Name')
    for r in rd:
        if len(r)>iv and r[im]=='gpu__time_duration.sum':
            rows.append((r[ik].replace('void ','').split('(')[0], float(r[iv].replace(',',''))))
step=[x for x in rows if 'fk_kernel' not in x[0]]
names=['all pairs, every leaf (value / roofline)','early-out, every leaf','library default (early-out + merged control points)','all pairs, merged control points']
out=["", "## per phase of the bench process (launch list in launch order; 5 steps = 20 launches per phase; us per launch under ncu, mean of the 5)", "",
     "| phase | frames | spheres | step (direct resolve fused) | resolve fallback | sum |", "|---|---|---|---|---|---|"]
for p in range(4):
    seg=step[20*p:20*p+20]
    agg=collections.OrderedDict()
    for k,t in seg: agg.setdefault(k,[]).append(t)
    vals=[sum(v)/len(v)/1e3 for v in agg.values()]
    out.append(f"| {names[p]} | "+" | ".join(f"{v:.1f}" for v in vals)+f" | {sum(vals):.1f} |")
d=json.loads(open(f'gpurun_out/{tag}_bench.json').read().strip().splitlines()[-1])
k=d['kernel_ms']; ld=d['library_default']
t0=json.load(open('profiles/r2_traffic.json'))['kernels']; t1=json.load(open('profiles/r2_traffic_library_default.json'))['kernels']
tot=lambda t: sum(v['dram_bytes_read']+v['dram_bytes_write'] for v in t.values())/1e9
out += ["", "CUDA-event times of the same phases in `bench.py` (200 timed steps, `profiles/r2_bench_config4.json`), frames / spheres / step / fallback: all pairs "
        + " / ".join(f"{k[n]['ms_per_step']:.4f}" for n in ('frames','spheres','step','resolve_fallback')) + f" ms (step {d['ms_per_step']:.4f} ms); library default "
        + " / ".join(f"{ld['kernel_ms'][n]:.4f}" for n in ('frames','spheres','step','resolve_fallback')) + f" ms (step {ld['ms_per_step']:.4f} ms). "
        f"DRAM traffic of one step (these captures): every leaf on its own {tot(t0):.2f} GB (`profiles/r2_traffic.json`), library default {tot(t1):.2f} GB (`profiles/r2_traffic_library_default.json`)."]
for f in ('profiles/r2_ncu_summary.md','profiles/r2_ncu_summary_library_default.md'):
    open(f,'a').write("\n".join(out)+"\n")
print("\n".join(out[-7:]))
print('value %.4g ms %.4f' % (d['value'], d['ms_per_step']), 'early', d['early_out']['ms_per_step'], d['early_out']['value'], 'default', ld['ms_per_step'], ld['value'], ld['speedup_over_all_pairs'], 'merged all pairs', ld['all_pairs_merged']['ms_per_step'])
print('latency', d['latency_b1_us']); print('rollout', {kk:(v['ms_per_rollout'], v['control_steps_per_s']) for kk,v in d['rollout'].items()})
print('e2e', d['e2e']['value'], d['e2e']['frac_of_h2d_peak'])
for kk,v in d['other_configs'].items(): print(kk, {a:b for a,b in v.items() if a in ('ms_per_step','value','value_early_out','value_library_default')})
print('roofline', d['roofline']['frac'], d['roofline_fp32']['frac'], d['roofline_fp32']['whole_step']['frac'], d['roofline']['kernel_ms_per_launch'], d['roofline']['traffic'])
print('parity', d['parity']['pass_a_strict_1e-5_vs_f32'], d['parity']['kept'], d['clocks'])
PY
