"""Turn an .ncu-rep (ncu --set full) and a launch list (ncu --metrics gpu__time_duration.sum --csv) into the
markdown summary committed under profiles/.  Runs in the build container (no GPU needed).

    python tools/summarize_ncu.py gpurun_out/prof_r1_v4.ncu-rep gpurun_out/launches_r1.csv profiles/r1_ncu_summary.md "<command>" \
        [profiles/r1_traffic.json <config> <envs>]

The optional traffic file (per-kernel dram__bytes_read/write of the captured launch) is what bench.py reads for
`roofline.traffic`.
"""
import json
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum", "sm__cycles_elapsed.avg.per_second",
]


def main():
    rep, launches, out, cmd = sys.argv[1:5]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = ["# ncu summary", "", f"command: `{cmd}`", "",
             "Captured with `ncu --set full --clock-control none --import-source on` on a B200 (one launch of each",
             "kernel, after warm-up; per-launch times under ncu are cold-cache and serialised -- compare shares).", ""]
    traffic = {}
    for r in data:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        try:
            traffic[d["Kernel Name"].replace("void ", "").split("(")[0]] = {
                "dram_bytes_read": float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]],
                "dram_bytes_write": float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]],
                "duration_ms_under_ncu": float(d["gpu__time_duration.sum"]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[u["gpu__time_duration.sum"]],
            }
        except (KeyError, ValueError):
            pass
        lines.append(f"## {d['Kernel Name']}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for m in METRICS:
            if m in d:
                lines.append(f"| {m} | {d[m]} | {u[m]} |")
        stalls = sorted(((float(v), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                         for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")),
                        reverse=True)
        lines.append("")
        lines.append("warp stall reasons (warps stalled per issue-active cycle): " + ", ".join(f"{n} {x:.2f}" for x, n in stalls[:8]))
        lines.append("")
    # launch list: aggregate by kernel name
    agg = {}
    with open(launches) as fh:
        rd = csv.reader(l for l in fh if not l.startswith("=="))
        lh = next(rd)
        ik, iv, im = lh.index("Kernel Name"), lh.index("Metric Value"), lh.index("Metric Name")
        for r in rd:
            if len(r) <= iv or r[im] != "gpu__time_duration.sum":
                continue
            t = float(r[iv].replace(",", ""))
            a = agg.setdefault(r[ik], [0, 0.0])
            a[0] += 1
            a[1] += t
    total = sum(a[1] for a in agg.values())
    lines += ["## launch list (whole process, `--metrics gpu__time_duration.sum`)", "",
              "| kernel | launches | total time | share |", "|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        lines.append(f"| `{k[:90]}` | {n} | {t/1e3:.1f} us | {100*t/total:.1f} % |")
    ours = {k: v for k, v in agg.items() if "rmp2_" in k}
    tot_ours = sum(v[1] for v in ours.values())
    lines += ["", "shares among the step's own kernels:", ""]
    for k, (n, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"* `{k.split('(')[0]}`: {n} launches, {t/n/1e3:.1f} us per launch, {100*t/tot_ours:.1f} % of the step")
    with open(out, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("wrote", out)
    if len(sys.argv) >= 8:
        tpath, config, envs = sys.argv[5], int(sys.argv[6]), int(sys.argv[7])
        with open(tpath, "w") as fh:
            json.dump({"command": cmd, "config": config, "envs": envs, "kernels": traffic,
                       "note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of each kernel from the "
                               "ncu --set full capture summarised in " + out}, fh, indent=1)
        print("wrote", tpath)


if __name__ == "__main__":
    main()
