"""Turn the raw artefacts of scripts/capture_profiles.sh (gpurun_out/<tag>_*) into the tracked summaries under
profiles/: launch list + per-kernel shares, the key `ncu --set full` metrics of the likelihood kernel, the
bench lines and the sanitizer log.   usage: python scripts/summarize_profiles.py <tag-in-gpurun_out> <tag-in-profiles>
"""
import collections
import csv
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src_tag, dst_tag = sys.argv[1], sys.argv[2]
G = os.path.join(REPO, "gpurun_out")
P = os.path.join(REPO, "profiles")

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_fp64.sum",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALLS = "smsp__average_warps_issue_stalled_"


def launches():
    path = os.path.join(G, f"{src_tag}_launches.csv")
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    shutil.copy(path, os.path.join(P, f"{dst_tag}_launches.csv"))
    tot = collections.OrderedDict()
    for r in rows:
        k = r[ik]
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += float(r[iv].replace(",", "")) * 1e-6
    allms = sum(v[1] for v in tot.values())
    step = {k: v for k, v in tot.items() if "fp64_peak" not in k}
    stepms = sum(v[1] for v in step.values())
    with open(os.path.join(P, f"{dst_tag}_launch_summary.csv"), "w") as fh:
        fh.write(f"# ncu launch list summary ({len(rows)} launches of `bench.py --steps 4 --warmup 3 --burn 20 --legs none`, C4; "
                 "gpu__time_duration.sum, --clock-control none)\n"
                 "# cold-cache, serialised: compare SHARES, not absolutes; share_of_step excludes the FP64 peak microbenchmark\n"
                 "kernel,launches,total_ms,share,share_of_step\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            sos = f"{v[1] / stepms:.4f}" if k in step else ""
            fh.write(f'"{k[:90]}",{v[0]},{v[1]:.3f},{v[1] / allms:.4f},{sos}\n')
    print("launch summary:", {k[:40]: round(v[1] / stepms, 4) for k, v in step.items()})


def kernel_metrics():
    path = os.path.join(G, f"{src_tag}_logl_raw.csv")
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = []
    for i, h in enumerate(hdr):
        if h in KEEP or (h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")):
            out.append((h, units[i], vals[i]))
    with open(os.path.join(P, f"{dst_tag}_logl_kernel_ncu.csv"), "w") as fh:
        fh.write("# ncu --set full --import-source on --clock-control none, emp::logl_rv_kernel<2, 2> (groups, feature mask), a launch of the\n"
                 "# warm-up/timed region of `bench.py --steps 4 --warmup 3 --burn 20 --legs none` (C4: N=10k, K=5, 4 ins, global MA(1); 32768\n"
                 f"# proposals per launch, ~78% inside the prior = evaluated).  source: profiles/{dst_tag}_logl.ncu-rep\n"
                 "metric,unit,value\n")
        for h, u, v in out:
            fh.write(f"{h},{u},{v}\n")
    d = {h: v for h, u, v in out}
    units_of = {h: u for h, u, v in out}

    def num(k):
        try:
            return float(str(d[k]).replace(",", ""))
        except (KeyError, ValueError):
            return None

    def scaled(k):  # ncu prints Kbyte / Mbyte / Gbyte, usecond / msecond ...
        v, u = num(k), units_of.get(k, "")
        if v is None:
            return None
        for pre, f in (("K", 1e3), ("M", 1e6), ("G", 1e9), ("T", 1e12)):
            if u.startswith(pre) and "byte" in u:
                return v * f
        if u in ("ns", "us", "ms", "s") or "second" in u:
            return v * {"n": 1e-9, "u": 1e-6, "m": 1e-3}.get(u[0], 1.0)
        return v
    dur = scaled("gpu__time_duration.sum")
    fl = [num("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"), num("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"),
          num("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum")]
    summary = {"source": f"profiles/{dst_tag}_logl_kernel_ncu.csv (ncu --set full --clock-control none, one launch of "
                         "emp::logl_rv_kernel<2, 2> on the burnt-in C4 ensemble; profiles/" + dst_tag + "_logl.ncu-rep)",
               "launch_ms_under_ncu": dur * 1e3 if dur else None,
               "fp64_pipe_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
               "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "xu_pct": num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
               "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
               "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"),
               "registers": num("launch__registers_per_thread")}
    xb, lts = scaled("l1tex__m_xbar2l1tex_read_bytes.sum"), scaled("lts__t_bytes.sum")
    if dur and xb:
        summary["l2_to_sm_gbs"] = xb / dur * 1e-9
        summary["l2_to_sm_bytes"] = xb
    if dur and lts:
        summary["l2_traffic_gbs"] = lts / dur * 1e-9
    if dur and all(x is not None for x in fl):
        ex = 2 * fl[0] + fl[1] + fl[2]
        summary["executed_fp64_tflops"] = ex / dur * 1e-12
        summary["frac_executed"] = ex / dur * 1e-12 / 33.9  # emp_fp64_peak of this pool's B200
        summary["executed_fp64_flops_per_launch"] = ex
    kj = os.path.join(P, "kernel_ncu.json")
    allk = json.load(open(kj)) if os.path.exists(kj) else {}
    allk["c4"] = summary
    json.dump(allk, open(kj, "w"), indent=1)
    dr, dw = scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum")
    if dr is not None:
        tj = os.path.join(P, "kernel_traffic.json")
        json.dump({"c4": {"dram_bytes_per_launch": int(dr + (dw or 0)),
                          "source": f"profiles/{dst_tag}_logl_kernel_ncu.csv: dram__bytes_read.sum {dr * 1e-6:.2f} MB + "
                                    f"dram__bytes_write.sum {(dw or 0) * 1e-6:.2f} MB"}}, open(tj, "w"), indent=1)
    rep = os.path.join(G, f"{src_tag}_logl.ncu-rep")
    if os.path.exists(rep):
        shutil.copy(rep, os.path.join(P, f"{dst_tag}_logl.ncu-rep"))
    print("kernel_ncu.json:", summary)
    print("kernel:", d.get("gpu__time_duration.sum"), "ms; fp64", d.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
          "issue", d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), "dram read", d.get("dram__bytes_read.sum"))
    return d


def benches():
    for name in ("c4", "c5", "c2", "reference"):
        p = os.path.join(G, f"{src_tag}_bench_{name}.json")
        if os.path.exists(p) and os.path.getsize(p) > 2:
            line = open(p).read().strip().splitlines()[-1]
            json.loads(line)
            open(os.path.join(P, f"{dst_tag}_bench_{name}.json"), "w").write(line + "\n")
    s = os.path.join(G, f"{src_tag}_sanitizer.log")
    if os.path.exists(s):
        shutil.copy(s, os.path.join(P, f"{dst_tag}_compute_sanitizer.log"))


if __name__ == "__main__":
    launches()
    kernel_metrics()
    benches()
