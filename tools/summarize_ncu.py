"""Summarise ncu artefacts brought back in gpurun_out/ into small tracked files under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_bench.csv profiles/r1_launches_bench.md
  python tools/summarize_ncu.py rep gpurun_out/prof_march.ncu-rep profiles/r1_ncu_march.md [top_kernel.json]
  python tools/summarize_ncu.py rep gpurun_out/prof_x_raw.csv profiles/r2_ncu_x.md      (raw page exported on the box)
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * m.get(unit, 1)


def rep(path, out, top_json=None):
    # a .ncu-rep, or the `ncu -i … --page raw --csv` export of one made on the GPU box (reports of many launches exceed
    # what travels back)
    if path.endswith(".csv"):
        txt = "".join(l for l in open(path, newline="") if not l.startswith("=="))
    else:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full summary of `{path}`", "",
             "(per launch; cold-cache, serialised replays — compare shares and ratios, not absolutes)", ""]
    agg = OrderedDict()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("bpltv::", "")
        agg.setdefault(name, []).append(r)
    top = None
    for name, rs in agg.items():
        lines += [f"## {name}  ({len(rs)} launch(es) captured)", "", "| metric | " + " | ".join(f"#{k}" for k in range(len(rs))) + " | unit |", "|---|" + "---|" * (len(rs) + 1)]
        for k in KEYS:
            if k in idx:
                lines.append(f"| {k} | " + " | ".join(r[idx[k]] for r in rs) + f" | {units[idx[k]]} |")
        lines.append("")
        r = rs[-1]
        if "dram__bytes_read.sum" in idx:
            tr = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            if top is None:
                top = {"kernel": name, "dram_bytes_per_launch": tr,
                       "duration_us": float(r[idx["gpu__time_duration.sum"]]), "source": path}
    open(out, "w").write("\n".join(lines) + "\n")
    if top_json and top:
        json.dump(top, open(top_json, "w"), indent=1)
    print("wrote", out)


def launches(path, out):
    rows = []
    for line in open(path, newline=""):
        if line.startswith("=="):
            continue
        rows.append(line)
    rd = list(csv.DictReader(io.StringIO("".join(rows))))
    agg = OrderedDict()
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("bpltv::", "")
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = [f"# ncu launch list summary of `{path}`", "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: the SHARE is the evidence)", "",
             "| kernel | launches | total µs | avg µs | share |", "|---|---|---|---|---|"]
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {name} | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.2f} % |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "rep":
        rep(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        launches(sys.argv[2], sys.argv[3])
