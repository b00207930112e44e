#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep + launches.csv into the committed text summaries under profiles/.
usage: python profiles/summarize.py <round-tag> <ncu-rep> [launches.csv]"""
import csv, io, json, subprocess, sys
from collections import defaultdict

tag, rep = sys.argv[1], sys.argv[2]
launches = sys.argv[3] if len(sys.argv) > 3 else None
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__inst_executed_op_global_red.sum", "smsp__inst_executed_op_global_atom.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = [f"# ncu --set full --clock-control none summary ({tag}); source report: {rep}", ""]
traffic = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    out.append(f"## {name}")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"  {w:85s} {r[i]} {units[i]}")
    try:
        rd = float(r[hdr.index('dram__bytes_read.sum')]); wr = float(r[hdr.index('dram__bytes_write.sum')])
        u = units[hdr.index('dram__bytes_read.sum')]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
        traffic[name] = (rd + wr) * scale
        out.append(f"  traffic (dram read+write per launch)                                                  {traffic[name]/1e9:.3f} GB")
    except Exception:
        pass
    out.append("")
if launches:
    d = defaultdict(list)
    rr = [x for x in csv.reader(open(launches)) if x and not x[0].startswith("==")]
    h = rr[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    for x in rr[1:]:
        if len(x) > vi:
            d[x[ki]].append(float(x[vi].replace(",", "")))
    out.append("## launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)")
    tot = sum(sum(v) for k, v in d.items() if "init" not in k and "pack" not in k)
    for k, v in d.items():
        out.append(f"  {k[:80]:80s} n={len(v):3d} mean={sum(v)/len(v)/1e6:8.3f} ms share={100*sum(v)/tot:5.1f}%")
open(f"profiles/{tag}_ncu_summary.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
json.dump({k: v for k, v in traffic.items()}, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
