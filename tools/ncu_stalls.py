#!/usr/bin/env python
"""Summarise an ncu report of one kernel: per-issue stall breakdown and the top stalled SASS
instructions with their dominant reasons.  Development tool.  usage: ncu_stalls.py report.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
d = dict(zip(hdr, vals))
print("duration_ms", d.get("gpu__time_duration.sum"), " issue_active", d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
      " fma_pipe", d.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
      " alu_pipe", d.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
      " lsu_wavefronts", d.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
      " regs", d.get("launch__registers_per_thread"))
st = {k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: float(v) for k, v in d.items()
      if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
print("warp-cycles per issue:", " ".join("%s=%.3f" % kv for kv in sorted(st.items(), key=lambda x: -x[1]) if kv[1] > 0.005), " total=%.2f" % sum(st.values()))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
base = int(rows[2][0], 16)
L = []
for r in rows[2:]:
    s = int(r[ix["# Samples"]] or 0)
    dd = sorted(((k[6:], int(r[ix[k]] or 0)) for k in stalls), key=lambda x: -x[1])[:3]
    L.append((int(r[0], 16) - base, r[1].strip(), s, dd))
tot = sum(x[2] for x in L)
print("samples", tot)
for x in sorted(L, key=lambda x: -x[2])[:top]:
    print("%05x %5d %4.1f%%  %-64s %s" % (x[0], x[2], 100.0 * x[2] / tot, x[1][:64], " ".join("%s=%d" % kv for kv in x[3] if kv[1])))
