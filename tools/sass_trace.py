#!/usr/bin/env python
"""Compressed trace of a kernel's SASS: runs of the same instruction class are collapsed, so the
placement of barriers / loads relative to the FP work is visible at a glance.  Development tool.
usage: sass_trace.py <object> <kernel-substring> [start_hex end_hex]"""
import re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
fn = None; ops = []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: fn = m.group(1); continue
    if fn is None or pat not in fn: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        a = int(m.group(1), 16)
        if lo <= a <= hi: ops.append((a, m.group(2).strip()))
def cls(t):
    t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t2.split()[0]
    base = op.split(".")[0]
    if base in ("FADD2", "FMUL2", "FFMA2"): return "FP2"
    if base in ("FFMA", "FADD", "FMUL"): return "FP1"
    if base in ("LDS", "STS", "PRMT", "LDG", "STG"): return op if base in ("LDS","STS") else base
    if base in ("SYNCS", "BAR", "UBLKCP", "BRA", "WARPSYNC", "ELECT", "MEMBAR", "FENCE", "BSSY", "BSYNC", "EXIT"): return "!" + t
    return "int"
run = None; n = 0; first = 0
for a, t in ops:
    c = cls(t)
    if c.startswith("!"):
        if run: print("  %04x  %-8s x%d" % (first, run, n)); run = None
        print("  %04x  %s" % (a, c[1:]))
        continue
    if c != run:
        if run: print("  %04x  %-8s x%d" % (first, run, n))
        run, n, first = c, 0, a
    n += 1
if run: print("  %04x  %-8s x%d" % (first, run, n))
