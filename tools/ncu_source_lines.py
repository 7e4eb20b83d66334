"""Aggregate an ncu source page (--print-source cuda,sass) by CUDA source line over all source files of a kernel:
instructions executed and stall samples.  usage: python tools/ncu_source_lines.py report.ncu-rep [kernel-index] [top]"""
import csv
import io
import os
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
# one block per (kernel launch, source file), each starting with "File Path" / "Function Name" header lines
blocks, cur = [], []
for line in out.splitlines():
    if line.startswith('"File Path"') and cur:
        blocks.append(cur); cur = []
    cur.append(line)
blocks.append(cur)
funcs = []
for b in blocks:
    fn = b[1] if len(b) > 1 else ""
    if fn not in funcs:
        funcs.append(fn)
ki = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = []
tot_i = tot_s = 0
for blk in blocks:
    if len(blk) < 3 or blk[1] != funcs[ki]:
        continue
    fname = os.path.basename(next(csv.reader([blk[0]]))[1])
    hdr_i = next(i for i, l in enumerate(blk) if l.startswith('"Line No"'))
    rows = list(csv.reader(io.StringIO("\n".join(blk[hdr_i:]))))
    hdr = rows[0]
    iL, iSamp, iInst = hdr.index("Line No"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    nI, nS = len(hdr) - iInst, len(hdr) - iSamp      # index from the end: source text may contain commas/quotes
    for r in rows[1:]:
        if r and r[iL].isdigit() and len(r) >= len(hdr):
            try:
                agg.append([f"{fname}:{r[iL]}", " ".join(r[1:len(r) - len(hdr) + 2]).strip()[:100], int(r[-nI] or 0), int(r[-nS] or 0)])
            except ValueError:
                continue
            tot_i += agg[-1][2]; tot_s += agg[-1][3]
print(funcs[ki])
print(f"total inst {tot_i}  samples {tot_s}")
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print(f"{a[0]:>24s} inst {100*a[2]/max(tot_i,1):5.1f}%  stall-samples {100*a[3]/max(tot_s,1):5.1f}%  {a[1]}")
