#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares from an ncu report:
   python profiles/src_lines.py gpurun_out/prof.ncu-rep [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
files = {}
cur = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if not r or r[0] in ("Function Name",) or cur is None: continue
    try: ln = int(r[0])
    except ValueError: continue
    if r[2] != "-": continue          # only the per-line aggregate rows
    iex = hdr.index("Instructions Executed"); ism = hdr.index("# Samples")
    files.setdefault(cur, []).append((ln, r[1], int(r[iex] or 0), int(r[ism] or 0)))
tot_i = sum(x[2] for v in files.values() for x in v); tot_s = sum(x[3] for v in files.values() for x in v)
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
for f, v in files.items():
    print("==", f)
    for ln, src, i, s in v:
        if i >= tot_i * minp / 100 or s >= tot_s * minp / 100:
            print("%5d  inst %5.2f%%  samples %5.2f%%  %s" % (ln, 100.0 * i / tot_i, 100.0 * s / tot_s, src.strip()[:100]))
