#!/usr/bin/env python
"""SASS listing of one kernel of an ncu report with executed-instruction and stall-sample shares:
   python profiles/sass_hot.py rep.ncu-rep kernel_regex [min_pct]   (rows >= min_pct of instructions or samples)
   python profiles/sass_hot.py rep.ncu-rep kernel_regex all          (full listing with cumulative shares)"""
import csv, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
mode = sys.argv[3] if len(sys.argv) > 3 else "0.5"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) > 6 and r[0].startswith("0x"): data.append(r)
ii, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
wi = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
we = hdr.index("L1 Wavefronts Shared Excessive") if wi is not None else None
tot_i = sum(int(r[ii] or 0) for r in data); tot_s = sum(int(r[si] or 0) for r in data)
print("kernel %s: %d SASS rows, %d warp-instructions, %d samples" % (kre, len(data), tot_i, tot_s))
ops = collections.Counter()
for r in data:
    op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
    ops[op.split(".")[0]] += int(r[ii] or 0)
print("opcode mix: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_i) for k, v in ops.most_common(24)))
minp = None if mode == "all" else float(mode)
for n, r in enumerate(data):
    i, s = int(r[ii] or 0), int(r[si] or 0)
    if minp is None or i >= tot_i * minp / 100 or s >= tot_s * minp / 100:
        lanes = (int(r[ti] or 0) / i) if i else 0
        print("%4d %-58s inst %5.2f%% smp %5.2f%% lanes %4.1f%s" % (n, r[1].strip()[:58], 100.0 * i / tot_i, 100.0 * s / max(tot_s, 1), lanes,
              ("  shwf %s(+%s)" % (r[wi], r[we])) if wi is not None and r[wi] not in ("0", "") else ""))
