#!/usr/bin/env python
"""Standalone pair search at the 100k-atom box (33,333 oxygens) for profiling: a few calls of wn_find_pairs."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import wn_pkg, wn_oracle_py as oracle
capi = wn_pkg.load_capi()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 33333
ctx = capi.Context(0)
xyz = oracle.synth_frames(oracle.synth_box(n), 0, 1)[0][::3].astype(np.float64)
for mode in (0, 1) if n <= 60000 else (0,):
    ctx.set_find_mode(mode)
    ctx.find_pairs(xyz)
    t0 = time.perf_counter()
    for _ in range(3):
        p, _t = ctx.find_pairs(xyz)
    print("mode", mode, "n", n, "pairs", len(p), "kernels_ms", ctx.timings()["total_ms"], "d2h_ms", ctx.timings()["d2h_ms"],
          "wall_ms", (time.perf_counter() - t0) / 3 * 1e3)
ctx.close()
