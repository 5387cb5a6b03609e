#!/usr/bin/env python
"""Finish-stage time of the two result layouts on device-resident frames (kernel experiments):
   python profiles/layout_times.py [n_waters] [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wn_pkg
capi = wn_pkg.load_capi()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 33333
F = int(sys.argv[2]) if len(sys.argv) > 2 else 250
ctx = capi.Context(0)
ctx.set_water_sites(n); ctx.set_batch_frames(F)
fb = 3 * n * 12
dev = ctx.dev_alloc(4 * F * fb)
for s in range(4):
    ctx.synth_frames_dev(n, s * F, F, dev + s * F * fb, seed=1234)
for layout in (capi.LAYOUT_EDGES, capi.LAYOUT_CSR16, capi.LAYOUT_EDGES, capi.LAYOUT_CSR16):
    tot = {}
    for s in range(4):
        ctx.hbond_batch_ex(dev + s * F * fb, F, 3 * n, on_device=True, layout=layout)
        tm = ctx.timings()
        if s:
            for k in ("bin_ms", "edges_ms", "finish_ms", "total_ms"):
                tot[k] = tot.get(k, 0.0) + tm[k] / 3
    print("layout", layout, {k: round(v, 3) for k, v in tot.items()})
