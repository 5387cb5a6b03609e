#!/usr/bin/env python
"""wn_components on frames too large for the shared-memory kernel (global path): python profiles/cc_prof_big.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, wn_pkg
capi = wn_pkg.load_capi()
n, F = 333333, 24
ctx = capi.Context(0); ctx.set_water_sites(n); ctx.set_batch_frames(F)
dev = ctx.dev_alloc(F * 3 * n * 12)
ctx.synth_frames_dev(n, 0, F, dev, seed=1234)
ctx.hbond_batch_ex(dev, F, 3 * n, on_device=True)
lab, st = ctypes.c_void_p(), ctypes.c_void_p()
for thr in (0.0, 0.0, 1.0):
    ctx.timer_start()
    ctx._ck(ctx.L.wn_components(ctx.h, thr, ctypes.byref(lab), ctypes.byref(st)))
    ms = ctx.timer_stop()
    s = np.zeros(F, capi.COMPONENT_DTYPE); ctx.d2h(s, st.value)
    print("thr", thr, "ms %.3f" % ms, s[0])
