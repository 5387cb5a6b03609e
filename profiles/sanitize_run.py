#!/usr/bin/env python
"""Workload for compute-sanitizer (profiles/sanitize.sh): small batches through every kernel family of the
library, checked against the CPU oracle so a sanitizer-clean run is also a correct one.
   216-water box (SPC, 6 frames) and 1,000-water box (3 frames): open boundary, rectangular PBC (all-pairs kernel
   and cell sweep), triclinic sweep, CSR16/CSR12 layouts, per-frame site subsets; wn_components on both paths
   (one CTA per frame in shared memory / multi-kernel global memory); wn_find_pairs all-pairs and cell-list;
   wn_format_edges."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import wn_pkg                      # noqa: E402
import wn_oracle_py as oracle      # noqa: E402

capi = wn_pkg.load_capi()
ctx = capi.Context(0)


def same(got, want):
    assert len(got) == len(want), (len(got), len(want))
    for k in ("i", "j"):
        assert np.array_equal(got[k], want[k])
    for k in ("length", "cosine"):
        assert np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32))
    assert np.allclose(got["energy"], want["energy"], rtol=1e-6, atol=0)


for n_waters, n_frames, model in ((216, 6, oracle.MODEL_SPC), (1000, 3, oracle.MODEL_TIP3P)):
    box = oracle.synth_box(n_waters, model=model)
    frames = oracle.synth_frames(box, 0, n_frames)
    sa, ha, ty = oracle.water_site_table(n_waters)
    ctx.set_water_sites(n_waters)
    res = ctx.hbond_batch(frames)
    fo = res["frame_offsets"]
    for f in range(n_frames):
        want = oracle.frame_hbonds(frames[f], sa, ha, ty)[0]
        same(res["edges"][int(fo[f]):int(fo[f + 1])], want)
    # components: shared-memory path, then the global-memory path
    for path in (None, "global"):
        if path:
            os.environ["WN_COMPONENTS_PATH"] = path
        else:
            os.environ.pop("WN_COMPONENTS_PATH", None)
        labels, cstats = ctx.components(0.5)
        for f in range(n_frames):
            lab, st = oracle.components(res["edges"][int(fo[f]):int(fo[f + 1])], n_waters, 0.5)
            assert np.array_equal(labels[f], lab) and int(cstats["n_components"][f]) == int(st[0])
    os.environ.pop("WN_COMPONENTS_PATH", None)
    text = ctx.format_edges(np.arange(n_frames))
    assert text.count(b"\n") == res["n_edges"] + n_frames
    # compact layouts
    for layout in (capi.LAYOUT_CSR16, capi.LAYOUT_CSR12):
        r2 = ctx.hbond_batch_ex(frames, layout=layout)
        assert r2["n_edges"] == res["n_edges"]
    # rectangular periodic box of the generator (216 waters: < 3 cells per dimension -> all-pairs kernel)
    L = float(box.box)
    boxes = np.tile(np.array([L, L, L], np.float32), (n_frames, 1))
    rp = ctx.hbond_batch(frames, boxes=boxes)
    fop = rp["frame_offsets"]
    for f in range(n_frames):
        want = oracle.frame_hbonds_pbc(frames[f], sa, ha, ty, boxes[f])[0]
        same(rp["edges"][int(fop[f]):int(fop[f + 1])], want)
    # triclinic box (cell sweep for the larger system)
    b9 = np.array([L, 0, 0, 0.2 * L, L, 0, -0.1 * L, 0.15 * L, L], np.float32)
    rt = ctx.hbond_batch(frames[:2], boxes=np.tile(b9, (2, 1)))
    fot = rt["frame_offsets"]
    for f in range(2):
        want = oracle.frame_hbonds_tric(frames[f], sa, ha, ty, b9)[0]
        same(rt["edges"][int(fot[f]):int(fot[f + 1])], want)
    # standalone pair search: all-pairs tile kernel and cell-list sweep
    xyz = frames[0][::3].astype(np.float64)
    want_p = oracle.brute_find_pairs(xyz)
    for mode in (1, 2):
        ctx.set_find_mode(mode)
        got_p, _ = ctx.find_pairs(xyz)
        assert np.array_equal(got_p, want_p)
    ctx.set_find_mode(0)
    print("sanitize_run: %d waters x %d frames ok (%d edges, %d pairs)" % (n_waters, n_frames, res["n_edges"], len(want_p)))

# device self-test of the shared-reciprocal division (a kernel of its own)
assert ctx.selftest_div(1 << 16, 7) == 0
ctx.close()
print("sanitize_run: done")
