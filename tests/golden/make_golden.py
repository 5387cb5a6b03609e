#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE's own sources.

Runs only where /root/reference exists: oracle/Makefile target `ref` compiles
src/brute-find-pairs.cpp and the computeEnergy / make_edges text of
src/waternetwork.cpp in place (against the CGAL arithmetic shim) into
oracle/_ref/libwn_ref.so; this script calls that library on seeded inputs and
stores inputs + outputs.  The fixtures travel to the GPU box; /root/reference and
this generation step do not.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import wn_oracle_py as o  # noqa: E402


def main():
    o.build_oracle()
    R = o.ref()
    assert R is not None, "oracle/_ref/libwn_ref.so missing: needs /root/reference"

    # 1. computeEnergy / switch known answers on float-representable inputs
    rs = np.array([2.0, 2.5, np.float32(2.8), 3.0, 3.5, 4.0, 5.0, 5.4999995, 5.5, 5.5000005, 5.75, 6.0, 6.25,
                   np.float32(6.49), 6.4999995, 6.5, 7.0], dtype=np.float64)
    cs = np.array([1.0, 0.5, 0.25, np.float32(0.1), 1e-3, 0.0, -0.0, -0.5, -1.0, np.nan], dtype=np.float64)
    grid_r, grid_c = np.meshgrid(rs, cs, indexing="ij")
    e = np.array([[R.wnref_compute_energy(float(r), float(c)) for c in cs] for r in rs])
    sr = np.array([R.wnref_switch_function_radius(float(r * r), 5.5 * 5.5, 6.5 * 6.5) for r in rs])
    sa_swapped = np.array([R.wnref_switch_function_angle(float(c * c), 0.0, 0.25) for c in cs])
    sa_plain = np.array([R.wnref_switch_function_angle(float(c), 0.25, 0.0301) for c in np.linspace(0, 1, 41)])
    np.savez(os.path.join(HERE, "energy_kat.npz"), r=rs, cos=cs, energy=e, switch_radius=sr,
             switch_angle_swapped=sa_swapped, switch_angle_plain=sa_plain)

    # 2. whole-frame edge lists on small synthetic boxes (SPC 64 and TIP3P 125 waters)
    for name, n, model, frame0, nf in (("spc64", 64, o.MODEL_SPC, 0, 3), ("tip3p125", 125, o.MODEL_TIP3P, 11, 2)):
        box = o.synth_box(n, model=model)
        frames = o.synth_frames(box, frame0, nf)
        edges, offs, npairs = [], [0], []
        for f in range(nf):
            ed, npr = o.ref_frame_hbonds_water(frames[f])
            edges.append(ed); offs.append(offs[-1] + len(ed)); npairs.append(npr)
        np.savez(os.path.join(HERE, "frames_%s.npz" % name), frames=frames, edges=np.concatenate(edges),
                 frame_offsets=np.array(offs, np.int64), n_pairs=np.array(npairs, np.int64))

    # 3. brute pair list + make_edges with mixed site types / hydrogen lists
    rng = np.random.default_rng(20261019)
    n_sites, hmax = 90, 3
    pos = rng.uniform(0, 1.6, size=(n_sites, 3)).astype(np.float32)
    nh = rng.integers(0, 4, size=n_sites)
    natoms = 4 * n_sites
    frame = np.zeros((natoms, 3), np.float32)
    frame[0::4] = pos
    for k in range(3):
        d = rng.normal(size=(n_sites, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
        frame[1 + k::4] = (pos + 0.1 * d).astype(np.float32)
    site_atom = (np.arange(n_sites) * 4).astype(np.int32)
    h_atom = -np.ones((n_sites, hmax), np.int32)
    for s in range(n_sites):
        for k in range(nh[s]):
            h_atom[s, k] = site_atom[s] + 1 + k
    for s in range(0, n_sites, 6):      # water-style lists [site atom itself, H1]
        h_atom[s] = [site_atom[s], site_atom[s] + 1, -1]
    site_type = rng.integers(0, 3, size=n_sites).astype(np.uint8)
    site_xyz, h_xyz, h_off = o.gather_sites(frame, site_atom, h_atom, hmax)
    pairs = o.ref_brute_find_pairs(site_xyz)
    ed = o.ref_make_edges(pairs, site_xyz, h_xyz, h_off, site_type)
    np.savez(os.path.join(HERE, "mixed_sites.npz"), frame=frame, site_atom=site_atom, h_atom=h_atom,
             site_type=site_type, pairs=pairs, edges=ed)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
