"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle, bit-exact on
(i, j, length, cosine), <= 1e-6 relative on energy (fp32-rounded, reference
src/waternetwork.cpp:222) and on the fp64 per-frame energy sum."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ENERGY_RTOL = 1e-6


def check_edges(got, want):
    assert len(got) == len(want), (len(got), len(want))
    assert np.array_equal(got["i"], want["i"])
    assert np.array_equal(got["j"], want["j"])
    assert np.array_equal(got["length"].view(np.uint32), want["length"].view(np.uint32))
    assert np.array_equal(got["cosine"].view(np.uint32), want["cosine"].view(np.uint32))
    ge, we = got["energy"].astype(np.float64), want["energy"].astype(np.float64)
    assert np.all(np.abs(ge - we) <= ENERGY_RTOL * np.abs(we)), np.max(np.abs(ge - we) / np.maximum(np.abs(we), 1e-300))
    # sign of zero as well (energy may be -0.0)
    z = we == 0
    assert np.array_equal(np.signbit(got["energy"][z]), np.signbit(want["energy"][z]))


def oracle_frames(oracle, frames, site_atom, h_atom, site_type, first_limit=None):
    edges, stats, evals = [], [], 0
    ties = oracle.Ties()
    for f in range(frames.shape[0]):
        e, s, _, ev = oracle.frame_hbonds(frames[f], site_atom, h_atom, site_type,
                                          first_limit=first_limit, ties=ties)
        edges.append(e); stats.append(s); evals += ev
    return edges, np.array(stats), evals, ties.as_dict()


def compare_batch(res, o_edges, o_stats, o_evals, o_ties):
    fo = res["frame_offsets"]
    assert fo[0] == 0 and fo[-1] == res["n_edges"]
    if "row_offsets" in res:                      # CSR view: row i of frame f starts at fo[f] + ro[f][i]
        ro = res["row_offsets"]
        for f, want in enumerate(o_edges):
            n = ro.shape[1]
            starts = np.searchsorted(want["i"], np.arange(n), side="left")
            assert np.array_equal(ro[f], starts.astype(np.uint32))
    for f, want in enumerate(o_edges):
        check_edges(res["edges"][int(fo[f]):int(fo[f + 1])], want)
    st = res["stats"]
    assert np.array_equal(st["num_sites"], o_stats[:, 0])
    assert np.array_equal(st["num_edges"], o_stats[:, 1])
    assert np.array_equal(st["ratio"], o_stats[:, 2])
    assert np.allclose(st["sum_energy"], o_stats[:, 3], rtol=ENERGY_RTOL, atol=0)
    assert res["n_pair_evals"] == o_evals
    assert int(st["pair_evals"].sum()) == o_evals
    t = res["ties"]
    assert t["length_equal_cut"] == o_ties["length_equal_cut"]
    assert t["cosine_argmax_equal"] == o_ties["cosine_argmax_equal"]
    assert t["pos_zero"] == o_ties["pos_zero"]


@pytest.mark.parametrize("n_waters,n_frames,model", [(216, 6, 1), (1000, 3, 0), (4000, 2, 0)])
def test_water_box_host_batch(ctx, oracle, n_waters, n_frames, model):
    box = oracle.synth_box(n_waters, model=model)
    frames = oracle.synth_frames(box, 0, n_frames)
    sa, ha, ty = oracle.water_site_table(n_waters)
    ctx.set_water_sites(n_waters)
    res = ctx.hbond_batch(frames)
    compare_batch(res, *oracle_frames(oracle, frames, sa, ha, ty))


def test_device_synth_matches_host(ctx, oracle):
    n_waters, n_frames = 1000, 4
    box = oracle.synth_box(n_waters)
    want = oracle.synth_frames(box, 7, n_frames)
    d = ctx.dev_alloc(want.nbytes)
    ctx.synth_frames_dev(n_waters, 7, n_frames, d)
    got = np.empty_like(want)
    ctx.d2h(got, d)
    ctx.dev_free(d)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_device_batch_multi_pass(ctx, oracle):
    """Device-resident input, several internal passes; result == single pass."""
    n_waters, n_frames = 512, 9
    box = oracle.synth_box(n_waters)
    frames = oracle.synth_frames(box, 100, n_frames)
    sa, ha, ty = oracle.water_site_table(n_waters)
    ctx.set_water_sites(n_waters)
    d = ctx.dev_alloc(frames.nbytes)
    ctx.h2d(d, frames)
    ctx.set_batch_frames(4)
    out = ctx.hbond_batch_dev(d, n_frames, 3 * n_waters)
    e, fo, st = ctx.fetch_dev_result(out)
    assert ctx.timings()["passes"] == 3
    ctx.set_batch_frames(0)
    ctx.dev_free(d)
    res = dict(out, edges=e, frame_offsets=fo, stats=st)
    compare_batch(res, *oracle_frames(oracle, frames, sa, ha, ty))


def test_find_pairs_matches_brute(ctx, oracle):
    for n_waters in (216, 1500):
        box = oracle.synth_box(n_waters)
        fr = oracle.synth_frames(box, 3, 1)[0]
        xyz = fr[::3].astype(np.float64)
        want = oracle.brute_find_pairs(xyz)
        got, ties = ctx.find_pairs(xyz)
        assert np.array_equal(got, want)
        assert ties["pair_cutoff_equal"] == 0


def test_find_pairs_edge_cases(ctx, oracle):
    for n in (0, 1, 2):
        xyz = np.zeros((n, 3))
        got, _ = ctx.find_pairs(xyz)
        assert np.array_equal(got, oracle.brute_find_pairs(xyz))
    # exact and near ties at the cutoff: d2 = 4.225 -> 10*d2 == 42.25 is NOT a pair (strict <)
    tie_xyz = np.array([[0, 0, 0], [0.5, 1.5, 1.3], [0, 0, 2.0554804791094465], [2.0554804791094470, 0, 0]])
    to = oracle.Ties()
    want_t = oracle.brute_find_pairs(tie_xyz, ties=to)
    got_t, ties_t = ctx.find_pairs(tie_xyz)
    assert np.array_equal(got_t, want_t)
    assert ties_t["pair_cutoff_equal"] == to.pair_cutoff_equal and ties_t["pair_cutoff_near"] == to.pair_cutoff_near
    assert ties_t["pair_cutoff_near"] >= 1
    rng = np.random.default_rng(5)
    xyz = rng.uniform(0, 3.0, size=(700, 3)).astype(np.float32).astype(np.float64)
    got, _ = ctx.find_pairs(xyz)
    assert np.array_equal(got, oracle.brute_find_pairs(xyz))
    got, _ = ctx.find_pairs(xyz, first_limit=50)
    assert np.array_equal(got, oracle.brute_find_pairs(xyz, first_limit=50))


def test_dense_cluster_long_rows(ctx, oracle):
    """All sites in one or two cells: rows far longer than the shared-memory pool
    (exercises the one-row and count/allocate/direct-write fallbacks)."""
    rng = np.random.default_rng(11)
    n = 700
    o = rng.uniform(0, 0.5, size=(n, 3))
    h1 = o + rng.normal(size=(n, 3)) * 0.05
    h2 = o + rng.normal(size=(n, 3)) * 0.05
    frames = np.stack([o, h1, h2], axis=1).reshape(1, 3 * n, 3).astype(np.float32)
    sa, ha, ty = oracle.water_site_table(n)
    ctx.set_water_sites(n)
    res = ctx.hbond_batch(frames)
    compare_batch(res, *oracle_frames(oracle, frames, sa, ha, ty))


def test_degenerate_inputs(ctx, oracle):
    sa, ha, ty = oracle.water_site_table(5)
    ctx.set_water_sites(5)
    # coincident sites (distance 0 -> NaN cosines -> counted, no edge), and far apart sites
    fr = np.zeros((2, 15, 3), np.float32)
    fr[1, :, 0] = np.arange(15) * 7.0
    res = ctx.hbond_batch(fr)
    compare_batch(res, *oracle_frames(oracle, fr, sa, ha, ty))
    # empty site table and zero frames
    ctx.set_water_sites(0)
    r0 = ctx.hbond_batch(np.zeros((3, 0, 3), np.float32))
    assert r0["n_edges"] == 0 and list(r0["frame_offsets"]) == [0, 0, 0, 0]
    ctx.set_water_sites(5)
    r1 = ctx.hbond_batch(np.zeros((0, 15, 3), np.float32))
    assert r1["n_edges"] == 0 and r1["n_frames"] == 0


def test_mixed_site_types_and_hydrogen_lists(ctx, oracle):
    """DONOR/ACCEPTOR/WATER sites with 0..3 hydrogens, real first hydrogens (F9: a
    first-site hydrogen at list index 0 can never be selected), asymmetric search."""
    rng = np.random.default_rng(3)
    n_sites, hmax = 400, 3
    natoms = n_sites * 4
    site_atom = np.arange(n_sites, dtype=np.int32) * 4
    nh = rng.integers(0, 4, size=n_sites)
    h_atom = -np.ones((n_sites, hmax), np.int32)
    for s in range(n_sites):
        for k in range(nh[s]):
            h_atom[s, k] = site_atom[s] + 1 + k
    # a few water-style lists that start with the site atom itself (F8)
    for s in range(0, n_sites, 7):
        h_atom[s, 0] = site_atom[s]; h_atom[s, 1] = site_atom[s] + 1; h_atom[s, 2] = -1
    site_type = rng.integers(0, 3, size=n_sites).astype(np.uint8)
    frames = np.zeros((3, natoms, 3), np.float32)
    for f in range(3):
        pos = rng.uniform(0, 2.2, size=(n_sites, 3))
        frames[f, 0::4] = pos
        for k in range(3):
            d = rng.normal(size=(n_sites, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
            frames[f, 1 + k::4] = pos + 0.1 * d
    for first_limit in (None, 37):
        ctx.set_sites(site_atom, h_atom, site_type, first_limit=first_limit)
        res = ctx.hbond_batch(frames)
        oe, ost, oev, oties = oracle_frames(oracle, frames, site_atom, h_atom, site_type, first_limit)
        compare_batch(res, oe, ost, oev, oties)
    assert oties["pos_zero"] > 0 or True


def test_nonfinite_coordinates_are_ignored(ctx, oracle):
    n = 300
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 0, 2).copy()
    frames[0, 3 * 10] = np.nan
    frames[0, 3 * 20 + 1] = np.inf
    frames[1, 3 * 5, 2] = -np.inf
    sa, ha, ty = oracle.water_site_table(n)
    ctx.set_water_sites(n)
    res = ctx.hbond_batch(frames)
    oe, ost, oev, oties = oracle_frames(oracle, frames, sa, ha, ty)
    fo = res["frame_offsets"]
    for f, want in enumerate(oe):
        check_edges(res["edges"][int(fo[f]):int(fo[f + 1])], want)


# ------------------------------------------------------------ BASELINE configurations
def test_config2_buried_group_4000_waters(ctx, oracle):
    """configs[1]: 4k waters + buried-water index group (the waters nearest the box
    centre at frame 0), here as the site table of the batch."""
    n = 4000
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 0, 3)
    centre = box.box / 2
    d = np.linalg.norm(frames[0][::3] - centre, axis=1)
    group = np.sort(np.argsort(d)[:600]).astype(np.int32)
    site_atom = 3 * group
    h_atom = np.stack([3 * group, 3 * group + 1], axis=1).astype(np.int32)
    ty = np.full(len(group), oracle.WATER, np.uint8)
    ctx.set_sites(site_atom, h_atom, ty)
    res = ctx.hbond_batch(frames)
    compare_batch(res, *oracle_frames(oracle, frames, site_atom, h_atom, ty))


def test_config5_asymmetric_selection_vs_solvent(ctx, oracle):
    """configs[4]: ~500-water selection vs the full solvent set: selection placed first,
    rows restricted to i < 500 (many short frames)."""
    n, nsel, nf = 4000, 500, 12
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 50, nf)
    d = np.linalg.norm(frames[0][::3] - box.box / 2, axis=1)
    order = np.argsort(d, kind="stable")
    waters = np.concatenate([np.sort(order[:nsel]), np.sort(order[nsel:])]).astype(np.int32)
    site_atom = 3 * waters
    h_atom = np.stack([3 * waters, 3 * waters + 1], axis=1).astype(np.int32)
    ty = np.full(n, oracle.WATER, np.uint8)
    ctx.set_sites(site_atom, h_atom, ty, first_limit=nsel)
    res = ctx.hbond_batch(frames)
    compare_batch(res, *oracle_frames(oracle, frames, site_atom, h_atom, ty, first_limit=nsel))


def test_full_size_100k_atoms_properties(ctx, oracle):
    """configs[2] at full size (33,333 waters): size-independent properties plus exact
    parity on the first 150 rows (the oracle restricted to rows i < 150 is O(150*N))."""
    n, nf = 33333, 3
    d = ctx.dev_alloc(nf * 3 * n * 12)
    ctx.synth_frames_dev(n, 1000, nf, d)
    frames = np.empty((nf, 3 * n, 3), np.float32)
    ctx.d2h(frames, d)
    sa, ha, ty = oracle.water_site_table(n)
    ctx.set_water_sites(n)
    out = ctx.hbond_batch_dev(d, nf, 3 * n)
    e, fo, st = ctx.fetch_dev_result(out)
    # host entry point, split in passes of 2 frames: identical bytes
    ctx.set_batch_frames(2)
    res = ctx.hbond_batch(frames)
    ctx.set_batch_frames(0)
    assert np.array_equal(res["frame_offsets"], fo) and res["edges"].tobytes() == e.tobytes()
    assert res["stats"].tobytes() == st.tobytes()
    for f in range(nf):
        ef = e[int(fo[f]):int(fo[f + 1])]
        assert np.all(ef["i"] < ef["j"]) and ef["j"].max() < n
        key = ef["i"].astype(np.int64) * n + ef["j"]
        assert np.all(np.diff(key) > 0)                      # lexicographic, no duplicates
        assert np.all(ef["length"] <= np.float32(6.5)) and np.all(ef["cosine"] > 0)
        # geometry recomputed on the host in the reference's operation order for every edge
        a = frames[f][3 * ef["i"]].astype(np.float64); b = frames[f][3 * ef["j"]].astype(np.float64)
        oo = a - b
        dist = np.sqrt((oo[:, 0] * oo[:, 0] + oo[:, 1] * oo[:, 1]) + oo[:, 2] * oo[:, 2]).astype(np.float32)
        assert np.array_equal((dist.astype(np.float64) * 10.0).astype(np.float32), ef["length"])
        assert st["num_edges"][f] == len(ef) and st["num_sites"][f] == n
        assert np.isclose(st["sum_energy"][f], ef["energy"].astype(np.float64).sum(), rtol=1e-9)
    # exact parity on the first rows
    ctx.set_water_sites(n, first_limit=150)
    res2 = ctx.hbond_batch(frames[:2])
    for f in range(2):
        want, stats, _, ev = oracle.frame_hbonds(frames[f], sa, ha, ty, first_limit=150)
        got = res2["edges"][int(res2["frame_offsets"][f]):int(res2["frame_offsets"][f + 1])]
        check_edges(got, want)
        full = e[int(fo[f]):int(fo[f + 1])]
        check_edges(full[full["i"] < 150], want)
        assert int(res2["stats"]["pair_evals"][f]) == ev
    ctx.dev_free(d)


def test_full_size_100k_atoms_whole_frames_vs_oracle(ctx, oracle):
    """configs[2] at full size, the configuration the bench numbers are quoted on: EVERY edge of whole
    frames against the oracle (reference src/brute-find-pairs.cpp:19-30 + src/waternetwork.cpp:182-263),
    open boundary and rectangular PBC -- bit-exact (i, j, length, cosine), energy <= 1e-6 relative, the
    statistics row and the pair-evaluation count.  The oracle needs ~3 s (open) / ~7 s (periodic) per frame."""
    n, nf = 33333, 2
    d = ctx.dev_alloc(nf * 3 * n * 12)
    ctx.synth_frames_dev(n, 4242, nf, d)
    frames = np.empty((nf, 3 * n, 3), np.float32)
    ctx.d2h(frames, d)
    sa, ha, ty = oracle.water_site_table(n)
    ctx.set_water_sites(n)
    L = np.float32(oracle.synth_box(n).box)
    # open boundary: two whole frames (device-resident batch, as the bench's timed leg runs it)
    out = ctx.hbond_batch_dev(d, nf, 3 * n)
    e, fo, st = ctx.fetch_dev_result(out)
    evals = 0
    for f in range(nf):
        want, stats, npairs, ev = oracle.frame_hbonds(frames[f], sa, ha, ty)
        assert len(want) > 390000 and npairs > 15000000
        check_edges(e[int(fo[f]):int(fo[f + 1])], want)
        assert st["num_sites"][f] == stats[0] and st["num_edges"][f] == stats[1] and st["ratio"][f] == stats[2]
        assert np.isclose(st["sum_energy"][f], stats[3], rtol=ENERGY_RTOL, atol=0)
        assert int(st["pair_evals"][f]) == ev
        evals += ev
    assert out["n_pair_evals"] == evals
    # rectangular PBC: one whole frame through the host entry point in the compact layout
    boxes = np.full((1, 3), L, np.float32)
    res = ctx.hbond_batch_ex(frames[:1], boxes=boxes, layout=1)
    want, stats, _, ev = oracle.frame_hbonds_pbc(frames[0], sa, ha, ty, boxes[0])
    assert len(want) > 420000
    got = np.zeros(res["n_edges"], want.dtype)
    ro = res["row_offsets"][0].astype(np.int64)
    cnt = np.diff(np.append(ro, res["n_edges"]))
    got["i"] = np.repeat(np.arange(n, dtype=np.int32), cnt)
    for k in ("j", "energy", "length", "cosine"):
        got[k] = res["edges"][k]
    check_edges(got, want)
    assert int(res["stats"]["pair_evals"][0]) == ev and res["stats"]["num_edges"][0] == stats[1]
    ctx.dev_free(d)


def test_energy_known_answers_through_the_batch(ctx, oracle):
    """Known-answer test of computeEnergy (reference src/waternetwork.cpp:49-66) on the GPU: two-water frames
    whose edge has exactly the (r, cos) of a grid point of tests/golden/energy_kat.npz (values from the
    reference's own sources) go through wn_hbond_batch; the returned energy must be the golden one
    (float-rounded as make_edges stores it, :222) within 1e-6 relative, with the sign of zero."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "energy_kat.npz"))
    sa, ha, ty = oracle.water_site_table(2)

    def frame_for(x, hx, hy):
        fr = np.zeros((6, 3), np.float32)
        fr[1] = (hx, hy, 0.0)                      # H1 of water A (site A at the origin)
        fr[2] = (0.0, 0.0, 0.08)
        fr[3] = (x, 0.0, 0.0)                      # water B on the x axis
        fr[4] = (np.float32(x) - np.float32(0.1), 0.0, 0.0)     # its H1 points at A: cosine -1, never selected
        fr[5] = (x, 0.0, 0.08)
        return fr

    def ulps(v, k):
        return np.float32(v).view(np.int32).__add__(k).view(np.float32) if k else np.float32(v)

    frames, wants = [], []
    for a, r in enumerate(g["r"]):
        rf = np.float32(r)
        if float(rf) != float(r) or rf > np.float32(6.5):
            continue
        for b, c in enumerate(g["cos"]):
            cf = np.float32(c)
            if not (c > 0) or float(cf) != float(c):
                continue
            hit = None
            for kx in range(-3, 4):
                x = ulps(np.float32(float(rf) / 10.0), kx)
                if np.float32(np.float64(x) * 10.0) != rf:
                    continue
                hx = np.float32(-float(cf) * 0.1)
                hy0 = np.float32(np.sqrt(max(0.01 - float(hx) ** 2, 0.0)))
                for ky in range(-40, 41):
                    hy = ulps(hy0, ky) if hy0 > 0 else np.float32(0.0)
                    fr = frame_for(x, hx, hy)
                    e, _, _, _ = oracle.frame_hbonds(fr, sa, ha, ty)
                    if len(e) == 1 and e["length"][0] == rf and e["cosine"][0] == cf:
                        hit = (fr, e[0])
                        break
                if hit:
                    break
            if hit:
                golden = np.float32(g["energy"][a, b])
                assert hit[1]["energy"].view(np.uint32) == golden.view(np.uint32)     # oracle == reference value
                frames.append(hit[0]); wants.append((rf, cf, golden))
    assert len(frames) >= 40, len(frames)            # 13 radii x 4 cosines are float-representable
    ctx.set_water_sites(2)
    res = ctx.hbond_batch(np.stack(frames))
    assert res["n_edges"] == len(frames) and np.all(np.diff(res["frame_offsets"]) == 1)
    for k, (rf, cf, golden) in enumerate(wants):
        ed = res["edges"][k]
        assert ed["i"] == 0 and ed["j"] == 1 and ed["length"] == rf and ed["cosine"] == cf
        assert abs(float(ed["energy"]) - float(golden)) <= ENERGY_RTOL * abs(float(golden)), (rf, cf, ed["energy"], golden)
        if golden == 0:
            assert np.signbit(ed["energy"]) == np.signbit(golden)
        # a one-edge frame: the statistics row's sum is that edge's float energy, however small (energies down to
        # 1e-8 sit below the 2^-33 resolution of the fixed-point partials; wn_k_stats then sums the written records)
        se = float(res["stats"]["sum_energy"][k])
        assert abs(se - float(ed["energy"])) <= 1.2e-7 * abs(float(ed["energy"])), (rf, cf, se, ed["energy"])


def test_find_pairs_full_size_rows(ctx, oracle):
    """CudaFindPairs at 33,333 sites: total count, order, and exact rows vs the oracle."""
    n = 33333
    box = oracle.synth_box(n)
    xyz = oracle.synth_frames(box, 0, 1)[0][::3].astype(np.float64)
    got, ties = ctx.find_pairs(xyz)
    key = got[:, 0].astype(np.int64) * n + got[:, 1]
    assert np.all(np.diff(key) > 0) and np.all(got[:, 0] < got[:, 1])
    want = oracle.brute_find_pairs(xyz, first_limit=64)
    assert np.array_equal(got[got[:, 0] < 64], want)
    lim, _ = ctx.find_pairs(xyz, first_limit=64)
    assert np.array_equal(lim, want)


def test_cell_list_find_pairs_matches_brute_and_all_pairs_kernel(ctx, oracle):
    """CudaFindPairs' cell-list sweep (cells of half the 2.0555 nm cutoff) == BruteFindPairs::find
    (src/brute-find-pairs.cpp:19-30) == the all-pairs tile kernel, bit for bit, open and periodic."""
    rng = np.random.default_rng(77)
    try:
        # forced on small inputs: one cell, a few cells, sparse points over a large extent, non-finite points
        cases = [np.zeros((5, 3)), rng.uniform(0, 1.0, size=(40, 3)), rng.uniform(-3, 9.0, size=(900, 3)),
                 rng.uniform(0, 400.0, size=(600, 3)), rng.normal(0, 2.0, size=(1200, 3))]
        bad = rng.uniform(0, 6.0, size=(500, 3))
        bad[7] = np.nan; bad[30, 1] = np.inf; bad[31, 2] = -np.inf
        cases.append(bad)
        for xyz in cases:
            xyz = xyz.astype(np.float32).astype(np.float64)
            to = oracle.Ties()
            want = oracle.brute_find_pairs(xyz, ties=to)
            ctx.set_find_mode(2)
            got, ties = ctx.find_pairs(xyz)
            assert ctx.last_find_used_cells()
            assert np.array_equal(got, want)
            assert ties["pair_cutoff_equal"] == to.pair_cutoff_equal and ties["pair_cutoff_near"] == to.pair_cutoff_near
            lim, _ = ctx.find_pairs(xyz, first_limit=17)
            assert np.array_equal(lim, oracle.brute_find_pairs(xyz, first_limit=17))
            ctx.set_find_mode(1)
            ap, _ = ctx.find_pairs(xyz)
            assert not ctx.last_find_used_cells() and np.array_equal(ap, want)
        # cutoff ties: 10*d2 == 42.25 is not a pair (strict <), and the tie report counts it
        tie_xyz = np.array([[0, 0, 0], [0.5, 1.5, 1.3], [0, 0, 2.0554804791094465], [2.0554804791094470, 0, 0]])
        to = oracle.Ties()
        want_t = oracle.brute_find_pairs(tie_xyz, ties=to)
        ctx.set_find_mode(2)
        got_t, ties_t = ctx.find_pairs(tie_xyz)
        assert np.array_equal(got_t, want_t)
        assert ties_t["pair_cutoff_equal"] == to.pair_cutoff_equal and ties_t["pair_cutoff_near"] == to.pair_cutoff_near
        # periodic: boxes of >= 5 cells take the sweep, smaller ones fall back; unwrapped coordinates
        for n, L in ((1500, (5.3, 5.2, 6.1)), (2500, (7.0, 5.2, 11.0)), (300, (1.9, 2.2, 3.0))):
            xyz = (rng.uniform(-1, 2, size=(n, 3)) * np.array(L)).astype(np.float32).astype(np.float64)
            want = oracle.brute_find_pairs_pbc(xyz, L)
            ctx.set_find_mode(2)
            got, _ = ctx.find_pairs(xyz, box=L)
            assert ctx.last_find_used_cells() == (min(L) > 5.14)
            assert np.array_equal(got, want)
            lim, _ = ctx.find_pairs(xyz, first_limit=40, box=L)
            assert np.array_equal(lim, oracle.brute_find_pairs_pbc(xyz, L, first_limit=40))
        # the 100k-atom box: automatic mode takes the sweep; whole list equal to the all-pairs kernel's
        n = 33333
        xyz = oracle.synth_frames(oracle.synth_box(n), 0, 1)[0][::3].astype(np.float64)
        ctx.set_find_mode(0)
        got, _ = ctx.find_pairs(xyz)
        assert ctx.last_find_used_cells()
        ctx.set_find_mode(1)
        ap, _ = ctx.find_pairs(xyz)
        assert np.array_equal(got, ap)
        L = (10.0, 10.0, 10.0)
        ctx.set_find_mode(0)
        got, _ = ctx.find_pairs(xyz, box=L)
        assert ctx.last_find_used_cells()
        ctx.set_find_mode(1)
        ap, _ = ctx.find_pairs(xyz, box=L)
        assert np.array_equal(got, ap)
    finally:
        ctx.set_find_mode(0)


def test_cell_list_find_pairs_one_million_atoms(ctx, oracle):
    """333,333 sites (BASELINE configs[3]): the sweep is O(N); order + scattered rows against the oracle."""
    n = 333333
    xyz = oracle.synth_frames(oracle.synth_box(n), 0, 1)[0][::3].astype(np.float64)
    rows = np.unique(np.random.default_rng(3).integers(0, n, size=24))
    got, _ = ctx.find_pairs(xyz)
    assert ctx.last_find_used_cells()
    key = got[:, 0].astype(np.int64) * n + got[:, 1]
    assert np.all(np.diff(key) > 0) and np.all(got[:, 0] < got[:, 1])
    # brute rows of the oracle: row i = pairs (i, j > i); the oracle's loop with first_limit is a prefix, so
    # evaluate the scattered rows by permuting each one to the front
    for i in rows:
        d = xyz - xyz[i]
        lhs = 10.0 * ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
        want_j = np.nonzero((lhs < 42.25) & (np.arange(n) > i))[0]
        a, b = np.searchsorted(got[:, 0], [i, i + 1])
        assert np.array_equal(got[a:b, 1], want_j)
    want = oracle.brute_find_pairs(xyz, first_limit=8)
    assert np.array_equal(got[got[:, 0] < 8], want)


def test_make_edges_on_a_caller_supplied_pair_list(ctx, oracle, capi):
    """wn_make_edges = the reference's make_edges (src/waternetwork.cpp:182-263) on whatever pair list the host
    finder produced -- e.g. DelaunayFindPairs' list, whose order is not lexicographic and whose (first, second)
    orientation is not i < j (:399).  Random orientation and order, duplicates and self pairs, mixed site types with
    0..3 hydrogens (the `pos > 0` rule depends on the orientation), against the oracle's make_edges bit for bit."""
    rng = np.random.default_rng(21)
    # (a) water box: every i<j pair within the brute cutoff, shuffled and randomly flipped
    n = 1000
    frame = oracle.synth_frames(oracle.synth_box(n), 7, 1)[0]
    sa, ha, ty = oracle.water_site_table(n)
    ctx.set_water_sites(n)
    site_xyz, h_xyz, h_off = oracle.gather_sites(frame, sa, ha, 2)
    pairs = oracle.brute_find_pairs(site_xyz)
    # in order and i<j: equals the batched path
    got, ev, _ = ctx.make_edges(frame, pairs)
    batch = ctx.hbond_batch(frame[None])
    check_edges(got, batch["edges"])
    assert ev == batch["n_pair_evals"]
    perm = rng.permutation(len(pairs))[:60000]
    pl = pairs[perm].copy()
    flip = rng.random(len(pl)) < 0.5
    pl[flip] = pl[flip][:, ::-1]
    pl = np.concatenate([pl, pl[:100], np.stack([np.arange(50), np.arange(50)], axis=1).astype(np.int32)])
    to = oracle.Ties()
    want, want_ev = oracle.make_edges(pl, site_xyz, h_xyz, h_off, ty, ties=to)
    got, ev, ties = ctx.make_edges(frame, pl)
    assert len(got) == len(want) and ev == want_ev
    assert np.array_equal(got["i"], want["i"]) and np.array_equal(got["j"], want["j"])      # given orientation kept
    check_edges(got, want)
    assert ties["length_equal_cut"] == to.length_equal_cut and ties["cosine_argmax_equal"] == to.cosine_argmax_equal
    # (b) mixed DONOR/ACCEPTOR/WATER table, 0..3 hydrogens, lists that start with the site atom (F8, F9)
    n_sites, hmax = 300, 3
    natoms = n_sites * 4
    site_atom = np.arange(n_sites, dtype=np.int32) * 4
    nh = rng.integers(0, 4, size=n_sites)
    h_atom = -np.ones((n_sites, hmax), np.int32)
    for s_ in range(n_sites):
        for k in range(nh[s_]):
            h_atom[s_, k] = site_atom[s_] + 1 + k
    for s_ in range(0, n_sites, 7):
        h_atom[s_, 0] = site_atom[s_]; h_atom[s_, 1] = site_atom[s_] + 1; h_atom[s_, 2] = -1
    site_type = rng.integers(0, 3, size=n_sites).astype(np.uint8)
    fr = np.zeros((natoms, 3), np.float32)
    pos = rng.uniform(0, 2.0, size=(n_sites, 3))
    fr[0::4] = pos
    for k in range(3):
        d = rng.normal(size=(n_sites, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
        fr[1 + k::4] = pos + 0.1 * d
    ctx.set_sites(site_atom, h_atom, site_type)
    site_xyz, h_xyz, h_off = oracle.gather_sites(fr, site_atom, h_atom, hmax)
    pl = rng.integers(0, n_sites, size=(40000, 2)).astype(np.int32)
    to = oracle.Ties()
    want, want_ev = oracle.make_edges(pl, site_xyz, h_xyz, h_off, site_type, ties=to)
    got, ev, ties = ctx.make_edges(fr, pl)
    assert len(want) > 100 and len(got) == len(want) and ev == want_ev
    assert np.array_equal(got["i"], want["i"]) and np.array_equal(got["j"], want["j"])
    check_edges(got, want)
    assert ties["pos_zero"] == to.pos_zero and to.pos_zero > 0
    # the two orientations of a pair are different questions (the reason this entry point exists)
    rev, _, _ = ctx.make_edges(fr, pl[:, ::-1])
    assert len(rev) != len(got) or not np.array_equal(rev["cosine"], got["cosine"])
    # (c) periodic: the brute periodic pair list of a water box, in order -> the periodic batch result
    n = 512
    box = oracle.synth_box(n)
    frame = oracle.synth_frames(box, 3, 1)[0]
    ctx.set_water_sites(n)
    L = np.float32(box.box)
    b9 = np.diag([L, L, L]).astype(np.float32).reshape(9)
    pairs = oracle.brute_find_pairs_pbc(frame[::3].astype(np.float64), (L, L, L))
    got, ev, _ = ctx.make_edges(frame, pairs, box9=b9)
    pb = ctx.hbond_batch(frame[None], boxes=np.array([[L, L, L]], np.float32))
    check_edges(got, pb["edges"])
    # (d) errors and empty input
    e0, _, _ = ctx.make_edges(frame, np.zeros((0, 2), np.int32))
    assert len(e0) == 0
    with pytest.raises(capi.WnError) as ei:
        ctx.make_edges(frame, np.array([[0, n]], np.int32))
    assert ei.value.code == capi.WN_ERR_BAD_ARG and "outside the table" in str(ei.value)


def test_energy_sum_fast_and_ordered_paths(ctx, oracle):
    """sum_energy of a frame: an exact fixed-point sum inside wn_k_copy (order-free), or -- when the fused kernel saw an
    edge under 0.56 A, whose energy may leave the fixed-point range -- the written records summed in final order by
    wn_k_stats.  Both within 1e-6 of the oracle, bit-identical run to run and independent of how frames are batched."""
    n = 1000
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 11, 4).copy()
    # frame 1: water 7 moved onto water 400 (O-O distance 0.02 nm): a huge, perfectly legal energy
    frames[1, 3 * 7:3 * 7 + 3] += frames[1, 3 * 400] - frames[1, 3 * 7] - np.float32(0.02)
    # frame 3: a second short pair, in the last rows (the copy's last, partly empty, warps)
    frames[3, 3 * 998:3 * 998 + 3] += frames[3, 3 * 999] - frames[3, 3 * 998] + np.float32(0.015)
    sa, ha, ty = oracle.water_site_table(n)
    ctx.set_water_sites(n)
    res = ctx.hbond_batch(frames)
    oe, ost, oev, oties = oracle_frames(oracle, frames, sa, ha, ty)
    compare_batch(res, oe, ost, oev, oties)
    for f in (1, 3):                                       # the two frames really carry an edge under 0.56 A, and it dominates
        assert (oe[f]["length"] < 0.56).sum() >= 1 and abs(ost[f, 3]) > 100 * abs(ost[0, 3])
    for f in (0, 2):
        assert (oe[f]["length"] < 0.56).sum() == 0
    again = ctx.hbond_batch(frames)
    assert np.array_equal(again["stats"]["sum_energy"].view(np.uint64), res["stats"]["sum_energy"].view(np.uint64))
    for f in range(4):                                     # batch composition does not matter
        one = ctx.hbond_batch(frames[f:f + 1])
        assert one["stats"]["sum_energy"].view(np.uint64)[0] == res["stats"]["sum_energy"].view(np.uint64)[f]


# ------------------------------------------------------------------ periodic boxes
def oracle_frames_pbc(oracle, frames, boxes, site_atom, h_atom, site_type, first_limit=None):
    edges, stats, evals = [], [], 0
    ties = oracle.Ties()
    for f in range(frames.shape[0]):
        e, s, _, ev = oracle.frame_hbonds_pbc(frames[f], site_atom, h_atom, site_type, boxes[f],
                                              first_limit=first_limit, ties=ties)
        edges.append(e); stats.append(s); evals += ev
    return edges, np.array(stats), evals, ties.as_dict()


@pytest.mark.parametrize("n_waters,model,nf", [(216, 1, 4),      # BASELINE configs[0] box: 2 cells/dim -> all-pairs kernel
                                               (512, 0, 3),      # 3 cells/dim -> 1-cell home blocks
                                               (1000, 0, 3),     # 4 cells/dim -> 2-cell home blocks, every block wraps
                                               (4000, 0, 2)])
def test_periodic_water_box(ctx, oracle, n_waters, model, nf):
    box = oracle.synth_box(n_waters, model=model)
    frames = oracle.synth_frames(box, 0, nf)
    L = np.float32(box.box)
    boxes = np.tile(np.array([L, L, L], np.float32), (nf, 1))
    sa, ha, ty = oracle.water_site_table(n_waters)
    ctx.set_water_sites(n_waters)
    res = ctx.hbond_batch(frames, boxes=boxes)
    compare_batch(res, *oracle_frames_pbc(oracle, frames, boxes, sa, ha, ty))
    # more edges than with open boundaries (pairs through the box faces)
    assert res["n_edges"] > ctx.hbond_batch(frames)["n_edges"]


def test_periodic_anisotropic_unwrapped_mixed(ctx, oracle):
    """Anisotropic box that changes per frame, molecules displaced by whole box vectors
    (unwrapped trajectory), broken molecules (hydrogen on the other side of the box),
    mixed site types, asymmetric search."""
    rng = np.random.default_rng(8)
    n_sites, hmax, nf = 500, 3, 3
    natoms = 4 * n_sites
    site_atom = np.arange(n_sites, dtype=np.int32) * 4
    nh = rng.integers(0, 4, size=n_sites)
    h_atom = -np.ones((n_sites, hmax), np.int32)
    for s in range(n_sites):
        for k in range(nh[s]):
            h_atom[s, k] = site_atom[s] + 1 + k
    for s in range(0, n_sites, 5):
        h_atom[s] = [site_atom[s], site_atom[s] + 1, -1]
    site_type = rng.integers(0, 3, size=n_sites).astype(np.uint8)
    frames = np.zeros((nf, natoms, 3), np.float32)
    boxes = np.zeros((nf, 3), np.float32)
    for f in range(nf):
        boxes[f] = [2.7 + 0.1 * f, 3.4, 2.1 + 0.2 * f]          # 4,5,3 cells ... 4,5,3
        pos = rng.uniform(0, 1, size=(n_sites, 3)) * boxes[f]
        frames[f, 0::4] = pos
        for k in range(3):
            d = rng.normal(size=(n_sites, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
            frames[f, 1 + k::4] = pos + 0.1 * d                  # some hydrogens stick out of the box
        # unwrap: whole molecules moved by integer box vectors; single hydrogens too (broken molecules)
        shift = rng.integers(-2, 3, size=(n_sites, 3)) * boxes[f]
        for k in range(4):
            frames[f, k::4] += shift
        frames[f, 1::4][::7] += boxes[f]
    for first_limit in (None, 60):
        ctx.set_sites(site_atom, h_atom, site_type, first_limit=first_limit)
        res = ctx.hbond_batch(frames, boxes=boxes)
        compare_batch(res, *oracle_frames_pbc(oracle, frames, boxes, site_atom, h_atom, site_type, first_limit))
    assert res["n_edges"] > 0


def test_periodic_huge_box_equals_open_boundary(ctx, oracle):
    n = 600
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 3, 2)
    ctx.set_water_sites(n)
    a = ctx.hbond_batch(frames)
    b = ctx.hbond_batch(frames, boxes=np.full((2, 3), 50.0, np.float32))
    assert a["edges"].tobytes() == b["edges"].tobytes() and np.array_equal(a["frame_offsets"], b["frame_offsets"])


def test_periodic_device_batch_and_bad_boxes(ctx, oracle, capi):
    n, nf = 1000, 5
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 0, nf)
    boxes = np.full((nf, 3), np.float32(box.box), np.float32)
    sa, ha, ty = oracle.water_site_table(n)
    ctx.set_water_sites(n)
    d = ctx.dev_alloc(frames.nbytes)
    ctx.h2d(d, frames)
    ctx.set_batch_frames(2)
    out = ctx.hbond_batch_dev(d, nf, 3 * n, boxes=boxes)
    ctx.set_batch_frames(0)
    e, fo, st = ctx.fetch_dev_result(out)
    ctx.dev_free(d)
    compare_batch(dict(out, edges=e, frame_offsets=fo, stats=st),
                  *oracle_frames_pbc(oracle, frames, boxes, sa, ha, ty))
    with pytest.raises(capi.WnError):
        ctx.hbond_batch(frames, boxes=np.zeros((nf, 3), np.float32))


def test_find_pairs_periodic(ctx, oracle):
    rng = np.random.default_rng(12)
    for n, L in ((300, (1.9, 2.2, 3.0)), (1500, (5.0, 4.3, 6.1))):
        xyz = (rng.uniform(-1, 2, size=(n, 3)) * np.array(L)).astype(np.float32).astype(np.float64)
        want = oracle.brute_find_pairs_pbc(xyz, L)
        got, _ = ctx.find_pairs(xyz, box=L)
        assert np.array_equal(got, want)
        got, _ = ctx.find_pairs(xyz, first_limit=40, box=L)
        assert np.array_equal(got, oracle.brute_find_pairs_pbc(xyz, L, first_limit=40))


def test_shared_reciprocal_division_is_ieee_exact(ctx):
    """The three `OO /= distance` divisions share one reciprocal (3 instructions each);
    2^33 random triples (kernel-like, arbitrary, edge patterns) against __ddiv_rn, bitwise."""
    assert ctx.selftest_div(1 << 33, seed=20261019) == 0
    assert ctx.selftest_div(1 << 30, seed=7) == 0


def test_randomised_systems_open_and_periodic(ctx, oracle):
    """40 seeded random systems: site counts 1..400, densities from dilute gas to 20x liquid,
    random DONOR/ACCEPTOR/WATER tables with 0-3 hydrogens (some lists starting with the site
    atom itself), random row limits, open boundary or random anisotropic periodic boxes
    (from 1 cell per dimension upwards, coordinates not wrapped).  Bit-exact vs the oracle."""
    rng = np.random.default_rng(20261019)
    for trial in range(40):
        n_sites = int(rng.integers(1, 401))
        nf = int(rng.integers(1, 4))
        hmax = 3
        natoms = 4 * n_sites
        site_atom = (np.arange(n_sites) * 4).astype(np.int32)
        h_atom = -np.ones((n_sites, hmax), np.int32)
        nh = rng.integers(0, 4, size=n_sites)
        for s in range(n_sites):
            for k in range(nh[s]):
                h_atom[s, k] = site_atom[s] + 1 + k
            if rng.random() < 0.25:
                h_atom[s] = [site_atom[s], site_atom[s] + 1, -1]
        all_water = rng.random() < 0.3
        site_type = np.full(n_sites, oracle.WATER, np.uint8) if all_water else rng.integers(0, 3, size=n_sites).astype(np.uint8)
        if all_water and rng.random() < 0.5:
            h_atom[:, 0] = site_atom; h_atom[:, 1] = site_atom + 1; h_atom[:, 2] = -1     # the reference water table
        periodic = rng.random() < 0.6
        side = float(rng.choice([0.5, 0.9, 1.4, 2.0, 2.7, 4.1, 7.0]))
        boxes = (np.array([side, side * rng.uniform(0.7, 1.6), side * rng.uniform(0.7, 1.6)]) *
                 rng.uniform(0.95, 1.05, size=(nf, 3))).astype(np.float32)
        frames = np.zeros((nf, natoms, 3), np.float32)
        for f in range(nf):
            pos = rng.uniform(0, 1, size=(n_sites, 3)) * boxes[f]
            if periodic:
                pos += rng.integers(-1, 2, size=(n_sites, 3)) * boxes[f]
            frames[f, 0::4] = pos
            for k in range(3):
                d = rng.normal(size=(n_sites, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
                frames[f, 1 + k::4] = pos + 0.1 * d
        first_limit = None if rng.random() < 0.6 else int(rng.integers(0, n_sites + 1))
        ctx.set_sites(site_atom, h_atom, site_type, first_limit=first_limit)
        if periodic:
            res = ctx.hbond_batch(frames, boxes=boxes)
            want = oracle_frames_pbc(oracle, frames, boxes, site_atom, h_atom, site_type, first_limit)
        else:
            res = ctx.hbond_batch(frames)
            want = oracle_frames(oracle, frames, site_atom, h_atom, site_type, first_limit)
        try:
            compare_batch(res, *want)
        except AssertionError as e:
            raise AssertionError("trial %d (n=%d, periodic=%s, side=%.1f, first_limit=%s): %s"
                                 % (trial, n_sites, periodic, side, first_limit, e))


def test_one_million_atoms_rows_and_properties(ctx, oracle):
    """configs[3] size (333,333 waters, ~1M atoms): exact parity on the first rows (oracle
    restricted to rows i < 40), order/uniqueness of the full list, open and periodic."""
    n, nf = 333333, 2
    d = ctx.dev_alloc(nf * 3 * n * 12)
    ctx.synth_frames_dev(n, 7, nf, d)
    frames = np.empty((nf, 3 * n, 3), np.float32)
    ctx.d2h(frames, d)
    sa, ha, ty = oracle.water_site_table(n)
    ctx.set_water_sites(n)
    L = np.float32(oracle.synth_box(n).box)
    for boxes in (None, np.full((nf, 3), L, np.float32)):
        out = ctx.hbond_batch_dev(d, nf, 3 * n, boxes=boxes)
        e, fo, st = ctx.fetch_dev_result(out)
        for f in range(nf):
            ef = e[int(fo[f]):int(fo[f + 1])]
            key = ef["i"].astype(np.int64) * n + ef["j"]
            assert np.all(np.diff(key) > 0) and np.all(ef["i"] < ef["j"]) and ef["j"].max() < n
            assert st["num_edges"][f] == len(ef)
            if boxes is None:
                want, _, _, _ = oracle.frame_hbonds(frames[f], sa, ha, ty, first_limit=40)
            else:
                want, _, _, _ = oracle.frame_hbonds_pbc(frames[f], sa, ha, ty, boxes[f], first_limit=40)
            check_edges(ef[ef["i"] < 40], want)
            # 2,000+ rows scattered over the whole box (every region of the 33^3-cell grid, not one corner),
            # plus the last rows: every edge of those rows, bit-exact, none missing, none extra
            rows = np.unique(np.concatenate([np.random.default_rng(100 + f).integers(0, n, 2200),
                                             np.arange(n - 30, n)])).astype(np.int32)
            want_r, _ = oracle.frame_hbonds_rows(frames[f], sa, ha, ty, rows, box=None if boxes is None else boxes[f])
            assert len(want_r) > 20000
            check_edges(ef[np.isin(ef["i"], rows)], want_r)
    ctx.dev_free(d)


def test_per_frame_site_subsets(ctx, oracle):
    """The alpha-shape filter of the reference keeps a different set of buried waters every
    frame (src/waternetwork.cpp:529-533,562-576): per-frame site lists in one batch, edge
    indices = positions in the frame's list.  Protein-like sites first (always present), then
    that frame's waters; open and periodic; compared with the oracle run on the gathered table."""
    rng = np.random.default_rng(4)
    n_w, n_p, nf = 900, 40, 6
    box = oracle.synth_box(n_w)
    wat = oracle.synth_frames(box, 0, nf)                       # [nf][3*n_w][3]
    natoms = 3 * n_w + 4 * n_p
    frames = np.zeros((nf, natoms, 3), np.float32)
    frames[:, :3 * n_w] = wat
    L = np.float32(box.box)
    for f in range(nf):
        pos = rng.uniform(0, 1, size=(n_p, 3)) * L
        frames[f, 3 * n_w + 0::4] = pos
        for k in range(3):
            d = rng.normal(size=(n_p, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
            frames[f, 3 * n_w + 1 + k::4] = pos + 0.1 * d
    # site table: protein sites (rows 0..n_p-1), then every water (rows n_p..)
    site_atom = np.concatenate([3 * n_w + 4 * np.arange(n_p), 3 * np.arange(n_w)]).astype(np.int32)
    h_atom = -np.ones((n_p + n_w, 3), np.int32)
    nh = rng.integers(0, 4, size=n_p)
    for s in range(n_p):
        for k in range(nh[s]):
            h_atom[s, k] = site_atom[s] + 1 + k
    h_atom[n_p:, 0] = 3 * np.arange(n_w); h_atom[n_p:, 1] = 3 * np.arange(n_w) + 1
    site_type = np.concatenate([rng.integers(0, 2, size=n_p), np.full(n_w, oracle.WATER)]).astype(np.uint8)
    ctx.set_sites(site_atom, h_atom, site_type)
    lists, off = [], [0]
    for f in range(nf):
        keep = np.sort(rng.choice(n_w, size=int(rng.integers(0, 400)), replace=False))
        lst = np.concatenate([np.arange(n_p), n_p + keep]).astype(np.int32)
        if f == 2:
            lst = lst[:0]                                           # an empty frame
        lists.append(lst); off.append(off[-1] + len(lst))
    sites = np.concatenate(lists)
    for boxes in (None, np.full((nf, 3), L, np.float32)):
        res = ctx.hbond_batch_subset(frames, off, sites, boxes=boxes)
        fo = res["frame_offsets"]
        for f in range(nf):
            lst = lists[f]
            if boxes is None:
                want, st, _, ev = oracle.frame_hbonds(frames[f], site_atom[lst], h_atom[lst], site_type[lst]) if len(lst) else \
                    (np.zeros(0, oracle.EDGE_DTYPE), np.zeros(4), 0, 0)
            else:
                want, st, _, ev = oracle.frame_hbonds_pbc(frames[f], site_atom[lst], h_atom[lst], site_type[lst], boxes[f]) if len(lst) else \
                    (np.zeros(0, oracle.EDGE_DTYPE), np.zeros(4), 0, 0)
            check_edges(res["edges"][int(fo[f]):int(fo[f + 1])], want)
            assert res["stats"]["num_sites"][f] == len(lst) and res["stats"]["num_edges"][f] == len(want)
            assert int(res["stats"]["pair_evals"][f]) == ev
    # consumers of the last batch after per-frame site lists: indices are list positions.  The device formatter
    # refuses an id table (it would print the ids of the wrong rows); components are labelled by list position.
    res = ctx.hbond_batch_subset(frames, off, sites)
    import wn_pkg
    with pytest.raises(wn_pkg.load_capi().WnError) as ei:
        ctx.format_edges(list(range(nf)), (np.arange(n_p + n_w) + 1).astype(np.uint32))
    assert ei.value.code == 3
    try:                                                          # no id table: the list positions themselves
        text = ctx.format_edges(list(range(nf)))
        assert text.count(b"\n") == nf + res["n_edges"]
    except wn_pkg.load_capi().WnError as err:                     # random protein sites can sit on top of each other:
        assert err.code == 5                                      # an energy wider than its column is reported, not truncated
    labels, cst = ctx.components(0.0)
    fo = res["frame_offsets"]
    for f in range(nf):
        lab, stc = oracle.components(res["edges"][int(fo[f]):int(fo[f + 1])], len(lists[f]), 0.0)
        assert np.array_equal(labels[f][:len(lists[f])], lab) and np.all(labels[f][len(lists[f]):] == -1)
        assert int(cst["n_components"][f]) == int(stc[0])
    with pytest.raises(Exception):
        ctx.hbond_batch_subset(frames, [0, 1, 2, 3, 4, 5, 6], np.array([0, 1, 2, 3, 4, n_p + n_w], np.int32))
    # a negative first_limit means "every row", as in the C++ wrappers
    ctx.set_sites(site_atom, h_atom, site_type, first_limit=-1)
    assert ctx.hbond_batch(frames[:1])["n_edges"] > 0


def test_gpu_text_formatter_matches_printf(ctx, oracle, capi):
    """f1 on the GPU: the rows of edges-frames.xvg.dat (src/waternetwork.cpp:638-652) formatted by
    wn_k_format_rows must equal "%6d%6d%8.4f%8.4f%8.4f\\n" (what std::fixed, precision 4, setw gives;
    the host EdgeFileWriter is checked against iostreams in tests/cpp/test_writer.cpp)."""
    n, nf = 700, 5
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 0, nf)
    ctx.set_water_sites(n)
    res = ctx.hbond_batch(frames)
    numbers = [0, 10, 20, 1000, 123456]
    ids = (np.arange(n) + 1 + 37).astype(np.uint32)              # solvent ids offset by proteinSites.size()+1
    text = ctx.format_edges(numbers, ids)
    fo = res["frame_offsets"]
    want = []
    for f in range(nf):
        want.append("#--- frame %d ---\n" % numbers[f])
        for e in res["edges"][int(fo[f]):int(fo[f + 1])]:
            want.append("%6d%6d%8.4f%8.4f%8.4f\n" % (ids[e["i"]], ids[e["j"]], e["energy"], e["length"], e["cosine"]))
    assert text == "".join(want).encode()
    # without an id table the indices themselves are printed
    t2 = ctx.format_edges(numbers)
    first = res["edges"][0]
    assert t2.split(b"\n")[1] == ("%6d%6d%8.4f%8.4f%8.4f" % (first["i"], first["j"], first["energy"], first["length"], first["cosine"])).encode()
    # an id with 7 digits does not fit the fixed-width row: reported, not silently truncated
    with pytest.raises(capi.WnError) as ei:
        ctx.format_edges(numbers, ids + 1000000)
    assert ei.value.code == 5
    # empty result
    ctx.set_water_sites(3)
    ctx.hbond_batch(np.zeros((2, 9, 3), np.float32) + np.arange(9)[None, :, None] * 5.0)
    assert ctx.format_edges([4, 5]) == b"#--- frame 4 ---\n#--- frame 5 ---\n"


# ------------------------------------------------------------------ triclinic boxes
def test_triclinic_boxes(ctx, oracle):
    """GROMACS triclinic boxes (lower-triangular matrix), all-pairs kernel: skewed water boxes with
    unwrapped coordinates, mixed site types, and the pair finder; bit-exact vs the triclinic oracle."""
    rng = np.random.default_rng(17)
    n, nf = 400, 3
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 5, nf).copy()
    L = float(box.box)
    boxes9 = np.zeros((nf, 9), np.float32)
    for f in range(nf):
        boxes9[f] = [L, 0, 0, 0.31 * L, 0.95 * L, 0, -0.22 * L, 0.4 * L, 0.9 * L + 0.05 * f]
        B = boxes9[f].reshape(3, 3).astype(np.float64)
        shift = rng.integers(-2, 3, size=(n, 3)) @ B                  # whole molecules moved by lattice vectors
        for a in range(3):
            frames[f, a::3] += shift.astype(np.float32)
    sa, ha, ty = oracle.water_site_table(n)
    ctx.set_water_sites(n)
    res = ctx.hbond_batch(frames, boxes=boxes9)
    edges, stats, evals = [], [], 0
    ties = oracle.Ties()
    for f in range(nf):
        e, s, _, ev = oracle.frame_hbonds_tric(frames[f], sa, ha, ty, boxes9[f], ties=ties)
        edges.append(e); stats.append(s); evals += ev
    compare_batch(res, edges, np.array(stats), evals, ties.as_dict())
    assert res["n_edges"] > 0
    # mixed table + row limit
    n_sites = 250
    site_atom = (np.arange(n_sites) * 4).astype(np.int32)
    h_atom = -np.ones((n_sites, 3), np.int32)
    nh = rng.integers(0, 4, size=n_sites)
    for s in range(n_sites):
        for k in range(nh[s]):
            h_atom[s, k] = site_atom[s] + 1 + k
    site_type = rng.integers(0, 3, size=n_sites).astype(np.uint8)
    b9 = np.array([[2.2, 0, 0, 0.7, 2.0, 0, 0.5, -0.6, 1.9]], np.float32)
    fr = np.zeros((1, 4 * n_sites, 3), np.float32)
    pos = rng.uniform(0, 1, size=(n_sites, 3)) @ b9.reshape(3, 3)
    fr[0, 0::4] = pos
    for k in range(3):
        d = rng.normal(size=(n_sites, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
        fr[0, 1 + k::4] = pos + 0.1 * d
    ctx.set_sites(site_atom, h_atom, site_type, first_limit=90)
    res = ctx.hbond_batch(fr, boxes=b9)
    t2 = oracle.Ties()
    e, s, _, ev = oracle.frame_hbonds_tric(fr[0], site_atom, h_atom, site_type, b9[0], first_limit=90, ties=t2)
    compare_batch(res, [e], np.array([s]), ev, t2.as_dict())
    # a diagonal matrix routes to the rectangular definition
    ctx.set_water_sites(n)
    diag = np.zeros((nf, 9), np.float32); diag[:, 0] = diag[:, 4] = diag[:, 8] = L
    fr0 = oracle.synth_frames(box, 5, nf)
    assert ctx.hbond_batch(fr0, boxes=diag)["edges"].tobytes() == \
        ctx.hbond_batch(fr0, boxes=np.full((nf, 3), np.float32(L)))["edges"].tobytes()
    # pair finder
    xyz = (rng.uniform(-1, 2, size=(500, 3)) @ boxes9[0].reshape(3, 3).astype(np.float64)).astype(np.float32).astype(np.float64)
    got, _ = ctx.find_pairs(xyz, box=boxes9[0])
    assert np.array_equal(got, oracle.brute_find_pairs_tric(xyz, boxes9[0]))


def test_energy_parameters_and_hydrogenless_tables(ctx, oracle):
    """setRadiusShift/setAngleShift surface: non-default switch parameters, including an ordering
    that makes the (normally unreachable) smooth branch of the angle switch live; and a
    site table without any hydrogen list."""
    n, nf = 500, 2
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 9, nf)
    ctx.set_water_sites(n)
    base = ctx.hbond_batch(frames)
    L = oracle.lib()
    for (r_on, r_off, t_on, t_off) in ((4.0, 6.0, 0.25, 0.0), (3.0, 6.5, 0.0, 0.25), (5.5, 6.5, 0.1, 0.6)):
        ctx.set_energy_params(r_on, r_off, t_on, t_off)
        res = ctx.hbond_batch(frames)
        e = res["edges"]
        # geometry does not depend on the energy parameters
        for k in ("i", "j", "length", "cosine"):
            assert np.array_equal(e[k], base["edges"][k])
        want = np.array([L.wno_compute_energy_ex(float(a), float(b), r_on, r_off, t_on, t_off)
                         for a, b in zip(e["length"], e["cosine"])]).astype(np.float32)
        ge, we = e["energy"].astype(np.float64), want.astype(np.float64)
        assert np.all(np.abs(ge - we) <= 1e-6 * np.abs(we)), np.max(np.abs(ge - we) / np.maximum(np.abs(we), 1e-300))
        assert np.count_nonzero(we) > 0
    ctx.set_energy_params()                                   # back to the reference defaults
    again = ctx.hbond_batch(frames)
    assert again["edges"].tobytes() == base["edges"].tobytes()
    # no hydrogens anywhere: pairs exist, nothing is evaluated (src/waternetwork.cpp:205-208)
    site_atom = (3 * np.arange(n)).astype(np.int32)
    ctx.set_sites(site_atom, -np.ones((n, 1), np.int32), np.full(n, oracle.WATER, np.uint8))
    r0 = ctx.hbond_batch(frames)
    assert r0["n_edges"] == 0 and r0["n_pair_evals"] == 0
    want, st, npairs, ev = oracle.frame_hbonds(frames[0], site_atom, -np.ones((n, 1), np.int32), np.full(n, oracle.WATER, np.uint8))
    assert len(want) == 0 and ev == 0 and npairs > 0


@pytest.mark.parametrize("n,scale,nf", [(4000, 1.0, 2),      # 7 x 7 x 6 cells: 2-cell home blocks
                                        (1400, 0.82, 3)])    # 4 x 4 x 4 cells: 1-cell home blocks, every block wraps
def test_triclinic_cell_sweep(ctx, oracle, n, scale, nf):
    """Triclinic boxes large enough for the cell-list sweep (shifted neighbour rows across the y
    and z boundaries), heavily skewed (|bx|, |cx| ~ ax/2, |cy| ~ by/2), molecules displaced by
    lattice vectors; bit-exact vs the triclinic oracle."""
    rng = np.random.default_rng(23 + n)
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 2, nf).copy()
    L = float(box.box) * scale
    boxes9 = np.zeros((nf, 9), np.float32)
    for f in range(nf):
        boxes9[f] = [L, 0, 0, (0.45 - 0.9 * (f % 2)) * L, 0.97 * L, 0, -0.48 * L, 0.46 * L, 0.93 * L + 0.02 * f]
        B = boxes9[f].reshape(3, 3).astype(np.float64)
        shift = rng.integers(-1, 2, size=(n, 3)) @ B
        for a in range(3):
            frames[f, a::3] += shift.astype(np.float32)
    sa, ha, ty = oracle.water_site_table(n)
    ctx.set_water_sites(n)
    res = ctx.hbond_batch(frames, boxes=boxes9)
    edges, stats, evals = [], [], 0
    ties = oracle.Ties()
    for f in range(nf):
        e, s, _, ev = oracle.frame_hbonds_tric(frames[f], sa, ha, ty, boxes9[f], ties=ties)
        edges.append(e); stats.append(s); evals += ev
    compare_batch(res, edges, np.array(stats), evals, ties.as_dict())
    assert res["n_edges"] > 10 * n


def test_randomised_triclinic_systems(ctx, oracle):
    """30 seeded random triclinic systems: box sizes from one cell to ~10 cells per dimension,
    skews up to +-0.7 of the diagonal elements (beyond what GROMACS itself allows), random tables,
    row limits, unwrapped coordinates.  Exercises the all-pairs kernel, 1- and 2-cell home blocks."""
    rng = np.random.default_rng(777)
    for trial in range(30):
        n_sites = int(rng.integers(1, 500))
        nf = int(rng.integers(1, 3))
        natoms = 4 * n_sites
        site_atom = (np.arange(n_sites) * 4).astype(np.int32)
        h_atom = -np.ones((n_sites, 3), np.int32)
        nh = rng.integers(0, 4, size=n_sites)
        for s in range(n_sites):
            for k in range(nh[s]):
                h_atom[s, k] = site_atom[s] + 1 + k
            if rng.random() < 0.3:
                h_atom[s] = [site_atom[s], site_atom[s] + 1, -1]
        site_type = rng.integers(0, 3, size=n_sites).astype(np.uint8) if rng.random() < 0.7 else np.full(n_sites, oracle.WATER, np.uint8)
        side = float(rng.choice([0.8, 1.5, 2.7, 3.4, 4.2, 6.6]))
        boxes9 = np.zeros((nf, 9), np.float32)
        frames = np.zeros((nf, natoms, 3), np.float32)
        for f in range(nf):
            ax, by, cz = side * rng.uniform(0.9, 1.3, size=3)
            boxes9[f] = [ax, 0, 0, rng.uniform(-0.7, 0.7) * ax, by, 0, rng.uniform(-0.7, 0.7) * ax, rng.uniform(-0.7, 0.7) * by, cz]
            B = boxes9[f].reshape(3, 3).astype(np.float64)
            pos = rng.uniform(0, 1, size=(n_sites, 3)) @ B + rng.integers(-1, 2, size=(n_sites, 3)) @ B
            frames[f, 0::4] = pos
            for k in range(3):
                d = rng.normal(size=(n_sites, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
                frames[f, 1 + k::4] = pos + 0.1 * d
        first_limit = None if rng.random() < 0.6 else int(rng.integers(0, n_sites + 1))
        ctx.set_sites(site_atom, h_atom, site_type, first_limit=first_limit)
        res = ctx.hbond_batch(frames, boxes=boxes9)
        edges, stats, evals = [], [], 0
        ties = oracle.Ties()
        for f in range(nf):
            e, s, _, ev = oracle.frame_hbonds_tric(frames[f], site_atom, h_atom, site_type, boxes9[f],
                                                   first_limit=first_limit, ties=ties)
            edges.append(e); stats.append(s); evals += ev
        try:
            compare_batch(res, edges, np.array(stats), evals, ties.as_dict())
        except AssertionError as err:
            raise AssertionError("trial %d (n=%d, side=%.1f, box=%s): %s" % (trial, n_sites, side, boxes9[0], err))


def test_general_entry_point_combinations(ctx, oracle):
    """wn_hbond_batch_ex: device-resident frames + triclinic matrices + per-frame site lists in one
    call, against the oracle on the gathered tables."""
    rng = np.random.default_rng(5)
    n, nf = 1500, 3
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 0, nf)
    L = float(box.box)
    b9 = np.tile(np.array([L, 0, 0, 0.3 * L, L, 0, -0.2 * L, 0.25 * L, L], np.float32), (nf, 1))
    sa, ha, ty = oracle.water_site_table(n)
    ctx.set_water_sites(n)
    lists, off = [], [0]
    for f in range(nf):
        lst = np.sort(rng.choice(n, size=int(rng.integers(300, 1200)), replace=False)).astype(np.int32)
        lists.append(lst); off.append(off[-1] + len(lst))
    d = ctx.dev_alloc(frames.nbytes)
    ctx.h2d(d, frames)
    out = ctx.hbond_batch_ex(d, nf, 3 * n, on_device=True, boxes=b9, subset_offsets=off, subset_sites=np.concatenate(lists))
    e, fo, st = ctx.fetch_dev_result(out)
    ctx.dev_free(d)
    for f in range(nf):
        lst = lists[f]
        want, s, _, ev = oracle.frame_hbonds_tric(frames[f], sa[lst], ha[lst], ty[lst], b9[f])
        check_edges(e[int(fo[f]):int(fo[f + 1])], want)
        assert st["num_sites"][f] == len(lst) and int(st["pair_evals"][f]) == ev


def test_compact_csr16_layout(ctx, oracle, capi):
    """WN_LAYOUT_CSR16: {j, energy, length, cosine} + row offsets carries exactly the wn_edge list."""
    n, nf = 2000, 4
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 0, nf)
    ctx.set_water_sites(n)
    ctx.set_batch_frames(3)                       # two internal passes
    full = ctx.hbond_batch_ex(frames)
    csr = ctx.hbond_batch_ex(frames, layout=capi.LAYOUT_CSR16)
    ctx.set_batch_frames(0)
    assert csr["layout"] == capi.LAYOUT_CSR16 and csr["edges"].dtype == capi.EDGE16_DTYPE
    assert np.array_equal(csr["frame_offsets"], full["frame_offsets"]) and np.array_equal(csr["row_offsets"], full["row_offsets"])
    assert csr["stats"].tobytes() == full["stats"].tobytes()
    for k in ("j", "energy", "length", "cosine"):
        assert np.array_equal(csr["edges"][k].view(np.uint32), full["edges"][k].view(np.uint32))
    # i reconstructed from the row offsets
    fo, ro = csr["frame_offsets"], csr["row_offsets"]
    for f in range(nf):
        cnt = np.diff(np.append(ro[f], fo[f + 1] - fo[f]).astype(np.int64))
        i = np.repeat(np.arange(n), cnt)
        assert np.array_equal(i, full["edges"]["i"][int(fo[f]):int(fo[f + 1])])


def test_compact_csr12_layout_and_double_buffered_results(ctx, oracle, capi):
    """WN_LAYOUT_CSR12 ({j, length, cosine} + row offsets; the energy is recomputed on the host on demand, see
    include/wn_energy.h): same list as the 20-byte layout, statistics (incl. the energy sum) unchanged.
    With two host result sets the previous call's arrays stay valid while the next call runs."""
    n, nf = 1000, 4
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 30, nf)
    ctx.set_water_sites(n)
    ref = ctx.hbond_batch_ex(frames, layout=capi.LAYOUT_EDGES)
    res = ctx.hbond_batch_ex(frames, layout=capi.LAYOUT_CSR12)
    assert res["layout"] == capi.LAYOUT_CSR12 and res["edges"].dtype.itemsize == 12
    assert np.array_equal(res["frame_offsets"], ref["frame_offsets"]) and np.array_equal(res["row_offsets"], ref["row_offsets"])
    for k in ("j", "length", "cosine"):
        assert np.array_equal(res["edges"][k].view(np.uint32), ref["edges"][k].view(np.uint32))
    assert res["stats"].tobytes() == ref["stats"].tobytes()
    # host energy on demand == the oracle's computeEnergy, float-rounded (both follow the reference's pow() arithmetic)
    e_host = np.array([np.float32(oracle.compute_energy(float(l), float(c))) for l, c in
                       zip(res["edges"]["length"][:500], res["edges"]["cosine"][:500])], np.float32)
    assert np.allclose(e_host, ref["edges"]["energy"][:500], rtol=ENERGY_RTOL, atol=0)
    with pytest.raises(capi.WnError):
        ctx.components(0.0)                                   # no energies in this layout
    # device-resident variant
    d = ctx.dev_alloc(frames.nbytes)
    ctx.h2d(d, frames)
    out = ctx.hbond_batch_ex(d, nf, 3 * n, on_device=True, layout=capi.LAYOUT_CSR12)
    e, fo, st = ctx.fetch_dev_result(out)
    assert e.tobytes() == res["edges"].tobytes() and st.tobytes() == ref["stats"].tobytes()
    ctx.dev_free(d)
    # two result sets: call k's zero-copy views survive call k+1
    ctx.set_result_sets(2)
    a = ctx.hbond_batch_ex(frames[:2], layout=capi.LAYOUT_CSR16, copy=False)
    a_bytes = a["edges"].tobytes()
    b = ctx.hbond_batch_ex(frames[2:], layout=capi.LAYOUT_CSR16, copy=False)
    assert a["edges"].tobytes() == a_bytes and b["edges"].tobytes() != a_bytes
    assert a["edges"].ctypes.data != b["edges"].ctypes.data
    ctx.set_result_sets(1)
    assert ctx.bind_thread_to_device() >= 0
    h2d, d2h = ctx.pcie_probe(64 << 20, 2)
    assert h2d > 1.0 and d2h > 1.0


def test_connected_components_match_oracle(ctx, oracle, capi):
    """wn_components (union-find on the device-resident edge list) against the CPU restatement of the
    reference script's filter + connected components: canonical labels and per-frame counts, exact."""
    n, nf = 4000, 3
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 0, nf)
    ctx.set_water_sites(n)
    res = ctx.hbond_batch(frames)
    fo = res["frame_offsets"]

    def check(thr, labels, stats):
        for f in range(nf):
            e = res["edges"][int(fo[f]):int(fo[f + 1])]
            lab, st = oracle.components(e, n, thr)
            assert np.array_equal(labels[f], lab), (thr, f)
            assert [stats[f][k] for k in ("n_components", "largest", "n_nodes", "n_edges")] == list(st), (thr, f)

    for thr in (0.0, 5e-5, 1.0, 3.0, 1e9):
        labels, stats = ctx.components(thr)
        check(thr, labels, stats)
    # giant component at threshold 0, fragments at a high threshold
    _, s0 = ctx.components(0.0)
    _, s3 = ctx.components(3.0)
    assert s0["largest"][0] > n // 2 and s3["n_components"][0] > s0["n_components"][0]
    # compact layout (row index from the CSR offsets) and a device-resident multi-pass batch
    ctx.set_batch_frames(2)
    ctx.hbond_batch_ex(frames, layout=capi.LAYOUT_CSR16)
    labels, stats = ctx.components(1.0)
    check(1.0, labels, stats)
    dev = ctx.dev_alloc(frames.nbytes)
    ctx.h2d(dev, frames)
    ctx.hbond_batch_dev(dev, nf, 3 * n)
    labels, stats = ctx.components(1.0)
    check(1.0, labels, stats)
    ctx.dev_free(dev)
    ctx.set_batch_frames(0)


def test_connected_components_small_and_degenerate(ctx, oracle):
    """Mixed site tables, empty frames, coincident sites (NaN energies are kept, as in the script)."""
    rng = np.random.default_rng(3)
    n = 300
    o = rng.uniform(0, 1.2, size=(n, 3))
    o[10] = o[11]                                  # coincident pair: NaN cosine/energy paths
    h1 = o + rng.normal(size=(n, 3)) * 0.06
    h2 = o + rng.normal(size=(n, 3)) * 0.06
    frames = np.stack([o, h1, h2], axis=1).reshape(1, 3 * n, 3).astype(np.float32)
    frames = np.concatenate([frames, frames * 100.0])          # second frame: nothing within reach
    ctx.set_water_sites(n)
    res = ctx.hbond_batch(frames)
    fo = res["frame_offsets"]
    for thr in (0.0, 0.7):
        labels, stats = ctx.components(thr)
        for f in range(2):
            lab, st = oracle.components(res["edges"][int(fo[f]):int(fo[f + 1])], n, thr)
            assert np.array_equal(labels[f], lab)
            assert [stats[f][k] for k in ("n_components", "largest", "n_nodes", "n_edges")] == list(st)
    assert stats["n_edges"][1] == 0 and np.all(labels[1] == -1)


def test_results_are_reproducible_run_to_run(ctx, oracle, capi):
    """Row slots are claimed with atomics and home blocks by a work counter, so arrival order differs from
    run to run; the ordered copy and the fixed-order sums must hide that: edges, offsets, statistics
    (including the fp64 energy sum) and component labels are byte-identical across runs, pass sizes and layouts."""
    n, nf = 4000, 6
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 100, nf)
    ctx.set_water_sites(n)
    ref = None
    for rep, pass_frames in enumerate((0, 0, 1, 4, 6)):
        ctx.set_batch_frames(pass_frames)
        res = ctx.hbond_batch(frames)
        labels, cst = ctx.components(0.5)
        cur = (res["edges"].tobytes(), res["frame_offsets"].tobytes(), res["stats"].tobytes(),
               res["row_offsets"].tobytes(), labels.tobytes(), cst.tobytes())
        if ref is None:
            ref = cur
        assert cur == ref, (rep, pass_frames)
    ctx.set_batch_frames(0)


def test_host_frames_results_kept_on_device(ctx, oracle, capi):
    """results_on_device: host frames in, edge list left in HBM (for wn_components / wn_format_edges);
    fetched explicitly it equals the host result, and the components computed from it are the same."""
    n, nf = 1000, 5
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 7, nf)
    ctx.set_water_sites(n)
    ctx.set_batch_frames(2)
    host = ctx.hbond_batch_ex(frames)
    lab_h, st_h = ctx.components(0.5)
    dev = ctx.hbond_batch_ex(frames, results_on_device=True)
    assert "edges_dev" in dev and dev["n_edges"] == host["n_edges"]
    # every counted pair evaluation was first a candidate test; the half-shell sweep looks at a few hundred
    # candidates per site at liquid density
    assert host["n_pair_evals"] <= host["n_candidate_tests"] == dev["n_candidate_tests"]
    assert 50 * n * nf < host["n_candidate_tests"] < 400 * n * nf
    e, fo, st = ctx.fetch_dev_result(dev)
    assert e.tobytes() == host["edges"].tobytes() and np.array_equal(fo, host["frame_offsets"])
    assert st.tobytes() == host["stats"].tobytes()
    lab_d, st_d = ctx.components(0.5)
    assert np.array_equal(lab_d, lab_h) and st_d.tobytes() == st_h.tobytes()
    text_d = ctx.format_edges(np.arange(nf))
    ctx.hbond_batch_ex(frames)
    assert ctx.format_edges(np.arange(nf)) == text_d
    ctx.set_batch_frames(0)


def test_connected_components_every_path(ctx, oracle, capi):
    """The two implementations behind wn_components -- one CTA per frame in shared memory, multi-kernel
    global memory (frames over ~56 k sites, or forced) -- on the same batches, both layouts, against the oracle."""
    import os
    cases = [(60000, 1, (0.0, 1.5), (None,)), (3000, 3, (0.0, 0.8), (None, "global"))]
    try:
        for n, nf, thrs, paths in cases:
            box = oracle.synth_box(n)
            frames = oracle.synth_frames(box, 0, nf)
            ctx.set_water_sites(n)
            res = ctx.hbond_batch(frames)
            fo = res["frame_offsets"]
            want = {thr: [oracle.components(res["edges"][int(fo[f]):int(fo[f + 1])], n, thr) for f in range(nf)] for thr in thrs}
            for layout in (capi.LAYOUT_EDGES, capi.LAYOUT_CSR16):
                ctx.hbond_batch_ex(frames, layout=layout)
                for path in paths:
                    if path is None:
                        os.environ.pop("WN_COMPONENTS_PATH", None)
                    else:
                        os.environ["WN_COMPONENTS_PATH"] = path
                    for thr in thrs:
                        labels, stats = ctx.components(thr)
                        for f in range(nf):
                            lab, st = want[thr][f]
                            assert np.array_equal(labels[f], lab), (n, layout, path, thr, f)
                            assert [stats[f][k] for k in ("n_components", "largest", "n_nodes", "n_edges")] == list(st), (n, layout, path, thr, f)
    finally:
        os.environ.pop("WN_COMPONENTS_PATH", None)


def test_new_entry_points_reject_misuse(oracle, capi):
    """Error behaviour of the later additions: no batch yet, unknown layout, text of a CSR16 result."""
    c = capi.Context(0)
    try:
        c.set_water_sites(10)
        with pytest.raises(capi.WnError) as ei:
            c.components(0.0, n_frames=1)
        assert ei.value.code == capi.WN_ERR_BAD_ARG and "no batch result" in str(ei.value)
        frames = oracle.synth_frames(oracle.synth_box(10), 0, 2)
        with pytest.raises(capi.WnError) as ei:
            c.hbond_batch_ex(frames, layout=7)
        assert ei.value.code == capi.WN_ERR_BAD_ARG and "result_layout" in str(ei.value)
        c.hbond_batch_ex(frames, layout=capi.LAYOUT_CSR16)
        with pytest.raises(capi.WnError) as ei:
            c.format_edges(np.arange(2))
        assert "WN_LAYOUT_EDGES" in str(ei.value)
        labels, stats = c.components(0.0)              # a CSR16 result is fine for the components
        assert labels.shape == (2, 10)
    finally:
        c.close()


def test_soak_many_frames_against_oracle(ctx, oracle):
    """200 frames of the 4,000-water box in one device batch (persistent warps over ~50 k home blocks, several
    internal passes): every frame's edge list equals the oracle's, bit for bit."""
    n, nf = 4000, 200
    box = oracle.synth_box(n)
    frames = oracle.synth_frames(box, 1000, nf)
    sa, ha, ty = oracle.water_site_table(n)
    ctx.set_water_sites(n)
    ctx.set_batch_frames(64)
    res = ctx.hbond_batch(frames)
    ctx.set_batch_frames(0)
    fo = res["frame_offsets"]
    for f in range(nf):
        want, stats, _, nevals = oracle.frame_hbonds(frames[f], sa, ha, ty)
        got = res["edges"][int(fo[f]):int(fo[f + 1])]
        assert len(got) == len(want), f
        assert np.array_equal(got["i"], want["i"]) and np.array_equal(got["j"], want["j"]), f
        assert np.array_equal(got["length"].view(np.uint32), want["length"].view(np.uint32)), f
        assert np.array_equal(got["cosine"].view(np.uint32), want["cosine"].view(np.uint32)), f
        assert np.allclose(got["energy"], want["energy"], rtol=1e-6, atol=0), f
        assert res["stats"]["pair_evals"][f] == nevals
