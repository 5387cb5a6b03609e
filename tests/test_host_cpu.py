"""CPU tests (-m "not gpu") of the boundary and the host logic: the C-ABI library
loads and exports every declared symbol, fails loudly without a GPU (no CPU
fallback), frame sharding and the statistics reduction pattern (gloo, world 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol(capi):
    import __graft_entry__ as g
    g.build()
    L = capi.load()
    header = open(os.path.join(ROOT, "include", "wn_b200.h")).read()
    declared = set(re.findall(r"\b(wn_[a-z0-9_]+)\s*\(", header))
    declared -= {"wn_ctx"}
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.wn_version().decode().endswith("sm_100a")


def test_struct_layouts_match_header(capi):
    assert ctypes.sizeof(capi.TieReport) == 40
    assert ctypes.sizeof(capi.BatchResult) == 8 * 3 + 8 * 2 + 40 + 8 + 8 + 8 + 8
    assert capi.EDGE_DTYPE.itemsize == 20 and capi.STATS_DTYPE.itemsize == 40 and capi.EDGE16_DTYPE.itemsize == 16
    assert ctypes.sizeof(capi.BatchArgs) == 8 + 4 * 4 + 8 * 3 + 8
    assert ctypes.sizeof(capi.Timings) == 32


def test_kernels_are_sm_100a_only():
    """The shipped object holds sm_100a SASS and nothing else (no multi-arch dispatch)."""
    lib = os.path.join(ROOT, "gmx-acqueduct_b200", "libwn_b200.so")
    out = subprocess.run(["cuobjdump", "--list-elf", lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_no_cpu_fallback_without_gpu(capi):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(capi.WnError) as ei:
        capi.Context(0)
    assert ei.value.code == 7 and "no CPU fallback" in str(ei.value)


def test_product_never_touches_the_oracle():
    """Nothing under gmx-acqueduct_b200/ or include/ may reference oracle/."""
    bad = []
    for base in ("gmx-acqueduct_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".o", ".so", ".pyc")):
                    continue
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"wn_oracle|wno_|oracle/|libwn_ref", txt):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_shard_frames_partition(capi):
    for F in (0, 1, 7, 10000, 12345):
        for G in (1, 2, 3, 4, 8):
            blocks = [capi.shard_frames(F, r, G) for r in range(G)]
            assert blocks[0][0] == 0 and blocks[-1][1] == F
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import wn_pkg, wn_oracle_py as o
capi = wn_pkg.load_capi()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
F, n = 7, 40
lo, hi = capi.shard_frames(F, rank, world)
box = o.synth_box(n)
sa, ha, ty = o.water_site_table(n)
stats = np.zeros(F, capi.STATS_DTYPE)
edges = []
for f in range(lo, hi):                       # this rank's shard, computed by the oracle (CPU stand-in)
    e, st, npairs, ev = o.frame_hbonds(o.synth_frames(box, f, 1)[0], sa, ha, ty)
    stats[f] = (st[0], st[1], st[2], st[3], ev)
    edges.append(e)
# the reduction pattern of wn_stats_reduce: zero-padded [F][5] fp64, SUM over ranks
t = torch.from_numpy(stats.view(np.float64).reshape(F, 5).copy())
dist.all_reduce(t, op=dist.ReduceOp.SUM)
full = t.numpy()
if rank == 0:
    want = np.zeros((F, 5))
    for f in range(F):
        e, st, npairs, ev = o.frame_hbonds(o.synth_frames(box, f, 1)[0], sa, ha, ty)
        want[f] = (st[0], st[1], st[2], st[3], ev)
    assert np.array_equal(full, want), (full, want)   # each row owned by exactly one rank: exact
    print("GLOO_OK")
dist.barrier()
dist.destroy_process_group()
"""


def test_frame_sharding_and_stats_reduction_world2_gloo(tmp_path, oracle):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout + r.stderr


def test_bench_reference_arm_contract(oracle):
    """`bench.py --impl reference` prints one JSON line with the required keys (tiny run)."""
    import json
    env = dict(os.environ, WN_BENCH_REF_WATERS="300")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--waters", "300"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_edge_file_writer_is_byte_identical_to_reference_iostream_formatting():
    """f1: edges-frames.xvg.dat / nodes-list.xvg.dat text, checked in C++ against std::ostream
    driven with the reference's manipulators (src/waternetwork.cpp:485-499, :638-652, :83-93)."""
    exe = os.path.join(ROOT, "gmx-acqueduct_b200", "test_writer")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "test_writer OK" in r.stdout, r.stdout + r.stderr


def test_frame_copy_pool_is_byte_exact():
    """include/frame-copy-pool.hpp (the helper threads FrameBatcher::push can share a frame copy with): odd sizes and
    offsets, back-to-back and spaced copies, 0-3 helpers, byte for byte; clean shutdown."""
    exe = os.path.join(ROOT, "gmx-acqueduct_b200", "test_copy_pool")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "test_copy_pool OK" in r.stdout, r.stdout + r.stderr


def test_make_sites_rules():
    """f3: makeSites typing rules (src/waternetwork.cpp:95-180) restated on plain arrays."""
    exe = os.path.join(ROOT, "gmx-acqueduct_b200", "test_sites")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "test_sites OK" in r.stdout, r.stdout + r.stderr


def test_python_status_codes_match_the_header():
    """status codes match the header (capi.py mirrors include/wn_b200.h)."""
    import re
    import wn_pkg
    capi = wn_pkg.load_capi()
    hdr = open(os.path.join(ROOT, "include", "wn_b200.h")).read()
    codes = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"#define (WN_(?:OK|ERR_[A-Z_]+))\s+(\d+)", hdr))
    assert len(codes) >= 8
    for name, value in codes.items():
        assert getattr(capi, name) == value, name


def test_gromacs_module_passes_the_front_end_against_the_api_shim():
    """GROMACS is absent: the compile-gated module is at least run through g++ -fsyntax-only against
    tests/shim/gromacs (declarations restated from the public API) and, ungated, compiles to nothing."""
    import subprocess
    src = os.path.join(ROOT, "gmx-acqueduct_b200", "host", "gromacs", "waternetwork-b200.cpp")
    inc = ["-I" + os.path.join(ROOT, "include")]
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-DWN_WITH_GROMACS",
                        "-I" + os.path.join(ROOT, "tests", "shim")] + inc + [src], capture_output=True, text=True)
    assert r.returncode == 0 and "error" not in r.stderr, r.stderr
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only"] + inc + [src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_host_energy_header_matches_the_oracle_bit_for_bit(tmp_path, oracle):
    """include/wn_energy.h (energy on demand for the 12-byte layout) == the oracle's computeEnergy restatement,
    float-rounded, on a 16k-point (r, cos) grid: both follow the reference's pow() arithmetic
    (src/waternetwork.cpp:12-66).  The oracle is linked here as the checker only."""
    import subprocess
    exe = str(tmp_path / "t_energy")
    src = os.path.join(ROOT, "tests", "c", "test_energy_header.c")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L" + os.path.join(ROOT, "oracle"), "-lwn_oracle", "-lm",
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("0 of "), r.stdout + r.stderr


# ---------------------------------------------------------------- device arithmetic restated on the CPU
def test_cell_list_reject_bound_is_a_superset_of_the_exact_pair_test():
    """wn_k_cp_sweep rejects candidates with an fp32 test on origin-relative float copies (csrc/wn_kernels_cellpairs.cu);
    the bound comes from cellpair_grid (csrc/wn_capi.cu): e = 2^-20 (amax + lmax + 2.1),
    bound = (4.225 + 2 sqrt(3 * 4.225) e + 3 e^2)(1 + 2^-21), rounded up to float.  Restated in numpy (float32 ops round
    like the device's; both a fused and an unfused accumulation are tried): no pair the reference's fp64 test accepts
    (10 * d2 < 42.25) may be rejected, for systems from 1 nm to 100 um across and pairs hugging the cutoff."""
    rng = np.random.default_rng(5)
    for amax in (1.0, 10.0, 100.0, 1.0e3, 1.0e4, 1.0e5):
        for lmax, pbc in ((0.0, False), (amax, True)):
            e = np.ldexp(amax + lmax + 2.1, -20)
            bd = (4.225 + 2.0 * np.sqrt(3.0 * 4.225) * e + 3.0 * e * e) * (1.0 + np.ldexp(1.0, -21))
            bf = np.float32(bd)
            if float(bf) < bd:
                bf = np.nextafter(bf, np.float32(np.inf))
            n = 20000
            a = rng.uniform(0.0, amax, size=(n, 3))
            u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
            # distances hugging the cutoff from below, plus a spread of smaller ones
            r = np.where(rng.random(n) < 0.7, np.sqrt(4.225) * (1.0 - rng.uniform(0, 1e-7, n)), rng.uniform(0, np.sqrt(4.225), n))
            b = a + u * r[:, None]
            d = a - b
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            accepted = 10.0 * d2 < 42.25
            assert accepted.sum() > n // 2
            af, bf32 = a.astype(np.float32), b.astype(np.float32)          # origin 0: the copies the kernel stores
            df = (af - bf32).astype(np.float32)
            if pbc:
                L = np.float32(lmax)
                df = (df - np.rint(df * (np.float32(1.0) / L)).astype(np.float32) * L).astype(np.float32)
            unfused = ((df[:, 0] * df[:, 0] + df[:, 1] * df[:, 1]).astype(np.float32) + df[:, 2] * df[:, 2]).astype(np.float32)
            d64 = df.astype(np.float64)
            fused = ((d64[:, 0] * d64[:, 0] + d64[:, 1] * d64[:, 1]) + d64[:, 2] * d64[:, 2]).astype(np.float32)   # one rounding
            for d2f in (unfused, fused):
                assert np.all(d2f[accepted] <= bf), (amax, pbc, float((d2f[accepted] - bf).max()))
            # and the bound stays tight where coordinates are small: few false positives
            if amax <= 100.0 and not pbc:
                far = 10.0 * d2 >= 42.25 * (1.0 + 1e-3)
                assert not np.any(unfused[far] <= bf)


def test_block_decode_reciprocals_give_exact_quotients():
    """wn_divmod (csrc/wn_kernels_edges.cu): q = umulhi(x, ceil(2^32 / d)) corrected by at most one, for the block
    index -> (bx, cy, cz) decode; exact for every divisor and every block index a pass can have."""
    rng = np.random.default_rng(9)
    ds = np.concatenate([np.arange(1, 600), rng.integers(600, 1 << 20, size=400)]).astype(np.uint64)
    for d in ds:
        m = np.uint64(0xffffffff) if d == 1 else np.uint64(((1 << 32) + int(d) - 1) // int(d))
        x = np.concatenate([np.arange(0, 2000), rng.integers(0, 1 << 31, size=3000),
                            np.array([int(d) - 1, int(d), int(d) + 1, (1 << 31) - 1])]).astype(np.uint64)
        q = ((x * m) >> np.uint64(32)).astype(np.int64)
        r = x.astype(np.int64) - q * int(d)
        q = np.where(r < 0, q - 1, np.where(r >= int(d), q + 1, q))
        r = x.astype(np.int64) - q * int(d)
        assert np.array_equal(q, x.astype(np.int64) // int(d)) and np.all((r >= 0) & (r < int(d)))


def test_fixed_point_energy_sum_range_and_accuracy():
    """wn_k_copy adds every float energy as a multiple of 2^-32 (csrc/wn_kernels_finish.cu).  With the oracle's
    computeEnergy: (1) no edge of at least 0.56 A (WN_SHORT_EDGE_A; shorter ones send the frame to the ordered path)
    reaches 2^17, so 32 rows x 512 slots stay inside 63 bits; (2) the fixed-point sum of a frame's energies equals the
    fp64 sum to ~1e-9 relative; (3) it does not depend on the order."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import wn_oracle_py as oracle
    r = np.concatenate([np.float32(0.56) + np.arange(0, 4000, dtype=np.float32) * np.float32(1e-4), np.linspace(0.96, 6.5, 3000).astype(np.float32)])
    worst = max(abs(oracle.compute_energy(float(x), 1.0)) for x in r)
    assert worst < 2.0 ** 17 and worst * 32 * 512 * 2.0 ** 32 < 2.0 ** 63
    n = 1000
    frame = oracle.synth_frames(oracle.synth_box(n), 0, 1)[0]
    e = oracle.frame_hbonds(frame, *oracle.water_site_table(n))[0]["energy"]
    fix = np.rint(e.astype(np.float64) * 2.0 ** 32).astype(np.int64)
    total = float(fix.sum()) * 2.0 ** -32
    ref = float(np.sum(e.astype(np.float64)))
    assert abs(total - ref) <= 1e-9 * abs(ref)
    perm = np.random.default_rng(1).permutation(len(fix))
    assert int(fix[perm].sum()) == int(fix.sum())
    # (4) wn_k_stats keeps the fixed-point sum only where |sum| * 1e-7 >= n_edges * 2^-33 and sums the float records
    # otherwise: with that rule every subset of a frame's edges -- down to single edges of 1e-8 at the cutoff -- stays
    # inside 1.2e-7 relative of the fp64 sum of its float energies (contract: 1e-6)
    rng = np.random.default_rng(2)
    order = np.argsort(np.abs(e))
    subsets = [order[:k] for k in (1, 2, 3, 5, 10, 40, 200)] + [np.array([k]) for k in order[:60]] + \
              [rng.choice(len(e), size=int(k), replace=False) for k in rng.integers(1, 30, size=200)]
    fallbacks = 0
    for idx in subsets:
        ee = e[idx].astype(np.float64)
        s_fix = float(np.rint(ee * 2.0 ** 32).sum()) * 2.0 ** -32
        keep = abs(s_fix) * 1e-7 >= len(ee) * 2.0 ** -33
        fallbacks += not keep
        got = s_fix if keep else float(ee.sum())
        assert abs(got - float(ee.sum())) <= 1.2e-7 * abs(float(ee.sum()))
    assert fallbacks > 0
