#!/usr/bin/env python
"""bench.py -- frames/s and H-bond pair evaluations/s of the fused hydrogen-bond path.

Workload: BASELINE.json configs[2], the configuration the metric is quoted on: a
periodic-sized 100k-atom box (33,333 TIP3P-like waters, 99,999 atoms), synthetic
frames (include/wn_synth.h, seed 1234), open boundary as in the reference's
BruteFindPairs (no PBC, src/brute-find-pairs.cpp:25).  A "step" is one pass of the
hot path over one batch of --frames-per-step distinct frames; every step reads
different frames (a step's input, 1.2 MB/frame, is far larger than the 126 MB L2).

  python bench.py --gpus N --steps K --warmup W            # this framework
  python bench.py --impl reference --steps K --warmup W    # reference CPU path, host cores

Own arm, per rank (one process per GPU, frame-sharded, weak scaling):
  value  : frames/s with the trajectory shard already resident in HBM
           (wn_hbond_batch_dev), CUDA-event time on the library's stream, max over ranks
  e2e    : frames/s through the host-buffer entry point wn_hbond_batch (pinned host
           frames -> H2D -> kernels -> D2H of edges, offsets, stats), same timing
  roofline: wn_k_edges (the dominant kernel), algorithmic bytes of SURVEY 8d
  cpu_baseline (rank 0, N=1): reference sources (oracle/_ref) or oracle port on host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_WATERS = 33333
SEED = 1234
METRIC = "frames/s (H-bond pair search + energy, 100k-atom water box)"


def synth_box_length(n_waters):
    """Cubic box edge of the synthetic generator (include/wn_synth.h): (N / 33.4 nm^-3)^(1/3),
    Newton iteration in the same order as wn_synth_make_box."""
    m = 1
    while m * m * m < n_waters:
        m += 1
    v = n_waters / 33.4
    x = m * 0.32
    for _ in range(60):
        x = x - (x * x * x - v) / (3.0 * x * x)
    return x


def workload_config(n_waters):
    """The workload both arms run (identical keys and values in the own and the reference line)."""
    return {"workload": "configs[2]: 100k-atom (33,333 TIP3P waters) box, open boundary (reference BruteFindPairs has no PBC)"
            if n_waters == N_WATERS else "custom: %d waters" % n_waters,
            "n_waters": n_waters, "natoms": 3 * n_waters, "seed": SEED,
            "generator": "include/wn_synth.h jittered-lattice water, 33.4 nm^-3, frames 0.. of seed 1234"}


def algorithmic_bytes_per_frame(natoms, edges_per_frame):
    """read the frame once as delivered, write compact edges + stats."""
    return 12.0 * natoms + 20.0 * edges_per_frame + 32.0


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smmax, reasons = [], None, set()
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                clk = float(parts[1]); smmax = float(parts[2])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:      # timed region shorter than a sample period: use every sample
            for ts, line in self.lines:
                parts = [p.strip() for p in line.split(",")]
                try:
                    sm.append(float(parts[1])); smmax = float(parts[2])
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smmax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------- reference (CPU) arm
def edge_digest(e):
    """Order-sensitive digest of the bit-exact part of an edge list: (i, j, length, cosine) bytes."""
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    for k in ("i", "j", "length", "cosine"):
        h.update(np.ascontiguousarray(e[k]).tobytes())
    return h.hexdigest()


def cpu_reference_run(frames_host, threads, keep_energy=False):
    """Reference path on the host: brute-find-pairs + make_edges per frame, frames
    fanned out over `threads` host threads (ctypes releases the GIL).  Uses the
    reference's own sources (oracle/_ref) when that build is present, else the
    oracle port.  Returns (seconds, kind, edges_total, evals_total)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import wn_oracle_py as oracle
    oracle.lib()
    ref = oracle.ref()
    kind = "reference" if ref is not None else "port"
    nf = frames_host.shape[0]
    n = frames_host.shape[1] // 3
    sa, ha, ty = oracle.water_site_table(n)
    results = [None] * nf

    def work(f):
        if ref is not None:
            e, npairs = oracle.ref_frame_hbonds_water(frames_host[f])
            results[f] = (len(e), None, edge_digest(e), e["energy"].copy() if keep_energy else None)
        else:
            e, st, npairs, ev = oracle.frame_hbonds(frames_host[f], sa, ha, ty)
            results[f] = (len(e), ev, edge_digest(e), e["energy"].copy() if keep_energy else None)

    t0 = time.perf_counter()
    pending = list(range(nf))
    lock = threading.Lock()

    def loop():
        while True:
            with lock:
                if not pending:
                    return
                f = pending.pop()
            work(f)

    ths = [threading.Thread(target=loop) for _ in range(min(threads, nf))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return dt, kind, sum(r[0] for r in results), results


def delaunay_proxy_run(frames_host, n_frames=2):
    """DelaunayFindPairs stand-in (the reference's CGAL finder cannot be built here: CGAL is
    absent): scipy.spatial.Delaunay (Qhull) edge set + oracle make_edges, one thread.  It
    reproduces python/edges.dat of the reference exactly but is NOT the
    reference's code path; reported for context only."""
    try:
        from scipy.spatial import Delaunay
    except Exception:
        return None
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import wn_oracle_py as oracle
    n = frames_host.shape[1] // 3
    sa, ha, ty = oracle.water_site_table(n)
    t0 = time.perf_counter()
    edges = 0
    for f in range(min(n_frames, frames_host.shape[0])):
        site_xyz, h_xyz, h_off = oracle.gather_sites(frames_host[f], sa, ha, 2)
        tri = Delaunay(site_xyz)
        s = tri.simplices
        pr = np.concatenate([s[:, [0, 1]], s[:, [0, 2]], s[:, [0, 3]], s[:, [1, 2]], s[:, [1, 3]], s[:, [2, 3]]])
        pr.sort(axis=1)
        pr = np.unique(pr, axis=0).astype(np.int32)
        e, _ = oracle.make_edges(pr, site_xyz, h_xyz, h_off, ty)
        edges += len(e)
    dt = time.perf_counter() - t0
    nf = min(n_frames, frames_host.shape[0])
    return {"value": nf / dt, "unit": "frames/s", "cores": 1, "kind": "proxy (scipy/Qhull Delaunay + oracle make_edges)",
            "sample": "%d frame(s), %.1f s; %.1f edges/site -- a different (sparser) edge list than the brute path"
                      % (nf, dt, edges / nf / n)}


def usable_threads():
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    # each in-flight frame holds the 17M-pair list twice (~0.5 GB): bound by memory
    try:
        with open("/proc/meminfo") as fh:
            avail_kb = [int(l.split()[1]) for l in fh if l.startswith("MemAvailable")][0]
        cores = max(1, min(cores, int(avail_kb / 1e6 / 1.0)))
    except Exception:
        pass
    return cores


def host_frames(n_frames, n_waters, frame0=0):
    """Synthetic frames on the host for the CPU arms (oracle-side generator)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import wn_oracle_py as oracle
    box = oracle.synth_box(n_waters, seed=SEED)
    return oracle.synth_frames(box, frame0, n_frames)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = usable_threads()
    per_step = threads                      # one frame per host thread per step
    frames = host_frames(per_step, args.waters)
    for _ in range(args.warmup):
        cpu_reference_run(frames[:min(per_step, threads)], threads)
    t = 0.0
    kind = "port"
    for _ in range(args.steps):
        dt, kind, _, _ = cpu_reference_run(frames, threads)
        t += dt
    fps = args.steps * per_step / t
    sample = "%d frame(s) per step, one per host thread, brute-find-pairs + make_edges" % per_step
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.waters),
        "run": {"frames_per_step": per_step, "frames_total": args.steps * per_step,
                "parallelism": "%d host threads, one frame each" % threads},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- own arm
def run_own(args):
    import wn_pkg
    capi = wn_pkg.load_capi()
    if not os.path.exists(capi.LIB_PATH):          # fresh checkout: compile (nvcc, in-tree); never a CPU fallback
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            import __graft_entry__ as g
            g.build()
        else:
            while not os.path.exists(capi.LIB_PATH):
                time.sleep(1.0)
            time.sleep(2.0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist      # control plane only: barrier + max of timings
        dist.init_process_group("gloo", rank=rank, world_size=world)

    ctx = capi.Context(local_rank)
    n_waters, natoms = args.waters, 3 * args.waters
    F, K, W = args.frames_per_step, args.steps, args.warmup
    ctx.set_water_sites(n_waters)
    ctx.set_batch_frames(args.pass_frames)      # 0: the library's own pass schedule inside a step

    if world > 1:
        ids = [ctx.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.nccl_init(ids[0], rank, world)

    # ---- trajectory shard resident in HBM: (W+K) distinct batches for this rank
    frame_bytes = natoms * 12
    n_local = (W + K) * F
    frames_dev = ctx.dev_alloc(n_local * frame_bytes)
    frame0 = rank * n_local                    # weak scaling: each rank owns its own frames
    for s in range(W + K):
        ctx.synth_frames_dev(n_waters, frame0 + s * F, F, frames_dev + s * F * frame_bytes, seed=SEED)

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- device-resident leg
    for s in range(W):
        ctx.hbond_batch_dev(frames_dev + s * F * frame_bytes, F, natoms)
    if world > 1:
        # warm-up of the collective as well (NCCL connects its channels on the first call)
        ctx.stats_reduce(np.zeros(world * F, capi.STATS_DTYPE))
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    edges_total = evals_total = tests_total = launches = passes = 0
    ties_total = {}
    edges_ms = bin_ms = finish_ms = 0.0
    stats_rows = []
    t0 = time.perf_counter()
    ctx.timer_start()
    for s in range(W, W + K):
        out = ctx.hbond_batch_dev(frames_dev + s * F * frame_bytes, F, natoms)
        edges_total += out["n_edges"]; evals_total += out["n_pair_evals"]; tests_total += out["n_candidate_tests"]
        for k, v in out["ties"].items():
            ties_total[k] = ties_total.get(k, 0) + v
        tm = ctx.timings()
        launches += tm["launches"]; passes += tm["passes"]
        edges_ms += tm["edges_ms"]; bin_ms += tm["bin_ms"]; finish_ms += tm["finish_ms"]
    # per-frame statistics of the last step -> NCCL sum over ranks: the path's only collective, INSIDE the timed
    # region (once per run, as in a trajectory job: the .xvg tables are written at the end)
    stats_reduce = None
    if world > 1:
        t_r0 = time.perf_counter()
        st = np.zeros(world * F, capi.STATS_DTYPE)
        mine = np.zeros(F, capi.STATS_DTYPE)
        ctx.d2h(mine, out["stats_dev"])
        st[rank * F:(rank + 1) * F] = mine
        ctx.stats_reduce(st)
        t_r1 = time.perf_counter()
    dev_ms = ctx.timer_stop()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    barrier()
    if world > 1:
        # every row is owned by exactly one rank, so the sum must return each rank's rows bit for bit
        own = st[rank * F:(rank + 1) * F]
        for name in capi.STATS_DTYPE.names:
            if not np.array_equal(own[name].view(np.uint64), mine[name].view(np.uint64)):
                raise SystemExit("bench: wn_stats_reduce changed this rank's own rows (%s)" % name)
        if not (np.all(st["num_sites"] == n_waters) and np.all(st["num_edges"] > 0) and np.all(st["sum_energy"] < 0)):
            raise SystemExit("bench: reduced statistics table has empty rows")
        # cross-rank check: all ranks hold the same table
        digest = [float(st["num_edges"].sum()), float(st["sum_energy"].sum())]
        alld = [None] * world
        dist.all_gather_object(alld, digest)
        if any(d != alld[0] for d in alld):
            raise SystemExit("bench: ranks disagree on the reduced statistics table")
        stats_reduce = {"rows": int(world * F), "bytes": int(st.nbytes), "ms_incl_d2h": (t_r1 - t_r0) * 1e3,
                        "inside_timed_region": True, "own_rows_bit_exact": True, "ranks_agree": True}

    # ---- sustained leg: the same step repeated for >= 1 s (>= 250 steps cycling over the resident frames: every
    # step still reads 300 MB of distinct input), with its own clock samples
    sustained = None
    if not args.no_sustained:
        K2 = max(250, int(1.2e3 / max(dev_ms / K, 1e-3)))
        sampler2 = ClockSampler(local_rank)
        sampler2.start()
        time.sleep(0.25)
        barrier()
        ts0 = time.perf_counter()
        ctx.timer_start()
        for s in range(K2):
            ctx.hbond_batch_dev(frames_dev + (s % (W + K)) * F * frame_bytes, F, natoms)
        sus_ms = ctx.timer_stop()
        ts1 = time.perf_counter()
        clocks2 = sampler2.stop(ts0, ts1)
        barrier()
        sus_ms_max = sus_ms
        if dist is not None:
            import torch
            tt = torch.tensor([sus_ms_max], dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sus_ms_max = float(tt[0])
        sustained = {"value": K2 * F * world / (sus_ms_max * 1e-3), "unit": "frames/s", "steps": K2,
                     "seconds": sus_ms_max * 1e-3, "ms_per_step": sus_ms_max / K2, "clocks": clocks2}

    # ---- periodic leg (north-star extension): same frames, rectangular PBC with the box of
    # the generator; fewer steps, reported beside the headline (the reference arm is open boundary)
    periodic = None
    if not args.no_pbc:
        Lbox = synth_box_length(n_waters)
        boxes = np.tile(np.array([Lbox, Lbox, Lbox], np.float32), (F, 1))
        kp = max(1, min(K, 5))
        ctx.hbond_batch_dev(frames_dev, F, natoms, boxes=boxes)           # warm-up (buffers, row capacity)
        ctx.hbond_batch_dev(frames_dev + F * frame_bytes, F, natoms, boxes=boxes)
        barrier()
        pe = pv = 0
        ctx.timer_start()
        for s in range(kp):
            o2 = ctx.hbond_batch_dev(frames_dev + ((W + s) % (W + K)) * F * frame_bytes, F, natoms, boxes=boxes)
            pe += o2["n_edges"]; pv += o2["n_pair_evals"]
        pms = ctx.timer_stop()
        periodic = {"steps": kp, "box_nm": float(Lbox), "ms": pms, "edges": pe, "evals": pv}

    def allmax(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allsum(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    dev_ms_max = allmax(dev_ms)
    frames_all = K * F * world
    fps = frames_all / (dev_ms_max * 1e-3)
    evals_all = allsum(float(evals_total))
    tests_all = allsum(float(tests_total))
    edges_all = allsum(float(edges_total))
    edges_per_frame = edges_all / frames_all

    # ---- standalone CudaFindPairs::find at the reference cutoff (the parity vehicle of the
    # FindPairs boundary): host points in, host pair list out; B_find = 24 N + 8 pairs
    find = None
    if rank == 0 and not args.no_find:
        one = np.empty((natoms, 3), np.float32)
        ctx.d2h(one, frames_dev)
        xyz = np.ascontiguousarray(one[::3], dtype=np.float64)
        ctx.find_pairs(xyz)                                   # warm-up (allocations)
        t_0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            pairs, _t = ctx.find_pairs(xyz)
            tmf = ctx.timings()
        wall = (time.perf_counter() - t_0) / reps
        used_cells = ctx.last_find_used_cells()
        ctx.set_find_mode(1)                                  # the exact all-pairs tile kernel, for comparison
        ctx.find_pairs(xyz)
        pairs_ap, _ = ctx.find_pairs(xyz)
        tm_ap = ctx.timings()
        ctx.set_find_mode(0)
        if not np.array_equal(pairs_ap, pairs):
            raise SystemExit("bench: the cell-list pair search and the all-pairs kernel disagree")
        bytes_find = 24.0 * n_waters + 8.0 * len(pairs)
        find = {"pairs": int(len(pairs)), "ties": _t if isinstance(_t, dict) else None,
                "kernels_ms": tmf["total_ms"], "d2h_ms": tmf["d2h_ms"],
                "wall_ms_incl_numpy_copy": wall * 1e3, "algorithmic_bytes": bytes_find,
                "kernel_GBps": bytes_find / (tmf["total_ms"] * 1e-3) / 1e9,
                "frac_of_hbm_peak": bytes_find / (tmf["total_ms"] * 1e-3) / 1e9 / float(
                    json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
                if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None,
                "cell_list": bool(used_cells), "all_pairs_kernels_ms": tm_ap["total_ms"],
                "note": "cell-list sweep at the reference's 2.0555 nm cutoff (bin + count + scan + fill, exact fp64 test "
                        "on every candidate); all_pairs_kernels_ms = the O(N^2) tile kernel on the same points, "
                        "pair lists compared equal; output D2H of 8 B/pair dominates a call"}

    # ---- edge-file text of a batch, formatted on the GPU (37 bytes per edge row, D2H included)
    text = None
    if rank == 0 and not args.no_find:
        nft = min(F, 50)
        ctx.hbond_batch_dev(frames_dev, nft, natoms)
        numbers = np.arange(nft, dtype=np.int32)
        ctx.format_edges_raw(numbers)                           # warm-up (buffers)
        t_0 = time.perf_counter()
        _, nbytes = ctx.format_edges_raw(numbers)
        dtx = time.perf_counter() - t_0
        text = {"frames": nft, "bytes": nbytes, "frames_per_s": nft / dtx, "GBps": nbytes / dtx / 1e9,
                "note": "rows of edges-frames.xvg.dat formatted by wn_k_format_rows, D2H into pinned host memory "
                        "included; host formatter (EdgeFileWriter): ~0.4 GB/s per core, iostream: ~0.03 GB/s"}

    # ---- connected components of the networks of a device-resident batch (the first step of the
    # reference's downstream script, python/network-analysis.py): union-find on the edge list in HBM
    comps = None
    if rank == 0 and not args.no_find:
        outc = ctx.hbond_batch_dev(frames_dev, F, natoms)
        import ctypes
        lab, stc = ctypes.c_void_p(), ctypes.c_void_p()
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peak_gbs = float(json.load(fh).get("hbm_gbs", 6650.0))
        except Exception:
            peak_gbs = 6650.0
        ctx._ck(ctx.L.wn_components(ctx.h, 0.0, ctypes.byref(lab), ctypes.byref(stc)))      # warm-up (buffers)
        ctx.timer_start()
        reps = 5
        for _ in range(reps):
            ctx._ck(ctx.L.wn_components(ctx.h, 0.0, ctypes.byref(lab), ctypes.byref(stc)))
        cms = ctx.timer_stop() / reps
        st_c = np.zeros(F, capi.COMPONENT_DTYPE)
        ctx.d2h(st_c, stc.value)
        cc_bytes = outc["n_edges"] * 12 + F * n_waters * 4      # i, j, energy of every edge + one label per site
        comps = {"frames": F, "ms": cms, "frames_per_s": F / (cms * 1e-3), "edges_per_s": outc["n_edges"] / (cms * 1e-3),
                 "algorithmic_bytes": cc_bytes, "GBps": cc_bytes / (cms * 1e-3) / 1e9,
                 "frac_of_hbm_peak": cc_bytes / (cms * 1e-3) / 1e9 / peak_gbs,
                 "components_frame0": int(st_c["n_components"][0]), "largest_frame0": int(st_c["largest"][0]),
                 "note": "wn_components at threshold 0.0 on the HBM-resident edge list (one CTA per frame, union-find in shared memory)"}

    if periodic is not None:
        pmax = allmax(periodic["ms"])
        periodic = {"value": periodic["steps"] * F * world / (pmax * 1e-3), "unit": "frames/s",
                    "steps": periodic["steps"], "box_nm": periodic["box_nm"],
                    "pair_evals_per_frame": allsum(float(periodic["evals"])) / (periodic["steps"] * F * world),
                    "edges_per_frame": allsum(float(periodic["edges"])) / (periodic["steps"] * F * world),
                    "note": "rectangular PBC, minimum image as defined in include/wn_b200.h; no reference counterpart"}

    # ---- end-to-end leg: pinned host frames -> wn_hbond_batch -> pinned host results
    ring = min(4, W + K)
    host_arr, host_ptr = ctx.pinned_array((ring, F, natoms, 3), np.float32)
    for r in range(ring):
        ctx.d2h(host_arr[r], frames_dev + r * F * frame_bytes)
    # one call per step; inside the call the library pipelines sub-passes (H2D of pass k+1 and
    # D2H of pass k-1 under the kernels of pass k), so give it several passes per step
    e2e_pass = max(F // 5, 1) if args.e2e_pass < 0 else args.e2e_pass      # 0 = the library's own schedule
    ctx.set_batch_frames(e2e_pass)
    e2e_layout = capi.LAYOUT_CSR16          # 16 bytes per edge over PCIe (row index via the CSR offsets)
    for s in range(min(W, 2)):
        ctx.hbond_batch_ex(None, F, natoms, host_ptr=host_ptr + (s % ring) * F * frame_bytes, layout=e2e_layout, copy=False)
    barrier()
    ctx.timer_start()
    e2e_t0 = time.perf_counter()
    d2h_bytes = 0
    e2e_launches = 0
    for s in range(K):
        res = ctx.hbond_batch_ex(None, F, natoms, host_ptr=host_ptr + (s % ring) * F * frame_bytes, layout=e2e_layout, copy=False)
        d2h_bytes += res["n_edges"] * 16 + (F + 1) * 8 + F * 40 + F * n_waters * 4
        e2e_launches += ctx.timings()["launches"]
        # the step's result is read on the host: total edge count of the last frame
        _ = int(res["frame_offsets"][-1])
    e2e_ms = ctx.timer_stop()
    e2e_wall = time.perf_counter() - e2e_t0
    e2e_ms_max = allmax(max(e2e_ms, e2e_wall * 1e3))
    e2e_fps = frames_all / (e2e_ms_max * 1e-3)

    # ---- PCIe ceiling: plain pinned copies on every rank AT THE SAME TIME (shared uplinks show here), and what
    # the e2e leg's D2H volume would take at that rate
    barrier()
    h2d_gbs, d2h_gbs = ctx.pcie_probe(256 << 20, 4)
    barrier()
    d2h_all = allsum(d2h_gbs)
    d2h_min = -allmax(-d2h_gbs)
    # the step's own traffic mix (results out, frames in, both directions at once), every rank at the same time
    barrier()
    mix_s = ctx.pcie_probe_mix(int(d2h_bytes / K), F * frame_bytes, 4)
    barrier()
    mix_s_max = allmax(mix_s)
    pcie = {"h2d_gbs_this_rank": h2d_gbs, "d2h_gbs_this_rank": d2h_gbs, "pcie_d2h_ceiling_gbs": d2h_all,
            "d2h_gbs_slowest_rank": d2h_min,
            "step_copy_ms_this_rank": mix_s * 1e3, "step_copy_ms_slowest_rank": mix_s_max * 1e3,
            "e2e_ceiling_frames_per_s": world * F / mix_s_max,
            "e2e_frac_of_ceiling": e2e_fps / (world * F / mix_s_max),
            "note": "ceiling = a step's D2H and H2D volumes as plain concurrent pinned copies, all ranks at once, "
                    "slowest rank (shared PCIe uplinks show here); *_gbs = one direction at a time"}

    # ---- the C++ drop-in itself: FrameBatcher (include/hydrogen-bond-edges.hpp) fed frame by frame, its sink called
    # per frame -- the binary a maintainer's analyzeFrame patch would behave like (host/wn-bench.cpp)
    e2e_cpp = None
    if rank == 0 and world == 1 and not args.no_cpp:
        exe = os.path.join(ROOT, "gmx-acqueduct_b200", "wn_bench")
        if os.path.exists(exe):
            ctx.set_batch_frames(0)
            try:
                pr = subprocess.run([exe, "-waters", str(n_waters), "-batch", str(F), "-batches", "12", "-warmup", "2",
                                     "-layout", "16", "-find", "0"], capture_output=True, text=True, timeout=600)
                if pr.returncode == 0 and pr.stdout.strip():
                    e2e_cpp = json.loads(pr.stdout.strip().splitlines()[-1])
                else:
                    e2e_cpp = {"error": (pr.stderr or "no output")[-300:]}
            except Exception as exc:        # noqa: BLE001 -- the leg is informative, not the headline
                e2e_cpp = {"error": str(exc)[-300:]}
            # the same drop-in with what PCIe allows at best: the 12-byte result layout ({j, length, cosine}; the energy
            # is a function of the last two, include/wn_energy.h recomputes it on demand on the host) and a producer that
            # decodes straight into the pinned staging buffer (stage() + commit(), no host copy per frame).  Its own
            # process with its own time limit: nothing here can disturb the numbers above.
            try:
                pr = subprocess.run([exe, "-waters", str(n_waters), "-batch", str(F), "-batches", "12", "-warmup", "2",
                                     "-layout", "12", "-mode", "stage", "-find", "0"], capture_output=True, text=True, timeout=300)
                if pr.returncode == 0 and pr.stdout.strip():
                    d12 = json.loads(pr.stdout.strip().splitlines()[-1])
                    e2e_cpp["csr12_stage"] = {k: d12.get(k) for k in ("e2e_cpp", "unit", "mode", "layout_bytes_per_edge", "timed_frames",
                                                                      "timed_s", "d2h_bytes_per_frame", "in_frame_order", "last_call")}
                else:
                    e2e_cpp["csr12_stage"] = {"error": (pr.stderr or "no output")[-300:]}
            except Exception as exc:        # noqa: BLE001
                if isinstance(e2e_cpp, dict):
                    e2e_cpp["csr12_stage"] = {"error": str(exc)[-300:]}

    # ---- host frames in -> component labels out: the edge list never crosses PCIe (results_on_device),
    # only 4 bytes per site and 16 per frame come back.  A different result than the headline e2e (labels and
    # component statistics instead of the edge list), reported beside it.
    e2e_cc = None
    if rank == 0 and not args.no_find:
        import ctypes
        lab_h, lab_hp = ctx.pinned_array((F, n_waters), np.int32)
        st_h, st_hp = ctx.pinned_array((F,), capi.COMPONENT_DTYPE)
        lab, stc = ctypes.c_void_p(), ctypes.c_void_p()

        def one(s):
            ctx.hbond_batch_ex(None, F, natoms, host_ptr=host_ptr + (s % ring) * F * frame_bytes, results_on_device=True, copy=False)
            ctx._ck(ctx.L.wn_components(ctx.h, 0.0, ctypes.byref(lab), ctypes.byref(stc)))
            ctx._ck(ctx.L.wn_memcpy_d2h(ctx.h, lab_hp, lab, F * n_waters * 4))
            ctx._ck(ctx.L.wn_memcpy_d2h(ctx.h, st_hp, stc, F * 16))
            return int(st_h["n_components"][-1])

        one(0)
        kc = max(1, min(K, 5))
        ctx.timer_start()
        t_0 = time.perf_counter()
        for s in range(kc):
            one(s + 1)
        cc_ms = max(ctx.timer_stop(), (time.perf_counter() - t_0) * 1e3)
        e2e_cc = {"value": kc * F / (cc_ms * 1e-3), "unit": "frames/s", "steps": kc,
                  "h2d_bytes_per_step": F * frame_bytes, "d2h_bytes_per_step": F * n_waters * 4 + F * 16 + (F + 1) * 8 + F * 40,
                  "note": "pinned host frames -> wn_hbond_batch_ex(results_on_device) -> wn_components(0.0) -> labels + "
                          "component statistics in pinned host memory; the edge list stays in HBM"}
        ctx.pinned_free(lab_hp); ctx.pinned_free(st_hp)


    # ---- roofline of the dominant kernel (wn_k_edges), this rank
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    bytes_per_launch = algorithmic_bytes_per_frame(natoms, edges_total / (K * F)) * (K * F / max(passes, 1))
    k_ms = edges_ms / max(passes, 1)
    achieved = bytes_per_launch / (k_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        if tj.get("n_waters") == n_waters and tj.get("frames_per_launch") == F:
            traffic = tj.get("wn_k_edges_dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "wn_k_edges", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "frac_of_nominal_8000": achieved / 8000.0,
                "bytes_per_launch": bytes_per_launch, "kernel_ms_per_launch": k_ms,
                "frames_per_launch": K * F / max(passes, 1),
                "pipeline_frac": (algorithmic_bytes_per_frame(natoms, edges_per_frame) * fps / world / 1e9) / peak,
                "stage_ms_per_step": {"bin": bin_ms / K, "edges": edges_ms / K, "finish": finish_ms / K}}

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same frames
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = usable_threads()
        nfr = max(1, min(threads, 64))
        sample = np.empty((nfr, natoms, 3), np.float32)
        ctx.d2h(sample, frames_dev)
        dt, kind, _, cres = cpu_reference_run(sample, threads, keep_energy=True)
        dt1, _, _, _ = cpu_reference_run(sample[:1], 1)
        # the frames the CPU arm just processed, through the GPU path: same edge lists?
        outp = ctx.hbond_batch_dev(frames_dev, nfr, natoms)
        ge, gfo, _gst = ctx.fetch_dev_result(outp)
        bad = []
        for f in range(nfr):
            gf = ge[int(gfo[f]):int(gfo[f + 1])]
            ok = len(gf) == cres[f][0] and edge_digest(gf) == cres[f][2]
            if ok and len(gf):
                we = cres[f][3].astype(np.float64)
                ok = bool(np.all(np.abs(gf["energy"].astype(np.float64) - we) <= 1e-6 * np.abs(we)))
            if not ok:
                bad.append(f)
        parity = {"frames": nfr, "edges": int(sum(r[0] for r in cres)), "ok": not bad,
                  "checked": "edge count and the bytes of (i, j, length, cosine) per frame vs the %s CPU result "
                             "of the same frames; energy <= 1e-6 relative" % kind}
        if bad:
            parity["bad_frames"] = bad
        chunked = None
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import wn_oracle_py as oracle
        if oracle.ref() is not None and hasattr(oracle.ref(), "wnref_frame_hbonds_water_chunked"):
            t_c = time.perf_counter()
            ec, _npairs, secs = oracle.ref_frame_hbonds_water_chunked(sample[0], threads)
            t_c = time.perf_counter() - t_c
            chunked = {"value": 1.0 / t_c, "unit": "frames/s", "cores": threads,
                       "find_s": secs[0], "make_edges_s": secs[1],
                       "same_edges_as_frame_per_thread": bool(len(ec) == cres[0][0] and edge_digest(ec) == cres[0][2]),
                       "note": "the reference's own per-frame schedule (src/waternetwork.cpp:583-633): single-threaded "
                               "brute find, then hardware_concurrency() std::async chunks of make_edges joined in order"}
        cpu = {"value": nfr / dt, "unit": "frames/s", "cores": min(threads, nfr), "kind": kind,
               "sample": "%d frame(s) of the same 100k-atom workload, one per host thread "
                         "(brute-find-pairs O(N^2) + make_edges), %.1f s wall" % (nfr, dt),
               "single_thread_frames_per_s": 1.0 / dt1,
               "intra_frame_chunked": chunked,
               "delaunay_find_pairs": delaunay_proxy_run(sample)}

    ctx.pinned_free(host_ptr)
    ctx.dev_free(frames_dev)
    # per-rank view (diagnostics: a slow rank shows here, the headline uses the max)
    mine_rank = {"rank": rank, "ms_per_step": dev_ms / K, "wall_ms_per_step": (t1 - t0) * 1e3 / K,
                 "edges_ms": edges_ms / K, "bin_ms": bin_ms / K, "finish_ms": finish_ms / K,
                 "e2e_ms_per_step": e2e_ms / K, "sm_mhz": clocks.get("sm_mhz") if isinstance(clocks, dict) else None,
                 "reasons": clocks.get("reasons") if isinstance(clocks, dict) else None}
    per_rank = [mine_rank]
    if dist is not None:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine_rank)
    if rank == 0:
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(n_waters),
            "run": {"frames_per_step": F, "frames_total": frames_all,
                    "l2_policy": "inputs larger than L2: every step reads %d distinct frames (%.0f MB)" % (F, F * frame_bytes / 1e6),
                    "parallelism": "frame-sharded x%d" % world},
            "pair_evals_per_s": evals_all / (dev_ms_max * 1e-3),
            "candidate_tests_per_s": tests_all / (dev_ms_max * 1e-3),
            "candidate_tests_per_frame": tests_all / frames_all,
            "ties_rank0": ties_total,      # exact-comparison ties met in the timed frames (SURVEY 8d evidence)
            "edges_per_s": edges_all / (dev_ms_max * 1e-3),
            "pair_evals_per_frame": evals_all / frames_all, "edges_per_frame": edges_per_frame,
            "wall_ms_per_step": (t1 - t0) * 1e3 / K,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": F * frame_bytes,
                    "d2h_bytes_per_step": d2h_bytes / K, "host_ring_batches": ring,
                    "frames_per_internal_pass": e2e_pass,
                    "result_layout": "WN_LAYOUT_CSR16 (16 B/edge + 4 B/site row offsets; wn_edge AoS is 20 B/edge)"},
            "gpu_launches": launches,
            "sustained": sustained,
            "stats_reduce": stats_reduce,
            "pcie": pcie,
            "e2e_cpp": e2e_cpp,
            "clocks": clocks,
            "per_rank": per_rank if world > 1 else None,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity_in_bench": parity,
            "periodic": periodic,
            "find_pairs": find,
            "edge_file_text": text,
            "components": comps,
            "e2e_components": e2e_cc,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    if parity is not None and not parity["ok"]:
        print("bench.py: GPU edge lists differ from the CPU reference on frames %s" % parity.get("bad_frames"), file=sys.stderr)
        return 3
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--waters", type=int, default=N_WATERS)
    ap.add_argument("--frames-per-step", type=int, default=250)
    ap.add_argument("--e2e-pass", type=int, default=0, help="frames per internal pass of the e2e leg (0: library default)")
    ap.add_argument("--pass-frames", type=int, default=0,
                    help="frames per internal pass of the device-resident leg (0: library default, a quarter of ~8M site-frames)")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 1 s sustained leg")
    ap.add_argument("--no-cpp", action="store_true", help="skip the C++ drop-in leg (wn_bench)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-pbc", action="store_true", help="skip the periodic-boundary leg")
    ap.add_argument("--no-find", action="store_true", help="skip the standalone CudaFindPairs::find leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
