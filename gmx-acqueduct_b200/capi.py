"""ctypes binding of the C ABI in include/wn_b200.h (libwn_b200.so).

Used by tests/, bench.py and __graft_entry__.py to drive the library exactly as a
host program would: plain pointers and sizes.  There is no CPU fallback here or in
the library -- if the shared object is missing, or no B200 is present, calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# WN_B200_LIB: load another build of the same library (kernel experiments); default = the in-tree build
LIB_PATH = os.environ.get("WN_B200_LIB") or os.path.join(HERE, "libwn_b200.so")

WN_OK = 0
# status codes of include/wn_b200.h
WN_ERR_CUDA, WN_ERR_OOM, WN_ERR_BAD_ARG, WN_ERR_NO_SITES, WN_ERR_OVERFLOW, WN_ERR_NCCL, WN_ERR_NO_DEVICE = 1, 2, 3, 4, 5, 6, 7
DONOR, ACCEPTOR, WATER = 0, 1, 2
MODEL_TIP3P, MODEL_SPC = 0, 1

EDGE_DTYPE = np.dtype([("i", "<i4"), ("j", "<i4"), ("energy", "<f4"),
                       ("length", "<f4"), ("cosine", "<f4")])
EDGE16_DTYPE = np.dtype([("j", "<i4"), ("energy", "<f4"), ("length", "<f4"), ("cosine", "<f4")])
EDGE12_DTYPE = np.dtype([("j", "<i4"), ("length", "<f4"), ("cosine", "<f4")])
COMPONENT_DTYPE = np.dtype([("n_components", "<i4"), ("largest", "<i4"), ("n_nodes", "<i4"), ("n_edges", "<i4")])
LAYOUT_EDGES, LAYOUT_CSR16, LAYOUT_CSR12 = 0, 1, 2
STATS_DTYPE = np.dtype([("num_sites", "<f8"), ("num_edges", "<f8"), ("ratio", "<f8"),
                        ("sum_energy", "<f8"), ("pair_evals", "<f8")])
assert EDGE_DTYPE.itemsize == 20 and STATS_DTYPE.itemsize == 40

# every symbol include/wn_b200.h declares
SYMBOLS = [
    "wn_create", "wn_destroy", "wn_last_error", "wn_version", "wn_find_pairs", "wn_find_pairs_pbc",
    "wn_hbond_batch_pbc", "wn_hbond_batch_pbc_dev", "wn_hbond_batch_subset", "wn_hbond_batch_ex", "wn_hbond_batch_box9", "wn_find_pairs_box9", "wn_set_find_mode", "wn_last_find_used_cells", "wn_make_edges", "wn_format_edges", "wn_components", "wn_set_sites",
    "wn_set_water_sites", "wn_set_energy_params", "wn_hbond_batch", "wn_hbond_batch_dev",
    "wn_set_batch_frames", "wn_dev_alloc", "wn_dev_free", "wn_host_alloc_pinned",
    "wn_host_free_pinned", "wn_memcpy_d2h", "wn_memcpy_h2d", "wn_synth_frames_dev",
    "wn_device_count", "wn_set_result_sets", "wn_bind_thread_to_device", "wn_pcie_probe", "wn_pcie_probe_mix",
    "wn_get_timings", "wn_selftest_div", "wn_timer_start", "wn_timer_stop", "wn_nccl_unique_id", "wn_nccl_init", "wn_stats_reduce",
]


class TieReport(C.Structure):
    _fields_ = [("pair_cutoff_equal", C.c_uint64), ("length_equal_cut", C.c_uint64),
                ("cosine_argmax_equal", C.c_uint64), ("pos_zero", C.c_uint64),
                ("pair_cutoff_near", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class BatchResult(C.Structure):
    _fields_ = [("edges", C.c_void_p), ("frame_offsets", C.c_void_p), ("stats", C.c_void_p),
                ("n_edges", C.c_uint64), ("n_pair_evals", C.c_uint64), ("ties", TieReport),
                ("n_frames", C.c_int32), ("on_device", C.c_int32),
                ("row_offsets", C.c_void_p), ("n_sites", C.c_int32), ("layout", C.c_int32),
                ("n_candidate_tests", C.c_uint64)]


class BatchArgs(C.Structure):
    _fields_ = [("frames", C.c_void_p), ("frames_on_device", C.c_int32), ("n_frames", C.c_int32),
                ("natoms", C.c_int32), ("box_stride", C.c_int32), ("boxes", C.c_void_p),
                ("subset_offsets", C.c_void_p), ("subset_sites", C.c_void_p),
                ("result_layout", C.c_int32), ("results_on_device", C.c_int32)]


class Timings(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("bin_ms", C.c_float), ("edges_ms", C.c_float),
                ("finish_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
                ("launches", C.c_int32), ("passes", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class WnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("wn_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """dlopen libwn_b200.so (no compute).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            LIB_PATH + " not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, u64, i64 = C.c_void_p, C.c_int32, C.c_uint64, C.c_int64
    L.wn_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.wn_destroy.argtypes = [vp]
    L.wn_destroy.restype = None
    L.wn_last_error.argtypes = [vp]
    L.wn_last_error.restype = C.c_char_p
    L.wn_version.restype = C.c_char_p
    L.wn_find_pairs.argtypes = [vp, vp, i32, i32, C.POINTER(vp), C.POINTER(u64), C.POINTER(TieReport)]
    L.wn_find_pairs_pbc.argtypes = [vp, vp, i32, i32, vp, C.POINTER(vp), C.POINTER(u64), C.POINTER(TieReport)]
    L.wn_hbond_batch_pbc.argtypes = [vp, vp, vp, i32, i32, C.POINTER(BatchResult)]
    L.wn_hbond_batch_pbc_dev.argtypes = [vp, vp, vp, i32, i32, C.POINTER(BatchResult)]
    L.wn_hbond_batch_subset.argtypes = [vp, vp, vp, i32, i32, vp, vp, C.POINTER(BatchResult)]
    L.wn_hbond_batch_ex.argtypes = [vp, C.POINTER(BatchArgs), C.POINTER(BatchResult)]
    L.wn_hbond_batch_box9.argtypes = [vp, vp, vp, i32, i32, C.POINTER(BatchResult)]
    L.wn_find_pairs_box9.argtypes = [vp, vp, i32, i32, vp, C.POINTER(vp), C.POINTER(u64), C.POINTER(TieReport)]
    L.wn_pcie_probe_mix.argtypes = [vp, C.c_size_t, C.c_size_t, i32, C.POINTER(C.c_double)]
    L.wn_make_edges.argtypes = [vp, vp, i32, vp, vp, u64, C.POINTER(vp), C.POINTER(u64), C.POINTER(u64), C.POINTER(TieReport)]
    L.wn_set_find_mode.argtypes = [vp, i32]
    L.wn_last_find_used_cells.argtypes = [vp]
    L.wn_format_edges.argtypes = [vp, vp, vp, i32, C.POINTER(vp), C.POINTER(u64)]
    L.wn_set_sites.argtypes = [vp, i32, vp, vp, i32, vp, i32]
    L.wn_set_water_sites.argtypes = [vp, i32, i32]
    L.wn_set_energy_params.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double]
    L.wn_hbond_batch.argtypes = [vp, vp, i32, i32, C.POINTER(BatchResult)]
    L.wn_hbond_batch_dev.argtypes = [vp, vp, i32, i32, C.POINTER(BatchResult)]
    L.wn_set_batch_frames.argtypes = [vp, i32]
    L.wn_device_count.argtypes = []
    L.wn_set_result_sets.argtypes = [vp, i32]
    L.wn_bind_thread_to_device.argtypes = [vp, C.POINTER(i32)]
    L.wn_pcie_probe.argtypes = [vp, C.c_size_t, i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.wn_dev_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.wn_dev_free.argtypes = [vp, vp]
    L.wn_host_alloc_pinned.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.wn_host_free_pinned.argtypes = [vp, vp]
    L.wn_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.wn_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.wn_synth_frames_dev.argtypes = [vp, i32, u64, i32, i64, i32, vp]
    L.wn_get_timings.argtypes = [vp, C.POINTER(Timings)]
    L.wn_selftest_div.argtypes = [vp, u64, u64, C.POINTER(u64)]
    L.wn_components.argtypes = [vp, C.c_float, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.wn_timer_start.argtypes = [vp]
    L.wn_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.wn_nccl_unique_id.argtypes = [vp]
    L.wn_nccl_init.argtypes = [vp, vp, i32, i32]
    L.wn_stats_reduce.argtypes = [vp, vp, i64]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("wn_destroy", "wn_last_error", "wn_version"):
            fn.restype = C.c_int
    _lib = L
    return L


class Context:
    """One wn_ctx: single caller, one GPU (include/wn_b200.h)."""

    def __init__(self, device=0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.wn_create(device, C.byref(h))
        if rc != WN_OK:
            raise WnError(rc, self.L.wn_last_error(None).decode())
        self.h = h
        self.n_sites = None

    def close(self):
        if getattr(self, "h", None):
            self.L.wn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != WN_OK:
            raise WnError(rc, self.L.wn_last_error(self.h).decode())

    # ---- FindPairs::find
    def find_pairs(self, xyz, first_limit=None, box=None):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        n = xyz.shape[0]
        fl = n if first_limit is None else int(first_limit)
        p, m, ties = C.c_void_p(), C.c_uint64(), TieReport()
        if box is None:
            self._ck(self.L.wn_find_pairs(self.h, xyz.ctypes.data, n, fl, C.byref(p), C.byref(m), C.byref(ties)))
        elif np.size(box) == 9:
            box = np.ascontiguousarray(box, dtype=np.float32).reshape(9)
            self._ck(self.L.wn_find_pairs_box9(self.h, xyz.ctypes.data, n, fl, box.ctypes.data, C.byref(p),
                                               C.byref(m), C.byref(ties)))
        else:
            box = np.ascontiguousarray(box, dtype=np.float32).reshape(3)
            self._ck(self.L.wn_find_pairs_pbc(self.h, xyz.ctypes.data, n, fl, box.ctypes.data, C.byref(p),
                                              C.byref(m), C.byref(ties)))
        cnt = int(m.value)
        if cnt == 0:
            return np.zeros((0, 2), np.int32), ties.as_dict()
        buf = (C.c_int32 * (2 * cnt)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.int32).reshape(cnt, 2).copy(), ties.as_dict()

    def make_edges(self, frame, pairs, box9=None):
        """make_edges of the reference on a caller-supplied pair list, in the given order and orientation:
        (edges, n_pair_evals, ties)."""
        frame = np.ascontiguousarray(frame, dtype=np.float32).reshape(-1, 3)
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        b = None if box9 is None else np.ascontiguousarray(box9, dtype=np.float32).reshape(9)
        e, n, ev, ties = C.c_void_p(), C.c_uint64(), C.c_uint64(), TieReport()
        self._ck(self.L.wn_make_edges(self.h, frame.ctypes.data, frame.shape[0], None if b is None else b.ctypes.data,
                                      pairs.ctypes.data if len(pairs) else None, len(pairs), C.byref(e), C.byref(n),
                                      C.byref(ev), C.byref(ties)))
        cnt = int(n.value)
        if cnt == 0:
            return np.zeros(0, EDGE_DTYPE), int(ev.value), ties.as_dict()
        buf = (C.c_char * (cnt * EDGE_DTYPE.itemsize)).from_address(e.value)
        return np.frombuffer(buf, dtype=EDGE_DTYPE).copy(), int(ev.value), ties.as_dict()

    def set_find_mode(self, mode):
        """0: automatic, 1: all-pairs tile kernel, 2: cell-list sweep (wn_b200.h WN_FIND_*)."""
        self._ck(self.L.wn_set_find_mode(self.h, int(mode)))

    def last_find_used_cells(self):
        return bool(self.L.wn_last_find_used_cells(self.h))

    # ---- site tables
    def set_sites(self, site_atom, h_atom, site_type, first_limit=None):
        site_atom = np.ascontiguousarray(site_atom, dtype=np.int32)
        n = len(site_atom)
        h_atom = np.ascontiguousarray(h_atom, dtype=np.int32).reshape(n, -1) if n else np.zeros((0, 1), np.int32)
        hmax = h_atom.shape[1]
        site_type = np.ascontiguousarray(site_type, dtype=np.uint8)
        fl = n if first_limit is None else int(first_limit)
        self._ck(self.L.wn_set_sites(self.h, n, site_atom.ctypes.data, h_atom.ctypes.data, hmax,
                                     site_type.ctypes.data, fl))
        self.n_sites = n

    def set_water_sites(self, n_waters, first_limit=None):
        fl = n_waters if first_limit is None else int(first_limit)
        self._ck(self.L.wn_set_water_sites(self.h, n_waters, fl))
        self.n_sites = n_waters

    def set_energy_params(self, r_on=5.5, r_off=6.5, theta_on=0.25, theta_off=0.0):
        self._ck(self.L.wn_set_energy_params(self.h, r_on, r_off, theta_on, theta_off))

    def set_batch_frames(self, n):
        self._ck(self.L.wn_set_batch_frames(self.h, int(n)))

    def set_result_sets(self, n):
        self._ck(self.L.wn_set_result_sets(self.h, int(n)))

    def bind_thread_to_device(self):
        n = C.c_int32()
        self._ck(self.L.wn_bind_thread_to_device(self.h, C.byref(n)))
        return int(n.value)

    def pcie_probe(self, nbytes=256 << 20, repeats=4):
        """(h2d GB/s, d2h GB/s) of plain pinned copies on this context's GPU."""
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.wn_pcie_probe(self.h, int(nbytes), int(repeats), C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def pcie_probe_mix(self, d2h_bytes, h2d_bytes, repeats=4):
        """Seconds per repeat of d2h_bytes out and h2d_bytes in, concurrently (pinned memory)."""
        sec = C.c_double()
        self._ck(self.L.wn_pcie_probe_mix(self.h, int(d2h_bytes), int(h2d_bytes), int(repeats), C.byref(sec)))
        return float(sec.value)

    # ---- batches
    def _wrap(self, res, copy=True):
        nf, ne = res.n_frames, int(res.n_edges)
        out = {"n_edges": ne, "n_pair_evals": int(res.n_pair_evals), "n_candidate_tests": int(res.n_candidate_tests),
               "ties": res.ties.as_dict(),
               "n_frames": nf}
        out["n_sites"] = res.n_sites
        out["layout"] = res.layout
        self._last_frames, self._last_on_device = nf, bool(res.on_device)
        edt = EDGE16_DTYPE if res.layout == LAYOUT_CSR16 else EDGE12_DTYPE if res.layout == LAYOUT_CSR12 else EDGE_DTYPE
        if res.on_device:
            out["edges_dev"] = res.edges
            out["frame_offsets_dev"] = res.frame_offsets
            out["stats_dev"] = res.stats
            out["row_offsets_dev"] = res.row_offsets
            return out
        if ne:
            e = np.frombuffer((C.c_char * (ne * edt.itemsize)).from_address(res.edges), dtype=edt)
        else:
            e = np.zeros(0, edt)
        fo = np.frombuffer((C.c_uint64 * (nf + 1)).from_address(res.frame_offsets), dtype=np.uint64)
        st = np.frombuffer((C.c_char * (nf * 40)).from_address(res.stats), dtype=STATS_DTYPE) if nf else \
            np.zeros(0, STATS_DTYPE)
        out["edges"] = e.copy() if copy else e
        out["frame_offsets"] = fo.copy() if copy else fo
        out["stats"] = st.copy() if copy else st
        if copy and nf and res.n_sites and res.row_offsets:
            ro = np.frombuffer((C.c_uint32 * (nf * res.n_sites)).from_address(res.row_offsets), dtype=np.uint32)
            out["row_offsets"] = ro.reshape(nf, res.n_sites).copy()
        return out

    def hbond_batch(self, frames, copy=True, boxes=None):
        """frames: host float32 [F][natoms][3] (numpy) or (ptr, n_frames, natoms);
        boxes: None (open boundary, the reference) or float32 [F][3] (rectangular PBC)."""
        if isinstance(frames, tuple):
            ptr, nf, natoms = frames
        else:
            frames = np.ascontiguousarray(frames, dtype=np.float32)
            if frames.ndim == 2:
                frames = frames[None]
            nf, natoms = frames.shape[0], frames.shape[1]
            ptr = frames.ctypes.data
        res = BatchResult()
        if boxes is None:
            self._ck(self.L.wn_hbond_batch(self.h, ptr, nf, natoms, C.byref(res)))
        elif np.size(boxes) == 9 * nf:
            boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(nf, 9)
            self._ck(self.L.wn_hbond_batch_box9(self.h, ptr, boxes.ctypes.data, nf, natoms, C.byref(res)))
        else:
            boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(nf, 3)
            self._ck(self.L.wn_hbond_batch_pbc(self.h, ptr, boxes.ctypes.data, nf, natoms, C.byref(res)))
        return self._wrap(res, copy)

    def hbond_batch_subset(self, frames, subset_offsets, subset_sites, boxes=None):
        """Per-frame site lists: frame f uses table rows subset_sites[off[f]:off[f+1]]."""
        frames = np.ascontiguousarray(frames, dtype=np.float32)
        nf, natoms = frames.shape[0], frames.shape[1]
        off = np.ascontiguousarray(subset_offsets, dtype=np.int64)
        sites = np.ascontiguousarray(subset_sites, dtype=np.int32)
        bptr = None
        if boxes is not None:
            boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(nf, 3)
            bptr = boxes.ctypes.data
        res = BatchResult()
        self._ck(self.L.wn_hbond_batch_subset(self.h, frames.ctypes.data, bptr, nf, natoms, off.ctypes.data,
                                              sites.ctypes.data if len(sites) else None, C.byref(res)))
        return self._wrap(res, True)

    def hbond_batch_ex(self, frames, n_frames=None, natoms=None, on_device=False, boxes=None,
                       subset_offsets=None, subset_sites=None, layout=LAYOUT_EDGES, copy=True, host_ptr=None,
                       results_on_device=False):
        """General form (wn_hbond_batch_ex): host array or device pointer, open / (L) / matrix boxes,
        optional per-frame site lists."""
        a = BatchArgs()
        keep = []
        a.result_layout = layout
        a.results_on_device = 1 if results_on_device else 0
        if on_device:
            a.frames, a.n_frames, a.natoms, a.frames_on_device = frames, n_frames, natoms, 1
        elif host_ptr is not None:
            a.frames, a.n_frames, a.natoms, a.frames_on_device = host_ptr, n_frames, natoms, 0
        else:
            frames = np.ascontiguousarray(frames, dtype=np.float32); keep.append(frames)
            a.frames, a.n_frames, a.natoms, a.frames_on_device = frames.ctypes.data, frames.shape[0], frames.shape[1], 0
        if boxes is not None:
            boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(a.n_frames, -1); keep.append(boxes)
            a.box_stride, a.boxes = boxes.shape[1], boxes.ctypes.data
        if subset_offsets is not None:
            off = np.ascontiguousarray(subset_offsets, dtype=np.int64); keep.append(off)
            sites = np.ascontiguousarray(subset_sites, dtype=np.int32); keep.append(sites)
            a.subset_offsets, a.subset_sites = off.ctypes.data, (sites.ctypes.data if len(sites) else None)
        res = BatchResult()
        self._ck(self.L.wn_hbond_batch_ex(self.h, C.byref(a), C.byref(res)))
        return self._wrap(res, copy)

    def hbond_batch_dev(self, frames_dev, n_frames, natoms, boxes=None):
        res = BatchResult()
        if boxes is None:
            self._ck(self.L.wn_hbond_batch_dev(self.h, frames_dev, n_frames, natoms, C.byref(res)))
        else:
            boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(n_frames, 3)
            self._ck(self.L.wn_hbond_batch_pbc_dev(self.h, frames_dev, boxes.ctypes.data, n_frames, natoms,
                                                   C.byref(res)))
        return self._wrap(res)

    def fetch_dev_result(self, out):
        """Copy a device-resident batch result to numpy arrays."""
        nf, ne = out["n_frames"], out["n_edges"]
        e = np.zeros(ne, EDGE16_DTYPE if out.get("layout") == LAYOUT_CSR16 else
                     EDGE12_DTYPE if out.get("layout") == LAYOUT_CSR12 else EDGE_DTYPE)
        fo = np.zeros(nf + 1, np.uint64)
        st = np.zeros(nf, STATS_DTYPE)
        if ne:
            self.d2h(e, out["edges_dev"])
        self.d2h(fo, out["frame_offsets_dev"])
        if nf:
            self.d2h(st, out["stats_dev"])
        return e, fo, st

    def components(self, threshold=0.0, n_frames=None):
        """Connected components of the last batch (wn_components): (labels [F][N] int32, stats [F] structured).
        For a device-resident batch the arrays are fetched to the host here."""
        lab, st = C.c_void_p(), C.c_void_p()
        self._ck(self.L.wn_components(self.h, C.c_float(threshold), C.byref(lab), C.byref(st)))
        F, N = self._last_frames if n_frames is None else n_frames, self.n_sites
        labels = np.zeros((F, N), np.int32)
        stats = np.zeros(F, COMPONENT_DTYPE)
        if self._last_on_device:
            if labels.size:
                self.d2h(labels, lab.value)
            if F:
                self.d2h(stats, st.value)
        else:
            if labels.size:
                labels[...] = np.frombuffer((C.c_char * (labels.size * 4)).from_address(lab.value), dtype=np.int32).reshape(F, N)
            if F:
                stats[...] = np.frombuffer((C.c_char * (F * 16)).from_address(st.value), dtype=COMPONENT_DTYPE)
        return labels, stats

    def format_edges(self, frame_numbers, site_ids=None):
        """Text of the edge file for the last batch result (bytes), formatted on the GPU."""
        fn = np.ascontiguousarray(frame_numbers, dtype=np.int32)
        ids = None if site_ids is None else np.ascontiguousarray(site_ids, dtype=np.uint32)
        t, n = C.c_void_p(), C.c_uint64()
        self._ck(self.L.wn_format_edges(self.h, fn.ctypes.data, None if ids is None else ids.ctypes.data,
                                        0 if ids is None else len(ids), C.byref(t), C.byref(n)))
        return C.string_at(t.value, int(n.value)) if n.value else b""

    def format_edges_raw(self, frame_numbers, site_ids=None):
        """As format_edges, without copying: (address of the library-owned pinned text, bytes)."""
        fn = np.ascontiguousarray(frame_numbers, dtype=np.int32)
        ids = None if site_ids is None else np.ascontiguousarray(site_ids, dtype=np.uint32)
        t, n = C.c_void_p(), C.c_uint64()
        self._ck(self.L.wn_format_edges(self.h, fn.ctypes.data, None if ids is None else ids.ctypes.data,
                                        0 if ids is None else len(ids), C.byref(t), C.byref(n)))
        return t.value, int(n.value)

    def timings(self):
        t = Timings()
        self._ck(self.L.wn_get_timings(self.h, C.byref(t)))
        return t.as_dict()

    def selftest_div(self, n_samples, seed=1):
        bad = C.c_uint64()
        self._ck(self.L.wn_selftest_div(self.h, int(n_samples), int(seed), C.byref(bad)))
        return int(bad.value)

    def timer_start(self):
        self._ck(self.L.wn_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.L.wn_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    # ---- plumbing
    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self._ck(self.L.wn_dev_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, ptr):
        self._ck(self.L.wn_dev_free(self.h, ptr))

    def pinned_array(self, shape, dtype):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._ck(self.L.wn_host_alloc_pinned(self.h, n, C.byref(p)))
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        return arr, p.value

    def pinned_free(self, ptr):
        self._ck(self.L.wn_host_free_pinned(self.h, ptr))

    def d2h(self, dst_np, src_dev):
        self._ck(self.L.wn_memcpy_d2h(self.h, dst_np.ctypes.data, src_dev, dst_np.nbytes))

    def h2d(self, dst_dev, src_np):
        src_np = np.ascontiguousarray(src_np)
        self._ck(self.L.wn_memcpy_h2d(self.h, dst_dev, src_np.ctypes.data, src_np.nbytes))

    def synth_frames_dev(self, n_waters, frame0, n_frames, frames_dev, seed=1234, model=MODEL_TIP3P):
        self._ck(self.L.wn_synth_frames_dev(self.h, n_waters, seed, model, frame0, n_frames, frames_dev))

    # ---- multi-GPU statistics
    def nccl_unique_id(self):
        buf = C.create_string_buffer(128)
        rc = self.L.wn_nccl_unique_id(buf)
        if rc != WN_OK:
            raise WnError(rc, self.L.wn_last_error(None).decode())
        return buf.raw

    def nccl_init(self, id128, rank, nranks):
        self._ck(self.L.wn_nccl_init(self.h, id128, rank, nranks))

    def stats_reduce(self, stats):
        assert stats.dtype == STATS_DTYPE and stats.flags.c_contiguous
        self._ck(self.L.wn_stats_reduce(self.h, stats.ctypes.data, len(stats)))
        return stats


def shard_frames(n_frames, rank, world):
    """Contiguous frame block of `rank`: [floor(r*F/G), floor((r+1)*F/G))."""
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world
