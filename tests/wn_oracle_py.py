"""ctypes bindings for the CPU oracle (oracle/libwn_oracle.so) and, when built,
the reference-source build (oracle/_ref/libwn_ref.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

EDGE_DTYPE = np.dtype([("i", "<i4"), ("j", "<i4"), ("energy", "<f4"),
                       ("length", "<f4"), ("cosine", "<f4")])
assert EDGE_DTYPE.itemsize == 20

DONOR, ACCEPTOR, WATER = 0, 1, 2
MODEL_TIP3P, MODEL_SPC = 0, 1


class Ties(C.Structure):
    _fields_ = [("pair_cutoff_equal", C.c_uint64), ("length_equal_cut", C.c_uint64),
                ("cosine_argmax_equal", C.c_uint64), ("pos_zero", C.c_uint64),
                ("pair_cutoff_near", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class SynthBox(C.Structure):
    _fields_ = [("n_waters", C.c_int32), ("m", C.c_int32), ("box", C.c_double),
                ("a", C.c_double), ("seed", C.c_uint64), ("model", C.c_int32)]


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "all"])
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


_lib = None
_ref = None


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "libwn_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        L.wno_switch_function_radius.restype = C.c_double
        L.wno_switch_function_radius.argtypes = [C.c_double] * 3
        L.wno_switch_function_angle.restype = C.c_double
        L.wno_switch_function_angle.argtypes = [C.c_double] * 3
        L.wno_compute_energy.restype = C.c_double
        L.wno_compute_energy.argtypes = [C.c_double] * 2
        L.wno_compute_energy_ex.restype = C.c_double
        L.wno_compute_energy_ex.argtypes = [C.c_double] * 6
        L.wno_brute_find_pairs_limited.restype = C.c_size_t
        L.wno_brute_find_pairs_limited.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                                   C.c_size_t, C.c_void_p]
        L.wno_make_edges.restype = C.c_size_t
        L.wno_make_edges.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.wno_gather_sites.restype = None
        L.wno_gather_sites.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.wno_water_site_table.restype = None
        L.wno_water_site_table.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.wno_frame_hbonds.restype = C.c_size_t
        L.wno_frame_hbonds.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.wno_synth_box.restype = None
        L.wno_synth_box.argtypes = [C.c_int32, C.c_uint64, C.c_int32, C.POINTER(SynthBox)]
        L.wno_synth_frames.restype = None
        L.wno_synth_frames.argtypes = [C.POINTER(SynthBox), C.c_int64, C.c_int32, C.c_void_p]
        L.wno_frame_hbonds_pbc.restype = C.c_size_t
        L.wno_frame_hbonds_pbc.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.wno_brute_find_pairs_pbc.restype = C.c_size_t
        L.wno_brute_find_pairs_pbc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_size_t, C.c_void_p]
        L.wno_wrap_frame.restype = None
        L.wno_wrap_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.wno_frame_hbonds_tric.restype = C.c_size_t
        L.wno_frame_hbonds_tric.argtypes = L.wno_frame_hbonds_pbc.argtypes
        L.wno_brute_find_pairs_tric.restype = C.c_size_t
        L.wno_brute_find_pairs_tric.argtypes = L.wno_brute_find_pairs_pbc.argtypes
        _lib = L
    return _lib


def ref():
    """The reference's own sources compiled against the CGAL shim, or None."""
    global _ref
    if _ref is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libwn_ref.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.wnref_switch_function_radius.restype = C.c_double
        R.wnref_switch_function_radius.argtypes = [C.c_double] * 3
        R.wnref_switch_function_angle.restype = C.c_double
        R.wnref_switch_function_angle.argtypes = [C.c_double] * 3
        R.wnref_compute_energy.restype = C.c_double
        R.wnref_compute_energy.argtypes = [C.c_double] * 2
        R.wnref_brute_find_pairs.restype = C.c_size_t
        R.wnref_brute_find_pairs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        R.wnref_make_edges.restype = C.c_size_t
        R.wnref_make_edges.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        R.wnref_frame_hbonds_water.restype = C.c_size_t
        R.wnref_frame_hbonds_water.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        if hasattr(R, "wnref_frame_hbonds_water_chunked"):
            R.wnref_frame_hbonds_water_chunked.restype = C.c_size_t
            R.wnref_frame_hbonds_water_chunked.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                                           C.c_void_p, C.c_void_p]
        _ref = R
    return _ref


# ---------------------------------------------------------------- oracle calls
def compute_energy(r, cosine):
    return lib().wno_compute_energy(float(r), float(cosine))


def brute_find_pairs(xyz, first_limit=None, ties=None):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
    n = xyz.shape[0]
    fl = n if first_limit is None else int(first_limit)
    tp = C.byref(ties) if ties is not None else None
    m = lib().wno_brute_find_pairs_limited(xyz.ctypes.data, n, fl, None, 0, None)
    pairs = np.empty((m, 2), dtype=np.int32)
    lib().wno_brute_find_pairs_limited(xyz.ctypes.data, n, fl, pairs.ctypes.data, m, tp)
    return pairs


def make_edges(pairs, site_xyz, h_xyz, h_off, site_type, ties=None):
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    site_xyz = np.ascontiguousarray(site_xyz, dtype=np.float64).reshape(-1, 3)
    h_xyz = np.ascontiguousarray(h_xyz, dtype=np.float64).reshape(-1, 3)
    h_off = np.ascontiguousarray(h_off, dtype=np.int32)
    site_type = np.ascontiguousarray(site_type, dtype=np.uint8)
    out = np.empty(max(len(pairs), 1), dtype=EDGE_DTYPE)
    ev = C.c_uint64(0)
    tp = C.byref(ties) if ties is not None else None
    ne = lib().wno_make_edges(pairs.ctypes.data, len(pairs), site_xyz.ctypes.data,
                              site_xyz.shape[0], h_xyz.ctypes.data, h_off.ctypes.data,
                              site_type.ctypes.data, out.ctypes.data, C.byref(ev), tp)
    return out[:ne].copy(), int(ev.value)


def gather_sites(frame, site_atom, h_atom, hmax):
    frame = np.ascontiguousarray(frame, dtype=np.float32).reshape(-1, 3)
    site_atom = np.ascontiguousarray(site_atom, dtype=np.int32)
    h_atom = np.ascontiguousarray(h_atom, dtype=np.int32).reshape(-1, hmax)
    n = len(site_atom)
    site_xyz = np.empty((n, 3), dtype=np.float64)
    h_xyz = np.empty((max(n * hmax, 1), 3), dtype=np.float64)
    h_off = np.empty(n + 1, dtype=np.int32)
    lib().wno_gather_sites(frame.ctypes.data, frame.shape[0], site_atom.ctypes.data,
                           h_atom.ctypes.data, hmax, n, site_xyz.ctypes.data,
                           h_xyz.ctypes.data, h_off.ctypes.data)
    return site_xyz, h_xyz[:h_off[n]].copy(), h_off


def water_site_table(n_waters):
    site_atom = np.empty(n_waters, dtype=np.int32)
    h_atom = np.empty((n_waters, 2), dtype=np.int32)
    site_type = np.empty(n_waters, dtype=np.uint8)
    lib().wno_water_site_table(n_waters, site_atom.ctypes.data, h_atom.ctypes.data,
                               site_type.ctypes.data)
    return site_atom, h_atom, site_type


def frame_hbonds(frame, site_atom, h_atom, site_type, first_limit=None, ties=None,
                 cap=None):
    """One frame end to end (gather + brute pairs + make_edges + stats)."""
    frame = np.ascontiguousarray(frame, dtype=np.float32).reshape(-1, 3)
    site_atom = np.ascontiguousarray(site_atom, dtype=np.int32)
    n = len(site_atom)
    h_atom = np.ascontiguousarray(h_atom, dtype=np.int32).reshape(n, -1) if n else \
        np.zeros((0, 1), np.int32)
    hmax = h_atom.shape[1]
    site_type = np.ascontiguousarray(site_type, dtype=np.uint8)
    fl = n if first_limit is None else int(first_limit)
    if cap is None:
        cap = max(n * (n - 1) // 2, 1) if n <= 4096 else 64 * n
    edges = np.empty(cap, dtype=EDGE_DTYPE)
    stats = np.zeros(4, dtype=np.float64)
    npairs, nevals = C.c_uint64(0), C.c_uint64(0)
    tp = C.byref(ties) if ties is not None else None
    ne = lib().wno_frame_hbonds(frame.ctypes.data, frame.shape[0], site_atom.ctypes.data,
                                h_atom.ctypes.data, hmax, site_type.ctypes.data, n, fl,
                                edges.ctypes.data, cap, stats.ctypes.data,
                                C.byref(npairs), C.byref(nevals), tp)
    if ne == C.c_size_t(-1).value:
        raise RuntimeError("oracle edge capacity overflow")
    return edges[:ne].copy(), stats, int(npairs.value), int(nevals.value)


def frame_hbonds_rows(frame, site_atom, h_atom, site_type, rows, box=None, cap=None):
    """Edges (i, j > i) of the listed rows of one frame, in list order; box=None: open boundary,
    else rectangular PBC (Lx, Ly, Lz).  Returns (edges, n_pair_evals)."""
    frame = np.ascontiguousarray(frame, dtype=np.float32).reshape(-1, 3)
    site_atom = np.ascontiguousarray(site_atom, dtype=np.int32)
    n = len(site_atom)
    h_atom = np.ascontiguousarray(h_atom, dtype=np.int32).reshape(n, -1) if n else np.zeros((0, 1), np.int32)
    hmax = h_atom.shape[1]
    site_type = np.ascontiguousarray(site_type, dtype=np.uint8)
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    bptr = None
    if box is not None:
        box = np.ascontiguousarray(box, dtype=np.float32).reshape(3)
        bptr = box.ctypes.data
    if cap is None:
        cap = 256 * max(len(rows), 1) + 1024
    edges = np.empty(cap, dtype=EDGE_DTYPE)
    nevals = C.c_uint64(0)
    fn = lib().wno_frame_hbonds_rows
    fn.restype = C.c_size_t
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                   C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
    ne = fn(frame.ctypes.data, frame.shape[0], site_atom.ctypes.data, h_atom.ctypes.data, hmax,
            site_type.ctypes.data, n, rows.ctypes.data, len(rows), bptr, edges.ctypes.data, cap, C.byref(nevals))
    if ne == C.c_size_t(-1).value:
        raise RuntimeError("oracle edge capacity overflow")
    return edges[:ne].copy(), int(nevals.value)


def components(edges, n_sites, threshold=0.0):
    """Connected components of one frame's edge list: (labels[n_sites], stats[4])."""
    edges = np.ascontiguousarray(edges, dtype=EDGE_DTYPE)
    labels = np.empty(max(n_sites, 1), dtype=np.int32)
    stats = np.zeros(4, dtype=np.int32)
    fn = lib().wno_components
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    fn(edges.ctypes.data, len(edges), n_sites, threshold, labels.ctypes.data, stats.ctypes.data)
    return labels[:n_sites].copy(), stats


# ------------------------------------------------------------- synthetic input
def synth_box(n_waters, seed=1234, model=MODEL_TIP3P):
    b = SynthBox()
    lib().wno_synth_box(n_waters, seed, model, C.byref(b))
    return b


def synth_frames(box, frame0, n_frames):
    out = np.empty((n_frames, 3 * box.n_waters, 3), dtype=np.float32)
    lib().wno_synth_frames(C.byref(box), frame0, n_frames, out.ctypes.data)
    return out


# --------------------------------------------------- reference-source build calls
def ref_brute_find_pairs(xyz):
    R = ref()
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
    n = xyz.shape[0]
    cap = n * (n - 1) // 2 + 1
    pairs = np.empty((cap, 2), dtype=np.int32)
    m = R.wnref_brute_find_pairs(xyz.ctypes.data, n, pairs.ctypes.data, cap)
    return pairs[:m].copy()


def ref_make_edges(pairs, site_xyz, h_xyz, h_off, site_type):
    R = ref()
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    site_xyz = np.ascontiguousarray(site_xyz, dtype=np.float64).reshape(-1, 3)
    h_xyz = np.ascontiguousarray(h_xyz, dtype=np.float64).reshape(-1, 3)
    h_off = np.ascontiguousarray(h_off, dtype=np.int32)
    site_type = np.ascontiguousarray(site_type, dtype=np.uint8)
    out = np.empty(max(len(pairs), 1), dtype=EDGE_DTYPE)
    ne = R.wnref_make_edges(pairs.ctypes.data, len(pairs), site_xyz.ctypes.data,
                            site_xyz.shape[0], h_xyz.ctypes.data, h_off.ctypes.data,
                            site_type.ctypes.data, out.ctypes.data)
    return out[:ne].copy()


def ref_frame_hbonds_water(frame, cap=None):
    """Whole-frame reference path for waters (reference sources): edges, n_pairs."""
    R = ref()
    frame = np.ascontiguousarray(frame, dtype=np.float32).reshape(-1, 3)
    n = frame.shape[0] // 3
    if cap is None:
        cap = 64 * max(n, 1)
    out = np.empty(cap, dtype=EDGE_DTYPE)
    npairs = C.c_uint64(0)
    ne = R.wnref_frame_hbonds_water(frame.ctypes.data, n, out.ctypes.data, cap, C.byref(npairs))
    if ne > cap:
        raise RuntimeError("reference edge capacity overflow")
    return out[:ne].copy(), int(npairs.value)


def ref_frame_hbonds_water_chunked(frame, n_chunks, cap=None):
    """The reference's intra-frame schedule (src/waternetwork.cpp:583-633): single-threaded find, then
    n_chunks std::async(make_edges) ranges joined in order.  Returns (edges, n_pairs, (find_s, edges_s))."""
    R = ref()
    frame = np.ascontiguousarray(frame, dtype=np.float32).reshape(-1, 3)
    n = frame.shape[0] // 3
    if cap is None:
        cap = 64 * max(n, 1)
    out = np.empty(cap, dtype=EDGE_DTYPE)
    npairs = C.c_uint64(0)
    secs = (C.c_double * 2)()
    ne = R.wnref_frame_hbonds_water_chunked(frame.ctypes.data, n, int(n_chunks), out.ctypes.data, cap,
                                            C.byref(npairs), secs)
    if ne > cap:
        raise RuntimeError("reference edge capacity overflow")
    return out[:ne].copy(), int(npairs.value), (secs[0], secs[1])


# ------------------------------------------------------------------ periodic mode
def frame_hbonds_pbc(frame, site_atom, h_atom, site_type, box, first_limit=None, ties=None, cap=None):
    """One frame under the rectangular-PBC definition of oracle/wn_oracle.h."""
    frame = np.ascontiguousarray(frame, dtype=np.float32).reshape(-1, 3)
    site_atom = np.ascontiguousarray(site_atom, dtype=np.int32)
    n = len(site_atom)
    h_atom = np.ascontiguousarray(h_atom, dtype=np.int32).reshape(n, -1) if n else np.zeros((0, 1), np.int32)
    hmax = h_atom.shape[1]
    site_type = np.ascontiguousarray(site_type, dtype=np.uint8)
    box = np.ascontiguousarray(box, dtype=np.float32).reshape(3)
    fl = n if first_limit is None else int(first_limit)
    if cap is None:
        cap = max(n * (n - 1) // 2, 1) if n <= 4096 else 64 * n
    edges = np.empty(cap, dtype=EDGE_DTYPE)
    stats = np.zeros(4, dtype=np.float64)
    npairs, nevals = C.c_uint64(0), C.c_uint64(0)
    tp = C.byref(ties) if ties is not None else None
    ne = lib().wno_frame_hbonds_pbc(frame.ctypes.data, frame.shape[0], site_atom.ctypes.data,
                                    h_atom.ctypes.data, hmax, site_type.ctypes.data, n, fl,
                                    box.ctypes.data, edges.ctypes.data, cap, stats.ctypes.data,
                                    C.byref(npairs), C.byref(nevals), tp)
    if ne == C.c_size_t(-1).value:
        raise RuntimeError("oracle edge capacity overflow")
    return edges[:ne].copy(), stats, int(npairs.value), int(nevals.value)


def brute_find_pairs_pbc(xyz, box, first_limit=None, ties=None):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
    box = np.ascontiguousarray(box, dtype=np.float32).reshape(3)
    n = xyz.shape[0]
    fl = n if first_limit is None else int(first_limit)
    tp = C.byref(ties) if ties is not None else None
    m = lib().wno_brute_find_pairs_pbc(xyz.ctypes.data, n, fl, box.ctypes.data, None, 0, None)
    pairs = np.empty((m, 2), dtype=np.int32)
    lib().wno_brute_find_pairs_pbc(xyz.ctypes.data, n, fl, box.ctypes.data, pairs.ctypes.data, m, tp)
    return pairs


def frame_hbonds_tric(frame, site_atom, h_atom, site_type, box9, first_limit=None, ties=None, cap=None):
    """One frame under the triclinic-PBC definition (box9 = GROMACS matrix, row-major)."""
    frame = np.ascontiguousarray(frame, dtype=np.float32).reshape(-1, 3)
    site_atom = np.ascontiguousarray(site_atom, dtype=np.int32)
    n = len(site_atom)
    h_atom = np.ascontiguousarray(h_atom, dtype=np.int32).reshape(n, -1) if n else np.zeros((0, 1), np.int32)
    hmax = h_atom.shape[1]
    site_type = np.ascontiguousarray(site_type, dtype=np.uint8)
    box9 = np.ascontiguousarray(box9, dtype=np.float32).reshape(9)
    fl = n if first_limit is None else int(first_limit)
    if cap is None:
        cap = max(n * (n - 1) // 2, 1) if n <= 4096 else 64 * n
    edges = np.empty(cap, dtype=EDGE_DTYPE)
    stats = np.zeros(4, dtype=np.float64)
    npairs, nevals = C.c_uint64(0), C.c_uint64(0)
    tp = C.byref(ties) if ties is not None else None
    ne = lib().wno_frame_hbonds_tric(frame.ctypes.data, frame.shape[0], site_atom.ctypes.data,
                                     h_atom.ctypes.data, hmax, site_type.ctypes.data, n, fl,
                                     box9.ctypes.data, edges.ctypes.data, cap, stats.ctypes.data,
                                     C.byref(npairs), C.byref(nevals), tp)
    if ne == C.c_size_t(-1).value:
        raise RuntimeError("oracle edge capacity overflow")
    return edges[:ne].copy(), stats, int(npairs.value), int(nevals.value)


def brute_find_pairs_tric(xyz, box9, first_limit=None, ties=None):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
    box9 = np.ascontiguousarray(box9, dtype=np.float32).reshape(9)
    n = xyz.shape[0]
    fl = n if first_limit is None else int(first_limit)
    tp = C.byref(ties) if ties is not None else None
    m = lib().wno_brute_find_pairs_tric(xyz.ctypes.data, n, fl, box9.ctypes.data, None, 0, None)
    pairs = np.empty((m, 2), dtype=np.int32)
    lib().wno_brute_find_pairs_tric(xyz.ctypes.data, n, fl, box9.ctypes.data, pairs.ctypes.data, m, tp)
    return pairs
