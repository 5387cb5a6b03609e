/*
 * wn_kernels_edges.cu -- fused pair search + donor-H...acceptor geometry (sm_100a).
 *
 * One kernel replaces, for every frame of a batch at once,
 *   mp_->find(sitePoints)   reference src/waternetwork.cpp:583
 *                           (BruteFindPairs::find, src/brute-find-pairs.cpp:19-30)
 *   make_edges              reference src/waternetwork.cpp:182-263
 * Every pair that make_edges can turn into an edge has a float length <= 6.5 A
 * (:204), i.e. a squared distance far inside the brute test 10*d2 < 42.25
 * (src/brute-find-pairs.cpp:25), so searching the 27 neighbouring cells of a grid
 * with edge >= 0.65 nm finds exactly the pairs that matter.
 *
 * Work decomposition: one warp per (frame, home cell).  The warp sweeps the
 * candidates of the 27 neighbouring cells (9 contiguous runs of the cell-sorted
 * site array), lane = candidate, inner loop = home sites (broadcast from shared
 * memory).  An fp32 distance prefilter (superset, margin 1e-5) plus `j > i`
 * selects candidates; their (home, candidate) codes are ballot-compacted into a
 * per-warp queue.  Whenever 32 codes are queued the warp switches to the exact
 * fp64 geometry with all lanes busy (the reference's operation order, explicit
 * round-to-nearest intrinsics, no FMA contraction).  Accepted edges are pooled in
 * shared memory, grouped by owning row at the end of the cell, and written to a
 * batch-wide temp array; wn_k_copy then places every row at its final
 * lexicographic position (edge order == pair order, :621-633).
 *
 * Rows that do not fit the pool fall back to one-row-at-a-time processing and, if
 * a single row is still too long, to a count -> allocate -> direct-write path, so
 * arbitrarily dense inputs stay correct.
 */
#include "wn_internal.cuh"

namespace {

constexpr int WARPS_PER_CTA = 8;
constexpr int HC   = 16;     /* home sites per chunk                     */
constexpr int PCAP = 256;    /* pooled edges per chunk                   */
constexpr int QCAP = 32 + 32 * HC;  /* queue: < 32 carried + one round of pushes */
constexpr int MAXRUN = 9;
constexpr unsigned FULL = 0xffffffffu;

/* jh = (home row slot << 27) | partner site index (n_sites < 2^27) */
struct PoolEntry { uint32_t jh; float length; float cosine; };

/* fixed part of the per-warp shared memory */
struct WarpShared {
    float4    hpos[HC];
    uint32_t  hmeta[HC];
    int       hrow[HC];      /* real site index of the home row            */
    int       cnt[HC];
    int       rbase[HC];
    int       fill[HC];
    uint32_t  queue[QCAP];
    int       run_start[MAXRUN + 1];
    int       run_pref[MAXRUN + 1];
    PoolEntry pool[PCAP];
};

enum { MODE_POOL = 0, MODE_COUNT = 1, MODE_DIRECT = 2 };

struct ChunkState {
    int pool_len;
    bool overflow;
    unsigned int evals, tie_len, tie_cos, tie_pos0;
    /* MODE_COUNT / MODE_DIRECT */
    unsigned long long direct_base;
    unsigned int direct_k;
};

/* Exact geometry of make_edges for `nb` queued codes (lane < nb active). */
template <bool TYPED>
__device__ __forceinline__ void
wn_eval_batch(const WnEdgeParams& p, size_t fbase, WarpShared& ws, const double* __restrict__ hdh,
              int qbase, int nb, int mode, ChunkState& cs, int lane)
{
    const bool act = lane < nb;
    bool counted = false, edge = false;
    bool t_len = false, t_cos = false, t_pos0 = false;
    float len_f = 0.f, cosmax = 0.f;
    int h = 0, jidx = 0;
    if (act) {
        const uint32_t code = ws.queue[qbase + lane];
        h = (int)(code >> 27);
        const uint32_t c = code & 0x07ffffffu;
        const float4 a4 = ws.hpos[h];
        const float4 b4 = __ldg(p.spos + fbase + c);
        jidx = __float_as_int(b4.w);
        /* Vector_3 OO = sitePoints.at(first) - sitePoints.at(second)      (:194) */
        const double ox = __dsub_rn((double)a4.x, (double)b4.x);
        const double oy = __dsub_rn((double)a4.y, (double)b4.y);
        const double oz = __dsub_rn((double)a4.z, (double)b4.z);
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ox, ox), __dmul_rn(oy, oy)), __dmul_rn(oz, oz));
        /* float distance = CGAL::sqrt(OO*OO)                               (:195) */
        const float dist_f = __double2float_rn(__dsqrt_rn(d2));
        const double dd = (double)dist_f;
        /* distance *= 10.0                                                 (:197) */
        len_f = __double2float_rn(__dmul_rn(dd, 10.0));
        /* if (distance > 6.5) isedge = false                               (:204) */
        counted = len_f <= WN_LEN_CUT_F;
        if (counted) {
            t_len = (len_f == WN_LEN_CUT_F);
            /* OO /= distance                                               (:196) */
            const double ux = __ddiv_rn(ox, dd), uy = __ddiv_rn(oy, dd), uz = __ddiv_rn(oz, dd);
            const uint32_t ma = ws.hmeta[h];
            const uint32_t mb = __ldg(p.smeta + fbase + c);
            const int nva = (int)((ma >> WN_META_NV_SHIFT) & WN_META_NV_MASK);
            const int nvb = (int)((mb >> WN_META_NV_SHIFT) & WN_META_NV_MASK);
            bool win = false, win0 = false;
            /* hydrogens of the first site: proj = DH*OO                 (:217-225) */
            for (int v = 0; v < nva; ++v) {
                const double* d = hdh + ((size_t)h * p.hs + v) * 3;
                const float cv = __double2float_rn(
                    __dadd_rn(__dadd_rn(__dmul_rn(d[0], ux), __dmul_rn(d[1], uy)), __dmul_rn(d[2], uz)));
                if (cv > cosmax) {                                      /* :242-247 */
                    cosmax = cv; win = true;
                    win0 = (v == 0) && (ma & WN_META_FIRST0);
                } else if (cosmax > 0.f && cv == cosmax) {
                    t_cos = true;
                }
            }
            /* hydrogens of the second site: proj = DH*(-1.0*OO)         (:229-237);
             * negation commutes with every rounding, so proj = -(DH*OO). */
            const double* db = p.sdh + ((fbase + c) * (size_t)p.hs) * 3;
            for (int v = 0; v < nvb; ++v) {
                const double bx = __ldg(db + 3 * v), by = __ldg(db + 3 * v + 1), bz = __ldg(db + 3 * v + 2);
                const float cv = -__double2float_rn(
                    __dadd_rn(__dadd_rn(__dmul_rn(bx, ux), __dmul_rn(by, uy)), __dmul_rn(bz, uz)));
                if (cv > cosmax) {
                    /* list index of this hydrogen is nH(first) + its own position: it is
                     * 0 only when the first site has no hydrogen list at all */
                    cosmax = cv; win = true;
                    win0 = (v == 0) && (mb & WN_META_FIRST0) && !(ma & WN_META_HASH);
                } else if (cosmax > 0.f && cv == cosmax) {
                    t_cos = true;
                }
            }
            t_pos0 = win && win0;
            edge = win && !win0;                                   /* if (pos > 0)  :251 */
        }
    }
    const unsigned m_cnt = __ballot_sync(FULL, counted);
    const unsigned m_edge = __ballot_sync(FULL, edge);
    cs.evals += __popc(m_cnt);
    if (__any_sync(FULL, t_len || t_cos || t_pos0)) {          /* rare */
        cs.tie_len += __popc(__ballot_sync(FULL, t_len));
        cs.tie_cos += __popc(__ballot_sync(FULL, t_cos));
        cs.tie_pos0 += __popc(__ballot_sync(FULL, t_pos0));
    }
    const int ne = __popc(m_edge);
    if (ne == 0) return;
    const int k = __popc(m_edge & ((1u << lane) - 1u));
    if (mode == MODE_POOL) {
        if (cs.pool_len + ne > PCAP) { cs.overflow = true; return; }
        if (edge) {
            PoolEntry e; e.jh = ((uint32_t)h << 27) | (uint32_t)jidx; e.length = len_f; e.cosine = cosmax;
            ws.pool[cs.pool_len + k] = e;
            atomicAdd(&ws.cnt[h], 1);
        }
        cs.pool_len += ne;
    } else if (mode == MODE_COUNT) {
        cs.direct_k += ne;
    } else {
        if (edge) {
            const unsigned long long pos = cs.direct_base + cs.direct_k + k;
            if (pos < p.temp_cap) {
                WnTempEdge t; t.j = jidx; t.length = len_f; t.cosine = cosmax; t.pad = 0;
                p.temp[pos] = t;
            }
        }
        cs.direct_k += ne;
    }
}

/* Sweep all candidates of the neighbourhood against home sites [hb, hb+hc). */
template <bool TYPED>
__device__ __forceinline__ void
wn_sweep(const WnEdgeParams& p, size_t fbase, WarpShared& ws, double* hdh, int hb, int hc,
         int n_runs, int total, int mode, ChunkState& cs, int lane)
{
    /* stage the home sites; pad the chunk to a multiple of 4 with unreachable dummies */
    const int hc4 = (hc + 3) & ~3;
    __syncwarp();
    if (lane < hc4) {
        float4 a = make_float4(1.0e30f, 1.0e30f, 1.0e30f, __int_as_float(0x7fffffff));
        uint32_t m = 0;
        int row = -1;
        if (lane < hc) {
            a = __ldg(p.spos + fbase + hb + lane);
            row = __float_as_int(a.w);
            /* rows >= first_limit own nothing: an index no candidate can exceed */
            if (row >= p.first_limit) a.w = __int_as_float(0x7fffffff);
            m = __ldg(p.smeta + fbase + hb + lane);
        }
        ws.hpos[lane] = a; ws.hrow[lane] = row; ws.hmeta[lane] = m;
    }
    if (lane < HC) { ws.cnt[lane] = 0; ws.fill[lane] = 0; }
    for (int t = lane; t < hc * p.hs * 3; t += 32)
        hdh[t] = __ldg(p.sdh + (fbase + hb) * (size_t)p.hs * 3 + t);
    __syncwarp();

    const float tf = (float)WN_PREFILTER_D2;
    int qlen = 0;
    int run = 0;
    const int rounds = (total + 31) >> 5;
    for (int r = 0; r < rounds; ++r) {
        const int v = r * 32 + lane;
        float cx = -1.0e30f, cy = -1.0e30f, cz = -1.0e30f;
        int cj = -1;
        uint32_t code_c = 0, cm = 0;
        if (v < total) {
            while (run + 1 < n_runs && v >= ws.run_pref[run + 1]) ++run;
            const int c = ws.run_start[run] + (v - ws.run_pref[run]);
            const float4 b4 = __ldg(p.spos + fbase + c);
            cx = b4.x; cy = b4.y; cz = b4.z; cj = __float_as_int(b4.w);
            code_c = (uint32_t)c;
            if (TYPED) cm = __ldg(p.smeta + fbase + c);
        }
        /* lane-private hit mask: bit h = home row h is a candidate partner */
        uint32_t hits = 0;
#pragma unroll
        for (int g = 0; g < HC / 4; ++g) {
            if (g * 4 < hc4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int h = g * 4 + u;
                    const float4 a4 = ws.hpos[h];
                    const float dx = a4.x - cx, dy = a4.y - cy, dz = a4.z - cz;
                    const float d2 = dx * dx + dy * dy + dz * dz;
                    bool pass = (d2 <= tf) && (cj > __float_as_int(a4.w));
                    if (TYPED) {
                        /* type rule (:199-203) and the no-hydrogen rule (:205-208) */
                        const uint32_t am = ws.hmeta[h];
                        const uint32_t ta = am & WN_META_TYPE_MASK, tb = cm & WN_META_TYPE_MASK;
                        const bool typ = (ta != tb) || (ta == WN_WATER && tb == WN_WATER);
                        const bool anyh = ((am | cm) & WN_META_HASH) != 0;
                        pass = pass && typ && anyh;
                    }
                    if (pass) hits |= (1u << h);
                }
            }
        }
        /* warp scan of the hit counts -> dense queue positions */
        const int cnt = __popc(hits);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        const int tot = __shfl_sync(FULL, inc, 31);
        if (tot) {
            int pos = qlen + inc - cnt;
            while (hits) {
                const int h = __ffs(hits) - 1;
                hits &= hits - 1;
                ws.queue[pos++] = ((uint32_t)h << 27) | code_c;
            }
            qlen += tot;
            __syncwarp();
            while (qlen >= 32) {
                wn_eval_batch<TYPED>(p, fbase, ws, hdh, qlen - 32, 32, mode, cs, lane);
                qlen -= 32;
                if (cs.overflow) return;
            }
            __syncwarp();
        }
    }
    if (qlen > 0) {
        __syncwarp();
        wn_eval_batch<TYPED>(p, fbase, ws, hdh, 0, qlen, mode, cs, lane);
    }
    __syncwarp();
}

/* Publish a pooled chunk: allocate temp space, group the pool by row. */
__device__ __forceinline__ void
wn_commit_pool(const WnEdgeParams& p, size_t fbase, WarpShared& ws, int hc, const ChunkState& cs, int lane)
{
    const int total = cs.pool_len;
    unsigned long long base = 0;
    if (lane == 0 && total > 0) base = atomicAdd(p.temp_cursor, (unsigned long long)total);
    base = __shfl_sync(FULL, base, 0);
    /* exclusive scan of the row counts (hc <= 16 <= 32 lanes) */
    const int c = (lane < hc) ? ws.cnt[lane] : 0;
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane < hc) {
        ws.rbase[lane] = inc - c;
        const int idx = ws.hrow[lane];
        p.row_pos[fbase + idx] = (uint32_t)(base + (unsigned long long)(inc - c));
        p.row_cnt[fbase + idx] = (uint32_t)c;
    }
    __syncwarp();
    for (int e = lane; e < total; e += 32) {
        const PoolEntry pe = ws.pool[e];
        const int ph = (int)(pe.jh >> 27);
        const int k = atomicAdd(&ws.fill[ph], 1);
        const unsigned long long pos = base + (unsigned long long)(ws.rbase[ph] + k);
        if (pos < p.temp_cap) {
            WnTempEdge t; t.j = (int32_t)(pe.jh & 0x07ffffffu); t.length = pe.length; t.cosine = pe.cosine; t.pad = 0;
            p.temp[pos] = t;
        }
    }
    __syncwarp();
}

template <bool TYPED>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4)
wn_k_edges(const WnEdgeParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    const int cell = blockIdx.x * WARPS_PER_CTA + warp;
    const WnGrid g = p.grid[f];
    if (cell >= g.ncells) return;

    const size_t per_warp = sizeof(WarpShared) + (size_t)HC * p.hs * 3 * sizeof(double);
    WarpShared& ws = *reinterpret_cast<WarpShared*>(smem_raw + (size_t)warp * per_warp);
    double* hdh = reinterpret_cast<double*>(smem_raw + (size_t)warp * per_warp + sizeof(WarpShared));

    const size_t fbase = (size_t)f * p.n_sites;
    const int* cs_f = p.cell_start + (size_t)f * (p.cell_cap + 1);
    const int h0 = __ldg(cs_f + cell), h1 = __ldg(cs_f + cell + 1);
    if (h0 == h1) return;

    /* the 9 candidate runs: cells (cx-1..cx+1, cy+dy, cz+dz) are contiguous in x */
    const int cx = cell % g.nx, cy = (cell / g.nx) % g.ny, cz = cell / (g.nx * g.ny);
    int rs = 0, rl = 0;
    if (lane < 9) {
        const int y = cy + (lane % 3) - 1, z = cz + (lane / 3) - 1;
        if (y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
            const int xlo = max(cx - 1, 0), xhi = min(cx + 1, g.nx - 1);
            const int rowc = (z * g.ny + y) * g.nx;
            rs = __ldg(cs_f + rowc + xlo);
            rl = __ldg(cs_f + rowc + xhi + 1) - rs;
        }
    }
    /* compact the non-empty runs and prefix their lengths */
    const unsigned nz = __ballot_sync(FULL, rl > 0);
    const int n_runs = __popc(nz);
    const int slot = __popc(nz & ((1u << lane) - 1u));
    int inc = rl;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    const int total = __shfl_sync(FULL, inc, 8);
    if (rl > 0) { ws.run_start[slot] = rs; ws.run_pref[slot] = inc - rl; }
    __syncwarp();

    unsigned long long evals = 0, tie_len = 0, tie_cos = 0, tie_pos0 = 0;
    /* One sweep call site, driven as a small state machine:
     *   POOL on a chunk of <= HC rows -> on pool overflow, POOL one row at a time
     *   -> if a single row still overflows: COUNT it, allocate, then DIRECT-write it. */
    int hb = h0, hc = min(HC, h1 - h0), mode = MODE_POOL, split_end = -1;
    unsigned long long direct_base = 0;
    while (hb < h1) {
        ChunkState cs = {};
        cs.direct_base = direct_base;
        wn_sweep<TYPED>(p, fbase, ws, hdh, hb, hc, n_runs, total, mode, cs, lane);
        bool advance = false;
        if (mode == MODE_POOL) {
            if (!cs.overflow) {
                wn_commit_pool(p, fbase, ws, hc, cs, lane);
                evals += cs.evals; tie_len += cs.tie_len; tie_cos += cs.tie_cos; tie_pos0 += cs.tie_pos0;
                advance = true;
            } else if (hc > 1) {
                split_end = hb + hc; hc = 1;
            } else {
                mode = MODE_COUNT;
            }
        } else if (mode == MODE_COUNT) {
            evals += cs.evals; tie_len += cs.tie_len; tie_cos += cs.tie_cos; tie_pos0 += cs.tie_pos0;
            unsigned long long base = 0;
            if (lane == 0) {
                base = atomicAdd(p.temp_cursor, (unsigned long long)cs.direct_k);
                p.row_pos[fbase + ws.hrow[0]] = (uint32_t)base;
                p.row_cnt[fbase + ws.hrow[0]] = cs.direct_k;
            }
            direct_base = __shfl_sync(FULL, base, 0);
            mode = MODE_DIRECT;
        } else {
            mode = MODE_POOL;
            advance = true;
        }
        if (advance) {
            hb += hc;
            hc = (hb < split_end) ? 1 : min(HC, h1 - hb);
        }
    }
    if (lane == 0) {
        WnFrameCounters* fc = p.counters + f;
        if (evals) atomicAdd(&fc->evals, evals);
        if (tie_len) atomicAdd(&fc->tie_len, tie_len);
        if (tie_cos) atomicAdd(&fc->tie_cos, tie_cos);
        if (tie_pos0) atomicAdd(&fc->tie_pos0, tie_pos0);
    }
}

}  // namespace

size_t wn_edges_smem_bytes(int hs)
{
    return (size_t)WARPS_PER_CTA * (sizeof(WarpShared) + (size_t)HC * hs * 3 * sizeof(double));
}

void wn_launch_edges(const WnEdgeParams& p, int n_frames, cudaStream_t st)
{
    const size_t smem = wn_edges_smem_bytes(p.hs);
    dim3 grid((p.cell_cap + WARPS_PER_CTA - 1) / WARPS_PER_CTA, n_frames);
    if (p.typed) {
        cudaFuncSetAttribute(wn_k_edges<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        wn_k_edges<true><<<grid, WARPS_PER_CTA * 32, smem, st>>>(p);
    } else {
        cudaFuncSetAttribute(wn_k_edges<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        wn_k_edges<false><<<grid, WARPS_PER_CTA * 32, smem, st>>>(p);
    }
}
