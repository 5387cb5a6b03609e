/*
 * wn_kernels_cellpairs.cu -- CudaFindPairs::find as a cell-list cutoff search (sm_100a).
 *
 * Same result as BruteFindPairs::find (reference src/brute-find-pairs.cpp:19-30): every i < j with
 *     10.0 * CGAL::squared_distance(p_i, p_j) < 42.25      (fp64, (dx*dx + dy*dy) + dz*dz, no FMA)
 * in lexicographic (i, j) order -- but in O(N) instead of O(N^2): the cutoff is a distance of 2.0555 nm
 * (the unit slip of :19-20), so the sites are binned on a grid of edge >= 2.0555/2 nm and every
 * site only meets the 5 x 5 x 5 cells around its own.
 *
 *   wn_k_cp_bin      cell of every site (fp64, clamped; non-finite sites land in a cell and never pair,
 *                    as every comparison with NaN is false in the reference) + population count
 *   (scan)           cell starts
 *   wn_k_cp_scatter  cell-sorted records {x, y, z, original index} in fp64, and origin-relative fp32 copies
 *   wn_k_cp_sweep<0> one warp per site i over the 25 contiguous runs of its neighbourhood (run bounds fetched by
 *                    25 lanes at once), 64 candidates per step with the next 64 already in flight: fp32 reject
 *                    (strict superset, bound from the host) -> ballot/popc queue -> the reference's EXACT fp64
 *                    test on full batches of 32 queued candidates; counts the partners j > i (and the cutoff
 *                    ties of the tie report)
 *   (scan)           row offsets: rows in index order = the reference's outer loop
 *   wn_k_cp_sweep<1> same sweep; partners set a bit in a per-warp bitmap over j (two levels: words, and a
 *                    summary of the non-empty words), then the bits are read back in ascending order --
 *                    the reference's inner loop order -- and written as (i, j).  No sort; every pair that is
 *                    written went through the reference's own arithmetic.
 * Measured (33,333 oxygens of the 100k-atom box, 15.9 M pairs): count 0.23 ms + fill 0.40 ms + scans, 0.68 ms
 * against 1.49 ms for the all-pairs tile kernel; 333,333 sites (183 M pairs): 22 ms.  The sweep is issue-bound
 * (~33 instructions per 32 candidates at 4,550 candidates per site), not memory-bound (L1 hit rate 90 %).
 * Periodic (rectangular) boxes: cells on the box, neighbour cells modulo the grid, displacement = the
 * fp64 minimum image of include/wn_b200.h ("Periodic boundaries").  Boxes with fewer than 5 cells in a
 * dimension, triclinic boxes and inputs beyond the bitmap's shared-memory budget use the exact O(N^2) tile
 * kernel (wn_kernels_pairs.cu).
 */
#include "wn_internal.cuh"

#include <algorithm>

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CP_R = 2;                       /* neighbourhood radius in cells */

__device__ __forceinline__ int cp_cell(double x, double o, double inv, int n, int pbc, double L, double invL)
{
    if (pbc) x = __dsub_rn(x, __dmul_rn(floor(__dmul_rn(x, invL)), L));        /* into [0, L) (up to rounding) */
    const int c = __double2int_rd(__dmul_rn(__dsub_rn(x, o), inv));            /* saturates; NaN -> 0 */
    return min(max(c, 0), n - 1);
}

__global__ void __launch_bounds__(256)
wn_k_cp_bin(const double* __restrict__ xyz, int n, const WnCellPairGrid g, int* __restrict__ cell_cnt,
            int* __restrict__ site_cell, int* __restrict__ site_rank)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = cp_cell(xyz[3 * (size_t)i], g.o[0], g.inv[0], g.n[0], g.pbc, g.L[0], g.invL[0]);
    const int cy = cp_cell(xyz[3 * (size_t)i + 1], g.o[1], g.inv[1], g.n[1], g.pbc, g.L[1], g.invL[1]);
    const int cz = cp_cell(xyz[3 * (size_t)i + 2], g.o[2], g.inv[2], g.n[2], g.pbc, g.L[2], g.invL[2]);
    const int cid = (cz * g.n[1] + cy) * g.n[0] + cx;
    site_cell[i] = cid;
    site_rank[i] = atomicAdd(cell_cnt + cid, 1);
}

/* exclusive scan of the cell populations (single block; <= a few million cells) */
__global__ void __launch_bounds__(1024)
wn_k_cp_cellscan(int* __restrict__ cell_cnt, int ncells)
{
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < ncells; base += blockDim.x) {
        const int c = base + threadIdx.x;
        const int v = c < ncells ? cell_cnt[c] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int w = s_w[lane];
            int iw = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, iw, o);
                if (lane >= o) iw += t;
            }
            s_w[lane] = iw - w;
        }
        __syncthreads();
        const int excl = s_carry + s_w[warp] + inc - v;
        if (c < ncells) cell_cnt[c] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) cell_cnt[ncells] = s_carry;
}

__global__ void __launch_bounds__(256)
wn_k_cp_scatter(const double* __restrict__ xyz, int n, const WnCellPairGrid g, const int* __restrict__ cell_start,
                const int* __restrict__ site_cell, const int* __restrict__ site_rank, double4* __restrict__ srec,
                float4* __restrict__ srecf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int dst = cell_start[site_cell[i]] + site_rank[i];
    const double x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
    srec[dst] = make_double4(x, y, z, __longlong_as_double((long long)i));
    /* origin-relative float copy for the reject test (the error bound in g.bound_f assumes exactly this rounding) */
    srecf[dst] = make_float4(__double2float_rn(__dsub_rn(x, g.o[0])), __double2float_rn(__dsub_rn(y, g.o[1])),
                             __double2float_rn(__dsub_rn(z, g.o[2])), __int_as_float(i));
}

/* One warp per cell-sorted site p (row i = its original index).  FILL = 0: count; FILL = 1: ordered output.
 *
 * The 125 neighbour cells hold ~4,500 candidates per site at water density, a quarter of them inside the cutoff
 * and half of those with j > i.  Each lane looks at one candidate per step with a cheap fp32 reject (origin-relative
 * float copies of the coordinates, the bound widened by the worst-case float error: a strict superset of the exact
 * test); survivors are queued per warp (ballot + popc, arrival order is free) and the reference's fp64 arithmetic runs
 * on full batches of 32 queued candidates, so the fp64 pipe only ever sees candidates that are (almost) pairs. */
constexpr int CP_QCAP = 96;                   /* queue: < 32 carried + <= 64 new per double step */

template <int FILL, int PBC>
__device__ __forceinline__ void
cp_exact_batch(const double4* __restrict__ srec, const double4& me, int i, const WnCellPairGrid& g, const int* queue, int nq, int lane,
               uint32_t& count, unsigned long long& tie, unsigned long long& near, uint32_t* bm, uint32_t* sm)
{
    if (lane < nq) {
        const double4 c = srec[queue[lane]];
        const int j = (int)__double_as_longlong(c.w);
        WN_CHECK(j > i);
        /* squared_distance(points.at(i), points.at(j)) (:25): p_i - p_j component-wise */
        double dx = __dsub_rn(me.x, c.x), dy = __dsub_rn(me.y, c.y), dz = __dsub_rn(me.z, c.z);
        if (PBC) {
            dx = __dsub_rn(dx, __dmul_rn(rint(__dmul_rn(dx, g.invL[0])), g.L[0]));
            dy = __dsub_rn(dy, __dmul_rn(rint(__dmul_rn(dy, g.invL[1])), g.L[1]));
            dz = __dsub_rn(dz, __dmul_rn(rint(__dmul_rn(dz, g.invL[2])), g.L[2]));
        }
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        const double lhs = __dmul_rn(10.0, d2);
        if (lhs < 42.25) {
            if (FILL) {
                WN_CHECK((j >> 5) < (int)(sm - bm));
                const uint32_t old = atomicOr(bm + (j >> 5), 1u << (j & 31));
                if (old == 0u) atomicOr(sm + (j >> 10), 1u << ((j >> 5) & 31));
            } else {
                ++count;
            }
        } else if (!FILL && lhs == 42.25) {
            ++tie;
        }
        /* near-ties: within 4 ulp(42.25) = 4 * 2^-47 of the cutoff, either side */
        if (!FILL && fabs(lhs - 42.25) <= 2.8421709430404007e-14) ++near;
    }
}

constexpr int CP_MAXRUN = 50;                 /* 25 rows of the neighbourhood, each possibly split by the periodic wrap */

template <int FILL, int PBC>
__global__ void __launch_bounds__(256)
wn_k_cp_sweep(const double4* __restrict__ srec, const float4* __restrict__ srecf, int n, int first_limit, const WnCellPairGrid g,
              const int* __restrict__ cell_start, const int* __restrict__ site_cell_sorted_src,
              uint32_t* __restrict__ row_cnt, const unsigned long long* __restrict__ row_off,
              int32_t* __restrict__ pairs, unsigned long long* __restrict__ ties, int words, int swords)
{
    extern __shared__ uint32_t cp_smem[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    /* per warp: queue[CP_QCAP], run starts [CP_MAXRUN], run prefixes [CP_MAXRUN + 2], then (FILL) the bitmap over j
     * [words], its summary [swords], the word list [1024 x u16] */
    const int per_warp = CP_QCAP + 2 * CP_MAXRUN + 2 + (FILL ? words + swords + 512 : 0);
    int* queue = reinterpret_cast<int*>(cp_smem + (size_t)warp * per_warp);
    int* run_start = queue + CP_QCAP;
    int* run_pref = run_start + CP_MAXRUN;
    uint32_t* bm = cp_smem + (size_t)warp * per_warp + CP_QCAP + 2 * CP_MAXRUN + 2;
    uint32_t* sm = bm + words;
    unsigned short* wl = reinterpret_cast<unsigned short*>(sm + swords);
    if (FILL) {
        for (int w = lane; w < words + swords; w += 32) bm[w] = 0u;
        __syncwarp();
    }
    const float bound = g.bound_f, Lx = g.Lf[0], Ly = g.Lf[1], Lz = g.Lf[2], iLx = g.invLf[0], iLy = g.invLf[1], iLz = g.invLf[2];
    unsigned long long tie = 0, near = 0;
    for (int p = blockIdx.x * warps + warp; p < n; p += gridDim.x * warps) {
        const double4 me = srec[p];
        const int i = (int)__double_as_longlong(me.w);
        if (i >= first_limit) { if (!FILL && lane == 0) row_cnt[i] = 0u; continue; }
        const float4 mf = __ldg(srecf + p);
        const int cid = site_cell_sorted_src[i];
        const int cx = cid % g.n[0], cy = (cid / g.n[0]) % g.n[1], cz = cid / (g.n[0] * g.n[1]);
        /* The neighbourhood as contiguous runs of the cell-sorted array: 25 rows (dy, dz) of 5 x-adjacent cells, two
         * pieces when a periodic row wraps.  One lane per (row, piece): all cell_start reads are in flight at once. */
        int total = 0;
        __syncwarp();
        for (int t0 = 0; t0 < (PBC ? 64 : 32); t0 += 32) {
            const int t = t0 + lane;
            const int r = PBC ? (t >> 1) : t, piece = PBC ? (t & 1) : 0;
            int rs = 0, rl = 0;
            if (r < (2 * CP_R + 1) * (2 * CP_R + 1)) {
                int y = cy + r % (2 * CP_R + 1) - CP_R, z = cz + r / (2 * CP_R + 1) - CP_R;
                bool ok = true;
                if (PBC) {
                    y = (y % g.n[1] + g.n[1]) % g.n[1];
                    z = (z % g.n[2] + g.n[2]) % g.n[2];
                } else {
                    ok = y >= 0 && y < g.n[1] && z >= 0 && z < g.n[2];
                }
                int xlo = cx - CP_R, xhi = cx + CP_R;
                if (PBC) {
                    /* piece 0: the in-range part; piece 1: the part that wraps around (cells taken modulo the grid) */
                    if (piece == 0) { xlo = max(xlo, 0); xhi = min(xhi, g.n[0] - 1); }
                    else if (xlo < 0) { xhi = g.n[0] - 1; xlo = xlo + g.n[0]; }
                    else if (xhi >= g.n[0]) { xlo = 0; xhi = xhi - g.n[0]; }
                    else ok = false;
                } else {
                    xlo = max(xlo, 0); xhi = min(xhi, g.n[0] - 1);
                }
                if (ok) {
                    const int rowc = (z * g.n[1] + y) * g.n[0];
                    rs = __ldg(cell_start + rowc + xlo);
                    rl = __ldg(cell_start + rowc + xhi + 1) - rs;
                }
            }
            int inc = rl;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += v;
            }
            if (t < CP_MAXRUN) { run_start[t] = rs; run_pref[t] = total + inc - rl; }
            total += __shfl_sync(FULL, inc, 31);
        }
        const int n_runs = PBC ? CP_MAXRUN : (2 * CP_R + 1) * (2 * CP_R + 1);
        __syncwarp();
        if (lane == 0) { run_pref[n_runs] = total; run_pref[n_runs + 1] = 0x7fffffff; }
        __syncwarp();
        uint32_t count = 0;
        int qlen = 0;
        /* run by run, two steps (64 candidates) per queue push, the next two records loaded before the current two
         * are tested */
        for (int r = 0; r < n_runs; ++r) {
            const int rb = run_start[r], re = rb + (run_pref[r + 1] - run_pref[r]);
            if (rb >= re) continue;
            int q = rb + lane;
            float4 c0 = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)), c1 = c0;       /* index -1: never > i */
            if (q < re) c0 = __ldg(srecf + q);
            if (q + 32 < re) c1 = __ldg(srecf + q + 32);
            for (; q - lane < re; q += 64) {
                float4 n0 = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)), n1 = n0;
                if (q + 64 < re) n0 = __ldg(srecf + q + 64);
                if (q + 96 < re) n1 = __ldg(srecf + q + 96);
                bool p0 = false, p1 = false;
                if (__float_as_int(c0.w) > i) {
                    float dx = mf.x - c0.x, dy = mf.y - c0.y, dz = mf.z - c0.z;
                    if (PBC) { dx -= rintf(dx * iLx) * Lx; dy -= rintf(dy * iLy) * Ly; dz -= rintf(dz * iLz) * Lz; }
                    p0 = dx * dx + dy * dy + dz * dz <= bound;                     /* false for NaN, as the exact test */
                }
                if (__float_as_int(c1.w) > i) {
                    float dx = mf.x - c1.x, dy = mf.y - c1.y, dz = mf.z - c1.z;
                    if (PBC) { dx -= rintf(dx * iLx) * Lx; dy -= rintf(dy * iLy) * Ly; dz -= rintf(dz * iLz) * Lz; }
                    p1 = dx * dx + dy * dy + dz * dz <= bound;
                }
                const unsigned m0 = __ballot_sync(FULL, p0), m1 = __ballot_sync(FULL, p1);
                if (m0 | m1) {
                    const unsigned lt = (1u << lane) - 1u;
                    WN_CHECK(qlen >= 0 && qlen + __popc(m0) + __popc(m1) <= CP_QCAP && (!p0 || q < n) && (!p1 || q + 32 < n));
                    if (p0) queue[qlen + __popc(m0 & lt)] = q;
                    if (p1) queue[qlen + __popc(m0) + __popc(m1 & lt)] = q + 32;
                    qlen += __popc(m0) + __popc(m1);
                    __syncwarp();
                    while (qlen >= 32) {
                        qlen -= 32;
                        cp_exact_batch<FILL, PBC>(srec, me, i, g, queue + qlen, 32, lane, count, tie, near, bm, sm);
                    }
                    __syncwarp();
                }
                c0 = n0; c1 = n1;
            }
        }
        if (qlen > 0) {
            cp_exact_batch<FILL, PBC>(srec, me, i, g, queue, qlen, lane, count, tie, near, bm, sm);
            __syncwarp();
        }
        if (!FILL) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(FULL, count, o);
            if (lane == 0) row_cnt[i] = count;
        } else {
            __syncwarp();
            /* Read the bits back in ascending j.  32 summary words (32,768 values of j) per step: the non-empty
             * bitmap words are listed in order (lane = summary word, warp scan of the counts), then the list is
             * consumed 32 words at a time -- lane = one word, warp scan of the popcounts -- so the expansion is
             * spread over the lanes even when all partners sit in a narrow index range. */
            unsigned long long w = row_off[i];
            for (int s0 = 0; s0 < swords; s0 += 32) {
                const int s = s0 + lane;
                const uint32_t sw = s < swords ? sm[s] : 0u;
                uint32_t inc = __popc(sw);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, inc, o);
                    if (lane >= o) inc += t;
                }
                const int n_list = (int)__shfl_sync(FULL, inc, 31);
                if (n_list == 0) continue;
                {
                    int at = (int)(inc - __popc(sw));
                    for (uint32_t t = sw; t; t &= t - 1) { WN_CHECK(at < 1024); wl[at++] = (unsigned short)(lane * 32 + __ffs(t) - 1); }
                    if (s < swords) sm[s] = 0u;
                }
                __syncwarp();
                for (int k0 = 0; k0 < n_list; k0 += 32) {
                    const int k = k0 + lane;
                    int wi = 0;
                    uint32_t bits = 0;
                    if (k < n_list) {
                        wi = s0 * 32 + (int)wl[k];
                        bits = bm[wi];
                        bm[wi] = 0u;                                 /* leave the bitmap clean for the next row */
                    }
                    uint32_t pinc = __popc(bits);
                    const uint32_t mine = pinc;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(FULL, pinc, o);
                        if (lane >= o) pinc += t;
                    }
                    unsigned long long pos = w + (pinc - mine);
                    for (; bits; bits &= bits - 1) {
                        WN_CHECK(wi < words && pos < row_off[i + 1]);
                        reinterpret_cast<int2*>(pairs)[pos++] = make_int2(i, wi * 32 + __ffs(bits) - 1);
                    }
                    w += __shfl_sync(FULL, pinc, 31);
                }
                __syncwarp();
            }
            __syncwarp();
        }
    }
    if (!FILL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { tie += __shfl_xor_sync(FULL, tie, o); near += __shfl_xor_sync(FULL, near, o); }
        if (lane == 0) {
            if (tie) atomicAdd(ties, tie);
            if (near) atomicAdd(ties + 1, near);
        }
    }
}

}  // namespace

/* shared memory per warp of the fill kernel: bitmap over j + summary (+ queue and word list, see the kernel) */
static size_t cp_bitmap_words(int n, int* words, int* swords)
{
    *words = (n + 31) / 32;
    *swords = (*words + 31) / 32;
    *words = *swords * 32;                       /* whole summary words */
    return (size_t)(*words + *swords);
}

static size_t cp_fill_words_per_warp(int n, int* words, int* swords)
{
    return cp_bitmap_words(n, words, swords) + CP_QCAP + 2 * CP_MAXRUN + 2 + 512;
}

int wn_cellpairs_supported(int n)
{
    int w, s;
    return cp_fill_words_per_warp(n, &w, &s) * sizeof(uint32_t) <= (size_t)200 * 1024;
}

void wn_launch_cellpairs_bin(const double* xyz, int n, const WnCellPairGrid& g, int* cell_cnt, int* site_cell,
                             int* site_rank, double4* srec, float4* srecf, cudaStream_t st)
{
    cudaMemsetAsync(cell_cnt, 0, ((size_t)g.ncells + 1) * sizeof(int), st);
    wn_k_cp_bin<<<(n + 255) / 256, 256, 0, st>>>(xyz, n, g, cell_cnt, site_cell, site_rank);
    wn_k_cp_cellscan<<<1, 1024, 0, st>>>(cell_cnt, g.ncells);
    wn_k_cp_scatter<<<(n + 255) / 256, 256, 0, st>>>(xyz, n, g, cell_cnt, site_cell, site_rank, srec, srecf);
}

void wn_launch_cellpairs_count(const double4* srec, const float4* srecf, int n, int first_limit, const WnCellPairGrid& g,
                               const int* cell_start, const int* site_cell, uint32_t* row_cnt, unsigned long long* ties,
                               cudaStream_t st)
{
    const int warps = 8;
    const int blocks = std::max(1, std::min((n + warps - 1) / warps, 148 * 32));
    const size_t smem = (size_t)warps * (CP_QCAP + 2 * CP_MAXRUN + 2) * sizeof(uint32_t);
    if (g.pbc)
        wn_k_cp_sweep<0, 1><<<blocks, warps * 32, smem, st>>>(srec, srecf, n, first_limit, g, cell_start, site_cell, row_cnt, nullptr, nullptr, ties, 0, 0);
    else
        wn_k_cp_sweep<0, 0><<<blocks, warps * 32, smem, st>>>(srec, srecf, n, first_limit, g, cell_start, site_cell, row_cnt, nullptr, nullptr, ties, 0, 0);
}

void wn_launch_cellpairs_fill(const double4* srec, const float4* srecf, int n, int first_limit, const WnCellPairGrid& g,
                              const int* cell_start, const int* site_cell, const unsigned long long* row_off, int32_t* pairs,
                              cudaStream_t st)
{
    int words, swords;
    const size_t per_warp = cp_fill_words_per_warp(n, &words, &swords) * sizeof(uint32_t);
    /* as many warps per CTA as the bitmaps allow (>= 1 by wn_cellpairs_supported), at most 8 */
    const int warps = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)200 * 1024) / per_warp));
    const size_t smem = per_warp * warps;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)220 * 1024) / (smem + 1024)));
    const int blocks = std::max(1, std::min((n + warps - 1) / warps, 148 * per_sm));
    if (g.pbc) {
        cudaFuncSetAttribute(wn_k_cp_sweep<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        wn_k_cp_sweep<1, 1><<<blocks, warps * 32, smem, st>>>(srec, srecf, n, first_limit, g, cell_start, site_cell, nullptr, row_off, pairs, nullptr, words, swords);
    } else {
        cudaFuncSetAttribute(wn_k_cp_sweep<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        wn_k_cp_sweep<1, 0><<<blocks, warps * 32, smem, st>>>(srec, srecf, n, first_limit, g, cell_start, site_cell, nullptr, row_off, pairs, nullptr, words, swords);
    }
}
