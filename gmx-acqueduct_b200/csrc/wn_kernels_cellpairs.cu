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
 *   wn_k_cp_scatter  cell-sorted records {x, y, z, original index}
 *   wn_k_cp_sweep<0> one warp per site i: EXACT fp64 test against the 25 contiguous runs of the
 *                    neighbourhood, counts the partners j > i (and the cutoff ties of the tie report)
 *   (scan)           row offsets: rows in index order = the reference's outer loop
 *   wn_k_cp_sweep<1> same sweep; partners set a bit in a per-warp bitmap over j (two levels: words, and a
 *                    summary of the non-empty words), then the bits are read back in ascending order --
 *                    the reference's inner loop order -- and written as (i, j).  No sort, no fp32 pre-test:
 *                    every candidate goes through the reference's own arithmetic.
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
wn_k_cp_scatter(const double* __restrict__ xyz, int n, const int* __restrict__ cell_start,
                const int* __restrict__ site_cell, const int* __restrict__ site_rank, double4* __restrict__ srec)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int dst = cell_start[site_cell[i]] + site_rank[i];
    srec[dst] = make_double4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], __longlong_as_double((long long)i));
}

/* One warp per cell-sorted site p (row i = its original index).  FILL = 0: count; FILL = 1: ordered output. */
template <int FILL, int PBC>
__global__ void __launch_bounds__(256)
wn_k_cp_sweep(const double4* __restrict__ srec, int n, int first_limit, const WnCellPairGrid g,
              const int* __restrict__ cell_start, const int* __restrict__ site_cell_sorted_src,
              uint32_t* __restrict__ row_cnt, const unsigned long long* __restrict__ row_off,
              int32_t* __restrict__ pairs, unsigned long long* __restrict__ ties, int words, int swords)
{
    extern __shared__ uint32_t cp_smem[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* bm = cp_smem + (size_t)warp * (words + swords);      /* [words] bitmap over j, then [swords] summary */
    uint32_t* sm = bm + words;
    if (FILL) {
        for (int w = lane; w < words + swords; w += 32) bm[w] = 0u;
        __syncwarp();
    }
    unsigned long long tie = 0, near = 0;
    for (int p = blockIdx.x * warps + warp; p < n; p += gridDim.x * warps) {
        const double4 me = srec[p];
        const int i = (int)__double_as_longlong(me.w);
        if (i >= first_limit) { if (!FILL && lane == 0) row_cnt[i] = 0u; continue; }
        const int cid = site_cell_sorted_src[i];
        const int cx = cid % g.n[0], cy = (cid / g.n[0]) % g.n[1], cz = cid / (g.n[0] * g.n[1]);
        uint32_t count = 0;
        /* 25 rows (dy, dz) of 5 x-adjacent cells: contiguous in the cell-sorted array (two pieces when a periodic row wraps) */
        for (int r = 0; r < (2 * CP_R + 1) * (2 * CP_R + 1); ++r) {
            int y = cy + r % (2 * CP_R + 1) - CP_R, z = cz + r / (2 * CP_R + 1) - CP_R;
            if (PBC) {
                y = (y % g.n[1] + g.n[1]) % g.n[1];
                z = (z % g.n[2] + g.n[2]) % g.n[2];
            } else if (y < 0 || y >= g.n[1] || z < 0 || z >= g.n[2]) {
                continue;
            }
            const int rowc = (z * g.n[1] + y) * g.n[0];
            for (int piece = 0; piece < (PBC ? 2 : 1); ++piece) {
                int xlo = cx - CP_R, xhi = cx + CP_R;
                if (PBC) {
                    /* piece 0: the in-range part; piece 1: the part that wraps around (cells taken modulo the grid) */
                    if (piece == 0) { xlo = max(xlo, 0); xhi = min(xhi, g.n[0] - 1); }
                    else if (xlo < 0) { xhi = g.n[0] - 1; xlo = xlo + g.n[0]; }
                    else if (xhi >= g.n[0]) { xlo = 0; xhi = xhi - g.n[0]; }
                    else continue;
                } else {
                    xlo = max(xlo, 0); xhi = min(xhi, g.n[0] - 1);
                }
                const int b = __ldg(cell_start + rowc + xlo), e = __ldg(cell_start + rowc + xhi + 1);
                for (int q = b + lane; q < e; q += 32) {
                    const double4 c = srec[q];
                    const int j = (int)__double_as_longlong(c.w);
                    if (j <= i) continue;
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
                            const uint32_t old = atomicOr(bm + (j >> 5), 1u << (j & 31));
                            if (old == 0u) atomicOr(sm + (j >> 10), 1u << ((j >> 5) & 31));
                        } else {
                            ++count;
                        }
                    } else if (!FILL && lhs == 42.25) {
                        ++tie;
                    }
                    if (!FILL && fabs(lhs - 42.25) <= 2.8421709430404007e-14) ++near;
                }
            }
        }
        if (!FILL) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(FULL, count, o);
            if (lane == 0) row_cnt[i] = count;
        } else {
            __syncwarp();
            /* read the bits back in ascending j: lane = one summary word (1,024 values of j) per step */
            unsigned long long w = row_off[i];
            for (int s0 = 0; s0 < swords; s0 += 32) {
                const int s = s0 + lane;
                const uint32_t sw = s < swords ? sm[s] : 0u;
                uint32_t mine = 0;
                for (uint32_t t = sw; t; t &= t - 1) mine += __popc(bm[s * 32 + __ffs(t) - 1]);
                uint32_t inc = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, inc, o);
                    if (lane >= o) inc += t;
                }
                unsigned long long pos = w + (inc - mine);
                for (uint32_t t = sw; t; t &= t - 1) {
                    const int wi = s * 32 + __ffs(t) - 1;
                    uint32_t bits = bm[wi];
                    bm[wi] = 0u;                                    /* leave the bitmap clean for the next row */
                    for (; bits; bits &= bits - 1)
                        reinterpret_cast<int2*>(pairs)[pos++] = make_int2(i, wi * 32 + __ffs(bits) - 1);
                }
                if (s < swords) sm[s] = 0u;
                w += __shfl_sync(FULL, inc, 31);
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

/* shared memory per warp of the fill kernel: bitmap over j + summary */
static size_t cp_bitmap_words(int n, int* words, int* swords)
{
    *words = (n + 31) / 32;
    *swords = (*words + 31) / 32;
    *words = *swords * 32;                       /* whole summary words */
    return (size_t)(*words + *swords);
}

int wn_cellpairs_supported(int n)
{
    int w, s;
    return cp_bitmap_words(n, &w, &s) * sizeof(uint32_t) <= (size_t)200 * 1024;
}

void wn_launch_cellpairs_bin(const double* xyz, int n, const WnCellPairGrid& g, int* cell_cnt, int* site_cell,
                             int* site_rank, double4* srec, cudaStream_t st)
{
    cudaMemsetAsync(cell_cnt, 0, ((size_t)g.ncells + 1) * sizeof(int), st);
    wn_k_cp_bin<<<(n + 255) / 256, 256, 0, st>>>(xyz, n, g, cell_cnt, site_cell, site_rank);
    wn_k_cp_cellscan<<<1, 1024, 0, st>>>(cell_cnt, g.ncells);
    wn_k_cp_scatter<<<(n + 255) / 256, 256, 0, st>>>(xyz, n, cell_cnt, site_cell, site_rank, srec);
}

void wn_launch_cellpairs_count(const double4* srec, int n, int first_limit, const WnCellPairGrid& g, const int* cell_start,
                               const int* site_cell, uint32_t* row_cnt, unsigned long long* ties, cudaStream_t st)
{
    const int warps = 8;
    const int blocks = std::max(1, std::min((n + warps - 1) / warps, 148 * 32));
    if (g.pbc)
        wn_k_cp_sweep<0, 1><<<blocks, warps * 32, 0, st>>>(srec, n, first_limit, g, cell_start, site_cell, row_cnt, nullptr, nullptr, ties, 0, 0);
    else
        wn_k_cp_sweep<0, 0><<<blocks, warps * 32, 0, st>>>(srec, n, first_limit, g, cell_start, site_cell, row_cnt, nullptr, nullptr, ties, 0, 0);
}

void wn_launch_cellpairs_fill(const double4* srec, int n, int first_limit, const WnCellPairGrid& g, const int* cell_start,
                              const int* site_cell, const unsigned long long* row_off, int32_t* pairs, cudaStream_t st)
{
    int words, swords;
    const size_t per_warp = cp_bitmap_words(n, &words, &swords) * sizeof(uint32_t);
    /* as many warps per CTA as the bitmaps allow (>= 1 by wn_cellpairs_supported), at most 8 */
    const int warps = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)200 * 1024) / per_warp));
    const size_t smem = per_warp * warps;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)220 * 1024) / (smem + 1024)));
    const int blocks = std::max(1, std::min((n + warps - 1) / warps, 148 * per_sm));
    if (g.pbc) {
        cudaFuncSetAttribute(wn_k_cp_sweep<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        wn_k_cp_sweep<1, 1><<<blocks, warps * 32, smem, st>>>(srec, n, first_limit, g, cell_start, site_cell, nullptr, row_off, pairs, nullptr, words, swords);
    } else {
        cudaFuncSetAttribute(wn_k_cp_sweep<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        wn_k_cp_sweep<1, 0><<<blocks, warps * 32, smem, st>>>(srec, n, first_limit, g, cell_start, site_cell, nullptr, row_off, pairs, nullptr, words, swords);
    }
}
