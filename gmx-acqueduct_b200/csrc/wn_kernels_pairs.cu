/*
 * wn_kernels_pairs.cu -- CudaFindPairs: the pair list of BruteFindPairs::find (sm_100a).
 *
 * Reference semantics (src/brute-find-pairs.cpp:19-30): every i<j with
 *     10.0 * CGAL::squared_distance(p_i, p_j) < 42.25        (fp64; 42.25 = (6.5f)^2)
 * in loop order.  squared_distance of the Cartesian kernel is
 * (dx*dx + dy*dy) + dz*dz with every product and sum rounded (no FMA).
 *
 * The reference cutoff is 2.0555 nm (unit slip), i.e. every site pairs
 * with ~3 % of the box at liquid density and the 8-byte-per-pair output dominates.
 * The kernel is an exact tile sweep: one thread owns one row i and walks j upwards
 * through shared-memory tiles of coordinates, so each row comes out in ascending j
 * with no sort; count -> scan -> fill gives the lexicographic order.  An fp32
 * test with a margin rejects the far pairs (97 % at 100k atoms) before the exact
 * fp64 comparison; fp32 never accepts on its own.  Row tiles are taken in mirror
 * pairs so that every block sweeps the same number of columns.
 */
#include "wn_internal.cuh"

#include <algorithm>

namespace {

constexpr int ROWS = 128;     /* rows (threads) per block  */
constexpr int TILE = 512;     /* j coordinates per tile    */

struct PairTile {
    double x[TILE], y[TILE], z[TILE];
    float4 f[TILE];            /* fp32 copy: one 16-byte broadcast load per test */
};

/* One thread owns (row i, column segment s) and walks the segment's j > i through
 * shared-memory tiles; blockIdx.y = segment.  Splitting the columns into `nseg` segments
 * gives the GPU enough blocks (a thread per row alone is 7 warps per SM at 33k rows) and
 * keeps the order: counts and offsets are indexed [row][segment], segments ascend in j. */
template <bool FILL, int PBC>
__global__ void __launch_bounds__(ROWS)
wn_k_pairs(const double* __restrict__ xyz, int n, int first_limit, int nseg, int seg_len, const WnPairBox box,
           uint32_t* __restrict__ row_cnt, const unsigned long long* __restrict__ row_off,
           int32_t* __restrict__ pairs, unsigned long long* __restrict__ ties)
{
    __shared__ PairTile t;
    __shared__ float s_amax;
    unsigned long long tie = 0, near = 0;
    {
        const int i0 = (int)blockIdx.x * ROWS;
        const int seg = (int)blockIdx.y;
        const int c0 = seg * seg_len, c1 = min(n, c0 + seg_len);
        if (c1 <= i0 + 1) {                      /* segment entirely left of the diagonal: nothing to pair */
            if (!FILL) { const int i = i0 + threadIdx.x; if (i < n) row_cnt[(size_t)i * nseg + seg] = 0; }
            return;
        }
        const int i = i0 + threadIdx.x;
        const bool own = i < n && i < first_limit;
        double xi = 0, yi = 0, zi = 0;
        if (i < n) { xi = xyz[3 * (size_t)i]; yi = xyz[3 * (size_t)i + 1]; zi = xyz[3 * (size_t)i + 2]; }
        const float fxi = (float)xi, fyi = (float)yi, fzi = (float)zi;
        const float amax = fmaxf(fmaxf(fabsf(fxi), fabsf(fyi)), fabsf(fzi));
        uint32_t count = 0;
        unsigned long long w = FILL && own ? row_off[(size_t)i * nseg + seg] : 0ull;

        for (int j0 = max(i0 + 1, c0); j0 < c1; j0 += TILE) {
            __syncthreads();
            if (threadIdx.x == 0) s_amax = 0.f;
            __syncthreads();
            float tmax = 0.f;
            for (int k = threadIdx.x; k < TILE; k += ROWS) {
                const int j = j0 + k;
                double x = 0, y = 0, z = 0;
                if (j < c1) { x = xyz[3 * (size_t)j]; y = xyz[3 * (size_t)j + 1]; z = xyz[3 * (size_t)j + 2]; }
                t.x[k] = x; t.y[k] = y; t.z[k] = z;
                const float4 ff = make_float4((float)x, (float)y, (float)z, 0.f);
                t.f[k] = ff;
                const float m = fmaxf(fmaxf(fabsf(ff.x), fabsf(ff.y)), fabsf(ff.z));
                if (m == m) tmax = fmaxf(tmax, m);
            }
            /* largest |coordinate| of the tile (non-negative floats order like their bit patterns) */
            atomicMax(reinterpret_cast<unsigned int*>(&s_amax), __float_as_uint(tmax));
            __syncthreads();
            if (!own) continue;
            /* fp32 reject bound: 10*d2 < 42.25 <=> d2 < 4.225.  d2f errs by <= ~1e-6 relative from
             * the arithmetic plus 2*|d|*eps*|coord| from rounding the coordinates to float; the
             * bound below covers both for every pair of this row and tile (never accepts alone).
             * periodic: the fp32 image shift adds <= ~1e-6 * L per component on top */
            const float ca = fmaxf(amax, s_amax);
            const float bound = 4.225f * 1.00002f + 1.0e-5f * ca * (1.0f + ca) +
                                (PBC ? 3.0e-5f * (box.Lf[0] + box.Lf[1] + box.Lf[2] + fabsf(box.trif[0]) + fabsf(box.trif[1]) + fabsf(box.trif[2])) : 0.0f);
            const int kend = min(TILE, c1 - j0);
            const int kbeg = max(0, i + 1 - j0);
            for (int k = kbeg; k < kend; ++k) {
                const float4 q = t.f[k];
                float dxf = fxi - q.x, dyf = fyi - q.y, dzf = fzi - q.z;
                if (PBC == 1) {
                    /* approximate minimum image; near +-L/2 either image has the same |d| */
                    dxf -= rintf(dxf * box.invLf[0]) * box.Lf[0];
                    dyf -= rintf(dyf * box.invLf[1]) * box.Lf[1];
                    dzf -= rintf(dzf * box.invLf[2]) * box.Lf[2];
                } else if (PBC == 2) {
                    const float kz = rintf(dzf * box.invLf[2]);
                    dxf -= kz * box.trif[1]; dyf -= kz * box.trif[2]; dzf -= kz * box.Lf[2];
                    const float jy = rintf(dyf * box.invLf[1]);
                    dxf -= jy * box.trif[0]; dyf -= jy * box.Lf[1];
                    dxf -= rintf(dxf * box.invLf[0]) * box.Lf[0];
                }
                const float d2f = dxf * dxf + dyf * dyf + dzf * dzf;
                if (d2f > bound) continue;              /* false for NaN: falls through to the exact test */
                double dx = __dsub_rn(xi, t.x[k]), dy = __dsub_rn(yi, t.y[k]), dz = __dsub_rn(zi, t.z[k]);
                if (PBC == 1) {
                    /* periodic definition (include/wn_b200.h, "Periodic boundaries"): d - rint(d * (1/L)) * L */
                    dx = __dsub_rn(dx, __dmul_rn(rint(__dmul_rn(dx, box.invL[0])), box.L[0]));
                    dy = __dsub_rn(dy, __dmul_rn(rint(__dmul_rn(dy, box.invL[1])), box.L[1]));
                    dz = __dsub_rn(dz, __dmul_rn(rint(__dmul_rn(dz, box.invL[2])), box.L[2]));
                } else if (PBC == 2) {
                    wn_minimg_tric(dx, dy, dz, box.L, box.tri);
                }
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                const double lhs = __dmul_rn(10.0, d2);
                if (lhs < 42.25) {
                    if (FILL) {
                        reinterpret_cast<int2*>(pairs)[w] = make_int2(i, j0 + k);
                        ++w;
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
        if (!FILL && i < n) row_cnt[(size_t)i * nseg + seg] = count;
    }
    if (!FILL) {
        if (tie) atomicAdd(ties, tie);
        if (near) atomicAdd(ties + 1, near);
    }
}

/* exclusive scan of n row counts into 64-bit offsets (+ total at [n]); single block */
__global__ void __launch_bounds__(1024)
wn_k_pairs_scan(const uint32_t* __restrict__ row_cnt, int n, unsigned long long* __restrict__ row_off)
{
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int b = min(threadIdx.x * per, n), e = min(b + per, n);
    unsigned long long sum = 0;
    for (int i = b; i < e; ++i) sum += row_cnt[i];
    __shared__ unsigned long long s_w[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long v = s_w[lane];
        unsigned long long iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += t;
        }
        s_w[lane] = iv - v;
        if (lane == 31) row_off[n] = iv;
    }
    __syncthreads();
    unsigned long long run = s_w[warp] + inc - sum;
    for (int i = b; i < e; ++i) { row_off[i] = run; run += row_cnt[i]; }
}

}  // namespace

/* column segments per row: enough blocks for ~8 per SM (half of them are left of the diagonal) */
int wn_pairs_segments(int n)
{
    const int nt = (n + ROWS - 1) / ROWS;
    int s = (148 * 16 + nt - 1) / std::max(nt, 1);
    return std::min(std::max(s, 1), 64);
}

void wn_launch_pairs_count(const double* xyz, int n, int first_limit, WnPairBox box, uint32_t* row_cnt,
                           unsigned long long* ties, cudaStream_t st)
{
    const int nseg = wn_pairs_segments(n), seg_len = (n + nseg - 1) / nseg;
    dim3 g((n + ROWS - 1) / ROWS, nseg);
    if (box.pbc == 2)
        wn_k_pairs<false, 2><<<g, ROWS, 0, st>>>(xyz, n, first_limit, nseg, seg_len, box, row_cnt, nullptr, nullptr, ties);
    else if (box.pbc == 1)
        wn_k_pairs<false, 1><<<g, ROWS, 0, st>>>(xyz, n, first_limit, nseg, seg_len, box, row_cnt, nullptr, nullptr, ties);
    else
        wn_k_pairs<false, 0><<<g, ROWS, 0, st>>>(xyz, n, first_limit, nseg, seg_len, box, row_cnt, nullptr, nullptr, ties);
}

/* exclusive scan over the n*nseg [row][segment] counts */
void wn_launch_pairs_scan(const uint32_t* row_cnt, int n, unsigned long long* row_off, cudaStream_t st)
{
    wn_k_pairs_scan<<<1, 1024, 0, st>>>(row_cnt, n * wn_pairs_segments(n), row_off);
}

void wn_launch_scan64(const uint32_t* cnt, int len, unsigned long long* off, cudaStream_t st)
{
    wn_k_pairs_scan<<<1, 1024, 0, st>>>(cnt, len, off);
}

void wn_launch_pairs_fill(const double* xyz, int n, int first_limit, WnPairBox box,
                          const unsigned long long* row_off, int32_t* pairs, cudaStream_t st)
{
    const int nseg = wn_pairs_segments(n), seg_len = (n + nseg - 1) / nseg;
    dim3 g((n + ROWS - 1) / ROWS, nseg);
    if (box.pbc == 2)
        wn_k_pairs<true, 2><<<g, ROWS, 0, st>>>(xyz, n, first_limit, nseg, seg_len, box, nullptr, row_off, pairs, nullptr);
    else if (box.pbc == 1)
        wn_k_pairs<true, 1><<<g, ROWS, 0, st>>>(xyz, n, first_limit, nseg, seg_len, box, nullptr, row_off, pairs, nullptr);
    else
        wn_k_pairs<true, 0><<<g, ROWS, 0, st>>>(xyz, n, first_limit, nseg, seg_len, box, nullptr, row_off, pairs, nullptr);
}
