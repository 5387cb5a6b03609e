/*
 * wn_kernels_finish.cu -- ordered edge list, energies and per-frame statistics.
 *
 * (wn_kernels_scan.cu) row scan + wn_k_frameoff : per-frame exclusive scan of the row lengths ->
 *     final position of every row, i.e. the lexicographic (i,j) order in which
 *     BruteFindPairs::find emits pairs (reference src/brute-find-pairs.cpp:21-27)
 *     and make_edges keeps them (src/waternetwork.cpp:621-633).
 * wn_k_copy  : one thread per row: rank the row's edges by j, evaluate
 *     computeEnergy (src/waternetwork.cpp:49-66, switches :12-47) for the selected
 *     hydrogen only (energies.at(pos), :257) and write the 20-byte edge records.
 * wn_k_stats : the hydrogen-bonds-num.xvg row of each frame (:659-662) plus the
 *     fp64 energy sum and the pair-evaluation count.
 */
#include "wn_internal.cuh"
#include "wn_geometry.cuh"

namespace {

constexpr int COPY_THREADS = 256;

/* ---- exclusive scan of the frame totals (single block) ---- */
__global__ void __launch_bounds__(1024)
wn_k_frameoff(const unsigned long long* __restrict__ frame_total, const uint32_t* __restrict__ frame_maxrow,
              int n_frames, unsigned long long* __restrict__ chain, unsigned long long* __restrict__ frame_off,
              unsigned long long* __restrict__ pass_info)
{
    __shared__ uint32_t s_max;
    /* the call's running edge total lives on the device: passes chain without a host round trip */
    const unsigned long long base = chain[0];
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    {
        uint32_t mx = 0;
        for (int i = threadIdx.x; i < n_frames; i += blockDim.x) mx = max(mx, frame_maxrow[i]);
        if (mx) atomicMax(&s_max, mx);
    }
    const int per = (n_frames + blockDim.x - 1) / blockDim.x;
    const int b = min(threadIdx.x * per, n_frames), e = min(b + per, n_frames);
    unsigned long long sum = 0;
    for (int i = b; i < e; ++i) sum += frame_total[i];
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
        if (lane == 31) { frame_off[n_frames] = base + iv; pass_info[0] = base + iv; chain[0] = base + iv; }
    }
    __syncthreads();
    if (threadIdx.x == 0) pass_info[1] = s_max;
    unsigned long long run = base + s_w[warp] + inc - sum;
    for (int i = b; i < e; ++i) { frame_off[i] = run; run += frame_total[i]; }
}

/* ---- ordered copy: one warp per 32 consecutive rows, one lane per edge ----
 * The 32 rows of a warp are contiguous in the final list, so the warp's output is
 * one contiguous span: lane e of a chunk takes the e-th temp entry of the span
 * (row found by a shuffle binary search over the row prefix), ranks it inside its
 * row by j, evaluates the energy and stores it at span_base + row_prefix + rank. */
#ifndef WN_COPY_MINBLOCKS
#define WN_COPY_MINBLOCKS 8
#endif
template <int LAYOUT>      /* WN_LAYOUT_EDGES / _CSR16 / _CSR12 */
__global__ void __launch_bounds__(COPY_THREADS, WN_COPY_MINBLOCKS)
wn_k_copy(const uint32_t* __restrict__ rowbuf, int row_cap,
          const uint32_t* __restrict__ row_cnt, const uint32_t* __restrict__ row_off,
          const unsigned long long* __restrict__ frame_off, const WnFrameCounters* __restrict__ counters, int n_sites,
          WnEnergyParams ep, wn_edge* __restrict__ edges, unsigned long long edge_cap, unsigned long long edge_base,
          float* __restrict__ energy_side, double* __restrict__ block_sums)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t o = (size_t)f * n_sites + (row < n_sites ? row : 0);
    const uint32_t cnt = row < n_sites ? row_cnt[o] : 0u;
    const uint32_t off = row < n_sites ? row_off[o] : 0u;
    /* exclusive prefix of the row lengths inside the warp */
    uint32_t inc = cnt;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, inc, s);
        if (lane >= s) inc += t;
    }
    const uint32_t pre = inc - cnt;
    const uint32_t total = __shfl_sync(FULL, inc, 31);
    /* frame_off carries the offset of this pass inside the call's edge array */
    const unsigned long long span = frame_off[f] - edge_base + __shfl_sync(FULL, off, 0);
    /* The warp's energy sum must not depend on the arrival order of the row entries.  Each (float) energy is added
     * as an integer multiple of 2^-32 -- integer addition is associative, so no second pass over the records just
     * written is needed (that re-read missed L2 often enough to cost 0.9 GB of DRAM reads per 250 frames).
     * Range: |E| <= 3855/r^6 < 2^17 for every edge of at least WN_SHORT_EDGE_A = 0.56 A (cos^4 and both switches are
     * <= 1), so a warp's 32 rows of up to 512 slots stay inside 63 bits; resolution 2.3e-10 per edge, i.e. <= 1e-9
     * relative on frame sums of 1e3..1e6 against a 1e-6 contract.  A frame with a shorter edge (the fused kernel
     * counts them) or a row capacity beyond 512 is summed by wn_k_stats from the written records in final order
     * instead; which path runs depends on the data only, so the sum stays run-to-run deterministic either way. */
    long long fix = 0;
    if (total && span + total <= edge_cap) {
        for (uint32_t e0 = 0; e0 < total; e0 += 32) {
            const uint32_t e = min(e0 + lane, total - 1);
            /* largest r with pre[r] <= e  (pre is non-decreasing over lanes) */
            int r = 0;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const uint32_t t = __shfl_sync(FULL, pre, r + s);
                if (t <= e) r += s;
            }
            const uint32_t rpre = __shfl_sync(FULL, pre, r);
            const uint32_t rcnt = __shfl_sync(FULL, cnt, r);
            if (e0 + lane < total) {
                const int irow = blockIdx.x * blockDim.x + warp * 32 + r;
                const uint4* rowp = reinterpret_cast<const uint4*>(rowbuf) + ((size_t)f * n_sites + irow) * (size_t)row_cap;
                WN_CHECK(e >= rpre && e - rpre < rcnt && rcnt <= (uint32_t)row_cap && irow < n_sites);
                const uint4 t = rowp[e - rpre];
                const int tj = (int)t.x;
                WN_CHECK(tj > irow && tj < n_sites);
                const float tlen = __uint_as_float(t.y), tcos = __uint_as_float(t.z);
                /* rank inside the row by j */
#ifndef WN_COPY_PLAIN_RANK
                /* entries with j' >= j counted through the carry of j' - j (0 <= j < 2^27: the unsigned borrow is the
                 * comparison): two instructions per entry instead of compare + add + select */
                uint32_t ge = 0;
                for (uint32_t k = 0; k < rcnt; ++k) {
                    uint32_t tmp;
                    asm("sub.cc.u32 %1, %2, %3;\n\taddc.u32 %0, %0, 0;" : "+r"(ge), "=r"(tmp) : "r"(rowp[k].x), "r"((uint32_t)tj));
                }
                const uint32_t rank = rcnt - ge;
#else
                uint32_t rank = 0;
                for (uint32_t k = 0; k < rcnt; ++k) rank += ((int)rowp[k].x < tj) ? 1u : 0u;
#endif
                WN_CHECK(rank < rcnt && span + rpre + rank < edge_cap);
                /* energies.push_back(computeEnergy(distance, proj)): float  (:222,:234) */
                const float en = __double2float_rn(wn_compute_energy((double)tlen, (double)tcos, ep));
                fix += __double2ll_rn((double)en * 4294967296.0);
                if (LAYOUT == WN_LAYOUT_CSR16) {
                    /* compact layout: {j, energy, length, cosine}, one aligned 16-byte store */
                    reinterpret_cast<uint4*>(edges)[span + rpre + rank] =
                        make_uint4((uint32_t)tj, __float_as_uint(en), __float_as_uint(tlen), __float_as_uint(tcos));
                } else if (LAYOUT == WN_LAYOUT_CSR12) {
                    /* {j, length, cosine}: the energy is a function of the last two and stays on the device
                     * (side array in final order, for the statistics row only) */
                    uint32_t* o = reinterpret_cast<uint32_t*>(edges) + (span + rpre + rank) * 3;
                    o[0] = (uint32_t)tj; o[1] = __float_as_uint(tlen); o[2] = __float_as_uint(tcos);
                    energy_side[span + rpre + rank] = en;
                } else {
                    wn_edge out;
                    out.i = irow; out.j = tj; out.energy = en; out.length = tlen; out.cosine = tcos;
                    edges[span + rpre + rank] = out;
                }
            }
        }
    }
    /* one partial per warp (no block barrier), summed in a fixed order by wn_k_stats; NaN = "this frame has energies the
     * fixed-point sum may not hold": wn_k_stats then sums the warp's records in final order */
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) fix += __shfl_down_sync(FULL, fix, s);
    if (lane == 0)
        block_sums[((size_t)f * gridDim.x + blockIdx.x) * (COPY_THREADS / 32) + warp] =
            (counters[f].short_edges || row_cap > 512) ? __longlong_as_double(0x7ff8000000000000ll)
                                                       : (double)fix * 2.3283064365386963e-10;   /* 2^-32 */
}

/* ---- per-frame statistics row: one warp per frame ----
 * The partial sums are added in a fixed order (lane-strided, then a shuffle tree), so the energy sum
 * does not depend on scheduling. */
__global__ void __launch_bounds__(128)
wn_k_stats(const double* __restrict__ block_sums, int n_blocks,
           const unsigned long long* __restrict__ frame_total,
           const WnFrameCounters* __restrict__ counters, int n_frames, int n_table, WnSubset sub,
           wn_frame_stats* __restrict__ stats,
           const wn_edge* __restrict__ edges, int layout, const float* __restrict__ energy_side,
           const uint32_t* __restrict__ row_off, const unsigned long long* __restrict__ frame_off)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (f >= n_frames) return;
    double s = 0.0;
    for (int b0 = 0; b0 < n_blocks; b0 += 32) {
        const int b = b0 + lane;
        double v = b < n_blocks ? block_sums[(size_t)f * n_blocks + b] : 0.0;
        /* partials flagged by wn_k_copy (energies beyond the fixed-point range): the 32 rows of that warp, summed by
         * this warp in final order with a fixed tree */
        unsigned flagged = __ballot_sync(FULL, v != v);
        while (flagged) {
            const int src = __ffs(flagged) - 1;
            flagged &= flagged - 1;
            const int row0 = (b0 + src) * 32;                                  /* first row of the flagged warp */
            unsigned long long lo = 0, hi = 0;                                 /* warps past the last row own nothing */
            if (row0 < n_table) {
                lo = frame_off[f] + row_off[(size_t)f * n_table + row0];
                hi = row0 + 32 < n_table ? frame_off[f] + row_off[(size_t)f * n_table + row0 + 32] : frame_off[f + 1];
            }
            double t = 0.0;
            for (unsigned long long e = lo + lane; e < hi; e += 32)
                t += layout == WN_LAYOUT_CSR16 ? (double)__uint_as_float(reinterpret_cast<const uint4*>(edges)[e].y)
                   : layout == WN_LAYOUT_CSR12 ? (double)energy_side[e]
                                               : (double)edges[e].energy;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(FULL, t, o);
            t = __shfl_sync(FULL, t, 0);
            if (lane == src) v = t;
        }
        s += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(FULL, s, o);
    /* The fixed-point partials carry up to 2^-33 of rounding per edge (only energies under 2^-8 round at all).  Where
     * that bound could reach 1e-7 of the frame's sum -- a frame of a few edges that all sit at the cutoff or at
     * cosine ~ 0, |E| ~ 1e-8 -- the sum is taken from the written float energies in final order instead, so the 1e-6
     * relative contract of sum_energy holds for every frame.  The choice depends on the frame's data only. */
    s = __shfl_sync(FULL, s, 0);
    const unsigned long long n_edges = frame_total[f];
    if (n_edges && !(fabs(s) * 1e-7 >= (double)n_edges * 1.1641532182693481e-10)) {      /* 2^-33 */
        const unsigned long long lo = frame_off[f], hi = frame_off[f + 1];
        double t = 0.0;
        for (unsigned long long e = lo + lane; e < hi; e += 32)
            t += layout == WN_LAYOUT_CSR16 ? (double)__uint_as_float(reinterpret_cast<const uint4*>(edges)[e].y)
               : layout == WN_LAYOUT_CSR12 ? (double)energy_side[e]
                                           : (double)edges[e].energy;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(FULL, t, o);
        s = t;
    }
    if (lane) return;
    const int n_sites = wn_frame_sites(sub, f, n_table);
    wn_frame_stats r;
    r.num_sites = (double)n_sites;                              /* :659 */
    r.num_edges = (double)frame_total[f];                       /* :660 */
    r.ratio = n_sites > 0 ? 1.0 * (double)frame_total[f] / (double)n_sites : 0.0;   /* :661 */
    r.sum_energy = s;
    r.pair_evals = (double)counters[f].evals;
    stats[f] = r;
}

}  // namespace

void wn_launch_frameoff(const unsigned long long* frame_total, const uint32_t* frame_maxrow, int n_frames,
                        unsigned long long* chain, unsigned long long* frame_off,
                        unsigned long long* pass_info, cudaStream_t st)
{
    wn_k_frameoff<<<1, 1024, 0, st>>>(frame_total, frame_maxrow, n_frames, chain, frame_off, pass_info);
}

void wn_launch_copy(const uint32_t* rowbuf, int row_cap, const uint32_t* row_cnt,
                    const uint32_t* row_off, const unsigned long long* frame_off, const WnFrameCounters* counters,
                    int n_frames, int n_sites, WnEnergyParams ep, wn_edge* edges,
                    unsigned long long edge_cap, int layout, float* energy_side, double* block_sums, int* n_blocks_out,
                    cudaStream_t st)
{
    dim3 g((n_sites + COPY_THREADS - 1) / COPY_THREADS, n_frames);
    *n_blocks_out = (int)g.x * (COPY_THREADS / 32);      /* partial sums per frame */
    /* edge_cap counts 20-byte records; the compact layouts use the same buffer */
    if (layout == WN_LAYOUT_CSR16)
        wn_k_copy<WN_LAYOUT_CSR16><<<g, COPY_THREADS, 0, st>>>(rowbuf, row_cap, row_cnt, row_off, frame_off, counters, n_sites, ep,
                                                               edges, edge_cap, 0ull, energy_side, block_sums);
    else if (layout == WN_LAYOUT_CSR12)
        wn_k_copy<WN_LAYOUT_CSR12><<<g, COPY_THREADS, 0, st>>>(rowbuf, row_cap, row_cnt, row_off, frame_off, counters, n_sites, ep,
                                                               edges, edge_cap, 0ull, energy_side, block_sums);
    else
        wn_k_copy<WN_LAYOUT_EDGES><<<g, COPY_THREADS, 0, st>>>(rowbuf, row_cap, row_cnt, row_off, frame_off, counters, n_sites, ep,
                                                               edges, edge_cap, 0ull, energy_side, block_sums);
}

void wn_launch_stats(const double* block_sums, int n_blocks, const unsigned long long* frame_total,
                     const WnFrameCounters* counters, int n_frames, int n_sites, WnSubset sub,
                     wn_frame_stats* stats, const wn_edge* edges, int layout, const float* energy_side,
                     const uint32_t* row_off, const unsigned long long* frame_off, cudaStream_t st)
{
    wn_k_stats<<<(n_frames + 3) / 4, 128, 0, st>>>(block_sums, n_blocks, frame_total, counters,
                                                       n_frames, n_sites, sub, stats, edges, layout, energy_side, row_off, frame_off);
}
