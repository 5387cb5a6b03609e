/*
 * wn_kernels_pairlist.cu -- make_edges over a CALLER-SUPPLIED pair list (sm_100a).
 *
 * The reference's make_edges (src/waternetwork.cpp:182-263) takes whatever pair list its finder
 * returned -- for the live module that is DelaunayFindPairs' list (src/waternetwork.cpp:399): Delaunay
 * edges in finite_edges order, with a (first, second) orientation that is NOT i < j.  The orientation
 * decides the sign of OO, which hydrogen list comes first in the arg-max, the `pos > 0` rule and which
 * site is printed first.  This file evaluates exactly that function on one frame: one thread per pair in
 * the given order and orientation, the same exact fp64 geometry as the fused kernel (wn_geometry.cuh),
 * and an order-preserving compaction of the pairs that become edges.
 *
 *   (wn_k_scatter)     site records in table order (identity permutation): float position, meta word,
 *                      unit D-H vectors -- the staging of the batched path, reused
 *   wn_k_pl_eval       pair -> {length, cosine} or "no edge"; per-block edge counts; tie/eval counters
 *   (scan)             block offsets
 *   wn_k_pl_write      edges in pair order: {first, second, energy, length, cosine}
 */
#include "wn_internal.cuh"
#include "wn_geometry.cuh"

namespace {

constexpr int PL_THREADS = 256;
constexpr unsigned FULL = 0xffffffffu;

__global__ void __launch_bounds__(256)
wn_k_pl_iota(int* __restrict__ rank, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rank[i] = i;
}

template <bool SIMPLE, int PBC>
__global__ void __launch_bounds__(PL_THREADS)
wn_k_pl_eval(const float4* __restrict__ spos, const uint32_t* __restrict__ smeta, const double* __restrict__ sdh, int hs,
             const WnGrid* __restrict__ grid, const int2* __restrict__ pairs, unsigned long long m, int n_sites,
             float2* __restrict__ tmp, uint32_t* __restrict__ blk_cnt, WnFrameCounters* __restrict__ counters,
             unsigned int* __restrict__ bad_index)
{
    const unsigned long long p = (unsigned long long)blockIdx.x * PL_THREADS + threadIdx.x;
    bool edge = false;
    PairOut r;
    r.counted = r.edge = r.t_len = r.t_cos = r.t_pos0 = false;
    r.len_f = r.cosmax = 0.f;
    if (p < m) {
        const int2 pr = pairs[p];
        if (pr.x < 0 || pr.x >= n_sites || pr.y < 0 || pr.y >= n_sites) {
            atomicExch(bad_index, 1u);           /* the reference throws std::out_of_range from .at() */
        } else {
            float4 a4 = __ldg(spos + pr.x);
            const float4 b4 = __ldg(spos + pr.y);
            const uint32_t ma = __ldg(smeta + pr.x), mb = __ldg(smeta + pr.y);
            /* type rule (:199-203) and the no-hydrogen rule (:205-208) */
            const uint32_t ta = ma & WN_META_TYPE_MASK, tb = mb & WN_META_TYPE_MASK;
            bool pass = (ta != tb) || (ta == WN_WATER && tb == WN_WATER);
            pass = pass && (((ma | mb) & WN_META_HASH) != 0);
            if (pass) {
                /* wn_pair_geometry gives the first role to the lower index bits: the pair's first site leads */
                a4.w = __int_as_float(-1);
                const WnGrid& g = grid[0];
                r = wn_pair_geometry<SIMPLE, PBC>(a4, b4, sdh + (size_t)pr.x * hs * 4, sdh + (size_t)pr.y * hs * 4, ma, mb, g);
                edge = r.edge;
            }
        }
        tmp[p] = edge ? make_float2(r.len_f, r.cosmax) : make_float2(-1.f, 0.f);
    }
    const int n_edge = __syncthreads_count(edge);
    const int n_cnt = __syncthreads_count(r.counted);
    const int n_tl = __syncthreads_count(r.t_len), n_tc = __syncthreads_count(r.t_cos), n_t0 = __syncthreads_count(r.t_pos0);
    if (threadIdx.x == 0) {
        blk_cnt[blockIdx.x] = (uint32_t)n_edge;
        if (n_cnt) atomicAdd(&counters->evals, (unsigned long long)n_cnt);
        if (n_tl) atomicAdd(&counters->tie_len, (unsigned long long)n_tl);
        if (n_tc) atomicAdd(&counters->tie_cos, (unsigned long long)n_tc);
        if (n_t0) atomicAdd(&counters->tie_pos0, (unsigned long long)n_t0);
    }
}

__global__ void __launch_bounds__(PL_THREADS)
wn_k_pl_write(const int2* __restrict__ pairs, const float2* __restrict__ tmp, unsigned long long m,
              const unsigned long long* __restrict__ blk_off, WnEnergyParams ep, wn_edge* __restrict__ edges,
              unsigned long long edge_cap)
{
    __shared__ int s_w[PL_THREADS / 32];
    const unsigned long long p = (unsigned long long)blockIdx.x * PL_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2 t = make_float2(-1.f, 0.f);
    if (p < m) t = tmp[p];
    const bool edge = t.x >= 0.f;
    const unsigned b = __ballot_sync(FULL, edge);
    if (lane == 0) s_w[warp] = __popc(b);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_w[w];
    if (edge) {
        const unsigned long long pos = blk_off[blockIdx.x] + (unsigned long long)before + __popc(b & ((1u << lane) - 1u));
        WN_CHECK(pos < edge_cap);
        if (pos < edge_cap) {
            const int2 pr = pairs[p];
            wn_edge out;
            out.i = pr.x; out.j = pr.y;
            /* energies.at(pos): computeEnergy(distance, proj) of the selected hydrogen (:222,:234,:257) */
            out.energy = __double2float_rn(wn_compute_energy((double)t.x, (double)t.y, ep));
            out.length = t.x; out.cosine = t.y;
            edges[pos] = out;
        }
    }
}

}  // namespace

void wn_launch_pl_iota(int* rank, int n, cudaStream_t st)
{
    if (n > 0) wn_k_pl_iota<<<(n + 255) / 256, 256, 0, st>>>(rank, n);
}

int wn_pl_blocks(unsigned long long m) { return (int)((m + PL_THREADS - 1) / PL_THREADS); }

void wn_launch_pl_eval(const float4* spos, const uint32_t* smeta, const double* sdh, int hs, int simple, int pbc,
                       const WnGrid* grid, const int2* pairs, unsigned long long m, int n_sites, float2* tmp,
                       uint32_t* blk_cnt, WnFrameCounters* counters, unsigned int* bad_index, cudaStream_t st)
{
    const int nb = wn_pl_blocks(m);
    if (nb == 0) return;
#define WN_PL(S, P) wn_k_pl_eval<S, P><<<nb, PL_THREADS, 0, st>>>(spos, smeta, sdh, hs, grid, pairs, m, n_sites, tmp, blk_cnt, counters, bad_index)
    if (simple) { if (pbc == 2) WN_PL(true, 2); else if (pbc == 1) WN_PL(true, 1); else WN_PL(true, 0); }
    else        { if (pbc == 2) WN_PL(false, 2); else if (pbc == 1) WN_PL(false, 1); else WN_PL(false, 0); }
#undef WN_PL
}

void wn_launch_pl_write(const int2* pairs, const float2* tmp, unsigned long long m, const unsigned long long* blk_off,
                        WnEnergyParams ep, wn_edge* edges, unsigned long long edge_cap, cudaStream_t st)
{
    const int nb = wn_pl_blocks(m);
    if (nb == 0) return;
    wn_k_pl_write<<<nb, PL_THREADS, 0, st>>>(pairs, tmp, m, blk_off, ep, edges, edge_cap);
}
