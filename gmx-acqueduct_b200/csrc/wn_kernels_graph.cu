/*
 * wn_kernels_graph.cu -- connected components of the per-frame hydrogen-bond networks (sm_100a).
 *
 * The first thing the reference's downstream analysis does with edges-frames.xvg.dat
 * (python/network-analysis.py:141-171) is, per frame: build a graph from the edge rows with
 * weight = |energy| (:169-171), drop the edges with weight <= threshold (filter_edge_by_weight,
 * :16-24), drop the nodes left without edges (remove_node_degree, :26-35) and take the connected
 * component of chosen nodes (nx.node_connected_component, :155).  Here the components of ALL frames
 * of a batch are labelled on the device, straight from the edge list the pipeline left in HBM
 * (lock-free union-find; an edge is kept unless |energy| <= threshold, so a NaN energy is kept, as
 * the Python comparison keeps it).  Frames whose parent array fits one SM's shared memory (up to ~56 k
 * sites) run in ONE CTA each, entirely on chip (wn_k_cc_smem, below); larger frames take the
 * multi-kernel global-memory path:
 *   wn_k_cc_iota      parent[v] = v
 *   wn_k_cc_sample    round k = 0, 1: every site joins the k-th kept edge of its own row, then
 *   wn_k_cc_compress  parent[v] = root(v)                      (neighbour sampling, as in Afforest)
 *   wn_k_cc_hook      one thread per edge: union of the two ends (nearly always already joined)
 *   wn_k_cc_flatten   parent[v] = root(v), or -1 for a site without a kept edge
 *   wn_k_cc_collect   smallest member of every component (atomicMin per warp-level group)
 *   wn_k_cc_relabel   label[v] = smallest member of v's component; component sizes
 *   wn_k_cc_stats     per frame: components, largest component, sites in components, kept edges
 * Trees are linked by a hashed priority of the root (larger hash under smaller hash): linking by the
 * site index itself builds chains as long as the index order of the input (a lattice-ordered box
 * gave 4.6 ms for one sampling round), while random priorities keep the expected depth logarithmic.
 * The labels are made canonical afterwards (smallest member), so the comparison with a CPU
 * union-find is exact.  Integer work on L2-resident arrays (132 KB of parents per 100k-atom
 * frame); 12 of the 20 (16) bytes of every edge are read three times (two sampling rounds, one hook).
 */
#include "wn_internal.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int CC_THREADS = 256;
#ifndef WN_CC_HOOK_SLICES
#define WN_CC_HOOK_SLICES 4
#endif
constexpr int CC_HOOK_SLICES = WN_CC_HOOK_SLICES;
constexpr int CC_SAMPLE_ROUNDS = 2;       /* measured at 250 x 33,333 sites, all edges kept: 0 -> 21.3 ms, 1 -> 7.8, 2 -> 4.1, 3 -> 4.5, 4 -> 4.7 */
constexpr unsigned FULL = 0xffffffffu;

/* bijective hash of a site index: the linking priority of a root */
__device__ __forceinline__ uint32_t wn_cc_prio(int v) { return (uint32_t)v * 0x9E3779B1u; }

/* root of v with path halving; concurrent halving writes store a valid ancestor */
__device__ __forceinline__ int wn_cc_find(volatile int* parent, int v)
{
    int p = parent[v];
    while (p != v) {
        const int gp = parent[p];
        if (gp != p) parent[v] = gp;
        v = p;
        p = gp;
    }
    return v;
}

/* lock-free union: the root with the larger priority value goes under the other one */
__device__ __forceinline__ void wn_cc_union(volatile int* parent, int a, int b)
{
    int ra = wn_cc_find(parent, a), rb = wn_cc_find(parent, b);
    while (ra != rb) {
        const bool a_low = wn_cc_prio(ra) < wn_cc_prio(rb);
        const int child = a_low ? rb : ra, top = a_low ? ra : rb;
        const int old = atomicCAS(const_cast<int*>(parent) + child, child, top);
        if (old == child) break;
        ra = wn_cc_find(parent, child);                    /* somebody else moved it: follow and retry */
        rb = wn_cc_find(parent, top);
    }
}

__device__ __forceinline__ void wn_cc_edge(const void* __restrict__ edges, bool csr16, unsigned long long e, int& j, float& en)
{
    if (csr16) {
        const uint4 r = reinterpret_cast<const uint4*>(edges)[e];
        j = (int)r.x; en = __uint_as_float(r.y);
    } else {
        const wn_edge* r = reinterpret_cast<const wn_edge*>(edges) + e;
        j = r->j; en = r->energy;
    }
}

__global__ void __launch_bounds__(CC_THREADS)
wn_k_cc_iota(int* __restrict__ parent, int* __restrict__ minidx, int n_sites)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n_sites) {
        parent[(size_t)blockIdx.y * n_sites + v] = v;
        minidx[(size_t)blockIdx.y * n_sites + v] = 0x7fffffff;
    }
}

/* Neighbour sampling: every site first joins the K-th kept edge of its own row.  With one union per site
 * few threads meet at the same root; after two rounds nearly all of a liquid's network is one tree, and the
 * pass over ALL edges finds both ends under the same root with two reads per edge.  (Hooking all ~12 edges
 * per site at once was measured 5x slower: every thread of the frame ends up at the giant component's root.) */
template <bool CSR16>
__global__ void __launch_bounds__(CC_THREADS)
wn_k_cc_sample(const void* __restrict__ edges, const unsigned long long* __restrict__ frame_off,
               const uint32_t* __restrict__ row_off, int n_sites, float threshold, int K, int* __restrict__ parent_all)
{
    const int f = blockIdx.y;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_sites) return;
    const unsigned long long b = frame_off[f], n = frame_off[f + 1] - b;
    const uint32_t* ro = row_off + (size_t)f * n_sites;
    const unsigned long long r0 = ro[v], r1 = (v + 1 < n_sites) ? (unsigned long long)ro[v + 1] : n;
    int seen = 0;
    for (unsigned long long e = r0; e < r1; ++e) {
        int j; float en;
        wn_cc_edge(edges, CSR16, b + e, j, en);
        if (fabsf(en) <= threshold) continue;              /* filter_edge_by_weight: removed iff weight <= threshold */
        if (seen++ == K) { wn_cc_union(parent_all + (size_t)f * n_sites, v, j); break; }
    }
}

/* parent[v] = root(v).  In place: no union runs concurrently, so roots are final, and a walk that passes
 * through v reads either v's old parent or v's root -- both ancestors of v on the way to the same root. */
__global__ void __launch_bounds__(CC_THREADS)
wn_k_cc_compress(int* __restrict__ parent_all, int n_sites)
{
    const int f = blockIdx.y;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_sites) return;
    volatile int* parent = parent_all + (size_t)f * n_sites;
    int r = v, p = parent[r];
    while (p != r) { r = p; p = parent[r]; }
    parent[v] = r;
}

template <bool CSR16>
__global__ void __launch_bounds__(CC_THREADS)
wn_k_cc_hook(const void* __restrict__ edges, const unsigned long long* __restrict__ frame_off,
             const uint32_t* __restrict__ row_off, int n_sites, float threshold, int slice, int n_slices,
             int* __restrict__ parent_all, int* __restrict__ has_all, unsigned int* __restrict__ kept)
{
    const int f = blockIdx.y;
    const unsigned long long b = frame_off[f], n_all = frame_off[f + 1] - b;
    /* this launch handles the slice-th part of the frame's edges (the trees are flattened between slices) */
    const unsigned long long per = (n_all + n_slices - 1) / n_slices;
    const unsigned long long e_lo = (unsigned long long)slice * per, n = min(n_all, e_lo + per);
    volatile int* parent = parent_all + (size_t)f * n_sites;
    int* has = has_all + (size_t)f * n_sites;             /* site has a kept edge (remove_node_degree keeps it) */
    unsigned int mine = 0;
    for (unsigned long long e = e_lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (unsigned long long)gridDim.x * blockDim.x) {
        int i, j;
        float en;
        wn_cc_edge(edges, CSR16, b + e, j, en);
        if (fabsf(en) <= threshold) continue;
        if (CSR16) {
            /* row of edge e: largest i with row_off[i] <= e */
            const uint32_t* ro = row_off + (size_t)f * n_sites;
            int lo = 0, hi = n_sites - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if ((unsigned long long)__ldg(ro + mid) <= e) lo = mid; else hi = mid - 1;
            }
            i = lo;
        } else {
            i = (reinterpret_cast<const wn_edge*>(edges) + (b + e))->i;
        }
        ++mine;
        has[i] = 1; has[j] = 1;
        wn_cc_union(parent, i, j);
    }
    /* kept-edge count of the frame */
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(FULL, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(kept + f, mine);
}

/* parent[v] = root(v), or -1 for a site without a kept edge (never visited by a walk) */
__global__ void __launch_bounds__(CC_THREADS)
wn_k_cc_flatten(int* __restrict__ parent_all, const int* __restrict__ has_all, int n_sites)
{
    const int f = blockIdx.y;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_sites) return;
    volatile int* parent = parent_all + (size_t)f * n_sites;
    int r = -1;
    if (has_all[(size_t)f * n_sites + v]) {
        r = v;
        int p = parent[r];
        while (p != r) { r = p; p = parent[r]; }
    }
    parent[v] = r;
}

/* smallest member of every component; lanes of a warp under the same root issue one atomic */
__global__ void __launch_bounds__(CC_THREADS)
wn_k_cc_collect(const int* __restrict__ root_all, int* __restrict__ minidx_all, int n_sites)
{
    const int f = blockIdx.y;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = v < n_sites ? root_all[(size_t)f * n_sites + v] : -1;
    const unsigned grp = __match_any_sync(FULL, r);
    if (r < 0) return;
    const int lowest = __reduce_min_sync(grp, v);
    if ((int)(threadIdx.x & 31) == __ffs(grp) - 1) atomicMin(minidx_all + (size_t)f * n_sites + r, lowest);
}

/* label[v] = smallest member of the component (in place over the roots); sizes collected at the label */
__global__ void __launch_bounds__(CC_THREADS)
wn_k_cc_relabel(int* __restrict__ label_all, const int* __restrict__ minidx_all, int* __restrict__ count_all, int n_sites)
{
    const int f = blockIdx.y;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    int l = -1;
    if (v < n_sites) {
        const int r = label_all[(size_t)f * n_sites + v];
        if (r >= 0) l = minidx_all[(size_t)f * n_sites + r];
    }
    const unsigned grp = __match_any_sync(FULL, l);
    if (l < 0) return;
    /* nobody reads label_all[v] as a root any more: every thread read its own slot above, and the
     * minidx lookups go to a different array */
    label_all[(size_t)f * n_sites + v] = l;
    if ((int)(threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(count_all + (size_t)f * n_sites + l, __popc(grp));
}

__global__ void __launch_bounds__(CC_THREADS)
wn_k_cc_stats(const int* __restrict__ labels_all, const int* __restrict__ count_all, const unsigned int* __restrict__ kept,
              int n_sites, wn_component_stats* __restrict__ stats)
{
    const int f = blockIdx.x;
    const int* labels = labels_all + (size_t)f * n_sites;
    const int* count = count_all + (size_t)f * n_sites;
    int comps = 0, largest = 0, nodes = 0;
    for (int v = threadIdx.x; v < n_sites; v += blockDim.x) {
        const int l = labels[v];
        if (l >= 0) ++nodes;
        if (l == v) { ++comps; largest = max(largest, count[v]); }
    }
    __shared__ int s_c[CC_THREADS / 32], s_l[CC_THREADS / 32], s_n[CC_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        comps += __shfl_down_sync(FULL, comps, o);
        nodes += __shfl_down_sync(FULL, nodes, o);
        largest = max(largest, __shfl_down_sync(FULL, largest, o));
    }
    if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = comps; s_l[threadIdx.x >> 5] = largest; s_n[threadIdx.x >> 5] = nodes; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int c = 0, l = 0, nn = 0;
        for (int w = 0; w < CC_THREADS / 32; ++w) { c += s_c[w]; l = max(l, s_l[w]); nn += s_n[w]; }
        wn_component_stats r;
        r.n_components = c; r.largest = l; r.n_nodes = nn; r.n_edges = (int32_t)kept[f];
        stats[f] = r;
    }
}

/* ---- one CTA per frame, union-find in shared memory ----
 * The parent array of a frame (4 bytes per site: 133 KB for the 100k-atom box) fits the 227 KB of one SM,
 * so the whole labelling of a frame runs in ONE CTA on shared memory: a hop of a find costs ~30 cycles
 * instead of an L2 round trip (the global version spends 98 % of its cycles waiting on those: issue
 * utilisation 2.4 %).  Same algorithm (hashed root priority, path halving, CAS hooking), then the smallest
 * member of every component is collected IN PLACE: roots switch their own slot to (member | FLAG) and the
 * members atomicMin into it, so no second array is needed.  Sizes go through the global count array. */
constexpr int CCS_THREADS = 1024;
constexpr int CC_FLAG = 1 << 30;
#ifndef WN_CCS_PHASES
#define WN_CCS_PHASES 8
#endif
constexpr int CCS_PHASES = WN_CCS_PHASES;

__device__ __forceinline__ int wn_ccs_find(volatile int* parent, int v)
{
    int p = parent[v];
    while (p != v) {
        const int gp = parent[p];
        if (gp != p) parent[v] = gp;
        v = p;
        p = gp;
    }
    return v;
}

__device__ __forceinline__ void wn_ccs_union(int* parent, int a, int b)
{
    int ra = wn_ccs_find(parent, a), rb = wn_ccs_find(parent, b);
    while (ra != rb) {
        const bool a_low = wn_cc_prio(ra) < wn_cc_prio(rb);
        const int child = a_low ? rb : ra, top = a_low ? ra : rb;
        const int old = atomicCAS(parent + child, child, top);
        if (old == child) break;
        ra = wn_ccs_find(parent, child);
        rb = wn_ccs_find(parent, top);
    }
}

template <bool CSR16>
__global__ void __launch_bounds__(CCS_THREADS, 1)
wn_k_cc_smem(const void* __restrict__ edges, const unsigned long long* __restrict__ frame_off,
             const uint32_t* __restrict__ row_off, int n_sites, float threshold,
             int* __restrict__ labels_all, int* __restrict__ count_all, wn_component_stats* __restrict__ stats)
{
    extern __shared__ int s_cc[];
    int* parent = s_cc;                                      /* [n_sites] */
    unsigned int* has = reinterpret_cast<unsigned int*>(s_cc + n_sites);   /* bit per site: has a kept edge */
    __shared__ int s_red[3][CCS_THREADS / 32];
    __shared__ unsigned int s_kept;
    const int f = blockIdx.x, tid = threadIdx.x;
    const unsigned long long b = frame_off[f], n = frame_off[f + 1] - b;
    const int n_words = (n_sites + 31) >> 5;
    for (int v = tid; v < n_sites; v += CCS_THREADS) parent[v] = v;
    for (int w = tid; w < n_words; w += CCS_THREADS) has[w] = 0u;
    if (tid == 0) s_kept = 0u;
    __syncthreads();

    unsigned int mine = 0;
    /* The edge list streams from HBM in CCS_PHASES slices; after every slice all sites are re-pointed at their
     * roots, so the finds of the next slice take one or two hops (random shared-memory reads with bank
     * conflicts are what this kernel is bound by: ~10 hops per find, 3.4 ms per 250 frames, without the
     * intermediate flattening; 1.5 ms with it). */
    auto flatten_all = [&]() {
        __syncthreads();
        for (int v = tid; v < n_sites; v += CCS_THREADS) {
            volatile int* vp = parent;
            int r = v, p = vp[r];
            while (p != r) { r = p; p = vp[r]; }
            vp[v] = r;
        }
        __syncthreads();
    };
    if (CSR16) {
        /* a warp takes 32 consecutive rows = one contiguous span of 16-byte records: lane per edge (coalesced),
         * the row index comes from a shuffle binary search over the 32 row offsets held by the lanes */
        const uint32_t* ro = row_off + (size_t)f * n_sites;
        const int lane = tid & 31, warp = tid >> 5, n_warps = CCS_THREADS / 32;
        const int n_blocks = (n_sites + 31) >> 5;
        const int slice = (n_blocks + CCS_PHASES - 1) / CCS_PHASES;
        for (int ph = 0; ph < CCS_PHASES; ++ph) {
            const int b1 = min(n_blocks, (ph + 1) * slice);
            for (int blk = ph * slice + warp; blk < b1; blk += n_warps) {
                const int v0 = blk << 5;
                const unsigned long long my = (v0 + lane < n_sites) ? (unsigned long long)ro[v0 + lane] : n;
                const unsigned long long first = __shfl_sync(FULL, my, 0);
                const unsigned long long last = (v0 + 32 < n_sites) ? (unsigned long long)ro[v0 + 32] : n;
                for (unsigned long long e0 = first; e0 < last; e0 += 64) {
                    uint4 rec[2];
                    unsigned long long ee[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        ee[u] = e0 + 32 * u + lane;
                        rec[u] = make_uint4(0u, 0u, 0u, 0u);
                        if (ee[u] < last) rec[u] = reinterpret_cast<const uint4*>(edges)[b + ee[u]];
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        /* largest r with ro[v0 + r] <= e (offsets are non-decreasing over the lanes) */
                        int r = 0;
#pragma unroll
                        for (int s = 16; s > 0; s >>= 1) {
                            const unsigned long long t = __shfl_sync(FULL, my, r + s);
                            if (t <= ee[u]) r += s;
                        }
                        if (ee[u] >= last || fabsf(__uint_as_float(rec[u].y)) <= threshold) continue;
                        const int i = v0 + r, j = (int)rec[u].x;
                        ++mine;
                        atomicOr(has + (i >> 5), 1u << (i & 31));
                        atomicOr(has + (j >> 5), 1u << (j & 31));
                        wn_ccs_union(parent, i, j);
                    }
                }
            }
            if (ph + 1 < CCS_PHASES) flatten_all();
        }
    } else {
        /* four edges per thread in flight hide the HBM latency with one CTA per SM */
        const wn_edge* ed = reinterpret_cast<const wn_edge*>(edges) + b;
        const unsigned long long slice = (n + CCS_PHASES - 1) / CCS_PHASES;
        for (int ph = 0; ph < CCS_PHASES; ++ph) {
            const unsigned long long s0 = (unsigned long long)ph * slice, s1 = min(n, s0 + slice);
            for (unsigned long long e0 = s0 + (unsigned long long)tid; e0 < s1; e0 += 4ull * CCS_THREADS) {
                int ei[4], ej[4];
                float en[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned long long e = e0 + (unsigned long long)u * CCS_THREADS;
                    ei[u] = -1; ej[u] = 0; en[u] = 0.f;
                    if (e < s1) { ei[u] = ed[e].i; ej[u] = ed[e].j; en[u] = ed[e].energy; }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (ei[u] < 0 || fabsf(en[u]) <= threshold) continue;    /* filter_edge_by_weight: removed iff weight <= threshold */
                    ++mine;
                    atomicOr(has + (ei[u] >> 5), 1u << (ei[u] & 31));
                    atomicOr(has + (ej[u] >> 5), 1u << (ej[u] & 31));
                    wn_ccs_union(parent, ei[u], ej[u]);
                }
            }
            if (ph + 1 < CCS_PHASES) flatten_all();
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(FULL, mine, o);
    if ((tid & 31) == 0 && mine) atomicAdd(&s_kept, mine);
    /* every member points at its root (in-place argument as in wn_k_cc_compress) */
    flatten_all();
    /* roots: own slot becomes (smallest member so far | FLAG); sites without an edge: -1 */
    for (int v = tid; v < n_sites; v += CCS_THREADS) {
        const bool member = (has[v >> 5] >> (v & 31)) & 1u;
        if (!member) parent[v] = -1;
        else if (parent[v] == v) parent[v] = v | CC_FLAG;
    }
    __syncthreads();
    for (int v = tid; v < n_sites; v += CCS_THREADS) {
        const int p = parent[v];
        if (p >= 0 && !(p & CC_FLAG)) atomicMin(parent + p, v | CC_FLAG);     /* p is a root slot: flagged values only */
    }
    __syncthreads();
    /* labels out, component sizes at the label's slot of the global count array (zeroed by the launcher) */
    int* labels = labels_all + (size_t)f * n_sites;
    int* count = count_all + (size_t)f * n_sites;
    int nodes = 0;
    for (int v0 = 0; v0 < n_sites; v0 += CCS_THREADS) {
        const int v = v0 + tid;
        int l = -1;
        if (v < n_sites) {
            const int p = parent[v];
            if (p >= 0) l = ((p & CC_FLAG) ? p : parent[p]) & ~CC_FLAG;
            labels[v] = l;
        }
        const unsigned grp = __match_any_sync(FULL, l);
        if (l >= 0) {
            ++nodes;
            if ((tid & 31) == __ffs(grp) - 1) atomicAdd(count + l, __popc(grp));
        }
    }
    __syncthreads();
    int comps = 0, largest = 0;
    for (int v = tid; v < n_sites; v += CCS_THREADS) {
        const int p = parent[v];
        /* the smallest member of a component is the site whose label is itself */
        if (p >= 0) {
            const int l = ((p & CC_FLAG) ? p : parent[p]) & ~CC_FLAG;
            if (l == v) { ++comps; largest = max(largest, (int)*reinterpret_cast<volatile int*>(count + v)); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        comps += __shfl_down_sync(FULL, comps, o);
        nodes += __shfl_down_sync(FULL, nodes, o);
        largest = max(largest, __shfl_down_sync(FULL, largest, o));
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = comps; s_red[1][tid >> 5] = largest; s_red[2][tid >> 5] = nodes; }
    __syncthreads();
    if (tid == 0) {
        int c = 0, l = 0, nn = 0;
        for (int w = 0; w < CCS_THREADS / 32; ++w) { c += s_red[0][w]; l = max(l, s_red[1][w]); nn += s_red[2][w]; }
        wn_component_stats r;
        r.n_components = c; r.largest = l; r.n_nodes = nn; r.n_edges = (int32_t)s_kept;
        stats[f] = r;
    }
}

}  // namespace

/* labels: [F][N] (the union-find parent array while the kernels run); work, count, minidx: [F][N] ints; kept: [F] */
void wn_launch_components(const void* edges, int csr16, const unsigned long long* frame_off, const uint32_t* row_off,
                          int n_frames, int n_sites, unsigned long long max_edges, float threshold,
                          int* labels, int* work, int* count, int* minidx, unsigned int* kept,
                          wn_component_stats* stats, cudaStream_t st)
{
    if (n_frames <= 0) return;
    const size_t n = (size_t)n_frames * (size_t)n_sites;
    cudaMemsetAsync(count, 0, n * sizeof(int), st);
    /* WN_COMPONENTS_PATH=global forces the multi-kernel path for frames that would fit shared memory (tests) */
    const char* force = getenv("WN_COMPONENTS_PATH");
    const bool allow_smem = !force || strcmp(force, "global") != 0;
    /* frames whose parent array fits one SM's shared memory: one CTA per frame, everything on chip */
    const size_t smem = (size_t)n_sites * sizeof(int) + (size_t)((n_sites + 31) / 32) * sizeof(unsigned int);
    if (allow_smem && n_sites > 0 && smem <= 220u * 1024u) {
        if (csr16) {
            cudaFuncSetAttribute(wn_k_cc_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            wn_k_cc_smem<true><<<n_frames, CCS_THREADS, smem, st>>>(edges, frame_off, row_off, n_sites, threshold, labels, count, stats);
        } else {
            cudaFuncSetAttribute(wn_k_cc_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            wn_k_cc_smem<false><<<n_frames, CCS_THREADS, smem, st>>>(edges, frame_off, row_off, n_sites, threshold, labels, count, stats);
        }
        return;
    }
    cudaMemsetAsync(work, 0, n * sizeof(int), st);
    cudaMemsetAsync(kept, 0, (size_t)n_frames * sizeof(unsigned int), st);
    if (n_sites > 0) {
        dim3 gv((n_sites + CC_THREADS - 1) / CC_THREADS, n_frames);
        wn_k_cc_iota<<<gv, CC_THREADS, 0, st>>>(labels, minidx, n_sites);
        if (max_edges > 0) {
            for (int k = 0; k < CC_SAMPLE_ROUNDS; ++k) {
                if (csr16) wn_k_cc_sample<true><<<gv, CC_THREADS, 0, st>>>(edges, frame_off, row_off, n_sites, threshold, k, labels);
                else wn_k_cc_sample<false><<<gv, CC_THREADS, 0, st>>>(edges, frame_off, row_off, n_sites, threshold, k, labels);
                wn_k_cc_compress<<<gv, CC_THREADS, 0, st>>>(labels, n_sites);
            }
            /* all edges in CC_HOOK_SLICES launches with a compression in between (shorter finds); a few edges per
             * thread: enough CTAs to fill the device, bounded grid for huge frames */
            const unsigned long long per = (max_edges + CC_HOOK_SLICES - 1) / CC_HOOK_SLICES;
            const unsigned long long want = (per + CC_THREADS * 4 - 1) / (CC_THREADS * 4);
            dim3 ge((unsigned int)std::min<unsigned long long>(std::max<unsigned long long>(want, 1), 65535), n_frames);
            for (int s = 0; s < CC_HOOK_SLICES; ++s) {
                if (csr16)
                    wn_k_cc_hook<true><<<ge, CC_THREADS, 0, st>>>(edges, frame_off, row_off, n_sites, threshold, s, CC_HOOK_SLICES, labels, work, kept);
                else
                    wn_k_cc_hook<false><<<ge, CC_THREADS, 0, st>>>(edges, frame_off, row_off, n_sites, threshold, s, CC_HOOK_SLICES, labels, work, kept);
                if (s + 1 < CC_HOOK_SLICES) wn_k_cc_compress<<<gv, CC_THREADS, 0, st>>>(labels, n_sites);
            }
        }
        wn_k_cc_flatten<<<gv, CC_THREADS, 0, st>>>(labels, work, n_sites);
        wn_k_cc_collect<<<gv, CC_THREADS, 0, st>>>(labels, minidx, n_sites);
        wn_k_cc_relabel<<<gv, CC_THREADS, 0, st>>>(labels, minidx, count, n_sites);
    }
    wn_k_cc_stats<<<n_frames, CC_THREADS, 0, st>>>(labels, count, kept, n_sites, stats);
}
