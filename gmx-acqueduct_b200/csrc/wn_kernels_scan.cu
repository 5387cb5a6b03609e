/*
 * wn_kernels_scan.cu -- per-frame exclusive scans and bounding boxes with many blocks per
 * frame (sm_100a).  The fused path needs, for every frame of a batch,
 *   * the bounding box of the sites (open boundary grid),
 *   * cell starts = exclusive scan of the cell populations (counting sort),
 *   * row offsets = exclusive scan of the row lengths (the lexicographic position of every
 *     row: edge order == pair order, reference src/waternetwork.cpp:621-633).
 * One block per frame is enough at 33k sites x 250 frames; at 1M atoms x 24 frames it left
 * most of the GPU idle (17 % of the pipeline), hence chunked three-phase versions:
 * partial sums per 2048-element chunk -> scan of the chunk sums per frame -> final offsets.
 */
#include "wn_internal.cuh"

#include <math.h>

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;     /* 2048 */

__device__ __forceinline__ int frame_len(const WnGrid* grid, int fixed_len, int f)
{
    return grid ? grid[f].ncells : fixed_len;
}

/* phase A: sum (and max) of every chunk */
__global__ void __launch_bounds__(SCAN_THREADS)
wn_k_scan_partial(const uint32_t* __restrict__ in, size_t stride, const WnGrid* __restrict__ grid, int fixed_len,
                  int n_chunks, uint32_t* __restrict__ part_sum, uint32_t* __restrict__ part_max)
{
    const int f = blockIdx.y, c = blockIdx.x;
    const int n = frame_len(grid, fixed_len, f);
    const uint32_t* src = in + (size_t)f * stride;
    uint32_t s = 0, m = 0;
    const int base = c * SCAN_CHUNK;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int i = base + k * SCAN_THREADS + threadIdx.x;         /* coalesced */
        if (i < n) { const uint32_t v = src[i]; s += v; m = max(m, v); }
    }
    __shared__ uint32_t sh_s[SCAN_THREADS / 32], sh_m[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    if (lane == 0) { sh_s[warp] = s; sh_m[warp] = m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t ts = 0, tm = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) { ts += sh_s[w]; tm = max(tm, sh_m[w]); }
        part_sum[(size_t)f * n_chunks + c] = ts;
        if (part_max) part_max[(size_t)f * n_chunks + c] = tm;
    }
}

/* phase B: exclusive scan of the chunk sums of one frame (one warp per frame) */
__global__ void __launch_bounds__(32)
wn_k_scan_chunks(uint32_t* __restrict__ part_sum, const uint32_t* __restrict__ part_max, int n_chunks,
                 unsigned long long* __restrict__ frame_total, uint32_t* __restrict__ frame_max)
{
    const int f = blockIdx.x, lane = threadIdx.x;
    uint32_t* ps = part_sum + (size_t)f * n_chunks;
    unsigned long long run = 0;
    uint32_t mx = 0;
    for (int c0 = 0; c0 < n_chunks; c0 += 32) {
        const int c = c0 + lane;
        const uint32_t v = c < n_chunks ? ps[c] : 0u;
        if (part_max && c < n_chunks) mx = max(mx, part_max[(size_t)f * n_chunks + c]);
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (c < n_chunks) ps[c] = (uint32_t)run + inc - v;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) {
        if (frame_total) frame_total[f] = run;
        if (frame_max) frame_max[f] = mx;
    }
}

/* phase C: exclusive scan inside every chunk + chunk offset; optionally writes the total at [n] */
__global__ void __launch_bounds__(SCAN_THREADS)
wn_k_scan_final(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t stride,
                const WnGrid* __restrict__ grid, int fixed_len, int n_chunks,
                const uint32_t* __restrict__ part_sum, int write_total, uint32_t total_value, WnSubset sub)
{
    const int f = blockIdx.y, c = blockIdx.x;
    const int n = frame_len(grid, fixed_len, f);
    const uint32_t* src = in + (size_t)f * stride;
    uint32_t* dst = out + (size_t)f * stride;
    const int base = c * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;      /* thread-contiguous items */
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = (base + k < n) ? src[base + k] : 0u; s += v[k]; }
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += sh[w];
    uint32_t run = part_sum[(size_t)f * n_chunks + c] + woff + inc - s;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < n) dst[base + k] = run; run += v[k]; }
    if (write_total && c == 0 && threadIdx.x == 0) dst[n] = (uint32_t)wn_frame_sites(sub, f, (int)total_value);
}

/* ---- bounding box: ordered-integer atomics ---- */
__device__ __forceinline__ uint32_t f2key(float v)
{
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);               /* monotone increasing */
}
__device__ __forceinline__ float key2f(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

/* scratch[f][0..2] = min keys, [3..5] = ~max keys; identity 0xFFFFFFFF for both (memset 0xFF) */
__global__ void __launch_bounds__(256)
wn_k_bbox_partial(const float* __restrict__ frames, int natoms, const int* __restrict__ site_atom,
                  int n_table, WnSubset sub, uint32_t* __restrict__ scratch)
{
    const int f = blockIdx.y;
    const int n_sites = wn_frame_sites(sub, f, n_table);
    const float* fr = frames + (size_t)f * natoms * 3;
    float lo[3] = {INFINITY, INFINITY, INFINITY};
    float hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    const int base = blockIdx.x * 4096;
    for (int k = threadIdx.x; k < 4096; k += 256) {
        const int i = base + k;
        if (i >= n_sites) break;
        const float* p = fr + 3 * (size_t)__ldg(site_atom + wn_table_row(sub, f, i));
        const float x = p[0], y = p[1], z = p[2];
        if (isfinite(x) && isfinite(y) && isfinite(z)) {
            lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x);
            lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y);
            lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z);
        }
    }
    __shared__ float s_lo[3][8], s_hi[3][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float a = lo[c], b = hi[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
            b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
        }
        if (lane == 0) { s_lo[c][warp] = a; s_hi[c][warp] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        float a = s_lo[c][0], b = s_hi[c][0];
        for (int w = 1; w < 8; ++w) { a = fminf(a, s_lo[c][w]); b = fmaxf(b, s_hi[c][w]); }
        if (a <= b) {                                                  /* at least one finite site */
            atomicMin(scratch + (size_t)f * 6 + c, f2key(a));
            atomicMin(scratch + (size_t)f * 6 + 3 + c, ~f2key(b));
        }
    }
}

__global__ void __launch_bounds__(128)
wn_k_bbox_finalize(const uint32_t* __restrict__ scratch, int n_frames, int cell_cap, WnGrid* __restrict__ grid)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const uint32_t* s = scratch + (size_t)f * 6;
    WnGrid g;
    g.pbc = 0;
    for (int c = 0; c < 3; ++c) { g.L[c] = 0.0; g.invL[c] = 0.0; g.tri[c] = 0.0; }
    if (s[0] == 0xffffffffu) {                /* no finite site */
        g.ox = g.oy = g.oz = 0.0; g.nx = g.ny = g.nz = 1; g.ncells = 1;
        for (int c = 0; c < 3; ++c) { g.inv[c] = 1.0; g.edge[c] = 1.0f; }
        g.tf = 1.0f;
    } else {
        double mn[3], mx[3];
        for (int c = 0; c < 3; ++c) { mn[c] = (double)key2f(s[c]); mx[c] = (double)key2f(~s[3 + c]); }
        double edge = WN_CELL_EDGE;
        const double ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
        double nx, ny, nz;
        for (;;) {
            nx = floor(ex / edge) + 1.0;
            ny = floor(ey / edge) + 1.0;
            nz = floor(ez / edge) + 1.0;
            if (nx * ny * nz <= (double)cell_cap) break;
            edge *= 1.1;                      /* coarser cells stay correct (edge >= cutoff) */
        }
        g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2];
        for (int c = 0; c < 3; ++c) { g.inv[c] = 1.0 / edge; g.edge[c] = (float)edge; }
        g.nx = (int)nx; g.ny = (int)ny; g.nz = (int)nz;
        g.ncells = g.nx * g.ny * g.nz;
        g.tf = __double2float_ru(WN_PREFILTER_BOUND(edge));
    }
    grid[f] = g;
}

}  // namespace

int wn_scan_chunks_for(int len) { return (len + SCAN_CHUNK - 1) / SCAN_CHUNK; }

void wn_launch_bbox(const float* frames, int n_frames, int natoms, const int* site_atom,
                    int n_sites, WnSubset sub, int cell_cap, WnGrid* grid, uint32_t* scratch, cudaStream_t st)
{
    cudaMemsetAsync(scratch, 0xff, (size_t)n_frames * 6 * sizeof(uint32_t), st);
    dim3 g((n_sites + 4095) / 4096, n_frames);
    wn_k_bbox_partial<<<g, 256, 0, st>>>(frames, natoms, site_atom, n_sites, sub, scratch);
    wn_k_bbox_finalize<<<(n_frames + 127) / 128, 128, 0, st>>>(scratch, n_frames, cell_cap, grid);
}

/* cell counts -> cell starts, in place; start[ncells] = n_sites */
void wn_launch_cellscan(int n_frames, int n_sites, WnSubset sub, int cell_cap, const WnGrid* grid, int* cell_cnt,
                        uint32_t* part, cudaStream_t st)
{
    const int nch = wn_scan_chunks_for(cell_cap);
    uint32_t* data = reinterpret_cast<uint32_t*>(cell_cnt);
    dim3 g(nch, n_frames);
    wn_k_scan_partial<<<g, SCAN_THREADS, 0, st>>>(data, (size_t)cell_cap + 1, grid, 0, nch, part, nullptr);
    wn_k_scan_chunks<<<n_frames, 32, 0, st>>>(part, nullptr, nch, nullptr, nullptr);
    wn_k_scan_final<<<g, SCAN_THREADS, 0, st>>>(data, data, (size_t)cell_cap + 1, grid, 0, nch, part, 1, (uint32_t)n_sites, sub);
}

/* row lengths -> row offsets (+ per-frame total and longest row) */
void wn_launch_rowscan(const uint32_t* row_cnt, int n_frames, int n_sites, uint32_t* row_off,
                       unsigned long long* frame_total, uint32_t* frame_maxrow, uint32_t* part,
                       cudaStream_t st)
{
    const int nch = wn_scan_chunks_for(n_sites);
    uint32_t* part_max = part + (size_t)n_frames * nch;
    dim3 g(nch, n_frames);
    wn_k_scan_partial<<<g, SCAN_THREADS, 0, st>>>(row_cnt, (size_t)n_sites, nullptr, n_sites, nch, part, part_max);
    wn_k_scan_chunks<<<n_frames, 32, 0, st>>>(part, part_max, nch, frame_total, frame_maxrow);
    wn_k_scan_final<<<g, SCAN_THREADS, 0, st>>>(row_cnt, row_off, (size_t)n_sites, nullptr, n_sites, nch, part, 0, 0u, WnSubset{nullptr, nullptr});
}
