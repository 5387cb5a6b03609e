/*
 * wn_capi.cu -- the C ABI of include/wn_b200.h: context, device memory, the batched
 * pipeline driver and the NCCL statistics reduction.  No kernel lives here.
 *
 * Pipeline of one pass (F frames resident per launch), mirroring the stages of
 * WaterNetwork::analyzeFrame (reference src/waternetwork.cpp:503-667):
 *   convert+gather (:516-523,:538-577)  -> wn_k_bbox, wn_k_bin, wn_k_cellscan, wn_k_scatter
 *   find + make_edges (:583,:590-633)   -> wn_k_edges
 *   edge order, energy, stats (:654-666)-> wn_k_rowscan, wn_k_frameoff, wn_k_copy, wn_k_stats
 */
#include "wn_internal.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>      /* header-only: ranges cost nothing unless a profiler injects a collector */
#include <sched.h>
#include <ctype.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

static thread_local std::string g_create_error;

struct wn_ctx {
    int device = 0;
    cudaStream_t st = nullptr;
    std::string err;

    /* site table */
    int n_sites = -1, hs = 1, first_limit = 0, simple = 0, max_atom = -1;
    int* d_site_atom = nullptr;
    int* d_hv_atom = nullptr;
    uint32_t* d_meta = nullptr;
    WnEnergyParams ep;

    /* per-pass buffers */
    int user_frames_per_pass = 0;
    int cap_frames = 0, cap_sites = 0, cap_hs = 0, cell_cap = 0;
    WnGrid* d_grid = nullptr;
    int* d_cell_cnt = nullptr;
    int* d_site_cell = nullptr;
    int* d_site_rank = nullptr;
    float4* d_spos = nullptr;
    uint32_t* d_smeta = nullptr;
    double* d_sdh = nullptr;
    uint32_t* d_row_off = nullptr;          /* call-wide [n_frames][n_sites] (CSR row offsets) */
    size_t row_off_cap = 0;
    int* d_blk_prefix = nullptr;             /* home-block prefix per frame [F+1], the work counter, division constants [F][2] */
    uint32_t* d_scan_part = nullptr;     /* chunk sums / maxima of the three-phase scans, bbox keys */
    /* Buffers a pass's front (staging + fused search + row scan, stream st) hands to its back (ordered copy +
     * statistics, stream st_fin).  Two sets: the back of pass k runs under the front of pass k+1. */
    struct PassSet {
        uint32_t* row_cnt = nullptr;
        unsigned long long* frame_total = nullptr;
        WnFrameCounters* counters = nullptr;
        uint32_t* frame_maxrow = nullptr;
        double* block_sums = nullptr;
        uint32_t* rowbuf = nullptr;
        size_t rowbuf_cap = 0;          /* 32-bit words */
        int* sub_off = nullptr;  size_t sub_off_cap = 0;        /* per-frame site lists of the pass (device) */
        int* sub_sites = nullptr;  size_t sub_sites_cap = 0;
        WnGrid* h_grid = nullptr;  size_t h_grid_cap = 0;        /* pinned: host-built periodic grids of the pass */
        cudaEvent_t back_done = nullptr;                         /* the back that last read this set has finished */
        bool used = false;
    } pset[2];
    cudaStream_t st_fin = nullptr;
    unsigned long long* d_chain = nullptr;    /* [1] edges of the call so far (advanced by wn_k_frameoff) */
    unsigned long long* h_pass = nullptr;     /* pinned [2 * passes]: {edge total after the pass, longest row} */
    size_t h_pass_cap = 0;
    std::vector<cudaEvent_t> ev_pass;         /* 6 timing events per pass */
    int row_cap = 32;               /* slots per row; grows (sticky) when a row overflows */
    float* d_stage[2] = {nullptr, nullptr};   /* host variant: frames of one pass, ping-pong */
    size_t stage_cap[2] = {0, 0};
    cudaStream_t st_h2d = nullptr, st_d2h = nullptr;   /* copy streams: PCIe both ways under the kernels */
    cudaEvent_t ev_stage[2] = {};
    std::vector<cudaEvent_t> ev_copy;         /* per-pass copy timing events (4 per pass) */

    /* call-wide outputs (device) */
    wn_edge* d_edges = nullptr;
    size_t edges_cap = 0;
    unsigned long long* d_frame_off = nullptr;
    wn_frame_stats* d_stats = nullptr;
    WnFrameCounters* d_counters_all = nullptr;
    size_t out_frames_cap = 0;

    /* call-wide outputs (pinned host).  Two sets: with wn_set_result_sets(ctx, 2) consecutive batch calls
     * alternate between them, so the result of call k stays valid while call k+1 runs (a caller can hand
     * batch k to its sink while the GPU works on batch k+1: FrameBatcher does). */
    struct HostSet {
        wn_edge* edges = nullptr;
        size_t edges_cap = 0;
        uint64_t* frame_off = nullptr;
        wn_frame_stats* stats = nullptr;
        WnFrameCounters* counters = nullptr;
        size_t frames_cap = 0;
        uint32_t* row_off = nullptr;
        size_t row_off_cap = 0;
    } hset[2];
    int n_sets = 1, cur = 0;
    float* d_eside = nullptr;                  /* WN_LAYOUT_CSR12: energies of the pass in final order (for the sum) */
    size_t eside_cap = 0;
    unsigned long long* h_scalars = nullptr;   /* pinned scratch for small D2H reads */

    /* text formatting of the last batch */
    int last_frames = -1;                      /* frames of the last wn_hbond_batch* call, -1: none */
    int last_on_device = 0;                    /* that call's results live on the device (*_dev) */
    int last_subset = 0;                       /* that call used per-frame site lists: edge indices are list positions */
    /* connected components of the last batch */
    int* d_cc_labels = nullptr; int* d_cc_work = nullptr; int* d_cc_count = nullptr; int* d_cc_minidx = nullptr;
    unsigned int* d_cc_kept = nullptr;
    wn_component_stats* d_cc_stats = nullptr;
    size_t cc_labels_cap = 0, cc_work_cap = 0, cc_count_cap = 0, cc_minidx_cap = 0, cc_kept_cap = 0, cc_stats_cap = 0;
    int* h_cc_labels = nullptr; wn_component_stats* h_cc_stats = nullptr;
    size_t h_cc_labels_cap = 0, h_cc_stats_cap = 0;
    unsigned long long last_edges = 0;
    char* d_text = nullptr;  size_t d_text_cap = 0;
    char* h_text = nullptr;  size_t h_text_cap = 0;
    unsigned long long* d_row_base = nullptr;  size_t row_base_cap = 0;
    unsigned int* d_site_ids = nullptr;  size_t site_ids_cap = 0;
    int* d_frame_flags = nullptr;  size_t frame_flags_cap = 0;

    /* pair path */
    double* d_xyz = nullptr;
    size_t xyz_cap = 0;
    uint32_t* d_prow_cnt = nullptr;
    unsigned long long* d_prow_off = nullptr;
    size_t prow_cap = 0;
    int32_t* d_pairs = nullptr;
    size_t d_pairs_cap = 0;
    int32_t* h_pairs = nullptr;
    size_t h_pairs_cap = 0;
    unsigned long long* d_pair_ties = nullptr;
    /* make_edges over a caller-supplied pair list */
    int2* d_pl_pairs = nullptr;  size_t pl_pairs_cap = 0;
    float2* d_pl_tmp = nullptr;  size_t pl_tmp_cap = 0;
    uint32_t* d_pl_blk = nullptr;  size_t pl_blk_cap = 0;
    unsigned long long* d_pl_off = nullptr;  size_t pl_off_cap = 0;
    WnFrameCounters* d_pl_counters = nullptr;   /* [1] counters, then one word: bad index flag */
    /* cell-list pair path */
    int find_mode = 0;                       /* WN_FIND_AUTO / WN_FIND_ALLPAIRS / WN_FIND_CELLS */
    int last_find_cells = 0;                 /* the last wn_find_pairs* call ran the cell-list sweep */
    int* d_cp_cell = nullptr;    size_t cp_cell_cap = 0;
    int* d_cp_site_cell = nullptr; int* d_cp_site_rank = nullptr; double4* d_cp_srec = nullptr; float4* d_cp_srecf = nullptr;
    size_t cp_site_cap = 0;

    int layout = 0;                          /* result layout of the current / last call */
    bool keep_dev = false;                   /* current call: host frames in, results stay on the device */
    int box_stride = 3;                      /* 3: (Lx,Ly,Lz) per frame; 9: triclinic matrices */
    const int64_t* sub_off_host = nullptr;   /* per-frame site subsets of the current call (host) */
    const int32_t* sub_sites_host = nullptr;
    std::vector<int> h_sub_off;

    /* timing */
    cudaEvent_t ev[8] = {};
    wn_timings tm = {};

    /* NCCL (loaded lazily) */
    void* nccl_lib = nullptr;
    ncclComm_t comm = nullptr;
    int nranks = 0, rank = 0;
    double* d_reduce = nullptr;
    size_t reduce_cap = 0;
};

namespace {

/* NVTX range over a scope (stage markers for nsys / ncu --nvtx) */
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

int fail(wn_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

int cuda_fail(wn_ctx* c, cudaError_t e, const char* what)
{
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    cudaGetLastError();
    return fail(c, e == cudaErrorMemoryAllocation ? WN_ERR_OOM : WN_ERR_CUDA, buf);
}

#define WN_CK(call)                                                     \
    do {                                                                \
        cudaError_t e_ = (call);                                        \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);        \
    } while (0)

template <typename T>
int dev_ensure(wn_ctx* ctx, T** p, size_t* cap, size_t need, bool preserve = false, size_t used = 0)
{
    if (need <= *cap && *p) return WN_OK;
    size_t ncap = std::max(need, *cap + *cap / 2);
    if (ncap == 0) ncap = 1;
    T* np = nullptr;
    WN_CK(cudaMalloc((void**)&np, ncap * sizeof(T)));
    if (*p) WN_CK(cudaDeviceSynchronize());                             /* rare: kernels and copies of other streams may be in flight */
    if (preserve && *p && used)
        WN_CK(cudaMemcpyAsync(np, *p, used * sizeof(T), cudaMemcpyDeviceToDevice, ctx->st));
    if (*p) { WN_CK(cudaStreamSynchronize(ctx->st)); WN_CK(cudaFree(*p)); }
    *p = np;
    *cap = ncap;
    return WN_OK;
}

template <typename T>
int dev_realloc_plain(wn_ctx* ctx, T** p, size_t count)
{
    if (*p) { WN_CK(cudaFree(*p)); *p = nullptr; }
    WN_CK(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
    return WN_OK;
}

template <typename T>
int host_ensure(wn_ctx* ctx, T** p, size_t* cap, size_t need, bool preserve = false, size_t used = 0)
{
    if (need <= *cap && *p) return WN_OK;
    size_t ncap = std::max(need, *cap + *cap / 2);
    if (ncap == 0) ncap = 1;
    T* np = nullptr;
    WN_CK(cudaMallocHost((void**)&np, ncap * sizeof(T)));
    if (*p) WN_CK(cudaDeviceSynchronize());               /* D2H into the old block may be in flight */
    if (preserve && *p && used) memcpy(np, *p, used * sizeof(T));
    if (*p) WN_CK(cudaFreeHost(*p));
    *p = np;
    *cap = ncap;
    return WN_OK;
}

/* pow(pow(r_off,2)-pow(r_on,2),3) of switch_function_radius (src/waternetwork.cpp:20),
 * with r_on/r_off the squared radii computeEnergy passes (:57) */
void set_radial_denominator(WnEnergyParams& ep)
{
    const double d = ep.off2 * ep.off2 - ep.on2 * ep.on2;
    ep.inv_denom_r = 1.0 / (d * d * d);
}

int default_frames_per_pass(int n_sites)
{
    const long long target = 8ll << 20;    /* site-frames per pass */
    long long f = target / std::max(n_sites, 1);
    return (int)std::min<long long>(std::max<long long>(f, 1), 8192);
}

/* (re)allocate the per-pass buffers for F frames of the current site table */
int ensure_pass_buffers(wn_ctx* ctx, int F)
{
    const int N = ctx->n_sites;
    if (F <= ctx->cap_frames && N <= ctx->cap_sites && ctx->hs <= ctx->cap_hs && ctx->d_grid) return WN_OK;
    WN_CK(cudaDeviceSynchronize());
    const int cf = std::max(F, ctx->cap_frames), cn = std::max(N, ctx->cap_sites), ch = std::max(ctx->hs, ctx->cap_hs);
    const int cell_cap = cn / 4 + 27;
    const size_t fn = (size_t)cf * cn;
    int rc;
    if ((rc = dev_realloc_plain(ctx, &ctx->d_grid, (size_t)cf))) return rc;
    if ((rc = dev_realloc_plain(ctx, &ctx->d_cell_cnt, (size_t)cf * (cell_cap + 1)))) return rc;
    if ((rc = dev_realloc_plain(ctx, &ctx->d_site_cell, fn))) return rc;
    if ((rc = dev_realloc_plain(ctx, &ctx->d_site_rank, fn))) return rc;
    if ((rc = dev_realloc_plain(ctx, &ctx->d_spos, fn))) return rc;
    if ((rc = dev_realloc_plain(ctx, &ctx->d_smeta, fn))) return rc;
    if ((rc = dev_realloc_plain(ctx, &ctx->d_sdh, fn * ch * 4))) return rc;
    if ((rc = dev_realloc_plain(ctx, &ctx->d_blk_prefix, 3 * (size_t)cf + 4))) return rc;     /* prefix[cf+1], counter, magic[cf] */
    if ((rc = dev_realloc_plain(ctx, &ctx->d_scan_part,
                                (size_t)cf * (2 * (size_t)wn_scan_chunks_for(std::max(cn, cell_cap + 1)) + 8)))) return rc;
    for (wn_ctx::PassSet& S : ctx->pset) {
        if ((rc = dev_realloc_plain(ctx, &S.row_cnt, fn))) return rc;
        if ((rc = dev_realloc_plain(ctx, &S.frame_total, (size_t)cf))) return rc;
        if ((rc = dev_realloc_plain(ctx, &S.counters, (size_t)cf))) return rc;
        if ((rc = dev_realloc_plain(ctx, &S.frame_maxrow, (size_t)cf))) return rc;
        if ((rc = dev_realloc_plain(ctx, &S.block_sums, (size_t)cf * ((cn + 255) / 256 + 1) * 8))) return rc;
    }
    ctx->cap_frames = cf; ctx->cap_sites = cn; ctx->cap_hs = ch; ctx->cell_cap = cell_cap;
    return WN_OK;
}

int ensure_out_frames(wn_ctx* ctx, size_t n_frames, bool host)
{
    {
        const size_t need = n_frames * (size_t)std::max(ctx->n_sites, 0) + 1;
        int rc0;
        if ((rc0 = dev_ensure(ctx, &ctx->d_row_off, &ctx->row_off_cap, need))) return rc0;
        if (host && (rc0 = host_ensure(ctx, &ctx->hset[ctx->cur].row_off, &ctx->hset[ctx->cur].row_off_cap, need))) return rc0;
    }
    if (n_frames + 1 > ctx->out_frames_cap || !ctx->d_frame_off) {
        const size_t cap = n_frames + 1;
        int rc;
        WN_CK(cudaStreamSynchronize(ctx->st));
        if ((rc = dev_realloc_plain(ctx, &ctx->d_frame_off, cap))) return rc;
        if ((rc = dev_realloc_plain(ctx, &ctx->d_stats, cap))) return rc;
        if ((rc = dev_realloc_plain(ctx, &ctx->d_counters_all, cap))) return rc;
        ctx->out_frames_cap = cap;
    }
    /* the small per-frame tables (offsets, statistics, counters) come back to the host for every batch */
    if (n_frames + 1 > ctx->hset[ctx->cur].frames_cap || !ctx->hset[ctx->cur].frame_off) {
        const size_t cap = n_frames + 1;
        if (ctx->hset[ctx->cur].frame_off) WN_CK(cudaFreeHost(ctx->hset[ctx->cur].frame_off));
        if (ctx->hset[ctx->cur].stats) WN_CK(cudaFreeHost(ctx->hset[ctx->cur].stats));
        if (ctx->hset[ctx->cur].counters) WN_CK(cudaFreeHost(ctx->hset[ctx->cur].counters));
        ctx->hset[ctx->cur].frame_off = nullptr; ctx->hset[ctx->cur].stats = nullptr; ctx->hset[ctx->cur].counters = nullptr;
        WN_CK(cudaMallocHost((void**)&ctx->hset[ctx->cur].frame_off, cap * sizeof(uint64_t)));
        WN_CK(cudaMallocHost((void**)&ctx->hset[ctx->cur].stats, cap * sizeof(wn_frame_stats)));
        WN_CK(cudaMallocHost((void**)&ctx->hset[ctx->cur].counters, cap * sizeof(WnFrameCounters)));
        ctx->hset[ctx->cur].frames_cap = cap;
    }
    return WN_OK;
}

/* Fronts queued ahead of the pass whose back is next.  Measured on the 100k-atom box (profiles/r02/overlap_sweep/):
 * with one front ahead the ordered copy of pass k shares the SMs with the search of pass k+1, and every split of a
 * 250-frame step into 2-10 overlapped passes was SLOWER than one pass (4.26-5.11 ms against 4.10 ms per step, for
 * any stream priority and with the search limited to 4 CTAs per SM): the persistent search kernel owns the register
 * file, the copy only gets the gaps, and both stages pay their launch tails once per pass.  So passes run back to
 * back; what the two streams still buy is that the D2H copy of a pass (host variant) starts the moment its back ends. */
#ifndef WN_PASS_LOOKAHEAD
#define WN_PASS_LOOKAHEAD 0
#endif

/* Device memory the row buffers of the two pass sets may take together (bytes). */
static const size_t WN_ROWBUF_BUDGET = (size_t)24 << 30;

struct PassPlan { int f0, nf; };

/* frames of [f, n_frames) -> passes of at most `unit` frames (and `nf_cap`, the row-buffer budget); `ramp`: the
 * first passes are short (unit/8, doubling) so that the first results reach the D2H engine early */
void build_plan(int f, int n_frames, int unit, int nf_cap, bool ramp, std::vector<PassPlan>& plan)
{
    unit = std::max(1, std::min(unit, nf_cap));
    if (ramp) {
        for (int nf = std::max(1, unit / 8); f < n_frames && nf < unit; nf *= 2) {
            const int take = std::min(nf, n_frames - f);
            plan.push_back(PassPlan{f, take});
            f += take;
        }
    }
    const int rest = n_frames - f;
    if (rest <= 0) return;
    const int np = (rest + unit - 1) / unit, each = rest / np, extra = rest % np;      /* equal passes: no short tail */
    for (int q = 0; q < np; ++q) {
        const int take = each + (q < extra ? 1 : 0);
        plan.push_back(PassPlan{f, take});
        f += take;
    }
}

/* FRONT of a pass, on ctx->st: staging (bbox, bin, cell scan, scatter), the fused search + geometry kernel, the
 * row scan and the frame offsets (chained on the device) for `nf` device-resident frames -> row buffers of set S.
 * ev[0..3]: start / staged / searched / front done. */
int pass_front(wn_ctx* ctx, wn_ctx::PassSet& S, const float* d_frames, const float* boxes, int nf, int natoms, size_t f0,
               unsigned long long* pass_info, cudaEvent_t* ev)
{
    const int N = ctx->n_sites;
    cudaStream_t st = ctx->st;
    int rc;
    const int cell_cap = ctx->cell_cap;
    /* the back of the pass that used this set two passes ago must have read everything */
    if (S.used) WN_CK(cudaStreamWaitEvent(st, S.back_done, 0));
    S.used = true;
    if ((rc = dev_ensure(ctx, &S.rowbuf, &S.rowbuf_cap, (size_t)nf * N * ctx->row_cap * WN_ROW_WORDS))) return rc;
    WnSubset sub = {nullptr, nullptr};
    if (ctx->sub_off_host) {
        /* the pass's slice of the per-frame site lists, offsets rebased to the slice */
        const int64_t base = ctx->sub_off_host[f0];
        const size_t cnt = (size_t)(ctx->sub_off_host[f0 + nf] - base);
        ctx->h_sub_off.resize((size_t)nf + 1);
        for (int k = 0; k <= nf; ++k) ctx->h_sub_off[k] = (int)(ctx->sub_off_host[f0 + k] - base);
        if ((rc = dev_ensure(ctx, &S.sub_off, &S.sub_off_cap, (size_t)nf + 1))) return rc;
        if ((rc = dev_ensure(ctx, &S.sub_sites, &S.sub_sites_cap, cnt + 1))) return rc;
        /* pageable sources: the calls return once the driver has staged them */
        WN_CK(cudaMemcpyAsync(S.sub_off, ctx->h_sub_off.data(), ((size_t)nf + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        if (cnt) WN_CK(cudaMemcpyAsync(S.sub_sites, ctx->sub_sites_host + base, cnt * sizeof(int), cudaMemcpyHostToDevice, st));
        WN_CK(cudaStreamSynchronize(st));
        sub.off = S.sub_off; sub.sites = S.sub_sites;
    }

    WN_CK(cudaEventRecord(ev[0], st));
    nvtxRangePushA("wn:bin (bbox, cells, sort)");
    struct PopOnExit { int n; ~PopOnExit() { while (n-- > 0) nvtxRangePop(); } } nvtx_guard{1};
    bool small_box = false;      /* periodic box with < 3 cells in some dimension: all-pairs kernel */
    int bx = 2;                  /* home block width; 1 when a periodic box has < 4 cells along x  */
    if (boxes) {
        if ((size_t)nf > S.h_grid_cap) {
            if (S.h_grid) { WN_CK(cudaDeviceSynchronize()); WN_CK(cudaFreeHost(S.h_grid)); S.h_grid = nullptr; }
            S.h_grid_cap = (size_t)nf + (size_t)nf / 2;
            WN_CK(cudaMallocHost((void**)&S.h_grid, S.h_grid_cap * sizeof(WnGrid)));
        }
        for (int f = 0; f < nf; ++f) {
            WnGrid& g = S.h_grid[f];
            if (ctx->box_stride == 9) {
                const float* b = boxes + 9 * (size_t)f;
                for (int k = 0; k < 9; ++k)
                    if (!std::isfinite(b[k])) return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch_box9: box must be finite");
                if (!(b[0] > 0.f && b[4] > 0.f && b[8] > 0.f) || b[1] != 0.f || b[2] != 0.f || b[5] != 0.f)
                    return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch_box9: box must be lower-triangular with a positive diagonal (GROMACS convention)");
                wn_host_tric_grid(b, cell_cap, &g);
                /* the shifted neighbour ranges span up to BX+3 cells in x and 4 rows in y */
                if (g.nx < 4 || g.ny < 4 || g.nz < 3) small_box = true;
                if (g.nx < 5) bx = 1;
                continue;
            }
            const float* b = boxes + 3 * (size_t)f;
            if (!(b[0] > 0.f && b[1] > 0.f && b[2] > 0.f) || !std::isfinite(b[0]) || !std::isfinite(b[1]) || !std::isfinite(b[2]))
                return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch_pbc: box lengths must be positive and finite");
            wn_host_pbc_grid(b, cell_cap, &g);
            if (g.nx < 3 || g.ny < 3 || g.nz < 3) small_box = true;
            if (g.nx < 4) bx = 1;
        }
        /* pinned and owned by the set: the copy is asynchronous, the set is not reused before its back has finished */
        WN_CK(cudaMemcpyAsync(ctx->d_grid, S.h_grid, (size_t)nf * sizeof(WnGrid), cudaMemcpyHostToDevice, st));
    } else {
        wn_launch_bbox(d_frames, nf, natoms, ctx->d_site_atom, N, sub, cell_cap, ctx->d_grid, ctx->d_scan_part, st);
    }
    WN_CK(cudaMemsetAsync(ctx->d_cell_cnt, 0, (size_t)nf * (cell_cap + 1) * sizeof(int), st));
    wn_launch_bin(d_frames, nf, natoms, ctx->d_site_atom, N, sub, cell_cap, ctx->d_grid, ctx->d_cell_cnt,
                  ctx->d_site_cell, ctx->d_site_rank, st);
    wn_launch_cellscan(nf, N, sub, cell_cap, ctx->d_grid, ctx->d_cell_cnt, ctx->d_scan_part, st);
    wn_launch_scatter(d_frames, nf, natoms, ctx->d_site_atom, ctx->d_hv_atom, ctx->d_meta, N, sub, ctx->hs,
                      cell_cap, ctx->d_cell_cnt, ctx->d_site_cell, ctx->d_site_rank, ctx->d_grid, ctx->d_spos,
                      ctx->d_smeta, ctx->d_sdh, st);
    ctx->tm.launches += boxes ? 5 : 7;     /* [bbox x2] bin, cell scan x3, scatter */
    WN_CK(cudaEventRecord(ev[1], st));
    nvtxRangePop(); nvtxRangePushA("wn:edges (fused search + geometry)");

    WN_CK(cudaMemsetAsync(S.row_cnt, 0, (size_t)nf * N * sizeof(uint32_t), st));
    WN_CK(cudaMemsetAsync(S.counters, 0, (size_t)nf * sizeof(WnFrameCounters), st));
    WnEdgeParams p;
    p.spos = ctx->d_spos; p.smeta = ctx->d_smeta; p.sdh = ctx->d_sdh;
    p.cell_start = ctx->d_cell_cnt; p.grid = ctx->d_grid;
    p.rowbuf = S.rowbuf; p.row_cnt = S.row_cnt; p.counters = S.counters;
    p.n_sites = N; p.cell_cap = cell_cap; p.hs = ctx->hs; p.first_limit = ctx->first_limit;
    p.row_cap = ctx->row_cap; p.simple = ctx->simple; p.pbc = boxes ? (ctx->box_stride == 9 ? 2 : 1) : 0; p.bx = bx; p.subset = sub;
    p.blk_prefix = ctx->d_blk_prefix; p.work_next = reinterpret_cast<unsigned int*>(ctx->d_blk_prefix + nf + 1); p.n_frames = nf;
    p.blk_magic = reinterpret_cast<uint2*>(ctx->d_blk_prefix + ((nf + 2 + 1) & ~1));
    if (small_box) wn_launch_edges_small(p, nf, st);
    else wn_launch_edges(p, nf, st);
    WN_CK(cudaGetLastError());
    WN_CK(cudaEventRecord(ev[2], st));
    nvtxRangePop(); nvtxRangePushA("wn:finish (row scan, frame offsets)");
    wn_launch_rowscan(S.row_cnt, nf, N, ctx->d_row_off + f0 * (size_t)N, S.frame_total, S.frame_maxrow, ctx->d_scan_part, st);
    /* the pass scalars (edge total of the call so far, longest row) are written by the kernel straight into pinned
     * host memory (UVA: the pointer is valid on the device): a cudaMemcpyAsync here would queue on the D2H copy
     * engine behind an earlier pass's result copy */
    wn_launch_frameoff(S.frame_total, S.frame_maxrow, nf, ctx->d_chain, ctx->d_frame_off + f0, pass_info, st);
    ctx->tm.launches += small_box ? 5 : 6;   /* [home-block prefix] fused search, row scan x3, frame offsets */
    WN_CK(cudaGetLastError());
    WN_CK(cudaEventRecord(ev[3], st));
    return WN_OK;
}

/* BACK of a pass, on ctx->st_fin: rows ordered by j at their lexicographic offset + energies (wn_k_copy), the
 * statistics rows, the pass's counters.  ev[3]: front done (waited for); ev[4], ev[5]: back start / done. */
int pass_back(wn_ctx* ctx, wn_ctx::PassSet& S, int nf, size_t f0, cudaEvent_t* ev)
{
    const int N = ctx->n_sites;
    cudaStream_t st = ctx->st_fin;
    WnSubset sub = {nullptr, nullptr};
    if (ctx->sub_off_host) { sub.off = S.sub_off; sub.sites = S.sub_sites; }
    WN_CK(cudaStreamWaitEvent(st, ev[3], 0));
    WN_CK(cudaEventRecord(ev[4], st));
    NvtxRange nvtx_back("wn:finish (ordered copy, energies, statistics)");
    int n_blocks = 0;
    wn_launch_copy(S.rowbuf, ctx->row_cap, S.row_cnt, ctx->d_row_off + f0 * (size_t)N, ctx->d_frame_off + f0, S.counters,
                   nf, N, ctx->ep, ctx->d_edges, ctx->edges_cap, ctx->layout, ctx->d_eside, S.block_sums, &n_blocks, st);
    wn_launch_stats(S.block_sums, n_blocks, S.frame_total, S.counters, nf, N, sub, ctx->d_stats + f0, ctx->d_edges, ctx->layout,
                    ctx->d_eside, ctx->d_row_off + f0 * (size_t)N, ctx->d_frame_off + f0, st);
    WN_CK(cudaMemcpyAsync(ctx->d_counters_all + f0, S.counters, (size_t)nf * sizeof(WnFrameCounters),
                          cudaMemcpyDeviceToDevice, st));
    ctx->tm.launches += 2;
    WN_CK(cudaGetLastError());
    WN_CK(cudaEventRecord(ev[5], st));
    WN_CK(cudaEventRecord(S.back_done, st));
    return WN_OK;
}

int batch_common(wn_ctx* ctx, const float* frames, const float* boxes, bool host_in, int n_frames, int natoms,
                 wn_batch_result* out)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    if (!out || n_frames < 0 || natoms < 0 || (!frames && n_frames > 0 && natoms > 0))
        return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch: bad argument");
    if (ctx->n_sites < 0) return fail(ctx, WN_ERR_NO_SITES, "wn_hbond_batch: call wn_set_sites first");
    if (ctx->max_atom >= natoms && ctx->n_sites > 0)
        return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch: site table refers to atoms beyond natoms");
    WN_CK(cudaSetDevice(ctx->device));
    ctx->err.clear();
    ctx->tm = wn_timings{};
    memset(out, 0, sizeof *out);
    if (ctx->n_sets == 2) ctx->cur ^= 1;          /* the previous call's host results stay untouched */
    int rc;
    /* results go where the frames came from, unless the caller keeps them on the device (results_on_device) */
    const bool host_out = host_in && !ctx->keep_dev;
    if ((rc = ensure_out_frames(ctx, (size_t)n_frames, host_out))) return rc;

    const int N = ctx->n_sites;
    unsigned long long edge_base = 0;
    if (N == 0 || n_frames == 0) {
        /* nothing to search: empty edge lists, zero rows */
        for (int f = 0; f <= n_frames; ++f) ctx->hset[ctx->cur].frame_off[f] = 0;
        for (int f = 0; f < n_frames; ++f) ctx->hset[ctx->cur].stats[f] = wn_frame_stats{0, 0, 0, 0, 0};
        if (!host_out && n_frames > 0) {
            WN_CK(cudaMemsetAsync(ctx->d_frame_off, 0, (size_t)(n_frames + 1) * sizeof(unsigned long long), ctx->st));
            WN_CK(cudaMemsetAsync(ctx->d_stats, 0, (size_t)n_frames * sizeof(wn_frame_stats), ctx->st));
            WN_CK(cudaStreamSynchronize(ctx->st));
        }
        out->edges = host_out ? ctx->hset[ctx->cur].edges : ctx->d_edges;
        out->frame_offsets = host_out ? ctx->hset[ctx->cur].frame_off : (const uint64_t*)ctx->d_frame_off;
        out->stats = host_out ? ctx->hset[ctx->cur].stats : ctx->d_stats;
        out->n_frames = n_frames;
        out->on_device = host_out ? 0 : 1;
        out->row_offsets = host_out ? ctx->hset[ctx->cur].row_off : ctx->d_row_off;
        out->n_sites = N;
        out->layout = ctx->layout;
        return WN_OK;
    }

    /* ---- the call as a pipeline of passes ----
     * st     : front of pass k   (staging, fused search, row scan)          -> row buffers of set k & 1
     * st_fin : back of pass k-1  (ordered copy + energies, statistics)      <- row buffers of set (k-1) & 1
     * st_h2d : frames of pass k+1 (host variant)      st_d2h : results of pass k-2 (host variant)
     * The host learns a pass's edge total and longest row from pinned memory after its front (one event wait, with
     * the next front already queued), sizes the edge array exactly, and queues the back.  A row longer than its
     * slots (dense input) drains the pipeline, doubles the (sticky) row capacity and repeats from that pass. */
    const int fpp = ctx->user_frames_per_pass > 0 ? ctx->user_frames_per_pass : default_frames_per_pass(N);
    const int unit = fpp;
    const size_t frame_floats = (size_t)natoms * 3;
    auto nf_cap_now = [&]() -> int {
        const size_t per_frame = (size_t)std::max(N, 1) * ctx->row_cap * WN_ROW_WORDS * sizeof(uint32_t);
        return (int)std::max<size_t>(1, std::min<size_t>((WN_ROWBUF_BUDGET / 2) / per_frame, (size_t)1 << 20));
    };
    std::vector<PassPlan> plan;
    build_plan(0, n_frames, unit, nf_cap_now(), host_in, plan);
    auto prepare = [&]() -> int {
        int rc2, max_nf = 0;
        for (const PassPlan& pp : plan) max_nf = std::max(max_nf, pp.nf);
        if ((rc2 = ensure_pass_buffers(ctx, max_nf))) return rc2;
        while (ctx->ev_pass.size() < 6 * plan.size()) { cudaEvent_t e; WN_CK(cudaEventCreate(&e)); ctx->ev_pass.push_back(e); }
        while (ctx->ev_copy.size() < 4 * plan.size()) { cudaEvent_t e; WN_CK(cudaEventCreate(&e)); ctx->ev_copy.push_back(e); }
        if ((rc2 = host_ensure(ctx, &ctx->h_pass, &ctx->h_pass_cap, 2 * plan.size()))) return rc2;
        if (host_in)
            for (int b = 0; b < 2; ++b)
                if ((rc2 = dev_ensure(ctx, &ctx->d_stage[b], &ctx->stage_cap[b], (size_t)max_nf * frame_floats))) return rc2;
        return WN_OK;
    };
    if ((rc = prepare())) return rc;
    WN_CK(cudaMemsetAsync(ctx->d_chain, 0, sizeof(unsigned long long), ctx->st));
    for (wn_ctx::PassSet& S : ctx->pset) S.used = false;      /* every earlier call ended drained */

    auto stage_in = [&](int k) -> int {
        const int f0 = plan[k].f0, nf = plan[k].nf, b = k & 1;
        NvtxRange nvtx_h2d("wn:h2d (frames of a pass)");
        /* the staging buffer was read by the front of pass k-2 */
        if (k >= 2) WN_CK(cudaStreamWaitEvent(ctx->st_h2d, ctx->ev_pass[6 * (k - 2) + 3], 0));
        WN_CK(cudaEventRecord(ctx->ev_copy[4 * k], ctx->st_h2d));
        WN_CK(cudaMemcpyAsync(ctx->d_stage[b], frames + (size_t)f0 * frame_floats,
                              (size_t)nf * frame_floats * sizeof(float), cudaMemcpyHostToDevice, ctx->st_h2d));
        WN_CK(cudaEventRecord(ctx->ev_copy[4 * k + 1], ctx->st_h2d));
        WN_CK(cudaEventRecord(ctx->ev_stage[b], ctx->st_h2d));
        return WN_OK;
    };
    int enq = 0, staged = -1, k = 0, grown_at = -1;
    while (k < (int)plan.size()) {
        const int n_pass = (int)plan.size();
        while (enq < n_pass && enq <= k + WN_PASS_LOOKAHEAD) {
            const int f0 = plan[enq].f0, nf = plan[enq].nf;
            const float* src = frames + (size_t)f0 * frame_floats;
            if (host_in) {
                if (staged < enq) { if ((rc = stage_in(enq))) return rc; staged = enq; }
                WN_CK(cudaStreamWaitEvent(ctx->st, ctx->ev_stage[enq & 1], 0));
                src = ctx->d_stage[enq & 1];
            }
            if ((rc = pass_front(ctx, ctx->pset[enq & 1], src, boxes ? boxes + ctx->box_stride * (size_t)f0 : nullptr, nf, natoms,
                                 (size_t)f0, ctx->h_pass + 2 * (size_t)enq, &ctx->ev_pass[6 * (size_t)enq]))) return rc;
            ++enq;
            if (host_in && enq < n_pass && staged < enq) { if ((rc = stage_in(enq))) return rc; staged = enq; }
        }
        const int f0 = plan[k].f0, nf = plan[k].nf;
        WN_CK(cudaEventSynchronize(ctx->ev_pass[6 * (size_t)k + 3]));
        const unsigned long long end_off = ctx->h_pass[2 * (size_t)k], max_row = ctx->h_pass[2 * (size_t)k + 1];
        if (max_row > (unsigned long long)ctx->row_cap) {
            /* a row overflowed its slots: drain, grow the (sticky) row capacity, repeat from this pass */
            if (grown_at == f0) return fail(ctx, WN_ERR_CUDA, "internal: row capacity still too small after growing");
            grown_at = f0;
            WN_CK(cudaDeviceSynchronize());
            int nc = ctx->row_cap;
            while ((unsigned long long)nc < max_row) nc *= 2;
            ctx->row_cap = nc;
            plan.resize((size_t)k);
            build_plan(f0, n_frames, unit, nf_cap_now(), false, plan);
            if ((rc = prepare())) return rc;
            ctx->h_scalars[8] = edge_base;
            WN_CK(cudaMemcpyAsync(ctx->d_chain, ctx->h_scalars + 8, sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->st));
            for (wn_ctx::PassSet& S : ctx->pset) S.used = false;
            enq = k; staged = k - 1;
            continue;
        }
        /* grow with head room: edge counts of equally sized batches differ by a fraction of a percent, and a
         * regrowth (allocation + copy + synchronisation) costs tens of milliseconds */
        if ((size_t)end_off > ctx->edges_cap || !ctx->d_edges) {
            const double per_frame = (double)end_off / (double)(f0 + nf);
            const size_t want = std::max((size_t)end_off + (size_t)end_off / 16, (size_t)(per_frame * n_frames * 1.02)) + 4096;
            if ((rc = dev_ensure(ctx, &ctx->d_edges, &ctx->edges_cap, want, true, (size_t)edge_base))) return rc;
        }
        if (ctx->layout == WN_LAYOUT_CSR12) {
            if ((rc = dev_ensure(ctx, &ctx->d_eside, &ctx->eside_cap, ctx->edges_cap))) return rc;
        }
        if ((rc = pass_back(ctx, ctx->pset[k & 1], nf, (size_t)f0, &ctx->ev_pass[6 * (size_t)k]))) return rc;
        if (host_in) {
            const unsigned long long before = edge_base;
            if (host_out && (size_t)end_off > ctx->hset[ctx->cur].edges_cap) {
                /* project the whole call from the density seen so far */
                const double per_frame = (double)end_off / (double)(f0 + nf);
                const size_t want = (size_t)(per_frame * n_frames * 1.05) + 1024;
                if ((rc = host_ensure(ctx, &ctx->hset[ctx->cur].edges, &ctx->hset[ctx->cur].edges_cap, std::max(want, (size_t)end_off), true,
                                      (size_t)before))) return rc;
            }
            NvtxRange nvtx_d2h("wn:d2h (edges + row offsets of a pass)");
            WN_CK(cudaStreamWaitEvent(ctx->st_d2h, ctx->ev_pass[6 * (size_t)k + 5], 0));
            WN_CK(cudaEventRecord(ctx->ev_copy[4 * k + 2], ctx->st_d2h));
            if (host_out && end_off > before) {
                /* 20-byte wn_edge records, or 16-byte wn_edge16 records in the same buffers */
                const size_t esz = ctx->layout == WN_LAYOUT_CSR16 ? sizeof(wn_edge16)
                                 : ctx->layout == WN_LAYOUT_CSR12 ? sizeof(wn_edge12) : sizeof(wn_edge);
                WN_CK(cudaMemcpyAsync(reinterpret_cast<char*>(ctx->hset[ctx->cur].edges) + (size_t)before * esz,
                                      reinterpret_cast<const char*>(ctx->d_edges) + (size_t)before * esz,
                                      (size_t)(end_off - before) * esz, cudaMemcpyDeviceToHost, ctx->st_d2h));
            }
            /* the row offsets of the pass travel with its edges */
            if (host_out && N > 0)
                WN_CK(cudaMemcpyAsync(ctx->hset[ctx->cur].row_off + (size_t)f0 * N, ctx->d_row_off + (size_t)f0 * N,
                                      (size_t)nf * N * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->st_d2h));
            WN_CK(cudaEventRecord(ctx->ev_copy[4 * k + 3], ctx->st_d2h));
        }
        edge_base = end_off;
        ++k;
    }
    /* everything the backs wrote is visible to st before the small tables are read */
    for (wn_ctx::PassSet& S : ctx->pset)
        if (S.used) WN_CK(cudaStreamWaitEvent(ctx->st, S.back_done, 0));
    /* small per-frame tables always come back to the host: totals and ties */
    WN_CK(cudaMemcpyAsync(ctx->hset[ctx->cur].frame_off, ctx->d_frame_off, (size_t)(n_frames + 1) * sizeof(uint64_t),
                          cudaMemcpyDeviceToHost, ctx->st));
    WN_CK(cudaMemcpyAsync(ctx->hset[ctx->cur].stats, ctx->d_stats, (size_t)n_frames * sizeof(wn_frame_stats),
                          cudaMemcpyDeviceToHost, ctx->st));
    WN_CK(cudaMemcpyAsync(ctx->hset[ctx->cur].counters, ctx->d_counters_all, (size_t)n_frames * sizeof(WnFrameCounters),
                          cudaMemcpyDeviceToHost, ctx->st));
    WN_CK(cudaStreamSynchronize(ctx->st));
    if (host_in) WN_CK(cudaStreamSynchronize(ctx->st_d2h));
    {
        /* stage times: sums over the passes (fronts and backs of neighbouring passes overlap);
         * total: first front start -> last front/back end */
        const int n_pass = (int)plan.size();
        float ms = 0, span = 0;
        for (int q = 0; q < n_pass; ++q) {
            cudaEvent_t* e = &ctx->ev_pass[6 * (size_t)q];
            cudaEventElapsedTime(&ms, e[0], e[1]); ctx->tm.bin_ms += ms;
            cudaEventElapsedTime(&ms, e[1], e[2]); ctx->tm.edges_ms += ms;
            cudaEventElapsedTime(&ms, e[2], e[3]); ctx->tm.finish_ms += ms;
            cudaEventElapsedTime(&ms, e[4], e[5]); ctx->tm.finish_ms += ms;
            cudaEventElapsedTime(&ms, ctx->ev_pass[0], e[3]); span = std::max(span, ms);
            cudaEventElapsedTime(&ms, ctx->ev_pass[0], e[5]); span = std::max(span, ms);
            if (host_in) {
                cudaEventElapsedTime(&ms, ctx->ev_copy[4 * q], ctx->ev_copy[4 * q + 1]); ctx->tm.h2d_ms += ms;
                cudaEventElapsedTime(&ms, ctx->ev_copy[4 * q + 2], ctx->ev_copy[4 * q + 3]); ctx->tm.d2h_ms += ms;
            }
        }
        ctx->tm.total_ms = span;
        ctx->tm.passes = n_pass;
    }
    if (ctx->hset[ctx->cur].frame_off[n_frames] != edge_base)
        return fail(ctx, WN_ERR_CUDA, "internal: edge totals disagree between scan and fused kernel");
    wn_tie_report ties = {0, 0, 0, 0, 0};
    unsigned long long evals = 0, tests = 0;
    for (int f = 0; f < n_frames; ++f) {
        evals += ctx->hset[ctx->cur].counters[f].evals;
        tests += ctx->hset[ctx->cur].counters[f].tests;
        ties.length_equal_cut += ctx->hset[ctx->cur].counters[f].tie_len;
        ties.cosine_argmax_equal += ctx->hset[ctx->cur].counters[f].tie_cos;
        ties.pos_zero += ctx->hset[ctx->cur].counters[f].tie_pos0;
    }
    out->edges = host_out ? ctx->hset[ctx->cur].edges : ctx->d_edges;
    out->frame_offsets = host_out ? ctx->hset[ctx->cur].frame_off : (const uint64_t*)ctx->d_frame_off;
    out->stats = host_out ? ctx->hset[ctx->cur].stats : ctx->d_stats;
    ctx->last_frames = n_frames;
    ctx->last_edges = edge_base;
    ctx->last_on_device = host_out ? 0 : 1;
    ctx->last_subset = ctx->sub_off_host ? 1 : 0;
    out->row_offsets = host_out ? ctx->hset[ctx->cur].row_off : ctx->d_row_off;
    out->n_sites = N;
    out->layout = ctx->layout;
    out->n_edges = edge_base;
    out->n_pair_evals = evals;
    out->n_candidate_tests = tests;
    out->ties = ties;
    out->n_frames = n_frames;
    out->on_device = host_out ? 0 : 1;
    return WN_OK;
}

}  // namespace

extern "C" {

const char* wn_version(void) { return "wn_b200 0.2 sm_100a"; }

int wn_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

int wn_create(int device, wn_ctx** out)
{
    if (!out) return fail(nullptr, WN_ERR_BAD_ARG, "wn_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, WN_ERR_NO_DEVICE, std::string("wn_create: no CUDA device (") +
                    (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                    "); this library has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail(nullptr, WN_ERR_BAD_ARG, "wn_create: device index out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(nullptr, WN_ERR_NO_DEVICE, std::string("wn_create: kernels are built for sm_100a only; device is ") +
                    prop.name + " (sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")");
    wn_ctx* ctx = new (std::nothrow) wn_ctx();
    if (!ctx) return fail(nullptr, WN_ERR_OOM, "wn_create: out of host memory");
    ctx->device = device;
    ctx->ep.on2 = 5.5 * 5.5; ctx->ep.off2 = 6.5 * 6.5;
    ctx->ep.a_on = 0.0; ctx->ep.a_off = 0.25;      /* (theta_off, theta_on): caller-side swap, :59-60 */
    set_radial_denominator(ctx->ep);
#define WN_CKC(call)                                                            \
    do {                                                                        \
        cudaError_t e2_ = (call);                                               \
        if (e2_ != cudaSuccess) { int rc_ = cuda_fail(nullptr, e2_, #call); delete ctx; return rc_; } \
    } while (0)
    WN_CKC(cudaSetDevice(device));
    {
        /* the back of a pass (ordered copy) goes first: its result is what the D2H engine waits for */
        int prio_lo = 0, prio_hi = 0;
        WN_CKC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        WN_CKC(cudaStreamCreateWithPriority(&ctx->st, cudaStreamNonBlocking, prio_lo));
        WN_CKC(cudaStreamCreateWithPriority(&ctx->st_fin, cudaStreamNonBlocking, prio_hi));
    }
    for (wn_ctx::PassSet& S : ctx->pset) WN_CKC(cudaEventCreateWithFlags(&S.back_done, cudaEventDisableTiming));
    WN_CKC(cudaMalloc((void**)&ctx->d_chain, sizeof(unsigned long long)));
    WN_CKC(cudaStreamCreateWithFlags(&ctx->st_h2d, cudaStreamNonBlocking));
    WN_CKC(cudaStreamCreateWithFlags(&ctx->st_d2h, cudaStreamNonBlocking));
    for (auto& ev : ctx->ev_stage) WN_CKC(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : ctx->ev) WN_CKC(cudaEventCreate(&ev));
    WN_CKC(cudaMallocHost((void**)&ctx->h_scalars, 64 * sizeof(unsigned long long)));
    WN_CKC(cudaMalloc((void**)&ctx->d_pair_ties, 2 * sizeof(unsigned long long)));
#undef WN_CKC
    *out = ctx;
    return WN_OK;
}

void wn_destroy(wn_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->comm && ctx->nccl_lib) {
        typedef ncclResult_t (*destroy_t)(ncclComm_t);
        destroy_t fn = (destroy_t)dlsym(ctx->nccl_lib, "ncclCommDestroy");
        if (fn) fn(ctx->comm);
    }
    void* dev[] = {ctx->d_site_atom, ctx->d_hv_atom, ctx->d_meta, ctx->d_grid, ctx->d_cell_cnt, ctx->d_site_cell,
                   ctx->d_site_rank, ctx->d_spos, ctx->d_smeta, ctx->d_sdh, ctx->d_row_off, ctx->d_blk_prefix, ctx->d_scan_part, ctx->d_chain,
                   ctx->pset[0].row_cnt, ctx->pset[0].frame_total, ctx->pset[0].counters, ctx->pset[0].frame_maxrow, ctx->pset[0].block_sums,
                   ctx->pset[0].rowbuf, ctx->pset[0].sub_off, ctx->pset[0].sub_sites,
                   ctx->pset[1].row_cnt, ctx->pset[1].frame_total, ctx->pset[1].counters, ctx->pset[1].frame_maxrow, ctx->pset[1].block_sums,
                   ctx->pset[1].rowbuf, ctx->pset[1].sub_off, ctx->pset[1].sub_sites,
                   ctx->d_stage[0], ctx->d_stage[1], ctx->d_edges, ctx->d_frame_off, ctx->d_stats, ctx->d_counters_all,
                   ctx->d_xyz, ctx->d_prow_cnt, ctx->d_prow_off, ctx->d_pairs, ctx->d_pair_ties, ctx->d_reduce,
                   ctx->d_text, ctx->d_row_base, ctx->d_site_ids, ctx->d_frame_flags,
                   ctx->d_cc_labels, ctx->d_cc_work, ctx->d_cc_count, ctx->d_cc_minidx, ctx->d_cc_kept, ctx->d_cc_stats, ctx->d_eside,
                   ctx->d_pl_pairs, ctx->d_pl_tmp, ctx->d_pl_blk, ctx->d_pl_off, ctx->d_pl_counters,
                   ctx->d_cp_cell, ctx->d_cp_site_cell, ctx->d_cp_site_rank, ctx->d_cp_srec, ctx->d_cp_srecf};
    for (void* p : dev) if (p) cudaFree(p);
    void* host[] = {ctx->hset[0].edges, ctx->hset[0].frame_off, ctx->hset[0].stats, ctx->hset[0].counters, ctx->hset[0].row_off,
                    ctx->hset[1].edges, ctx->hset[1].frame_off, ctx->hset[1].stats, ctx->hset[1].counters, ctx->hset[1].row_off,
                    ctx->h_scalars, ctx->h_pass, ctx->pset[0].h_grid, ctx->pset[1].h_grid, ctx->h_pairs, ctx->h_text, ctx->h_cc_labels, ctx->h_cc_stats};
    for (void* p : host) if (p) cudaFreeHost(p);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->ev_stage) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->ev_copy) cudaEventDestroy(ev);
    for (auto& ev : ctx->ev_pass) cudaEventDestroy(ev);
    for (wn_ctx::PassSet& S : ctx->pset) if (S.back_done) cudaEventDestroy(S.back_done);
    if (ctx->st_fin) cudaStreamDestroy(ctx->st_fin);
    if (ctx->st_h2d) cudaStreamDestroy(ctx->st_h2d);
    if (ctx->st_d2h) cudaStreamDestroy(ctx->st_d2h);
    if (ctx->st) cudaStreamDestroy(ctx->st);
    delete ctx;
}

const char* wn_last_error(const wn_ctx* ctx)
{
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int wn_set_sites(wn_ctx* ctx, int32_t n_sites, const int32_t* site_atom, const int32_t* h_atom,
                 int32_t hmax, const uint8_t* site_type, int32_t first_limit)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    if (n_sites < 0 || hmax < 0 || (n_sites > 0 && (!site_atom || !site_type)) ||
        (n_sites > 0 && hmax > 0 && !h_atom))
        return fail(ctx, WN_ERR_BAD_ARG, "wn_set_sites: bad argument");
    if (n_sites >= (1 << 27)) return fail(ctx, WN_ERR_OVERFLOW, "wn_set_sites: more than 2^27-1 sites per frame");
    WN_CK(cudaSetDevice(ctx->device));
    /* Statically invalid hydrogens (the site atom itself, SURVEY F8: DH = 0/0 = NaN,
     * never selected) are dropped here; list positions are kept in the meta word. */
    int hs = 1, max_atom = -1;
    bool simple = true;     /* every site WATER with exactly one valid hydrogen, not at list position 0 */
    std::vector<uint32_t> meta((size_t)n_sites);
    std::vector<std::vector<int>> valid((size_t)n_sites);
    for (int s = 0; s < n_sites; ++s) {
        if (site_atom[s] < 0) return fail(ctx, WN_ERR_BAD_ARG, "wn_set_sites: negative site atom index");
        if (site_type[s] > WN_WATER) return fail(ctx, WN_ERR_BAD_ARG, "wn_set_sites: unknown site type");
        max_atom = std::max(max_atom, site_atom[s]);
        int nh = 0;
        bool first0 = false;
        for (int h = 0; h < hmax; ++h) {
            const int at = h_atom[(size_t)s * hmax + h];
            if (at < 0) break;
            max_atom = std::max(max_atom, at);
            if (at != site_atom[s]) {
                if (valid[s].empty() && h == 0) first0 = true;
                valid[s].push_back(at);
            }
            ++nh;
        }
        if ((int)valid[s].size() > WN_MAX_VALID_H)
            return fail(ctx, WN_ERR_BAD_ARG, "wn_set_sites: more than WN_MAX_VALID_H hydrogens on one site");
        hs = std::max(hs, (int)valid[s].size());
        meta[s] = (uint32_t)site_type[s] | (nh > 0 ? WN_META_HASH : 0u) | (first0 ? WN_META_FIRST0 : 0u) |
                  ((uint32_t)valid[s].size() << WN_META_NV_SHIFT);
        if (site_type[s] != WN_WATER || valid[s].size() != 1 || first0) simple = false;
    }
    std::vector<int> hv((size_t)n_sites * hs, -1);
    for (int s = 0; s < n_sites; ++s)
        for (size_t v = 0; v < valid[s].size(); ++v) hv[(size_t)s * hs + v] = valid[s][v];

    WN_CK(cudaStreamSynchronize(ctx->st));
    int rc;
    if ((rc = dev_realloc_plain(ctx, &ctx->d_site_atom, (size_t)n_sites))) return rc;
    if ((rc = dev_realloc_plain(ctx, &ctx->d_hv_atom, (size_t)n_sites * hs))) return rc;
    if ((rc = dev_realloc_plain(ctx, &ctx->d_meta, (size_t)n_sites))) return rc;
    if (n_sites > 0) {
        WN_CK(cudaMemcpy(ctx->d_site_atom, site_atom, (size_t)n_sites * sizeof(int), cudaMemcpyHostToDevice));
        WN_CK(cudaMemcpy(ctx->d_hv_atom, hv.data(), hv.size() * sizeof(int), cudaMemcpyHostToDevice));
        WN_CK(cudaMemcpy(ctx->d_meta, meta.data(), meta.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    ctx->n_sites = n_sites; ctx->hs = hs; ctx->simple = simple ? 1 : 0; ctx->max_atom = max_atom;
    ctx->first_limit = first_limit < 0 ? n_sites : first_limit;     /* negative: every row, as the C++ wrappers say */
    ctx->row_cap = 32;      /* a new table is a new system: forget a grown row capacity */
    return WN_OK;
}

int wn_set_water_sites(wn_ctx* ctx, int32_t n_waters, int32_t first_limit)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    if (n_waters < 0) return fail(ctx, WN_ERR_BAD_ARG, "wn_set_water_sites: negative count");
    std::vector<int32_t> sa((size_t)n_waters), ha((size_t)n_waters * 2);
    std::vector<uint8_t> ty((size_t)n_waters, (uint8_t)WN_WATER);
    for (int w = 0; w < n_waters; ++w) { sa[w] = 3 * w; ha[2 * w] = 3 * w; ha[2 * w + 1] = 3 * w + 1; }
    return wn_set_sites(ctx, n_waters, sa.data(), ha.data(), 2, ty.data(), first_limit);
}

int wn_set_energy_params(wn_ctx* ctx, double r_on, double r_off, double theta_on, double theta_off)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    /* computeEnergy passes (r*r, r_on*r_on, r_off*r_off) to the radial switch and
     * (cos*cos, theta_off, theta_on) to the angle switch (src/waternetwork.cpp:57-60) */
    ctx->ep.on2 = r_on * r_on; ctx->ep.off2 = r_off * r_off;
    ctx->ep.a_on = theta_off; ctx->ep.a_off = theta_on;
    set_radial_denominator(ctx->ep);
    return WN_OK;
}

int wn_set_batch_frames(wn_ctx* ctx, int32_t frames_per_pass)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    if (frames_per_pass < 0) return fail(ctx, WN_ERR_BAD_ARG, "wn_set_batch_frames: negative");
    ctx->user_frames_per_pass = frames_per_pass;
    return WN_OK;
}

int wn_set_result_sets(wn_ctx* ctx, int32_t n_sets)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    if (n_sets != 1 && n_sets != 2) return fail(ctx, WN_ERR_BAD_ARG, "wn_set_result_sets: 1 or 2");
    ctx->n_sets = n_sets;
    if (n_sets == 1) ctx->cur = 0;
    return WN_OK;
}

/* CPUs next to the device, from sysfs (the PCI device's local_cpulist); empty when the platform does not say */
static int local_cpus_of_device(int device, cpu_set_t* set)
{
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return 0; }
    for (char* c = bus; *c; ++c) *c = (char)tolower((unsigned char)*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* fh = fopen(path.c_str(), "r");
    if (!fh) return 0;
    char line[4096] = {0};
    const bool got = fgets(line, sizeof line, fh) != nullptr;
    fclose(fh);
    if (!got) return 0;
    CPU_ZERO(set);
    int n = 0;
    for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k == 1) b = a;
        if (k < 1) continue;
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, set); ++n; }
    }
    return n;
}

int wn_bind_thread_to_device(wn_ctx* ctx, int32_t* n_cpus)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    if (n_cpus) *n_cpus = 0;
    cpu_set_t want, have, both;
    const int n = local_cpus_of_device(ctx->device, &want);
    if (n <= 0) return WN_OK;                              /* the platform does not expose the locality: nothing to do */
    if (sched_getaffinity(0, sizeof have, &have) != 0) return WN_OK;
    CPU_AND(&both, &want, &have);
    const int k = CPU_COUNT(&both);
    if (k <= 0 || k == CPU_COUNT(&have)) { if (n_cpus) *n_cpus = k == CPU_COUNT(&have) ? k : 0; return WN_OK; }
    if (sched_setaffinity(0, sizeof both, &both) != 0) return WN_OK;
    if (n_cpus) *n_cpus = k;
    return WN_OK;
}

int wn_pcie_probe(wn_ctx* ctx, size_t bytes, int32_t repeats, double* h2d_gbs, double* d2h_gbs)
{
    if (!ctx || bytes == 0 || repeats < 1) return WN_ERR_BAD_ARG;
    WN_CK(cudaSetDevice(ctx->device));
    void *h = nullptr, *d = nullptr;
    WN_CK(cudaMallocHost(&h, bytes));
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess) { cudaFreeHost(h); return cuda_fail(ctx, e, "cudaMalloc (wn_pcie_probe)"); }
    memset(h, 1, bytes);
    float ms_in = 0.f, ms_out = 0.f;
    cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->st_h2d);           /* warm-up */
    cudaStreamSynchronize(ctx->st_h2d);
    cudaEventRecord(ctx->ev[4], ctx->st_h2d);
    for (int r = 0; r < repeats; ++r) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->st_h2d);
    cudaEventRecord(ctx->ev[5], ctx->st_h2d);
    cudaEventSynchronize(ctx->ev[5]);
    cudaEventElapsedTime(&ms_in, ctx->ev[4], ctx->ev[5]);
    cudaEventRecord(ctx->ev[4], ctx->st_d2h);
    for (int r = 0; r < repeats; ++r) cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->st_d2h);
    cudaEventRecord(ctx->ev[5], ctx->st_d2h);
    cudaEventSynchronize(ctx->ev[5]);
    cudaEventElapsedTime(&ms_out, ctx->ev[4], ctx->ev[5]);
    e = cudaGetLastError();
    cudaFree(d); cudaFreeHost(h);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "wn_pcie_probe");
    if (h2d_gbs) *h2d_gbs = (double)bytes * repeats / (ms_in * 1e-3) / 1e9;
    if (d2h_gbs) *d2h_gbs = (double)bytes * repeats / (ms_out * 1e-3) / 1e9;
    return WN_OK;
}

int wn_pcie_probe_mix(wn_ctx* ctx, size_t d2h_bytes, size_t h2d_bytes, int32_t repeats, double* seconds)
{
    if (!ctx || !seconds || repeats < 1 || (d2h_bytes == 0 && h2d_bytes == 0)) return WN_ERR_BAD_ARG;
    WN_CK(cudaSetDevice(ctx->device));
    void *ho = nullptr, *hi = nullptr, *d_o = nullptr, *di = nullptr;
    cudaError_t e = cudaSuccess;
    if (d2h_bytes) { e = cudaMallocHost(&ho, d2h_bytes); if (e == cudaSuccess) e = cudaMalloc(&d_o, d2h_bytes); }
    if (e == cudaSuccess && h2d_bytes) { e = cudaMallocHost(&hi, h2d_bytes); if (e == cudaSuccess) e = cudaMalloc(&di, h2d_bytes); }
    if (e == cudaSuccess) {
        if (ho) memset(ho, 1, d2h_bytes);                      /* touch: first-touch placement, no page faults in the timed part */
        if (hi) memset(hi, 1, h2d_bytes);
        cudaDeviceSynchronize();
        cudaEventRecord(ctx->ev[4], ctx->st_d2h);
        cudaStreamWaitEvent(ctx->st_h2d, ctx->ev[4], 0);
        for (int r = 0; r < repeats; ++r) {
            if (d2h_bytes) cudaMemcpyAsync(ho, d_o, d2h_bytes, cudaMemcpyDeviceToHost, ctx->st_d2h);
            if (h2d_bytes) cudaMemcpyAsync(di, hi, h2d_bytes, cudaMemcpyHostToDevice, ctx->st_h2d);
        }
        cudaEventRecord(ctx->ev[5], ctx->st_h2d);
        cudaStreamWaitEvent(ctx->st_d2h, ctx->ev[5], 0);
        cudaEventRecord(ctx->ev[5], ctx->st_d2h);                /* both directions done */
        cudaEventSynchronize(ctx->ev[5]);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
        *seconds = (double)ms * 1e-3 / repeats;
        e = cudaGetLastError();
    }
    if (d_o) cudaFree(d_o);
    if (di) cudaFree(di);
    if (ho) cudaFreeHost(ho);
    if (hi) cudaFreeHost(hi);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "wn_pcie_probe_mix");
    return WN_OK;
}

int wn_hbond_batch_ex(wn_ctx* ctx, const wn_batch_args* a, wn_batch_result* out)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    if (!a || !out) return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch_ex: NULL argument");
    const int nf = a->n_frames;
    if (a->box_stride != 0 && a->box_stride != 3 && a->box_stride != 9)
        return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch_ex: box_stride must be 0, 3 or 9");
    if (a->box_stride != 0 && !a->boxes && nf > 0) return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch_ex: boxes is NULL");
    /* per-frame site lists */
    if (a->subset_offsets) {
        if (ctx->n_sites < 0) return fail(ctx, WN_ERR_NO_SITES, "wn_hbond_batch_ex: call wn_set_sites first");
        for (int f = 0; f < nf; ++f) {
            const int64_t lo = a->subset_offsets[f], hi = a->subset_offsets[f + 1];
            if (lo < 0 || hi < lo || hi - lo > (int64_t)ctx->n_sites)
                return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch_ex: subset offsets must ascend, at most n_sites entries per frame");
        }
        const int64_t total = nf > 0 ? a->subset_offsets[nf] : 0;
        if (total > 0 && !a->subset_sites) return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch_ex: subset_sites is NULL");
        for (int64_t k = nf > 0 ? a->subset_offsets[0] : 0; k < total; ++k)
            if (a->subset_sites[k] < 0 || a->subset_sites[k] >= ctx->n_sites)
                return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch_ex: site index outside the site table");
    }
    /* a matrix batch whose frames are all rectangular uses the rectangular definition and grid */
    const float* boxes = a->box_stride ? a->boxes : nullptr;
    int stride = a->box_stride == 9 ? 9 : 3;
    std::vector<float> b3;
    if (a->box_stride == 9) {
        bool rectangular = true;
        for (int f = 0; f < nf && rectangular; ++f) {
            const float* b = a->boxes + 9 * (size_t)f;
            if (b[1] != 0.f || b[2] != 0.f || b[3] != 0.f || b[5] != 0.f || b[6] != 0.f || b[7] != 0.f) rectangular = false;
        }
        if (rectangular) {
            b3.resize((size_t)nf * 3);
            for (int f = 0; f < nf; ++f) {
                b3[3 * (size_t)f] = a->boxes[9 * (size_t)f]; b3[3 * (size_t)f + 1] = a->boxes[9 * (size_t)f + 4];
                b3[3 * (size_t)f + 2] = a->boxes[9 * (size_t)f + 8];
            }
            boxes = b3.data();
            stride = 3;
        }
    }
    if (a->result_layout != WN_LAYOUT_EDGES && a->result_layout != WN_LAYOUT_CSR16 && a->result_layout != WN_LAYOUT_CSR12)
        return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch_ex: unknown result_layout");
    ctx->layout = a->result_layout;
    ctx->keep_dev = a->results_on_device != 0;
    ctx->box_stride = stride;
    ctx->sub_off_host = a->subset_offsets;
    ctx->sub_sites_host = a->subset_offsets ? a->subset_sites : nullptr;
    const int rc = batch_common(ctx, a->frames, boxes, a->frames_on_device == 0, nf, a->natoms, out);
    ctx->box_stride = 3;
    ctx->keep_dev = false;
    ctx->sub_off_host = nullptr;
    ctx->sub_sites_host = nullptr;
    return rc;
}

static wn_batch_args make_args(const float* frames, int on_device, int32_t n_frames, int32_t natoms, int stride,
                               const float* boxes, const int64_t* so, const int32_t* ss)
{
    wn_batch_args a;
    a.frames = frames; a.frames_on_device = on_device; a.n_frames = n_frames; a.natoms = natoms;
    a.box_stride = stride; a.boxes = boxes; a.subset_offsets = so; a.subset_sites = ss;
    a.result_layout = WN_LAYOUT_EDGES; a.results_on_device = 0;
    return a;
}

int wn_hbond_batch(wn_ctx* ctx, const float* frames_host, int32_t n_frames, int32_t natoms, wn_batch_result* out)
{
    const wn_batch_args a = make_args(frames_host, 0, n_frames, natoms, 0, nullptr, nullptr, nullptr);
    return wn_hbond_batch_ex(ctx, &a, out);
}

int wn_hbond_batch_dev(wn_ctx* ctx, const float* frames_dev, int32_t n_frames, int32_t natoms, wn_batch_result* out)
{
    const wn_batch_args a = make_args(frames_dev, 1, n_frames, natoms, 0, nullptr, nullptr, nullptr);
    return wn_hbond_batch_ex(ctx, &a, out);
}

int wn_hbond_batch_pbc(wn_ctx* ctx, const float* frames_host, const float* boxes_host, int32_t n_frames,
                       int32_t natoms, wn_batch_result* out)
{
    if (ctx && !boxes_host && n_frames > 0) return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch_pbc: boxes is NULL");
    const wn_batch_args a = make_args(frames_host, 0, n_frames, natoms, 3, boxes_host, nullptr, nullptr);
    return wn_hbond_batch_ex(ctx, &a, out);
}

int wn_hbond_batch_pbc_dev(wn_ctx* ctx, const float* frames_dev, const float* boxes_host, int32_t n_frames,
                           int32_t natoms, wn_batch_result* out)
{
    if (ctx && !boxes_host && n_frames > 0) return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch_pbc_dev: boxes is NULL");
    const wn_batch_args a = make_args(frames_dev, 1, n_frames, natoms, 3, boxes_host, nullptr, nullptr);
    return wn_hbond_batch_ex(ctx, &a, out);
}

int wn_hbond_batch_box9(wn_ctx* ctx, const float* frames_host, const float* boxes9_host, int32_t n_frames,
                        int32_t natoms, wn_batch_result* out)
{
    if (ctx && !boxes9_host && n_frames > 0) return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch_box9: boxes is NULL");
    const wn_batch_args a = make_args(frames_host, 0, n_frames, natoms, 9, boxes9_host, nullptr, nullptr);
    return wn_hbond_batch_ex(ctx, &a, out);
}

int wn_hbond_batch_subset(wn_ctx* ctx, const float* frames_host, const float* boxes_host, int32_t n_frames,
                          int32_t natoms, const int64_t* subset_offsets, const int32_t* subset_sites,
                          wn_batch_result* out)
{
    if (ctx && n_frames > 0 && !subset_offsets)
        return fail(ctx, WN_ERR_BAD_ARG, "wn_hbond_batch_subset: subset_offsets is NULL");
    const wn_batch_args a = make_args(frames_host, 0, n_frames, natoms, boxes_host ? 3 : 0, boxes_host, subset_offsets,
                                      subset_sites);
    return wn_hbond_batch_ex(ctx, &a, out);
}

int wn_get_timings(wn_ctx* ctx, wn_timings* out)
{
    if (!ctx || !out) return WN_ERR_BAD_ARG;
    *out = ctx->tm;
    return WN_OK;
}

int wn_timer_start(wn_ctx* ctx)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    WN_CK(cudaSetDevice(ctx->device));
    WN_CK(cudaEventRecord(ctx->ev[6], ctx->st));
    return WN_OK;
}

int wn_timer_stop(wn_ctx* ctx, float* elapsed_ms)
{
    if (!ctx || !elapsed_ms) return WN_ERR_BAD_ARG;
    WN_CK(cudaSetDevice(ctx->device));
    WN_CK(cudaEventRecord(ctx->ev[7], ctx->st));
    WN_CK(cudaEventSynchronize(ctx->ev[7]));
    WN_CK(cudaEventElapsedTime(elapsed_ms, ctx->ev[6], ctx->ev[7]));
    return WN_OK;
}

static int find_pairs_impl(wn_ctx* ctx, const double* xyz, int32_t n, int32_t first_limit, const float* box3,
                           const int32_t** pairs, uint64_t* m, wn_tie_report* ties);

int wn_find_pairs(wn_ctx* ctx, const double* xyz, int32_t n, int32_t first_limit,
                  const int32_t** pairs, uint64_t* m, wn_tie_report* ties)
{
    return find_pairs_impl(ctx, xyz, n, first_limit, nullptr, pairs, m, ties);
}

int wn_find_pairs_pbc(wn_ctx* ctx, const double* xyz, int32_t n, int32_t first_limit, const float box[3],
                      const int32_t** pairs, uint64_t* m, wn_tie_report* ties)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    if (!box || !(box[0] > 0.f && box[1] > 0.f && box[2] > 0.f))
        return fail(ctx, WN_ERR_BAD_ARG, "wn_find_pairs_pbc: box lengths must be positive");
    return find_pairs_impl(ctx, xyz, n, first_limit, box, pairs, m, ties);
}

int wn_find_pairs_box9(wn_ctx* ctx, const double* xyz, int32_t n, int32_t first_limit, const float box9[9],
                       const int32_t** pairs, uint64_t* m, wn_tie_report* ties)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    if (!box9 || !(box9[0] > 0.f && box9[4] > 0.f && box9[8] > 0.f) || box9[1] != 0.f || box9[2] != 0.f || box9[5] != 0.f)
        return fail(ctx, WN_ERR_BAD_ARG, "wn_find_pairs_box9: box must be lower-triangular with a positive diagonal");
    if (box9[3] == 0.f && box9[6] == 0.f && box9[7] == 0.f) {
        const float b3[3] = {box9[0], box9[4], box9[8]};
        return find_pairs_impl(ctx, xyz, n, first_limit, b3, pairs, m, ties);
    }
    ctx->box_stride = 9;
    const int rc = find_pairs_impl(ctx, xyz, n, first_limit, box9, pairs, m, ties);
    ctx->box_stride = 3;
    return rc;
}

/* Grid of the cell-list pair search.  The pair test 10*d2 < 42.25 accepts distances below 2.0555 nm (the unit
 * slip of src/brute-find-pairs.cpp:19-20); cells of at least half that, so partners sit within two cells.
 * Open boundary: cubic cells over the bounding box of the finite points.  Periodic: the box divided into
 * floor(L/edge) cells per dimension, at least 5 (otherwise the 5 neighbour cells alias: all-pairs kernel). */
static bool cellpair_grid(const double* xyz, int n, const WnPairBox& pb, WnCellPairGrid* g)
{
    const double edge_min = 0.5 * std::sqrt(4.225) * (1.0 + 1.0e-9);
    const double cell_budget = std::max(4096.0, 2.0 * (double)n);
    memset(g, 0, sizeof *g);
    g->pbc = pb.pbc;
    double lo_all[3] = {HUGE_VAL, HUGE_VAL, HUGE_VAL}, hi_all[3] = {-HUGE_VAL, -HUGE_VAL, -HUGE_VAL};
    for (size_t k = 0; k < (size_t)n; ++k)
        for (int c = 0; c < 3; ++c) {
            const double v = xyz[3 * k + c];
            if (std::isfinite(v)) { lo_all[c] = std::min(lo_all[c], v); hi_all[c] = std::max(hi_all[c], v); }
        }
    if (pb.pbc == 1) {
        double e = edge_min;
        for (;;) {
            double cells = 1.0;
            for (int c = 0; c < 3; ++c) {
                const double q = std::floor(pb.L[c] / e);
                if (!(q >= 5.0)) return false;
                g->n[c] = (int)std::min(q, 1.0e6);
                cells *= (double)g->n[c];
            }
            if (cells <= cell_budget) break;
            e *= 1.26;
        }
        for (int c = 0; c < 3; ++c) {
            g->o[c] = 0.0; g->L[c] = pb.L[c]; g->invL[c] = pb.invL[c];
            g->inv[c] = (double)g->n[c] / pb.L[c];
        }
    } else {
        const double* lo = lo_all;
        const double* hi = hi_all;
        double e = edge_min;
        for (;;) {
            double cells = 1.0;
            for (int c = 0; c < 3; ++c) {
                const double ext = hi[c] >= lo[c] ? hi[c] - lo[c] : 0.0;
                g->n[c] = (int)std::min(std::floor(ext / e) + 1.0, 1.0e6);
                cells *= (double)g->n[c];
            }
            if (cells <= cell_budget) break;
            e *= 1.26;
        }
        for (int c = 0; c < 3; ++c) { g->o[c] = hi[c] >= lo[c] ? lo[c] : 0.0; g->inv[c] = 1.0 / e; }
    }
    g->ncells = g->n[0] * g->n[1] * g->n[2];
    /* fp32 reject bound on the squared distance.  The kernel tests origin-relative float copies: each component of a
     * displacement is off by <= e = 2^-20 * (largest |coordinate - origin| + largest box edge + cutoff) (roundings of the
     * two copies, of their difference and of the float image shift, with head-room), so an exact d2 <= 4.225 shows up
     * as d2_f32 <= (4.225 + 2*sqrt(3*4.225)*e + 3*e^2) * (1 + 2^-21): a strict superset, false positives ~1e-5. */
    double amax = 0.0, lmax = 0.0;
    for (int c = 0; c < 3; ++c) {
        if (hi_all[c] >= lo_all[c]) amax = std::max(amax, std::max(std::fabs(hi_all[c] - g->o[c]), std::fabs(lo_all[c] - g->o[c])));
        lmax = std::max(lmax, g->L[c]);
        g->Lf[c] = (float)g->L[c];
        g->invLf[c] = g->pbc ? 1.0f / (float)g->L[c] : 0.0f;
    }
    if (!(amax < 1.0e30)) return false;              /* float copies would overflow: exact all-pairs kernel */
    const double e = std::ldexp(amax + lmax + 2.1, -20);
    const double bd = (4.225 + 2.0 * std::sqrt(3.0 * 4.225) * e + 3.0 * e * e) * (1.0 + std::ldexp(1.0, -21));
    float bf = (float)bd;
    if ((double)bf < bd) bf = std::nextafterf(bf, INFINITY);
    g->bound_f = bf;
    return true;
}

static int find_pairs_impl(wn_ctx* ctx, const double* xyz, int32_t n, int32_t first_limit, const float* box3,
                           const int32_t** pairs, uint64_t* m, wn_tie_report* ties)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    WnPairBox pb;
    memset(&pb, 0, sizeof pb);
    if (box3 && ctx->box_stride == 9) {
        const float* b = box3;
        pb.pbc = 2;
        const float d[3] = {b[0], b[4], b[8]}, t[3] = {b[3], b[6], b[7]};
        for (int c = 0; c < 3; ++c) {
            pb.L[c] = (double)d[c]; pb.invL[c] = 1.0 / pb.L[c]; pb.Lf[c] = d[c]; pb.invLf[c] = 1.0f / d[c];
            pb.tri[c] = (double)t[c]; pb.trif[c] = t[c];
        }
    } else if (box3) {
        pb.pbc = 1;
        for (int c = 0; c < 3; ++c) {
            pb.L[c] = (double)box3[c]; pb.invL[c] = 1.0 / pb.L[c];
            pb.Lf[c] = box3[c]; pb.invLf[c] = 1.0f / box3[c];
        }
    }
    if (n < 0 || (n > 0 && !xyz) || !pairs || !m) return fail(ctx, WN_ERR_BAD_ARG, "wn_find_pairs: bad argument");
    WN_CK(cudaSetDevice(ctx->device));
    ctx->err.clear();
    *pairs = nullptr; *m = 0;
    if (ties) *ties = wn_tie_report{0, 0, 0, 0, 0};
    if (n < 2 || first_limit <= 0) {
        int rc;
        if ((rc = host_ensure(ctx, &ctx->h_pairs, &ctx->h_pairs_cap, (size_t)2))) return rc;
        *pairs = ctx->h_pairs;
        return WN_OK;
    }
    cudaStream_t st = ctx->st;
    int rc;
    if ((rc = dev_ensure(ctx, &ctx->d_xyz, &ctx->xyz_cap, (size_t)n * 3))) return rc;
    const size_t n_cnt = (size_t)n * wn_pairs_segments(n);      /* [row][column segment] */
    if (n_cnt + 1 > ctx->prow_cap) {
        WN_CK(cudaStreamSynchronize(st));
        if ((rc = dev_realloc_plain(ctx, &ctx->d_prow_cnt, n_cnt + 1))) return rc;
        if ((rc = dev_realloc_plain(ctx, &ctx->d_prow_off, n_cnt + 1))) return rc;
        ctx->prow_cap = n_cnt + 1;
    }
    /* cell-list sweep (O(N)) or the exact all-pairs tile kernel (tiny inputs, triclinic boxes, periodic boxes
     * under 5 cells of half the cutoff in a dimension) */
    WnCellPairGrid cg;
    bool cells = ctx->find_mode != WN_FIND_ALLPAIRS && pb.pbc != 2 && wn_cellpairs_supported(n) &&
                 (ctx->find_mode == WN_FIND_CELLS || n >= 1024);
    if (cells) cells = cellpair_grid(xyz, n, pb, &cg);
    ctx->last_find_cells = cells ? 1 : 0;
    if (cells) {
        if ((rc = dev_ensure(ctx, &ctx->d_cp_cell, &ctx->cp_cell_cap, (size_t)cg.ncells + 1))) return rc;
        if ((size_t)n > ctx->cp_site_cap) {
            WN_CK(cudaStreamSynchronize(st));
            if ((rc = dev_realloc_plain(ctx, &ctx->d_cp_site_cell, (size_t)n))) return rc;
            if ((rc = dev_realloc_plain(ctx, &ctx->d_cp_site_rank, (size_t)n))) return rc;
            if ((rc = dev_realloc_plain(ctx, &ctx->d_cp_srec, (size_t)n))) return rc;
            if ((rc = dev_realloc_plain(ctx, &ctx->d_cp_srecf, (size_t)n))) return rc;
            ctx->cp_site_cap = (size_t)n;
        }
    }
    const size_t n_tot = cells ? (size_t)n : n_cnt;             /* where the scan leaves the pair total */
    ctx->tm = wn_timings{};
    WN_CK(cudaMemcpyAsync(ctx->d_xyz, xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    WN_CK(cudaMemsetAsync(ctx->d_pair_ties, 0, 2 * sizeof(unsigned long long), st));
    WN_CK(cudaEventRecord(ctx->ev[0], st));
    if (cells) {
        wn_launch_cellpairs_bin(ctx->d_xyz, n, cg, ctx->d_cp_cell, ctx->d_cp_site_cell, ctx->d_cp_site_rank, ctx->d_cp_srec,
                                ctx->d_cp_srecf, st);
        wn_launch_cellpairs_count(ctx->d_cp_srec, ctx->d_cp_srecf, n, first_limit, cg, ctx->d_cp_cell, ctx->d_cp_site_cell, ctx->d_prow_cnt,
                                  ctx->d_pair_ties, st);
        wn_launch_scan64(ctx->d_prow_cnt, n, ctx->d_prow_off, st);
        ctx->tm.launches = 5;
    } else {
        wn_launch_pairs_count(ctx->d_xyz, n, first_limit, pb, ctx->d_prow_cnt, ctx->d_pair_ties, st);
        wn_launch_pairs_scan(ctx->d_prow_cnt, n, ctx->d_prow_off, st);
        ctx->tm.launches = 2;
    }
    WN_CK(cudaGetLastError());
    WN_CK(cudaEventRecord(ctx->ev[1], st));
    WN_CK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_prow_off + n_tot, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    WN_CK(cudaMemcpyAsync(ctx->h_scalars + 1, ctx->d_pair_ties, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    WN_CK(cudaStreamSynchronize(st));
    const unsigned long long total = ctx->h_scalars[0];
    if (ties) { ties->pair_cutoff_equal = ctx->h_scalars[1]; ties->pair_cutoff_near = ctx->h_scalars[2]; }
    if ((rc = dev_ensure(ctx, &ctx->d_pairs, &ctx->d_pairs_cap, (size_t)total * 2 + 2))) return rc;
    if ((rc = host_ensure(ctx, &ctx->h_pairs, &ctx->h_pairs_cap, (size_t)total * 2 + 2))) return rc;
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->tm.total_ms = ms;
    if (total) {
        WN_CK(cudaEventRecord(ctx->ev[2], st));
        if (cells)
            wn_launch_cellpairs_fill(ctx->d_cp_srec, ctx->d_cp_srecf, n, first_limit, cg, ctx->d_cp_cell, ctx->d_cp_site_cell, ctx->d_prow_off,
                                     ctx->d_pairs, st);
        else
            wn_launch_pairs_fill(ctx->d_xyz, n, first_limit, pb, ctx->d_prow_off, ctx->d_pairs, st);
        WN_CK(cudaGetLastError());
        WN_CK(cudaEventRecord(ctx->ev[3], st));
        WN_CK(cudaMemcpyAsync(ctx->h_pairs, ctx->d_pairs, (size_t)total * 2 * sizeof(int32_t),
                              cudaMemcpyDeviceToHost, st));
        WN_CK(cudaEventRecord(ctx->ev[4], st));
        WN_CK(cudaStreamSynchronize(st));
        cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
        ctx->tm.total_ms += ms;
        cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]);
        ctx->tm.d2h_ms = ms;
        ctx->tm.launches += 1;
    }
    ctx->tm.passes = 1;
    *pairs = ctx->h_pairs;
    *m = total;
    return WN_OK;
}

/* ---- make_edges over a caller-supplied pair list (src/waternetwork.cpp:182-263) ---- */
int wn_make_edges(wn_ctx* ctx, const float* frame, int32_t natoms, const float* box9, const int32_t* pairs, uint64_t m,
                  const wn_edge** edges, uint64_t* n_edges, uint64_t* n_pair_evals, wn_tie_report* ties)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    if (!edges || !n_edges || natoms < 0 || (!frame && natoms > 0) || (!pairs && m > 0))
        return fail(ctx, WN_ERR_BAD_ARG, "wn_make_edges: bad argument");
    if (ctx->n_sites < 0) return fail(ctx, WN_ERR_NO_SITES, "wn_make_edges: call wn_set_sites first");
    if (ctx->max_atom >= natoms && ctx->n_sites > 0)
        return fail(ctx, WN_ERR_BAD_ARG, "wn_make_edges: site table refers to atoms beyond natoms");
    if (m >= ((uint64_t)1 << 40)) return fail(ctx, WN_ERR_OVERFLOW, "wn_make_edges: pair list too long");
    WN_CK(cudaSetDevice(ctx->device));
    ctx->err.clear();
    ctx->tm = wn_timings{};
    *edges = nullptr; *n_edges = 0;
    if (n_pair_evals) *n_pair_evals = 0;
    if (ties) *ties = wn_tie_report{0, 0, 0, 0, 0};
    const int N = ctx->n_sites;
    int pbc = 0;
    WnGrid g;
    memset(&g, 0, sizeof g);
    int rc;
    if ((rc = ensure_pass_buffers(ctx, 1))) return rc;
    if (box9) {
        for (int k = 0; k < 9; ++k)
            if (!std::isfinite(box9[k])) return fail(ctx, WN_ERR_BAD_ARG, "wn_make_edges: box must be finite");
        if (!(box9[0] > 0.f && box9[4] > 0.f && box9[8] > 0.f) || box9[1] != 0.f || box9[2] != 0.f || box9[5] != 0.f)
            return fail(ctx, WN_ERR_BAD_ARG, "wn_make_edges: box must be lower-triangular with a positive diagonal (GROMACS convention)");
        if (box9[3] == 0.f && box9[6] == 0.f && box9[7] == 0.f) {
            const float b3[3] = {box9[0], box9[4], box9[8]};
            wn_host_pbc_grid(b3, ctx->cell_cap, &g);
            pbc = 1;
        } else {
            wn_host_tric_grid(box9, ctx->cell_cap, &g);
            pbc = 2;
        }
    }
    if (N == 0 || m == 0) {
        if ((rc = host_ensure(ctx, &ctx->hset[ctx->cur].edges, &ctx->hset[ctx->cur].edges_cap, (size_t)1))) return rc;
        *edges = ctx->hset[ctx->cur].edges;
        return WN_OK;
    }
    cudaStream_t st = ctx->st;
    ctx->last_frames = -1;                       /* the call reuses the batch buffers: no "last batch" afterwards */
    const int nb = wn_pl_blocks(m);
    if ((rc = dev_ensure(ctx, &ctx->d_stage[0], &ctx->stage_cap[0], (size_t)natoms * 3))) return rc;
    if ((rc = dev_ensure(ctx, &ctx->d_pl_pairs, &ctx->pl_pairs_cap, (size_t)m))) return rc;
    if ((rc = dev_ensure(ctx, &ctx->d_pl_tmp, &ctx->pl_tmp_cap, (size_t)m))) return rc;
    if ((rc = dev_ensure(ctx, &ctx->d_pl_blk, &ctx->pl_blk_cap, (size_t)nb + 1))) return rc;
    if ((rc = dev_ensure(ctx, &ctx->d_pl_off, &ctx->pl_off_cap, (size_t)nb + 1))) return rc;
    if (!ctx->d_pl_counters) WN_CK(cudaMalloc((void**)&ctx->d_pl_counters, sizeof(WnFrameCounters) + 8));
    unsigned int* d_bad = reinterpret_cast<unsigned int*>(ctx->d_pl_counters + 1);
    /* pageable sources: each call returns once the driver has staged the data */
    WN_CK(cudaMemcpyAsync(ctx->d_stage[0], frame, (size_t)natoms * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    WN_CK(cudaMemcpyAsync(ctx->d_pl_pairs, pairs, (size_t)m * sizeof(int2), cudaMemcpyHostToDevice, st));
    WN_CK(cudaMemcpyAsync(ctx->d_grid, &g, sizeof g, cudaMemcpyHostToDevice, st));
    WN_CK(cudaMemsetAsync(ctx->d_pl_counters, 0, sizeof(WnFrameCounters) + 8, st));
    WN_CK(cudaEventRecord(ctx->ev[0], st));
    /* site records in table order: the batched path's staging kernel with the identity permutation
     * (one cell holding every site, rank = table row) */
    WN_CK(cudaMemsetAsync(ctx->d_cell_cnt, 0, 2 * sizeof(int), st));
    WN_CK(cudaMemsetAsync(ctx->d_site_cell, 0, (size_t)N * sizeof(int), st));
    wn_launch_pl_iota(ctx->d_site_rank, N, st);
    const WnSubset nosub = {nullptr, nullptr};
    wn_launch_scatter(ctx->d_stage[0], 1, natoms, ctx->d_site_atom, ctx->d_hv_atom, ctx->d_meta, N, nosub, ctx->hs,
                      ctx->cell_cap, ctx->d_cell_cnt, ctx->d_site_cell, ctx->d_site_rank, ctx->d_grid, ctx->d_spos,
                      ctx->d_smeta, ctx->d_sdh, st);
    wn_launch_pl_eval(ctx->d_spos, ctx->d_smeta, ctx->d_sdh, ctx->hs, ctx->simple, pbc, ctx->d_grid, ctx->d_pl_pairs, m, N,
                      ctx->d_pl_tmp, ctx->d_pl_blk, ctx->d_pl_counters, d_bad, st);
    wn_launch_scan64(ctx->d_pl_blk, nb, ctx->d_pl_off, st);
    WN_CK(cudaGetLastError());
    WN_CK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_pl_off + nb, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    WN_CK(cudaMemcpyAsync(ctx->h_scalars + 1, ctx->d_pl_counters, sizeof(WnFrameCounters) + 8, cudaMemcpyDeviceToHost, st));
    WN_CK(cudaStreamSynchronize(st));
    const unsigned long long total = ctx->h_scalars[0];
    WnFrameCounters fc;
    memcpy(&fc, ctx->h_scalars + 1, sizeof fc);
    unsigned int bad = 0;
    memcpy(&bad, reinterpret_cast<const char*>(ctx->h_scalars + 1) + sizeof fc, sizeof bad);
    if (bad) return fail(ctx, WN_ERR_BAD_ARG, "wn_make_edges: a pair refers to a site outside the table");
    if ((size_t)total > ctx->edges_cap || !ctx->d_edges)
        if ((rc = dev_ensure(ctx, &ctx->d_edges, &ctx->edges_cap, (size_t)total + (size_t)total / 16 + 4096))) return rc;
    if ((rc = host_ensure(ctx, &ctx->hset[ctx->cur].edges, &ctx->hset[ctx->cur].edges_cap, (size_t)total + 1))) return rc;
    wn_launch_pl_write(ctx->d_pl_pairs, ctx->d_pl_tmp, m, ctx->d_pl_off, ctx->ep, ctx->d_edges, ctx->edges_cap, st);
    WN_CK(cudaGetLastError());
    WN_CK(cudaEventRecord(ctx->ev[1], st));
    if (total)
        WN_CK(cudaMemcpyAsync(ctx->hset[ctx->cur].edges, ctx->d_edges, (size_t)total * sizeof(wn_edge), cudaMemcpyDeviceToHost, st));
    WN_CK(cudaEventRecord(ctx->ev[2], st));
    WN_CK(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->tm.total_ms = ms;
    cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); ctx->tm.d2h_ms = ms;
    ctx->tm.launches = 5; ctx->tm.passes = 1;
    *edges = ctx->hset[ctx->cur].edges;
    *n_edges = total;
    if (n_pair_evals) *n_pair_evals = fc.evals;
    if (ties) { ties->length_equal_cut = fc.tie_len; ties->cosine_argmax_equal = fc.tie_cos; ties->pos_zero = fc.tie_pos0; }
    return WN_OK;
}

int wn_set_find_mode(wn_ctx* ctx, int32_t mode)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    if (mode != WN_FIND_AUTO && mode != WN_FIND_ALLPAIRS && mode != WN_FIND_CELLS)
        return fail(ctx, WN_ERR_BAD_ARG, "wn_set_find_mode: unknown mode");
    ctx->find_mode = mode;
    return WN_OK;
}

int wn_last_find_used_cells(const wn_ctx* ctx) { return ctx ? ctx->last_find_cells : 0; }

/* ---- plumbing ---- */
int wn_dev_alloc(wn_ctx* ctx, size_t bytes, void** dev_ptr)
{
    if (!ctx || !dev_ptr) return WN_ERR_BAD_ARG;
    WN_CK(cudaSetDevice(ctx->device));
    WN_CK(cudaMalloc(dev_ptr, bytes ? bytes : 1));
    return WN_OK;
}
int wn_dev_free(wn_ctx* ctx, void* dev_ptr)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    WN_CK(cudaSetDevice(ctx->device));
    WN_CK(cudaStreamSynchronize(ctx->st));
    WN_CK(cudaFree(dev_ptr));
    return WN_OK;
}
int wn_host_alloc_pinned(wn_ctx* ctx, size_t bytes, void** host_ptr)
{
    if (!ctx || !host_ptr) return WN_ERR_BAD_ARG;
    WN_CK(cudaSetDevice(ctx->device));
    WN_CK(cudaMallocHost(host_ptr, bytes ? bytes : 1));
    return WN_OK;
}
int wn_host_free_pinned(wn_ctx* ctx, void* host_ptr)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    WN_CK(cudaFreeHost(host_ptr));
    return WN_OK;
}
int wn_memcpy_d2h(wn_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    WN_CK(cudaSetDevice(ctx->device));
    WN_CK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->st));
    WN_CK(cudaStreamSynchronize(ctx->st));
    return WN_OK;
}
int wn_memcpy_h2d(wn_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes)
{
    if (!ctx) return WN_ERR_BAD_ARG;
    WN_CK(cudaSetDevice(ctx->device));
    WN_CK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->st));
    WN_CK(cudaStreamSynchronize(ctx->st));
    return WN_OK;
}
int wn_synth_frames_dev(wn_ctx* ctx, int32_t n_waters, uint64_t seed, int32_t model,
                        int64_t frame0, int32_t n_frames, float* frames_dev)
{
    if (!ctx || n_waters < 0 || n_frames < 0 || (!frames_dev && n_waters > 0 && n_frames > 0))
        return ctx ? fail(ctx, WN_ERR_BAD_ARG, "wn_synth_frames_dev: bad argument") : WN_ERR_BAD_ARG;
    WN_CK(cudaSetDevice(ctx->device));
    wn_launch_synth(n_waters, seed, model, frame0, n_frames, frames_dev, ctx->st);
    WN_CK(cudaGetLastError());
    WN_CK(cudaStreamSynchronize(ctx->st));
    return WN_OK;
}

int wn_format_edges(wn_ctx* ctx, const int32_t* frame_numbers, const uint32_t* site_ids, int32_t n_site_ids,
                    const char** text, uint64_t* text_bytes)
{
    if (!ctx || !text || !text_bytes) return WN_ERR_BAD_ARG;
    *text = nullptr; *text_bytes = 0;
    if (ctx->last_frames < 0) return fail(ctx, WN_ERR_BAD_ARG, "wn_format_edges: no batch result to format");
    if (ctx->layout != WN_LAYOUT_EDGES) return fail(ctx, WN_ERR_BAD_ARG, "wn_format_edges: needs a WN_LAYOUT_EDGES result");
    /* per-frame site lists: edge indices are positions in each frame's list, not table rows, so a table of
     * ids cannot be applied on the device; map the indices on the host (EdgeFileWriter) instead */
    if (ctx->last_subset && site_ids)
        return fail(ctx, WN_ERR_BAD_ARG, "wn_format_edges: the last batch used per-frame site lists (edge indices are list "
                    "positions); pass site_ids = NULL or format on the host");
    const int F = ctx->last_frames;
    if (F > 0 && !frame_numbers) return fail(ctx, WN_ERR_BAD_ARG, "wn_format_edges: frame_numbers is NULL");
    if (site_ids && n_site_ids < ctx->n_sites)
        return fail(ctx, WN_ERR_BAD_ARG, "wn_format_edges: site_ids shorter than the site table");
    WN_CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->st;
    /* frame headers and row positions from the per-frame edge counts (already on the host) */
    std::vector<std::string> hdr((size_t)F);
    std::vector<unsigned long long> row_base((size_t)F + 1);
    unsigned long long pos = 0, max_edges = 0;
    for (int f = 0; f < F; ++f) {
        hdr[f] = "#--- frame " + std::to_string(frame_numbers[f]) + " ---\n";
        const unsigned long long cnt = ctx->hset[ctx->cur].frame_off[f + 1] - ctx->hset[ctx->cur].frame_off[f];
        row_base[f] = pos + hdr[f].size();
        pos = row_base[f] + 37ull * cnt;
        max_edges = std::max(max_edges, cnt);
    }
    row_base[F] = pos;
    int rc;
    if ((rc = dev_ensure(ctx, &ctx->d_text, &ctx->d_text_cap, (size_t)pos + 4))) return rc;
    if ((rc = host_ensure(ctx, &ctx->h_text, &ctx->h_text_cap, (size_t)pos + 4))) return rc;
    if ((rc = dev_ensure(ctx, &ctx->d_row_base, &ctx->row_base_cap, (size_t)F + 1))) return rc;
    if ((rc = dev_ensure(ctx, &ctx->d_frame_flags, &ctx->frame_flags_cap, (size_t)F + 1))) return rc;
    WN_CK(cudaMemcpyAsync(ctx->d_row_base, row_base.data(), ((size_t)F + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    WN_CK(cudaMemsetAsync(ctx->d_frame_flags, 0, ((size_t)F + 1) * sizeof(int), st));
    const unsigned int* d_ids = nullptr;
    if (site_ids) {
        if ((rc = dev_ensure(ctx, &ctx->d_site_ids, &ctx->site_ids_cap, (size_t)std::max(ctx->n_sites, 1)))) return rc;
        WN_CK(cudaMemcpyAsync(ctx->d_site_ids, site_ids, (size_t)ctx->n_sites * sizeof(unsigned int), cudaMemcpyHostToDevice, st));
        d_ids = ctx->d_site_ids;
    }
    WN_CK(cudaStreamSynchronize(st));             /* pageable sources above */
    wn_launch_format_rows(ctx->d_edges, ctx->d_frame_off, ctx->d_row_base, d_ids, F, max_edges, ctx->d_text,
                          ctx->d_frame_flags, st);
    WN_CK(cudaGetLastError());
    std::vector<int> flags((size_t)F + 1, 0);
    if (pos) WN_CK(cudaMemcpyAsync(ctx->h_text, ctx->d_text, (size_t)pos, cudaMemcpyDeviceToHost, st));
    WN_CK(cudaMemcpyAsync(flags.data(), ctx->d_frame_flags, ((size_t)F + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    WN_CK(cudaStreamSynchronize(st));
    for (int f = 0; f < F; ++f)
        if (flags[f]) return fail(ctx, WN_ERR_OVERFLOW, "wn_format_edges: a field does not fit the fixed-width row; format this batch on the host");
    for (int f = 0; f < F; ++f) memcpy(ctx->h_text + (row_base[f] - hdr[f].size()), hdr[f].data(), hdr[f].size());
    *text = ctx->h_text;
    *text_bytes = pos;
    return WN_OK;
}

int wn_components(wn_ctx* ctx, float weight_threshold, const int32_t** labels, const wn_component_stats** stats)
{
    if (!ctx || !labels || !stats) return WN_ERR_BAD_ARG;
    *labels = nullptr; *stats = nullptr;
    if (ctx->last_frames < 0) return fail(ctx, WN_ERR_BAD_ARG, "wn_components: no batch result to analyse");
    WN_CK(cudaSetDevice(ctx->device));
    const int F = ctx->last_frames, N = std::max(ctx->n_sites, 0);
    const size_t n = (size_t)F * N;
    int rc;
    if ((rc = dev_ensure(ctx, &ctx->d_cc_labels, &ctx->cc_labels_cap, n + 1))) return rc;
    if ((rc = dev_ensure(ctx, &ctx->d_cc_work, &ctx->cc_work_cap, n + 1))) return rc;
    if ((rc = dev_ensure(ctx, &ctx->d_cc_count, &ctx->cc_count_cap, n + 1))) return rc;
    if ((rc = dev_ensure(ctx, &ctx->d_cc_minidx, &ctx->cc_minidx_cap, n + 1))) return rc;
    if ((rc = dev_ensure(ctx, &ctx->d_cc_kept, &ctx->cc_kept_cap, (size_t)F + 1))) return rc;
    if ((rc = dev_ensure(ctx, &ctx->d_cc_stats, &ctx->cc_stats_cap, (size_t)F + 1))) return rc;
    unsigned long long max_edges = 0;
    for (int f = 0; f < F; ++f) max_edges = std::max<unsigned long long>(max_edges, ctx->hset[ctx->cur].frame_off[f + 1] - ctx->hset[ctx->cur].frame_off[f]);
    if (ctx->layout == WN_LAYOUT_CSR12)
        return fail(ctx, WN_ERR_BAD_ARG, "wn_components: a WN_LAYOUT_CSR12 result carries no energies; use WN_LAYOUT_EDGES or WN_LAYOUT_CSR16");
    wn_launch_components(ctx->d_edges, ctx->layout == WN_LAYOUT_CSR16 ? 1 : 0, ctx->d_frame_off, ctx->d_row_off, F, N,
                         max_edges, weight_threshold, ctx->d_cc_labels, ctx->d_cc_work, ctx->d_cc_count, ctx->d_cc_minidx,
                         ctx->d_cc_kept, ctx->d_cc_stats, ctx->st);
    WN_CK(cudaGetLastError());
    if (ctx->last_on_device) {
        WN_CK(cudaStreamSynchronize(ctx->st));
        *labels = ctx->d_cc_labels; *stats = ctx->d_cc_stats;
        return WN_OK;
    }
    if ((rc = host_ensure(ctx, &ctx->h_cc_labels, &ctx->h_cc_labels_cap, n + 1))) return rc;
    if ((rc = host_ensure(ctx, &ctx->h_cc_stats, &ctx->h_cc_stats_cap, (size_t)F + 1))) return rc;
    if (n) WN_CK(cudaMemcpyAsync(ctx->h_cc_labels, ctx->d_cc_labels, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    if (F) WN_CK(cudaMemcpyAsync(ctx->h_cc_stats, ctx->d_cc_stats, (size_t)F * sizeof(wn_component_stats), cudaMemcpyDeviceToHost, ctx->st));
    WN_CK(cudaStreamSynchronize(ctx->st));
    *labels = ctx->h_cc_labels; *stats = ctx->h_cc_stats;
    return WN_OK;
}

int wn_selftest_div(wn_ctx* ctx, uint64_t n_samples, uint64_t seed, uint64_t* mismatches)
{
    if (!ctx || !mismatches) return WN_ERR_BAD_ARG;
    WN_CK(cudaSetDevice(ctx->device));
    WN_CK(cudaMemsetAsync(ctx->d_pair_ties, 0, sizeof(unsigned long long), ctx->st));
    const int iters = 1024;
    const unsigned long long per_block = 256ull * iters;
    int blocks = (int)std::min<unsigned long long>((n_samples + per_block - 1) / per_block, 1u << 20);
    if (blocks < 1) blocks = 1;
    wn_launch_selftest_div(seed, blocks, iters, ctx->d_pair_ties, ctx->st);
    WN_CK(cudaGetLastError());
    WN_CK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_pair_ties, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->st));
    WN_CK(cudaStreamSynchronize(ctx->st));
    *mismatches = ctx->h_scalars[0];
    return WN_OK;
}

/* ---- NCCL statistics reduction ---- */
static void* nccl_open(std::string* why)
{
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        void* h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) return h;
    }
    const char* e = dlerror();                  /* one call: dlerror() clears the message it returns */
    if (why) *why = e ? e : "libnccl.so.2 not found";
    return nullptr;
}

int wn_nccl_unique_id(void* id128)
{
    if (!id128) return WN_ERR_BAD_ARG;
    std::string why;
    void* lib = nccl_open(&why);
    if (!lib) return fail(nullptr, WN_ERR_NCCL, "wn_nccl_unique_id: " + why);
    typedef ncclResult_t (*fn_t)(ncclUniqueId*);
    fn_t fn = (fn_t)dlsym(lib, "ncclGetUniqueId");
    if (!fn) return fail(nullptr, WN_ERR_NCCL, "ncclGetUniqueId not found");
    ncclUniqueId id;
    if (fn(&id) != ncclSuccess) return fail(nullptr, WN_ERR_NCCL, "ncclGetUniqueId failed");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(id128, &id, 128);
    return WN_OK;
}

int wn_nccl_init(wn_ctx* ctx, const void* id128, int32_t rank, int32_t nranks)
{
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return WN_ERR_BAD_ARG;
    WN_CK(cudaSetDevice(ctx->device));
    std::string why;
    if (!ctx->nccl_lib) ctx->nccl_lib = nccl_open(&why);
    if (!ctx->nccl_lib) return fail(ctx, WN_ERR_NCCL, "wn_nccl_init: " + why);
    typedef ncclResult_t (*fn_t)(ncclComm_t*, int, ncclUniqueId, int);
    fn_t fn = (fn_t)dlsym(ctx->nccl_lib, "ncclCommInitRank");
    if (!fn) return fail(ctx, WN_ERR_NCCL, "ncclCommInitRank not found");
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = fn(&ctx->comm, nranks, id, rank);
    if (r != ncclSuccess) return fail(ctx, WN_ERR_NCCL, "ncclCommInitRank failed: " + std::to_string((int)r));
    ctx->nranks = nranks; ctx->rank = rank;
    return WN_OK;
}

int wn_stats_reduce(wn_ctx* ctx, wn_frame_stats* stats_host, int64_t n_frames_total)
{
    if (!ctx || (!stats_host && n_frames_total > 0) || n_frames_total < 0) return WN_ERR_BAD_ARG;
    if (!ctx->comm) return fail(ctx, WN_ERR_NCCL, "wn_stats_reduce: call wn_nccl_init first");
    if (n_frames_total == 0) return WN_OK;
    WN_CK(cudaSetDevice(ctx->device));
    const size_t count = (size_t)n_frames_total * (sizeof(wn_frame_stats) / sizeof(double));
    int rc;
    if ((rc = dev_ensure(ctx, &ctx->d_reduce, &ctx->reduce_cap, count))) return rc;
    WN_CK(cudaMemcpyAsync(ctx->d_reduce, stats_host, count * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
    typedef ncclResult_t (*fn_t)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    fn_t fn = (fn_t)dlsym(ctx->nccl_lib, "ncclAllReduce");
    if (!fn) return fail(ctx, WN_ERR_NCCL, "ncclAllReduce not found");
    ncclResult_t r = fn(ctx->d_reduce, ctx->d_reduce, count, ncclDouble, ncclSum, ctx->comm, ctx->st);
    if (r != ncclSuccess) return fail(ctx, WN_ERR_NCCL, "ncclAllReduce failed: " + std::to_string((int)r));
    WN_CK(cudaMemcpyAsync(stats_host, ctx->d_reduce, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    WN_CK(cudaStreamSynchronize(ctx->st));
    return WN_OK;
}

}  /* extern "C" */
