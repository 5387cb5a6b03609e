/*
 * wn_internal.cuh -- shared declarations of the sm_100a kernels and the context.
 * Not part of the public boundary (that is include/wn_b200.h).
 */
#ifndef WN_INTERNAL_CUH
#define WN_INTERNAL_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "wn_b200.h"

/* ---- checked build (make -C gmx-acqueduct_b200 checked -> libwn_b200_checked.so, -DWN_CHECKS) ----
 * compute-sanitizer is not available on every pool; this build carries its own bounds checks on every shared-memory
 * queue, bitmap and row-slot index and on every computed global index, and traps (the call then fails with
 * cudaErrorAssert / an illegal-instruction error) where the release build would read or write out of bounds. */
#if defined(WN_CHECKS) && defined(__CUDACC__)
#include <cstdio>
#define WN_CHECK(cond)                                                                                  \
    do {                                                                                                \
        if (!(cond)) {                                                                                  \
            printf("WN_CHECK failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__,     \
                   (int)blockIdx.x, (int)threadIdx.x);                                                  \
            __trap();                                                                                   \
        }                                                                                               \
    } while (0)
#else
#define WN_CHECK(cond) ((void)0)
#endif

/* ---- per-frame cell grid ----
 * open boundary: bounding box of the finite sites (wn_k_bbox), cubic cells.
 * periodic     : the box itself, n = floor(L/edge_min) cells per dimension (host-built). */
struct WnGrid {
    double ox, oy, oz;   /* origin (minimum corner); 0 in periodic mode            */
    double inv[3];       /* 1 / cell edge, per dimension                            */
    double L[3];         /* periodic box lengths (fp64 of the float box), else 0    */
    double invL[3];      /* 1.0 / L                                                  */
    int nx, ny, nz;
    int ncells;          /* nx*ny*nz <= cell capacity                               */
    float edge[3];       /* cell edge (fp32 copy, for block-local frames)           */
    float tf;            /* fp32 prefilter bound on d^2 (superset of the exact test) */
    int pbc;             /* 0: open, 1: rectangular periodic box, 2: triclinic (L = ax,by,cz) */
    double tri[3];       /* triclinic off-diagonals bx, cx, cy (box rows a=(ax,0,0) b=(bx,by,0) c=(cx,cy,cz)) */
};

/* Minimum cell edge, and the fp32 prefilter bound for a given (largest) cell edge: exact
 * acceptance needs d2 <= 0.4225*(1+2.5e-7); the dot-product form on block-local
 * coordinates (|a|^2 <= 1.5 e^2, |c|^2 <= 8.5 e^2, six fp32 roundings on terms <= 8.5 e^2,
 * coordinates rounded at <= 6e-8 * 2e) errs by <= ~3.1e-6 e^2 + 5e-7 e; > 2x head-room.
 * Periodic grids add a term for the float image shift folded into the run origin
 * (wn_host_pbc_grid: 8 ulp of twice the longest box vector). */
#define WN_CELL_EDGE      (0.65 * (1.0 + 1.0e-5))
#define WN_PREFILTER_BOUND(e) (0.4225 * (1.0 + 1.0e-5) + 8.0e-6 * (e) * (e) + 2.0e-6 * (e))

/* ---- per-frame site subsets (e.g. the buried waters the alpha-shape filter keeps, which
 * change every frame: reference src/waternetwork.cpp:529-533,562-576).  Frame f uses the table
 * rows sites[off[f] .. off[f+1]) in that order; site index i of the frame = position in that
 * list.  off == nullptr: every frame uses the whole table. ---- */
struct WnSubset {
    const int* off;      /* [F+1], device */
    const int* sites;    /* table rows, device */
};
#if defined(__CUDACC__)
/* triclinic minimum image (include/wn_b200.h, "Periodic boundaries"): reduce along c, then b,
 * then a; every operation rounded.  L = (ax, by, cz), tri = (bx, cx, cy). */
__device__ __forceinline__ void wn_minimg_tric(double& x, double& y, double& z, const double* L, const double* tri)
{
    const double k = rint(__ddiv_rn(z, L[2]));
    x = __dsub_rn(x, __dmul_rn(k, tri[1])); y = __dsub_rn(y, __dmul_rn(k, tri[2])); z = __dsub_rn(z, __dmul_rn(k, L[2]));
    const double j = rint(__ddiv_rn(y, L[1]));
    x = __dsub_rn(x, __dmul_rn(j, tri[0])); y = __dsub_rn(y, __dmul_rn(j, L[1]));
    const double i = rint(__ddiv_rn(x, L[0]));
    x = __dsub_rn(x, __dmul_rn(i, L[0]));
}
__device__ __forceinline__ int wn_frame_sites(const WnSubset& s, int f, int n_table) { return s.off ? s.off[f + 1] - s.off[f] : n_table; }
__device__ __forceinline__ int wn_table_row(const WnSubset& s, int f, int i) { return s.off ? s.sites[s.off[f] + i] : i; }
#endif

/* ---- per-frame counters written by the fused kernel ---- */
struct WnFrameCounters {
    unsigned long long evals;      /* `if (isedge)` bodies                      */
    unsigned long long tie_len;    /* (float)dist == 6.5f                       */
    unsigned long long tie_cos;    /* equal positive cosines in the arg-max     */
    unsigned long long tie_pos0;   /* arg-max on list index 0                   */
    unsigned long long tests;      /* (home site, candidate) pairs the fp32 prefilter looked at (implementation-dependent) */
    unsigned long long short_edges;/* edges shorter than WN_SHORT_EDGE_A: their energy may leave the fixed-point range of
                                    * the ordered copy's energy sum (wn_k_copy), which then sums this frame in final order */
};
/* |E| <= 3855/r^6 (cos^4 <= 1, switches <= 1): below 0.56 A an edge's energy can reach 2^17 */
#define WN_SHORT_EDGE_A 0.56f

/* ---- site meta word (one per site) ---- */
#define WN_META_TYPE_MASK 3u
#define WN_META_HASH      (1u << 2)  /* hydrogen list not empty                       */
#define WN_META_FIRST0    (1u << 3)  /* first valid hydrogen sits at list position 0  */
#define WN_META_NV_SHIFT  4          /* number of valid hydrogens, 0..WN_MAX_VALID_H  */
#define WN_META_NV_MASK   0xFu

/* Edge-length cut of make_edges (src/waternetwork.cpp:204): reject when the float
 * length in Angstrom is > 6.5.  Any accepted pair has squared distance (nm^2)
 * <= 0.4225*(1+2.5e-7); the fp32 prefilter and the cell edge carry a margin. */
#define WN_LEN_CUT_F      6.5f

/* Energy parameters in the form the kernels consume (computeEnergy, :49-66). */
struct WnEnergyParams {
    double on2, off2;      /* r_on^2, r_off^2 (arguments of the radial switch)   */
    double a_on, a_off;    /* angle switch slots after the caller-side swap      */
    double inv_denom_r;    /* 1 / (off2^2 - on2^2)^3 of the radial switch          */
};

/* Row buffers produced by the fused kernel, consumed by the ordered copy: per row
 * row_cap entries of 4 32-bit words {j, length bits, cosine bits, unused}. */
#define WN_ROW_WORDS 4

/* ---- fused-path launch parameters ---- */
struct WnEdgeParams {
    const float4*   spos;       /* [F][N] sorted: x,y,z, site index bits          */
    const uint32_t* smeta;      /* [F][N] sorted meta                             */
    const double*   sdh;        /* [F][N][hs][4] sorted unit D-H vectors (x,y,z,pad) */
    const int*      cell_start; /* [F][cell_cap+1]                                */
    const WnGrid*   grid;       /* [F]                                            */
    uint32_t*       rowbuf;     /* [F][N][row_cap][4] edges of row i, arrival order */
    uint32_t*       row_cnt;    /* [F][N] edges owned by the row (zeroed before)  */
    WnFrameCounters* counters;  /* [F]                                            */
    int n_sites;
    int cell_cap;
    int hs;                     /* valid-hydrogen slots per site                  */
    int first_limit;
    int row_cap;
    int simple;                 /* 1: every site WATER with the list [site atom, H] */
    int pbc;                    /* 0 open, 1 rectangular, 2 triclinic (all-pairs kernel only) */
    int bx;                     /* home block width in cells (1 or 2)               */
    WnSubset subset;            /* per-frame site subsets (off == nullptr: whole table) */
    int*            blk_prefix; /* [F+1] exclusive prefix of the home blocks per frame (written by the launcher's kernel) */
    unsigned int*   work_next;  /* next unclaimed home block of the pass (dynamic scheduling)  */
    uint2*          blk_magic;  /* [F] ceil(2^32 / nbx), ceil(2^32 / ny): block index -> (bx, cy, cz) without divisions */
    int n_frames;
};

/* ---- kernel launchers (wn_kernels_*.cu).  All asynchronous on `st`. ---- */
void wn_launch_bbox(const float* frames, int n_frames, int natoms, const int* site_atom,
                    int n_sites, WnSubset sub, int cell_cap, WnGrid* grid, uint32_t* scratch /* [F][6] */, cudaStream_t st);
int wn_scan_chunks_for(int len);    /* chunks per frame of the three-phase scans */
void wn_launch_bin(const float* frames, int n_frames, int natoms, const int* site_atom,
                   int n_sites, WnSubset sub, int cell_cap, const WnGrid* grid, int* cell_cnt,
                   int* site_cell, int* site_rank, cudaStream_t st);
/* periodic grid of one frame, built on the host (uploaded with the pass) */
void wn_host_pbc_grid(const float box[3], int cell_cap, WnGrid* g);
/* triclinic frame (box9 = GROMACS matrix, row-major): cells on the brick [0,ax) x [0,by) x [0,cz) */
void wn_host_tric_grid(const float box9[9], int cell_cap, WnGrid* g);
void wn_launch_cellscan(int n_frames, int n_sites, WnSubset sub, int cell_cap, const WnGrid* grid,
                        int* cell_cnt, uint32_t* part /* [F][chunks(cell_cap)] */, cudaStream_t st);
void wn_launch_scatter(const float* frames, int n_frames, int natoms, const int* site_atom,
                       const int* hv_atom, const uint32_t* meta, int n_sites, WnSubset sub, int hs,
                       int cell_cap, const int* cell_start, const int* site_cell,
                       const int* site_rank, const WnGrid* grid, float4* spos, uint32_t* smeta, double* sdh,
                       cudaStream_t st);
void wn_launch_edges(const WnEdgeParams& p, int n_frames, cudaStream_t st);
/* all-pairs variant for periodic boxes with fewer than 3 cells in some dimension */
void wn_launch_edges_small(const WnEdgeParams& p, int n_frames, cudaStream_t st);
void wn_launch_rowscan(const uint32_t* row_cnt, int n_frames, int n_sites,
                       uint32_t* row_off, unsigned long long* frame_total, uint32_t* frame_maxrow,
                       uint32_t* part /* [2][F][chunks(n_sites)] */, cudaStream_t st);
/* chain[0] (device): edges of the call before this pass, advanced by the pass's total;
 * pass_info[0] = the new chain[0], pass_info[1] = longest row of the pass */
void wn_launch_frameoff(const unsigned long long* frame_total, const uint32_t* frame_maxrow, int n_frames,
                        unsigned long long* chain, unsigned long long* frame_off,
                        unsigned long long* pass_info, cudaStream_t st);
void wn_launch_copy(const uint32_t* rowbuf, int row_cap, const uint32_t* row_cnt,
                    const uint32_t* row_off, const unsigned long long* frame_off, const WnFrameCounters* counters,
                    int n_frames, int n_sites, WnEnergyParams ep, wn_edge* edges,
                    unsigned long long edge_cap, int layout, float* energy_side, double* block_sums, int* n_blocks_out,
                    cudaStream_t st);
void wn_launch_stats(const double* block_sums, int n_blocks, const unsigned long long* frame_total,
                     const WnFrameCounters* counters, int n_frames, int n_sites, WnSubset sub,
                     wn_frame_stats* stats, const wn_edge* edges, int layout, const float* energy_side,
                     const uint32_t* row_off /* of the pass */, const unsigned long long* frame_off /* of the pass */,
                     cudaStream_t st);
size_t wn_edges_smem_bytes(int hs);
void wn_launch_selftest_div(unsigned long long seed, int blocks, int iters, unsigned long long* mismatches,
                            cudaStream_t st);

/* text rows of the edge file (wn_kernels_format.cu) */
void wn_launch_components(const void* edges, int csr16, const unsigned long long* frame_off, const uint32_t* row_off,
                          int n_frames, int n_sites, unsigned long long max_edges, float threshold,
                          int* labels, int* work, int* count, int* minidx, unsigned int* kept,
                          wn_component_stats* stats, cudaStream_t st);
void wn_launch_format_rows(const wn_edge* edges, const unsigned long long* frame_off,
                           const unsigned long long* row_base, const unsigned int* site_ids, int n_frames,
                           unsigned long long max_edges_per_frame, char* text, int* frame_flags, cudaStream_t st);

/* brute pair path (wn_kernels_pairs.cu) */
struct WnPairBox { int pbc; /* 0 open, 1 rectangular, 2 triclinic */ double L[3], invL[3], tri[3]; float Lf[3], invLf[3], trif[3]; };
void wn_launch_pairs_count(const double* xyz, int n, int first_limit, WnPairBox box, uint32_t* row_cnt,
                           unsigned long long* ties, cudaStream_t st);
void wn_launch_pairs_scan(const uint32_t* row_cnt, int n, unsigned long long* row_off,
                          cudaStream_t st);
int wn_pairs_segments(int n);       /* column segments per row: counts/offsets are [n][segments] */
void wn_launch_pairs_fill(const double* xyz, int n, int first_limit, WnPairBox box,
                          const unsigned long long* row_off, int32_t* pairs, cudaStream_t st);

/* cell-list pair path (wn_kernels_cellpairs.cu): CudaFindPairs::find in O(N) */
struct WnCellPairGrid {
    double o[3];         /* grid origin (open: minimum corner of the finite sites; periodic: 0) */
    double inv[3];       /* 1 / cell edge                                                        */
    double L[3], invL[3];/* periodic box                                                         */
    int n[3];            /* cells per dimension                                                  */
    int ncells;
    int pbc;             /* 0 open, 1 rectangular periodic                                       */
    float Lf[3], invLf[3];  /* fp32 copies for the reject test                                   */
    float bound_f;       /* reject when the fp32 squared distance exceeds this (superset of 10*d2 < 42.25) */
};
int wn_cellpairs_supported(int n);     /* the per-warp bitmap over j fits in shared memory */
void wn_launch_cellpairs_bin(const double* xyz, int n, const WnCellPairGrid& g, int* cell_cnt, int* site_cell,
                             int* site_rank, double4* srec, float4* srecf, cudaStream_t st);
void wn_launch_cellpairs_count(const double4* srec, const float4* srecf, int n, int first_limit, const WnCellPairGrid& g,
                               const int* cell_start, const int* site_cell, uint32_t* row_cnt, unsigned long long* ties,
                               cudaStream_t st);
void wn_launch_cellpairs_fill(const double4* srec, const float4* srecf, int n, int first_limit, const WnCellPairGrid& g,
                              const int* cell_start, const int* site_cell, const unsigned long long* row_off, int32_t* pairs,
                              cudaStream_t st);
/* exclusive 64-bit scan of `len` 32-bit counts (+ total at [len]) */
void wn_launch_scan64(const uint32_t* cnt, int len, unsigned long long* off, cudaStream_t st);

/* make_edges over a caller-supplied pair list (wn_kernels_pairlist.cu) */
void wn_launch_pl_iota(int* rank, int n, cudaStream_t st);
int wn_pl_blocks(unsigned long long m);
void wn_launch_pl_eval(const float4* spos, const uint32_t* smeta, const double* sdh, int hs, int simple, int pbc,
                       const WnGrid* grid, const int2* pairs, unsigned long long m, int n_sites, float2* tmp,
                       uint32_t* blk_cnt, WnFrameCounters* counters, unsigned int* bad_index, cudaStream_t st);
void wn_launch_pl_write(const int2* pairs, const float2* tmp, unsigned long long m, const unsigned long long* blk_off,
                        WnEnergyParams ep, wn_edge* edges, unsigned long long edge_cap, cudaStream_t st);

/* synthetic frames (wn_synth.cu, compiled with -fmad=false) */
void wn_launch_synth(int n_waters, uint64_t seed, int model, int64_t frame0, int n_frames,
                     float* frames, cudaStream_t st);

#endif
