/*
 * wn_b200.h -- C ABI of the B200-native hydrogen-bond pair-search + energy path.
 *
 * This is the drop-in boundary.  The C++ host classes under include/ (.hpp)
 * (CudaFindPairs, HydrogenBondEdges, FrameBatcher) are thin wrappers over the
 * entry points below; a maintainer of the reference binds the same entry points
 * from WaterNetwork::analyzeFrame.  Plain pointers and
 * sizes only; no C++ or torch types; no exception crosses this boundary.
 *
 * What each entry point replaces in the reference (/root/reference):
 *   wn_find_pairs          BruteFindPairs::find            src/brute-find-pairs.cpp:7-32
 *                          behind FindPairs::find          include/find-pairs.hpp:11
 *   wn_set_sites           site gather tables              src/waternetwork.cpp:538-577
 *   wn_set_energy_params   setRadiusShift / setAngleShift  include/calculate-hydrogen-bond-energy.hpp:9-10
 *                          (defaults = computeEnergy's     src/waternetwork.cpp:51-54)
 *   wn_hbond_batch*        fromGmxtoCgalPosition + gather  src/waternetwork.cpp:68-81,538-577
 *                          + mp_->find + make_edges        src/waternetwork.cpp:583,182-263
 *                          + computeEnergy                 src/waternetwork.cpp:12-66
 *                          + per-frame stats rows          src/waternetwork.cpp:654-666
 *                          for MANY frames per call (frame batching of analyzeFrame)
 *   wn_stats_reduce        (new) NCCL sum of per-frame stats across frame shards
 *
 * There is no CPU fallback: every compute entry point runs CUDA kernels built for
 * sm_100a and returns a non-zero status if the device or the kernels are unusable.
 */
#ifndef WN_B200_H
#define WN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------- */
#define WN_OK               0
#define WN_ERR_CUDA         1   /* a CUDA runtime call or kernel failed            */
#define WN_ERR_OOM          2   /* device or pinned-host allocation failed         */
#define WN_ERR_BAD_ARG      3   /* null pointer, negative size, inconsistent table */
#define WN_ERR_NO_SITES     4   /* wn_hbond_batch* before wn_set_sites             */
#define WN_ERR_OVERFLOW     5   /* index space exceeded (pairs/edges > limits)     */
#define WN_ERR_NCCL         6   /* NCCL call failed / communicator not initialised */
#define WN_ERR_NO_DEVICE    7   /* no CUDA device / wrong architecture             */

/* include/utils.hpp:7  enum SiteType {DONOR, ACCEPTOR, WATER} */
#define WN_DONOR    0
#define WN_ACCEPTOR 1
#define WN_WATER    2

/* Most hydrogens per site whose position differs from the site atom itself. */
#define WN_MAX_VALID_H 8

typedef struct wn_ctx wn_ctx;

/* Compact HydrogenBond (include/utils.hpp:25-32).  The reference's three doubles
 * each hold a float-rounded value (src/waternetwork.cpp:195-197,211-212), so the
 * float fields are lossless.  i<j are 0-based site indices; direction is always
 * FORWARD in the reference (src/waternetwork.cpp:255) and is not stored. */
typedef struct wn_edge {
    int32_t i, j;
    float   energy;   /* energies.at(pos)            */
    float   length;   /* O..O distance, Angstrom     */
    float   cosine;   /* coss.at(pos)                */
} wn_edge;

/* Compact layout (WN_LAYOUT_CSR16): the same list without the repeated row index; i is given by
 * row_offsets (CSR).  16 bytes per edge instead of 20: fewer bytes over PCIe, aligned stores. */
typedef struct wn_edge16 {
    int32_t j;
    float   energy, length, cosine;
} wn_edge16;
/* Smallest layout (WN_LAYOUT_CSR12): {j, length, cosine}, rows via row_offsets.  The energy of an edge is a
 * pure function of its (length, cosine) -- computeEnergy(distance, proj), src/waternetwork.cpp:49-66,222 -- so a
 * host that needs it recomputes it on demand (include/wn_energy.h: wn_edge_energy, the reference's own
 * formula); 12 instead of 16/20 bytes per edge cross PCIe.  stats[f].sum_energy is still filled on the device. */
typedef struct wn_edge12 {
    int32_t j;
    float   length, cosine;
} wn_edge12;
#define WN_LAYOUT_EDGES 0   /* wn_edge[]   : {i, j, energy, length, cosine}                      */
#define WN_LAYOUT_CSR16 1   /* wn_edge16[] : {j, energy, length, cosine}, rows via row_offsets   */
#define WN_LAYOUT_CSR12 2   /* wn_edge12[] : {j, length, cosine}, rows via row_offsets           */

/* One row per frame.  First three = columns 0..2 of hydrogen-bonds-num.xvg
 * (src/waternetwork.cpp:659-662); sum_energy and pair_evals are extras.
 * sum_energy: the sum of the frame's float edge energies, independent of scheduling and of how frames are batched
 * (bit-identical run to run).  Accuracy: <= 1.2e-7 relative to the fp64 sum of those floats for every frame -- an
 * exact 2^-32 fixed-point sum where 2^-33 per edge is below 1e-7 of the total, else (and for frames with an edge
 * under 0.56 A) the written records summed in final order. */
typedef struct wn_frame_stats {
    double num_sites;
    double num_edges;
    double ratio;        /* 1.0*num_edges/num_sites                          */
    double sum_energy;   /* sum over edges of (double)energy, fp64            */
    double pair_evals;   /* executions of make_edges' `if (isedge)` body      */
} wn_frame_stats;

/* Exact-comparison ties seen by the kernels (reported next to parity claims). */
typedef struct wn_tie_report {
    uint64_t pair_cutoff_equal;    /* 10*d2 == 42.25 in the pair test          */
    uint64_t length_equal_cut;     /* (float)dist == 6.5f in the edge test     */
    uint64_t cosine_argmax_equal;  /* equal positive cosines in the arg-max    */
    uint64_t pos_zero;             /* arg-max on list index 0 (edge dropped)   */
    uint64_t pair_cutoff_near;     /* |10*d2 - 42.25| <= 4 ulp(42.25) (incl. ==) */
} wn_tie_report;

/* Result of one batch.  Pointers are library-owned and stay valid until the next
 * wn_hbond_batch* call on the same context or wn_destroy. */
typedef struct wn_batch_result {
    const wn_edge*        edges;          /* concatenated, frame by frame; wn_edge16* / wn_edge12* for the compact layouts */
    const uint64_t*       frame_offsets;  /* n_frames+1 offsets into edges       */
    const wn_frame_stats* stats;          /* n_frames rows                       */
    uint64_t              n_edges;        /* == frame_offsets[n_frames]          */
    uint64_t              n_pair_evals;   /* sum over frames                     */
    wn_tie_report         ties;           /* sum over frames                     */
    int32_t               n_frames;
    int32_t               on_device;      /* 1: edges/offsets/stats are device pointers */
    /* CSR view of the same list (hand-off to graph analytics, e.g. connected components of
     * the buried-water network): edges of site i in frame f are
     * edges[frame_offsets[f] + row_offsets[f*n_sites + i] .. + row_offsets[f*n_sites + i + 1])
     * (the last row of a frame ends at frame_offsets[f+1]); [n_frames][n_sites] uint32. */
    const uint32_t*       row_offsets;
    int32_t               n_sites;
    int32_t               layout;         /* WN_LAYOUT_EDGES or WN_LAYOUT_CSR16 */
    uint64_t              n_candidate_tests; /* (site, candidate) pairs the search examined, sum over frames:
                                              * implementation-dependent, reported next to n_pair_evals */
} wn_batch_result;

/* ---- context ------------------------------------------------------------ */
int  wn_create(int device, wn_ctx** out);
void wn_destroy(wn_ctx* ctx);
/* Text of the last error on this context ("" if none).  ctx may be NULL: then
 * the last error of a failed wn_create on this thread. */
const char* wn_last_error(const wn_ctx* ctx);
/* Library identification: "wn_b200 <version> sm_100a". */
const char* wn_version(void);
/* Number of CUDA devices visible to the process (0 when there is none or the driver is unusable). */
int wn_device_count(void);

/* ---- FindPairs::find (include/find-pairs.hpp:11) -------------------------
 * Semantics of BruteFindPairs::find (src/brute-find-pairs.cpp:19-30): all i<j
 * with 10.0*squared_distance < 42.25 (fp64, no FMA), lexicographic order.
 * xyz: host AoS double[n][3] (std::vector<Point> storage).  *pairs: library-owned
 * pinned host int32[2*m], valid until the next call.  first_limit restricts rows
 * to i < first_limit (pass n, or any value >= n, for the reference behaviour). */
int wn_find_pairs(wn_ctx* ctx, const double* xyz, int32_t n, int32_t first_limit,
                  const int32_t** pairs, uint64_t* m, wn_tie_report* ties);

/* ---- Periodic boundaries (north-star extension) ---------------------------
 * The reference has no PBC: analyzeFrame receives t_pbc* and never touches it
 * (src/waternetwork.cpp:503), BruteFindPairs uses plain Euclidean distances
 * (src/brute-find-pairs.cpp:25).  The periodic mode is therefore DEFINED here (and restated
 * by the test oracle).  Rectangular box L = (Lx,Ly,Lz), float32 as in a GROMACS frame:
 *   1. wn_hbond_batch_pbc*: every coordinate is first put in the box in fp64 and rounded
 *      back to float, pw = (float)(x - floor(x * (1.0/L)) * L)   (identity for 0 <= x < L);
 *      wn_find_pairs_pbc takes the caller's points as they are.
 *   2. every displacement the path forms (site - site, hydrogen - site) is the minimum
 *      image of the fp64 difference, d = pa - pb; d = d - rint(d * (1.0/L)) * L
 *      (every operation rounded, rint half-to-even: +-L/2 is kept, d(b,a) == -d(a,b)).
 *   3. everything else is unchanged.
 * Triclinic boxes (the *_box9 entry points; box9 = GROMACS' matrix, row-major: a = (ax,0,0),
 * b = (bx,by,0), c = (cx,cy,cz)) reduce along c, then b, then a, every operation rounded in fp64:
 *   put in box:  fz = floor(z/cz): (x,y,z) -= fz*c;  fy = floor(y/by): (x,y) -= fy*(bx,by);
 *                fx = floor(x/ax): x -= fx*ax;  result rounded to float
 *   min image:   k = rint(d.z/cz): d -= k*c;  j = rint(d.y/by): d -= j*b;  i = rint(d.x/ax): d.x -= i*ax
 * (the shortest image whenever the cutoff is small against the box, as GROMACS assumes).  A call
 * whose frames are all rectangular uses the rectangular definition above; one with any
 * triclinic frame uses the triclinic formulas for every frame.  Both run the cell-list sweep
 * (boxes with fewer than 3-4 cells of 0.65 nm in a dimension use an all-pairs kernel).
 *
 * How this relates to GROMACS' own rectangular rule (pbc_dx_aiuc: `dx = xi - xj` in FLOAT,
 * `while (dx > L/2) dx -= L; while (dx <= -L/2) dx += L`).  The two are the same geometry, not the
 * same bits -- deliberately:
 *   - here the difference of two float coordinates is formed in fp64 (exact) and only the image shift rounds, so a
 *     pair that does not cross the box is evaluated with EXACTLY the reference's open-boundary arithmetic (a box much
 *     larger than the system reproduces the non-periodic edge list bit for bit; tested).  The float rule rounds every
 *     difference to 24 bits first: d^2 of a good part of the pairs moves by ~1e-7 relative, lengths and cosines by a
 *     few float ulps, and a pair within ~1e-5 A of the 6.5 A cut (or a cosine within ~1e-6 of 0) may change sides;
 *   - exactly at +-L/2 (only reachable when a box edge is under twice the cut, 1.3 nm) this definition keeps the
 *     sign, d(b,a) == -d(a,b), so a pair is the same question in both orientations; the float rule maps both to +L/2;
 *   - boxes so small that a pair is within the cut through two images: both rules keep one image, one edge.
 * tests/test_oracle_cpu.py::test_pbc_contract_against_the_gromacs_float_minimum_image measures all three. */
int wn_find_pairs_pbc(wn_ctx* ctx, const double* xyz, int32_t n, int32_t first_limit,
                      const float box[3], const int32_t** pairs, uint64_t* m, wn_tie_report* ties);

int wn_find_pairs_box9(wn_ctx* ctx, const double* xyz, int32_t n, int32_t first_limit,
                       const float box9[9], const int32_t** pairs, uint64_t* m, wn_tie_report* ties);

/* How wn_find_pairs* searches.  WN_FIND_AUTO (default): the cell-list sweep (cells of half the 2.0555 nm
 * cutoff, 5x5x5 neighbourhood, exact fp64 test on every candidate, O(N)) for n >= 1024 points on open and
 * rectangular periodic boxes of at least 5 cells per dimension; the exact all-pairs tile kernel otherwise
 * (triclinic boxes, small periodic boxes, tiny inputs).  Both give the same pair list bit for bit; the other two
 * modes force one of them where it applies (tests).  wn_last_find_used_cells: which one the last call ran. */
#define WN_FIND_AUTO     0
#define WN_FIND_ALLPAIRS 1
#define WN_FIND_CELLS    2
int wn_set_find_mode(wn_ctx* ctx, int32_t mode);
int wn_last_find_used_cells(const wn_ctx* ctx);

/* ---- make_edges on a caller-supplied pair list (src/waternetwork.cpp:182-263) --------------
 * The reference calls make_edges on whatever list its pair finder returned.  The live module's finder is
 * DelaunayFindPairs (src/waternetwork.cpp:399): Delaunay edges in finite_edges order whose (first, second)
 * orientation is not i < j.  The orientation decides the sign of OO (:194), which site's hydrogen list is
 * scanned first in the arg-max (:215-249), the `pos > 0` rule (:251) and which site is printed first -- so the
 * batched path above (all i < j pairs within 0.65 nm, lexicographic) equals Brute + make_edges, not the live
 * Delaunay output.  This entry point closes that gap: the host keeps its own finder (Delaunay stays on the host,
 * out of scope here) and hands the list over; every pair is evaluated in the GIVEN order and orientation, and
 * the pairs that become edges come back in that order with i = first, j = second.
 * frame: host float[natoms][3]; box9: NULL (reference behaviour: plain Euclidean) or a GROMACS box matrix
 * (periodic extension as defined above); pairs: host int32[m][2] indexing the site table of wn_set_sites
 * (an index outside the table -> WN_ERR_BAD_ARG, where the reference's .at() throws).  first_limit is ignored.
 * *edges: library-owned pinned wn_edge[*n_edges], valid until the next call; afterwards the context holds no
 * "last batch" (wn_format_edges / wn_components need a wn_hbond_batch* call first). */
int wn_make_edges(wn_ctx* ctx, const float* frame, int32_t natoms, const float* box9, const int32_t* pairs, uint64_t m,
                  const wn_edge** edges, uint64_t* n_edges, uint64_t* n_pair_evals, wn_tie_report* ties);

/* ---- site tables (src/waternetwork.cpp:538-577) --------------------------
 * site_atom[n_sites]: atom index (into a frame) of each site.
 * h_atom[n_sites*hmax]: atom indices of its hydrogen list in list order, -1 padded.
 * site_type[n_sites]: WN_DONOR/WN_ACCEPTOR/WN_WATER.
 * first_limit: only rows i < first_limit own edges (asymmetric search); >= n_sites
 * for the reference behaviour.  The table is copied; call again to change it. */
int wn_set_sites(wn_ctx* ctx, int32_t n_sites, const int32_t* site_atom,
                 const int32_t* h_atom, int32_t hmax, const uint8_t* site_type,
                 int32_t first_limit);
/* The reference's water table: site w = atom 3w, hydrogens [3w, 3w+1]
 * (src/waternetwork.cpp:564-575), all WN_WATER. */
int wn_set_water_sites(wn_ctx* ctx, int32_t n_waters, int32_t first_limit);

/* Radial / angular switch parameters of computeEnergy.  Defaults (5.5, 6.5,
 * 0.25, 0.0) are what every live call in the reference uses. */
int wn_set_energy_params(wn_ctx* ctx, double r_on, double r_off,
                         double theta_on, double theta_off);

/* ---- batched frames ------------------------------------------------------
 * frames: float32 AoS [n_frames][natoms][3] (rvec layout, GMX_DOUBLE=OFF).
 * wn_hbond_batch      : host input, results copied to library-owned pinned host
 *                       memory (H2D + kernels + D2H).
 * wn_hbond_batch_dev  : device input, results left in device memory.
 * Edge order within a frame = the reference's (lexicographic i<j). */
int wn_hbond_batch(wn_ctx* ctx, const float* frames_host, int32_t n_frames,
                   int32_t natoms, wn_batch_result* out);
int wn_hbond_batch_dev(wn_ctx* ctx, const float* frames_dev, int32_t n_frames,
                       int32_t natoms, wn_batch_result* out);
/* Periodic variants: boxes = host float32 [n_frames][3] = (Lx,Ly,Lz) of every frame (the
 * diagonal of GROMACS' t_trxframe::box; rectangular boxes only).  Coordinates are first put
 * in the box, every displacement is the fp64 minimum image ("Periodic boundaries" above). */
int wn_hbond_batch_pbc(wn_ctx* ctx, const float* frames_host, const float* boxes_host,
                       int32_t n_frames, int32_t natoms, wn_batch_result* out);
int wn_hbond_batch_pbc_dev(wn_ctx* ctx, const float* frames_dev, const float* boxes_host,
                           int32_t n_frames, int32_t natoms, wn_batch_result* out);
/* General boxes: boxes9_host = float32 [n_frames][9], GROMACS' t_trxframe::box row-major. */
int wn_hbond_batch_box9(wn_ctx* ctx, const float* frames_host, const float* boxes9_host,
                        int32_t n_frames, int32_t natoms, wn_batch_result* out);
/* Per-frame site subsets: the buried waters the reference's alpha-shape filter keeps change
 * every frame (as_->locate, src/waternetwork.cpp:529-533; the node loop :562-576 then builds
 * sitePoints from that frame's list).  Frame f uses the rows
 * subset_sites[subset_offsets[f] .. subset_offsets[f+1]) of the site table, in that order; the
 * edge indices i<j of the frame are positions in that list (= indices into the frame's
 * sitePoints).  At most n_sites (table size) entries per frame.  boxes_host == NULL: open
 * boundary (the reference); otherwise [n_frames][3] as in wn_hbond_batch_pbc.
 * Result: row_offsets has stride n_sites; stats[f].num_sites is the frame's list length. */
int wn_hbond_batch_subset(wn_ctx* ctx, const float* frames_host, const float* boxes_host,
                          int32_t n_frames, int32_t natoms, const int64_t* subset_offsets,
                          const int32_t* subset_sites, wn_batch_result* out);
/* The general form: every combination of the options above in one call.  The other
 * wn_hbond_batch* entry points are shorthands for it. */
typedef struct wn_batch_args {
    const float*   frames;          /* [n_frames][natoms][3] float32                               */
    int32_t        frames_on_device;/* 0: host memory (results to pinned host); 1: device memory    */
    int32_t        n_frames;
    int32_t        natoms;
    int32_t        box_stride;      /* 0: open boundary; 3: boxes = (Lx,Ly,Lz); 9: GROMACS matrices  */
    const float*   boxes;           /* host, [n_frames][box_stride]; NULL iff box_stride == 0        */
    const int64_t* subset_offsets;  /* host, [n_frames+1], or NULL: every frame uses the whole table */
    const int32_t* subset_sites;    /* host, rows of the site table                                  */
    int32_t        result_layout;   /* WN_LAYOUT_EDGES (default), WN_LAYOUT_CSR16 or WN_LAYOUT_CSR12  */
    int32_t        results_on_device; /* 1: host frames in, but edges and row offsets stay in device memory (as for a
                                     * *_dev batch) for consumers that continue on the GPU: wn_components,
                                     * wn_format_edges; only the small per-frame tables come back.  0: default */
} wn_batch_args;
int wn_hbond_batch_ex(wn_ctx* ctx, const wn_batch_args* args, wn_batch_result* out);

/* Frames that one wn_hbond_batch* call processes per kernel pipeline pass
 * (larger inputs are split internally).  0 restores the default heuristic. */
int wn_set_batch_frames(wn_ctx* ctx, int32_t frames_per_pass);

/* Host result buffers of a context: 1 (default) -- the pointers of a batch result stay valid until the next
 * wn_hbond_batch* call; 2 -- consecutive calls alternate between two sets, so a result stays valid until the
 * call AFTER the next one (the caller can consume batch k on one thread while batch k+1 runs on another;
 * FrameBatcher does exactly that).  wn_format_edges / wn_components always refer to the latest call. */
int wn_set_result_sets(wn_ctx* ctx, int32_t n_sets);

/* Pin the CALLING thread to the CPUs next to the context's GPU (sysfs local_cpulist of the PCI device), so that
 * pinned buffers allocated afterwards (first touch) and the copies' host side sit on the GPU's NUMA node.
 * *n_cpus (may be NULL): CPUs in the resulting set, 0 when the platform exposes no locality (nothing changed). */
int wn_bind_thread_to_device(wn_ctx* ctx, int32_t* n_cpus);

/* Plain pinned-memory copy bandwidth of this context's GPU (GB/s, `repeats` back-to-back copies of `bytes`):
 * the ceiling of the host-buffer entry points; run it on all ranks at once to see what shared uplinks leave. */
int wn_pcie_probe(wn_ctx* ctx, size_t bytes, int32_t repeats, double* h2d_gbs, double* d2h_gbs);
/* The same with the traffic mix of a batch call: per repeat `d2h_bytes` device-to-host and `h2d_bytes` host-to-device
 * AT THE SAME TIME (two copy streams), *seconds = time per repeat until both directions are done.  Called on all ranks
 * at once with a step's volumes, it is the time no host-buffer batch call can beat on this machine. */
int wn_pcie_probe_mix(wn_ctx* ctx, size_t d2h_bytes, size_t h2d_bytes, int32_t repeats, double* seconds);

/* ---- edge file text (src/waternetwork.cpp:638-652), formatted on the GPU ---------------
 * Formats the edge list of the LAST wn_hbond_batch* call of this context as the rows of
 * edges-frames.xvg.dat: per frame "#--- frame N ---\n" (N = frame_numbers[f]) and one 37-byte
 * row per edge, "%6u%6u%8.4f%8.4f%8.4f\n" with id = site_ids[index] (NULL: the index itself;
 * the reference prints SiteInfo::id).  Byte-identical to the reference's iostream output.
 * *text: library-owned pinned host memory, valid until the next call on this context.
 * Returns WN_ERR_OVERFLOW (nothing usable in *text) if some id needs more than 6 digits or a
 * value more than 8 characters: format that batch with a general host formatter instead.
 * After a batch with per-frame site lists the edge indices are list positions, which a table of
 * ids cannot translate on the device: site_ids must be NULL then (WN_ERR_BAD_ARG otherwise). */
/* ---- connected components of the per-frame networks (python/network-analysis.py:141-171) ----
 * Works on the edge list of the LAST wn_hbond_batch* call of this context (either layout), on the
 * device.  Per frame, as the reference's downstream script does with the edge file: an edge is
 * dropped iff |energy| <= weight_threshold (filter_edge_by_weight, :16-24; the script uses 0.0 on
 * the 4-decimal text, i.e. 5e-5 on the binary value), sites left without an edge are dropped
 * (remove_node_degree, :26-35), the rest is split into connected components
 * (nx.node_connected_component, :155).
 * labels[f * n_sites + s]: smallest site index of the component of site s in frame f, or -1 if s has
 * no kept edge.  After a batch with per-frame site lists (wn_hbond_batch_subset) s and the labels are
 * positions in that frame's list, as the edge indices are; entries past the list's length are -1.  stats[f]: see wn_component_stats.  Both are library-owned and valid until the next
 * call on this context: pinned host memory when the batch came from host frames, device memory when
 * it was a *_dev batch (like the batch result itself). */
typedef struct wn_component_stats {
    int32_t n_components;   /* components with at least one edge                 */
    int32_t largest;        /* sites in the largest component                    */
    int32_t n_nodes;        /* sites with at least one kept edge                 */
    int32_t n_edges;        /* kept edges                                        */
} wn_component_stats;
int wn_components(wn_ctx* ctx, float weight_threshold, const int32_t** labels, const wn_component_stats** stats);

int wn_format_edges(wn_ctx* ctx, const int32_t* frame_numbers, const uint32_t* site_ids,
                    int32_t n_site_ids, const char** text, uint64_t* text_bytes);

/* ---- device buffers + synthetic frames (bench/test plumbing) ------------- */
int wn_dev_alloc(wn_ctx* ctx, size_t bytes, void** dev_ptr);
int wn_dev_free(wn_ctx* ctx, void* dev_ptr);
int wn_host_alloc_pinned(wn_ctx* ctx, size_t bytes, void** host_ptr);
int wn_host_free_pinned(wn_ctx* ctx, void* host_ptr);
int wn_memcpy_d2h(wn_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int wn_memcpy_h2d(wn_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
/* Frames [frame0, frame0+n_frames) of the synthetic water box of include/wn_synth.h
 * generated on the device, bit-identical to the host generator. */
int wn_synth_frames_dev(wn_ctx* ctx, int32_t n_waters, uint64_t seed, int32_t model,
                        int64_t frame0, int32_t n_frames, float* frames_dev);

/* ---- timing of the last wn_hbond_batch* call ----------------------------- */
typedef struct wn_timings {
    float total_ms;        /* CUDA events around the whole call's device work (wn_hbond_batch*:
                            * sum of the stages; wn_find_pairs*: count + scan + fill kernels) */
    float bin_ms;          /* bounding box + cell binning + staging kernels     */
    float edges_ms;        /* fused search + geometry kernel                    */
    float finish_ms;       /* row scan + ordered copy/energy + stats kernels    */
    float h2d_ms, d2h_ms;  /* host variant only                                 */
    int32_t launches;      /* kernels launched by the call                      */
    int32_t passes;        /* internal passes (sub-batches)                     */
} wn_timings;
int wn_get_timings(wn_ctx* ctx, wn_timings* out);
/* A CUDA-event stopwatch on the context's stream (the stream every kernel of this
 * context is launched on): start records an event, stop records a second one,
 * synchronises and returns the elapsed device time between them. */
int wn_timer_start(wn_ctx* ctx);
int wn_timer_stop(wn_ctx* ctx, float* elapsed_ms);

/* Device self-test: the fused kernel divides the O..O vector by its float-rounded length
 * with one shared reciprocal and an FMA correction per component (exact for a divisor with
 * <= 24 significant bits, see wn_kernels_edges.cu).  This compares that routine with the
 * IEEE division bit for bit on n_samples triples (kernel-like, arbitrary and edge-pattern
 * operands) and returns the number of differing triples (must be 0). */
int wn_selftest_div(wn_ctx* ctx, uint64_t n_samples, uint64_t seed, uint64_t* mismatches);

/* ---- multi-GPU: per-frame statistics reduction (NCCL) ---------------------
 * Frames are sharded across ranks with no data-path exchange; the only
 * collective is the sum of a zero-padded [n_frames_total][5] fp64 stats array.
 * id: 128 bytes from wn_nccl_unique_id on rank 0, distributed by the caller. */
int wn_nccl_unique_id(void* id128);
int wn_nccl_init(wn_ctx* ctx, const void* id128, int32_t rank, int32_t nranks);
/* stats_host: [n_frames_total] rows, rows not owned by this rank zero.
 * In-place sum over ranks (all-reduce); every rank gets the full table. */
int wn_stats_reduce(wn_ctx* ctx, wn_frame_stats* stats_host, int64_t n_frames_total);

#ifdef __cplusplus
}
#endif
#endif /* WN_B200_H */
