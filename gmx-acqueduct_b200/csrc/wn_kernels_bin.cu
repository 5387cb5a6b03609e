/*
 * wn_kernels_bin.cu -- frame staging for the fused hydrogen-bond path (sm_100a).
 *
 * Replaces, for a whole batch of frames per launch, the per-frame host plumbing of
 * WaterNetwork::analyzeFrame (reference src/waternetwork.cpp):
 *   fromGmxtoCgalPosition  :68-81    float rvec -> double Point (exact widening)
 *   site gather            :538-577  sitePoints / hydrogenPoints / siteInfos
 * and adds what an O(N) search needs: a per-frame cell grid (bounding box of the
 * sites, open boundary -- BruteFindPairs has no PBC, src/brute-find-pairs.cpp:25),
 * a counting (single-digit radix) sort of the sites by cell, and the per-site unit
 * donor-hydrogen vectors DH/|DH| that make_edges recomputes for every pair
 * (:219-220,:231-232) -- they depend on the site only, so they are computed once
 * per site per frame here, with the reference's exact fp64 operation order.
 */
#include "wn_internal.cuh"

#include <math.h>

namespace {

/* periodic mode, step 1 of the definition (include/wn_b200.h, "Periodic boundaries"): put the coordinate in the
 * box in fp64 and round back to float; identity for 0 <= x < L */
__device__ __forceinline__ float wn_wrap(float x, double L, double invL)
{
    const double xd = (double)x;
    return __double2float_rn(__dsub_rn(xd, __dmul_rn(floor(__dmul_rn(xd, invL)), L)));
}

/* triclinic: put the point in the brick [0,ax) x [0,by) x [0,cz) -- a fundamental domain of the
 * lower-triangular lattice -- reducing along c, then b, then a (GROMACS order); fp64, then float */
__device__ __forceinline__ void wn_wrap_tric(float& xf, float& yf, float& zf, const double* L, const double* tri)
{
    double x = (double)xf, y = (double)yf, z = (double)zf;
    const double fz = floor(__ddiv_rn(z, L[2]));
    x = __dsub_rn(x, __dmul_rn(fz, tri[1])); y = __dsub_rn(y, __dmul_rn(fz, tri[2])); z = __dsub_rn(z, __dmul_rn(fz, L[2]));
    const double fy = floor(__ddiv_rn(y, L[1]));
    x = __dsub_rn(x, __dmul_rn(fy, tri[0])); y = __dsub_rn(y, __dmul_rn(fy, L[1]));
    const double fx = floor(__ddiv_rn(x, L[0]));
    x = __dsub_rn(x, __dmul_rn(fx, L[0]));
    xf = __double2float_rn(x); yf = __double2float_rn(y); zf = __double2float_rn(z);
}

__device__ __forceinline__ int wn_cell_coord(double x, double o, double inv, int n)
{
    /* __double2int_rd saturates and maps NaN to 0; the clamp keeps every site,
     * finite or not, inside the table.  Monotone in x, and two sites closer than
     * the cell edge differ by at most one cell. */
    int c = __double2int_rd((x - o) * inv);
    return min(max(c, 0), n - 1);
}

/* The staging kernels wait on gathers and (wn_k_bin) on returning atomics: resident warps matter more than registers
 * (at the compiler's own 58-64 registers only half of the warp slots were filled: issue utilisation 16-20 %; measured
 * 8 / 6 / 5 / 4 blocks per SM: 0.503 / 0.482 / 0.482 / 0.484 ms for the staging of 250 frames). */
#ifndef WN_STAGE_MINBLOCKS
#define WN_STAGE_MINBLOCKS 6
#endif

/* ---- cell id + arrival rank per site ---- */
__global__ void __launch_bounds__(256, WN_STAGE_MINBLOCKS)
wn_k_bin(const float* __restrict__ frames, int natoms, const int* __restrict__ site_atom,
         int n_sites, WnSubset sub, int cell_cap, const WnGrid* __restrict__ grid, int* __restrict__ cell_cnt,
         int* __restrict__ site_cell, int* __restrict__ site_rank)
{
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= wn_frame_sites(sub, f, n_sites)) return;
    const WnGrid& g = grid[f];           /* fields are read where they are needed (a by-value copy held 40 registers) */
    const float* p = frames + ((size_t)f * natoms + (size_t)__ldg(site_atom + wn_table_row(sub, f, i))) * 3;
    float x = p[0], y = p[1], z = p[2];
    if (g.pbc == 1) { x = wn_wrap(x, g.L[0], g.invL[0]); y = wn_wrap(y, g.L[1], g.invL[1]); z = wn_wrap(z, g.L[2], g.invL[2]); }
    else if (g.pbc == 2) wn_wrap_tric(x, y, z, g.L, g.tri);
    const int cx = wn_cell_coord((double)x, g.ox, g.inv[0], g.nx);
    const int cy = wn_cell_coord((double)y, g.oy, g.inv[1], g.ny);
    const int cz = wn_cell_coord((double)z, g.oz, g.inv[2], g.nz);
    const int cid = (cz * g.ny + cy) * g.nx + cx;
    const int rank = atomicAdd(cell_cnt + (size_t)f * (cell_cap + 1) + cid, 1);
    site_cell[(size_t)f * n_sites + i] = cid;
    site_rank[(size_t)f * n_sites + i] = rank;
}

/* ---- write the cell-sorted site records ---- */
__global__ void __launch_bounds__(256, WN_STAGE_MINBLOCKS)
wn_k_scatter(const float* __restrict__ frames, int natoms, const int* __restrict__ site_atom,
             const int* __restrict__ hv_atom, const uint32_t* __restrict__ meta, int n_sites, WnSubset sub,
             int hs, int cell_cap, const int* __restrict__ cell_start,
             const int* __restrict__ site_cell, const int* __restrict__ site_rank,
             const WnGrid* __restrict__ grid,
             float4* __restrict__ spos, uint32_t* __restrict__ smeta, double* __restrict__ sdh)
{
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= wn_frame_sites(sub, f, n_sites)) return;
    const int row = wn_table_row(sub, f, i);          /* row of the site table; i = site index of this frame */
    const float* fr = frames + (size_t)f * natoms * 3;
    const float* p = fr + 3 * (size_t)__ldg(site_atom + row);
    float x = p[0], y = p[1], z = p[2];
    const WnGrid& g = grid[f];           /* box fields are read inside the periodic branches only (registers) */
    const int pbc = g.pbc;
    const double* L = g.L;
    const double* invL = g.invL;
    const double* tri = g.tri;
    if (pbc) {
        if (pbc == 1) { x = wn_wrap(x, L[0], invL[0]); y = wn_wrap(y, L[1], invL[1]); z = wn_wrap(z, L[2], invL[2]); }
        else wn_wrap_tric(x, y, z, L, tri);
    }
    const int cid = site_cell[(size_t)f * n_sites + i];
    const int dst = cell_start[(size_t)f * (cell_cap + 1) + cid] + site_rank[(size_t)f * n_sites + i];
    const size_t o = (size_t)f * n_sites + dst;
    spos[o] = make_float4(x, y, z, __int_as_float(i));
    smeta[o] = __ldg(meta + row);
    /* DH = *hydrogen - sitePoint; DH /= sqrt(DH*DH)   (src/waternetwork.cpp:219-220) */
    const double sx = (double)x, sy = (double)y, sz = (double)z;
    for (int v = 0; v < hs; ++v) {
        const int at = __ldg(hv_atom + (size_t)row * hs + v);
        double dx = NAN, dy = NAN, dz = NAN;
        if (at >= 0) {
            const float* h = fr + 3 * (size_t)at;
            if (pbc == 2) {
                float hx = h[0], hy = h[1], hz = h[2];
                wn_wrap_tric(hx, hy, hz, L, tri);
                dx = __dsub_rn((double)hx, sx); dy = __dsub_rn((double)hy, sy); dz = __dsub_rn((double)hz, sz);
                wn_minimg_tric(dx, dy, dz, L, tri);
            } else if (pbc == 1) {
                /* wrapped hydrogen, then the minimum image of (hydrogen - site) */
                dx = __dsub_rn((double)wn_wrap(h[0], L[0], invL[0]), sx);
                dy = __dsub_rn((double)wn_wrap(h[1], L[1], invL[1]), sy);
                dz = __dsub_rn((double)wn_wrap(h[2], L[2], invL[2]), sz);
                dx = __dsub_rn(dx, __dmul_rn(rint(__dmul_rn(dx, invL[0])), L[0]));
                dy = __dsub_rn(dy, __dmul_rn(rint(__dmul_rn(dy, invL[1])), L[1]));
                dz = __dsub_rn(dz, __dmul_rn(rint(__dmul_rn(dz, invL[2])), L[2]));
            } else {
                dx = __dsub_rn((double)h[0], sx);
                dy = __dsub_rn((double)h[1], sy);
                dz = __dsub_rn((double)h[2], sz);
            }
            const double l = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)),
                                                  __dmul_rn(dz, dz)));
            dx = __ddiv_rn(dx, l); dy = __ddiv_rn(dy, l); dz = __ddiv_rn(dz, l);
        }
        double* d = sdh + (o * hs + v) * 4;
        reinterpret_cast<double2*>(d)[0] = make_double2(dx, dy);
        reinterpret_cast<double2*>(d)[1] = make_double2(dz, 0.0);
    }
}

}  // namespace

void wn_launch_bin(const float* frames, int n_frames, int natoms, const int* site_atom,
                   int n_sites, WnSubset sub, int cell_cap, const WnGrid* grid, int* cell_cnt,
                   int* site_cell, int* site_rank, cudaStream_t st)
{
    dim3 g((n_sites + 255) / 256, n_frames);
    wn_k_bin<<<g, 256, 0, st>>>(frames, natoms, site_atom, n_sites, sub, cell_cap, grid, cell_cnt,
                                site_cell, site_rank);
}

void wn_launch_scatter(const float* frames, int n_frames, int natoms, const int* site_atom,
                       const int* hv_atom, const uint32_t* meta, int n_sites, WnSubset sub, int hs,
                       int cell_cap, const int* cell_start, const int* site_cell,
                       const int* site_rank, const WnGrid* grid, float4* spos, uint32_t* smeta, double* sdh,
                       cudaStream_t st)
{
    dim3 g((n_sites + 255) / 256, n_frames);
    wn_k_scatter<<<g, 256, 0, st>>>(frames, natoms, site_atom, hv_atom, meta, n_sites, sub, hs,
                                    cell_cap, cell_start, site_cell, site_rank, grid, spos, smeta, sdh);
}

/* Periodic grid: n = floor(L / edge_min) cells per dimension (>= 1), coarsened until the
 * table fits; the cell edge L/n is >= edge_min in every dimension. */
void wn_host_pbc_grid(const float box[3], int cell_cap, WnGrid* g)
{
    double L[3], emin = WN_CELL_EDGE;
    int n[3];
    for (int c = 0; c < 3; ++c) L[c] = (double)box[c];
    for (;;) {
        double tot = 1.0;
        for (int c = 0; c < 3; ++c) {
            double nd = floor(L[c] / emin);
            if (!(nd >= 1.0)) nd = 1.0;
            if (nd > 1.0e6) nd = 1.0e6;
            n[c] = (int)nd;
            tot *= nd;
        }
        if (tot <= (double)cell_cap) break;
        emin *= 1.1;
    }
    double emax = 0.0;
    g->ox = g->oy = g->oz = 0.0;
    for (int c = 0; c < 3; ++c) {
        const double e = L[c] / n[c];
        g->L[c] = L[c]; g->invL[c] = 1.0 / L[c];
        g->inv[c] = (double)n[c] / L[c];
        g->edge[c] = (float)e;
        if (e > emax) emax = e;
    }
    g->nx = n[0]; g->ny = n[1]; g->nz = n[2];
    g->ncells = n[0] * n[1] * n[2];
    /* the periodic image shift is folded into a float run origin (orx -+ L, or ii*ax + jj*bx + k*cx):
     * each component is rounded at up to ulp(|shift|)/2 + ulp(|origin|)/2, and a d^2 of ~0.42 with
     * |d| <= 1.2 nm moves by <= 2*1.2*3*ulp -- negligible for everyday boxes, decisive from ~30 nm on */
    double lmax = 0.0;
    for (int c = 0; c < 3; ++c) lmax = fmax(lmax, 2.0 * L[c]);      /* triclinic: |ii*ax + jj*bx + k*cx| < 2 max L */
    const double ulp_shift = (double)(nextafterf((float)lmax, 3.0e38f) - (float)lmax);
    const double tfd = WN_PREFILTER_BOUND(emax) + 8.0 * ulp_shift;
    float tf = (float)tfd;
    if ((double)tf < tfd) tf = nextafterf(tf, 3.0e38f);
    g->tf = tf;
    g->pbc = 1;
    g->tri[0] = g->tri[1] = g->tri[2] = 0.0;
}

void wn_host_tric_grid(const float box9[9], int cell_cap, WnGrid* g)
{
    /* cells on the brick [0,ax) x [0,by) x [0,cz): same rule as the rectangular grid */
    const float diag[3] = {box9[0], box9[4], box9[8]};
    wn_host_pbc_grid(diag, cell_cap, g);
    g->tri[0] = (double)box9[3]; g->tri[1] = (double)box9[6]; g->tri[2] = (double)box9[7];
    g->pbc = 2;
}
