/*
 * wn_oracle.c -- CPU oracle (test infrastructure only; see wn_oracle.h).
 *
 * Plain C restatement of the reference hot path.  Every comparison-feeding
 * value is an IEEE binary64 computed in the reference's association order;
 * build with -ffp-contract=off so the compiler cannot fuse a*b+c.
 */
#include "wn_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- CGAL Cartesian-kernel semantics (restated; CGAL is not in the tree) ---- */

typedef struct { double x, y, z; } v3;

static inline v3 v3_sub(const double* p, const double* q)      /* Point - Point */
{
    v3 r; r.x = p[0] - q[0]; r.y = p[1] - q[1]; r.z = p[2] - q[2]; return r;
}
static inline double v3_dot(v3 a, v3 b)                         /* Vector * Vector */
{
    return a.x * b.x + a.y * b.y + a.z * b.z;                   /* ((xx+yy)+zz) */
}
static inline v3 v3_div(v3 a, double c)                         /* Vector / scalar */
{
    v3 r; r.x = a.x / c; r.y = a.y / c; r.z = a.z / c; return r;
}
static inline v3 v3_scale(double c, v3 a)                       /* scalar * Vector */
{
    v3 r; r.x = c * a.x; r.y = c * a.y; r.z = c * a.z; return r;
}

/* ---- src/waternetwork.cpp:12-28 ---- */
double wno_switch_function_radius(double r, double r_on, double r_off)
{
    double sw = 0.0;
    if (r_off > r && r > r_on) {
        sw = pow(pow(r, 2) - pow(r_off, 2), 2);
        sw *= pow(r_off, 2) + 2 * pow(r, 2) - 3 * pow(r_on, 2);
        sw /= pow(pow(r_off, 2) - pow(r_on, 2), 3);
    } else if (r < r_on) {
        sw = 1.0;
    }
    return sw;
}

/* ---- src/waternetwork.cpp:30-47 ---- */
double wno_switch_function_angle(double a, double a_on, double a_off)
{
    double sw = 0.0;
    if (a_off < a && a < a_on) {
        sw = pow(pow(a, 2) - pow(a_off, 2), 2);
        sw *= pow(a_off, 2) + 2 * pow(a, 2) - 3 * pow(a_on, 2);
        sw /= pow(pow(a_off, 2) - pow(a_on, 2), 3);
        sw *= -1.0;
    } else if (a > a_on) {
        sw = 1.0;
    }
    return sw;
}

/* ---- src/waternetwork.cpp:49-66 ----
 * C and D are `static float` in the reference, widened in the divisions.
 * Note the caller-side argument swap (:59-60): the angle switch receives
 * (theta_off, theta_on) in its (a_on, a_off) slots. */
double wno_compute_energy_ex(double r, double cosine, double r_on, double r_off,
                             double theta_on, double theta_off)
{
    static float C = 3855;
    static float D = 738;
    double radius_switch = wno_switch_function_radius(r * r, r_on * r_on, r_off * r_off);
    double angle_switch = wno_switch_function_angle(cosine * cosine, theta_off, theta_on);
    double energy = ((C / pow(r, 6.0)) - (D / pow(r, 4.0)));
    energy *= pow(cosine, 4.0);
    energy *= radius_switch;
    energy *= angle_switch;
    return energy;
}

double wno_compute_energy(double r, double cosine)
{
    return wno_compute_energy_ex(r, cosine, 5.5, 6.5, 0.25, 0.0);
}

/* ---- src/brute-find-pairs.cpp:19-30 ----
 * float cutoff = 6.5; cutoff *= cutoff;  -> 42.25f, compared as double.
 * CGAL::squared_distance(p,q) = (dx*dx + dy*dy) + dz*dz. */
size_t wno_brute_find_pairs_limited(const double* xyz, int n, int first_limit,
                                    int32_t* pairs, size_t cap, wno_ties* ties)
{
    float cutoff = 6.5f;
    size_t count = 0;
    int i, j;
    cutoff *= cutoff;
    if (first_limit > n) first_limit = n;
    for (i = 0; i < first_limit; ++i) {
        const double* p = xyz + 3 * (size_t)i;
        for (j = i + 1; j < n; ++j) {
            const double* q = xyz + 3 * (size_t)j;
            v3 d = v3_sub(p, q);
            double d2 = v3_dot(d, d);
            double lhs = 10.0 * d2;
            if (lhs < cutoff) {
                if (pairs && count < cap) {
                    pairs[2 * count] = i;
                    pairs[2 * count + 1] = j;
                }
                ++count;
            } else if (ties && lhs == (double)cutoff) {
                ties->pair_cutoff_equal++;
            }
            if (ties && fabs(lhs - (double)cutoff) <= 2.8421709430404007e-14) ties->pair_cutoff_near++;
        }
    }
    return count;
}

size_t wno_brute_find_pairs(const double* xyz, int n, int32_t* pairs, size_t cap,
                            wno_ties* ties)
{
    return wno_brute_find_pairs_limited(xyz, n, n, pairs, cap, ties);
}

/* ---- src/waternetwork.cpp:182-263 ---- */
size_t wno_make_edges(const int32_t* pairs, size_t m,
                      const double* site_xyz, int n,
                      const double* h_xyz, const int32_t* h_off,
                      const uint8_t* site_type,
                      wno_edge* out, uint64_t* n_evals, wno_ties* ties)
{
    size_t count = 0, k;
    uint64_t evals = 0;
    float* coss = NULL;
    float* energies = NULL;
    int maxh = 0, s;
    (void)n;
    for (s = 0; s < n; ++s) {
        int nh = h_off[s + 1] - h_off[s];
        if (nh > maxh) maxh = nh;
    }
    coss = (float*)malloc(sizeof(float) * (size_t)(2 * maxh + 1));
    energies = (float*)malloc(sizeof(float) * (size_t)(2 * maxh + 1));

    for (k = 0; k < m; ++k) {
        const int a = pairs[2 * k], b = pairs[2 * k + 1];
        const double* pa = site_xyz + 3 * (size_t)a;
        const double* pb = site_xyz + 3 * (size_t)b;
        const int nha = h_off[a + 1] - h_off[a];
        const int nhb = h_off[b + 1] - h_off[b];
        int isedge = 0;
        v3 OO = v3_sub(pa, pb);                                   /* :194 */
        float distance = (float)sqrt(v3_dot(OO, OO));             /* :195 */
        OO = v3_div(OO, (double)distance);                        /* :196 */
        distance = (float)((double)distance * 10.0);              /* :197 */

        if (site_type[a] != site_type[b]) isedge = 1;             /* :199 */
        if (site_type[a] == WNO_WATER && site_type[b] == WNO_WATER) isedge = 1;
        /* a tie at the cut is reported where it is consequential: the pair passes the
         * type rule and the hydrogen rule, so the comparison below decides its fate */
        if (ties && distance == 6.5f && isedge && !(nha == 0 && nhb == 0)) ties->length_equal_cut++;
        if ((double)distance > 6.5) isedge = 0;                   /* :204 */
        if (nha == 0 && nhb == 0) isedge = 0;                     /* :205-208 */

        if (isedge) {
            int nc = 0, h, pos = -1, i;
            float cosmax = 0.0f;
            ++evals;
            for (h = 0; h < nha; ++h) {                           /* :215-225 */
                v3 DH = v3_sub(h_xyz + 3 * (size_t)(h_off[a] + h), pa);
                float proj;
                DH = v3_div(DH, sqrt(v3_dot(DH, DH)));
                proj = (float)v3_dot(DH, OO);
                energies[nc] = (float)wno_compute_energy((double)distance, (double)proj);
                coss[nc] = proj;
                ++nc;
            }
            for (h = 0; h < nhb; ++h) {                           /* :227-237 */
                v3 DH = v3_sub(h_xyz + 3 * (size_t)(h_off[b] + h), pb);
                float proj;
                DH = v3_div(DH, sqrt(v3_dot(DH, DH)));
                proj = (float)v3_dot(DH, v3_scale(-1.0, OO));
                energies[nc] = (float)wno_compute_energy((double)distance, (double)proj);
                coss[nc] = proj;
                ++nc;
            }
            for (i = 0; i < nc; ++i) {                            /* :240-249 */
                if (coss[i] > cosmax) {
                    cosmax = coss[i];
                    pos = i;
                } else if (ties && cosmax > 0.0f && coss[i] == cosmax) {
                    ties->cosine_argmax_equal++;
                }
            }
            if (ties && pos == 0) ties->pos_zero++;
            if (pos > 0) {                                        /* :251-258 */
                out[count].i = a;
                out[count].j = b;
                out[count].energy = energies[pos];
                out[count].length = distance;
                out[count].cosine = coss[pos];
                ++count;
            }
        }
    }
    free(coss);
    free(energies);
    if (n_evals) *n_evals = evals;
    return count;
}

/* ---- src/waternetwork.cpp:68-81 (widening) and :538-577 (gather) ---- */
void wno_gather_sites(const float* frame_xyz, int natoms,
                      const int32_t* site_atom, const int32_t* h_atom, int hmax, int n,
                      double* site_xyz, double* h_xyz, int32_t* h_off)
{
    int s, h, c, nh_total = 0;
    (void)natoms;
    for (s = 0; s < n; ++s) {
        const float* p = frame_xyz + 3 * (size_t)site_atom[s];
        for (c = 0; c < 3; ++c) site_xyz[3 * (size_t)s + c] = (double)p[c];
        h_off[s] = nh_total;
        for (h = 0; h < hmax; ++h) {
            int at = h_atom[(size_t)s * hmax + h];
            if (at < 0) break;
            for (c = 0; c < 3; ++c)
                h_xyz[3 * (size_t)nh_total + c] = (double)frame_xyz[3 * (size_t)at + c];
            ++nh_total;
        }
    }
    h_off[n] = nh_total;
}

/* ---- src/waternetwork.cpp:564-575: hydrogens.at(i) = solventPoints.at(3*index+i),
 * i < nbHydrogen (=2 for water, makeSites :155-161), site = solventPoints.at(3*index) */
void wno_water_site_table(int n_waters, int32_t* site_atom, int32_t* h_atom,
                          uint8_t* site_type)
{
    int w;
    for (w = 0; w < n_waters; ++w) {
        site_atom[w] = 3 * w;
        h_atom[2 * w] = 3 * w;
        h_atom[2 * w + 1] = 3 * w + 1;
        site_type[w] = WNO_WATER;
    }
}

size_t wno_frame_hbonds(const float* frame_xyz, int natoms,
                        const int32_t* site_atom, const int32_t* h_atom, int hmax,
                        const uint8_t* site_type, int n, int first_limit,
                        wno_edge* edges, size_t cap, double* stats,
                        uint64_t* n_pairs, uint64_t* n_evals, wno_ties* ties)
{
    double* site_xyz = (double*)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
    double* h_xyz = (double*)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1) * (size_t)(hmax > 0 ? hmax : 1));
    int32_t* h_off = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n + 1));
    int32_t* pairs = NULL;
    wno_edge* tmp = NULL;
    size_t m, ne, k;
    double sum_e = 0.0;

    wno_gather_sites(frame_xyz, natoms, site_atom, h_atom, hmax, n, site_xyz, h_xyz, h_off);
    m = wno_brute_find_pairs_limited(site_xyz, n, first_limit, NULL, 0, NULL);
    pairs = (int32_t*)malloc(sizeof(int32_t) * 2 * (m ? m : 1));
    wno_brute_find_pairs_limited(site_xyz, n, first_limit, pairs, m, ties);
    tmp = (wno_edge*)malloc(sizeof(wno_edge) * (m ? m : 1));
    ne = wno_make_edges(pairs, m, site_xyz, n, h_xyz, h_off, site_type, tmp, n_evals, ties);
    if (n_pairs) *n_pairs = m;
    for (k = 0; k < ne; ++k) sum_e += (double)tmp[k].energy;
    if (stats) {                                     /* :654-666 + extra sum E */
        stats[0] = (double)n;
        stats[1] = (double)ne;
        stats[2] = n > 0 ? 1.0 * (double)ne / (double)n : 0.0;
        stats[3] = sum_e;
    }
    if (ne <= cap && edges) memcpy(edges, tmp, sizeof(wno_edge) * ne);
    free(tmp); free(pairs); free(h_off); free(h_xyz); free(site_xyz);
    return ne <= cap ? ne : (size_t)-1;
}

/* ======================= periodic boundaries (defined here, see wn_oracle.h) ======================= */

static inline double minimg(double d, double L, double invL)
{
    return d - rint(d * invL) * L;
}

void wno_wrap_frame(const float* frame_xyz, int natoms, const float box[3], float* out_xyz)
{
    int a, c;
    for (a = 0; a < natoms; ++a)
        for (c = 0; c < 3; ++c) {
            const double L = (double)box[c], invL = 1.0 / L;
            const double x = (double)frame_xyz[3 * (size_t)a + c];
            out_xyz[3 * (size_t)a + c] = (float)(x - floor(x * invL) * L);
        }
}

/* the pair test of src/brute-find-pairs.cpp:19-30 on minimum-image displacements */
size_t wno_brute_find_pairs_pbc(const double* xyz, int n, int first_limit, const float box[3],
                                int32_t* pairs, size_t cap, wno_ties* ties)
{
    float cutoff = 6.5f;
    size_t count = 0;
    int i, j, c;
    double L[3], invL[3];
    cutoff *= cutoff;
    for (c = 0; c < 3; ++c) { L[c] = (double)box[c]; invL[c] = 1.0 / L[c]; }
    if (first_limit > n) first_limit = n;
    for (i = 0; i < first_limit; ++i)
        for (j = i + 1; j < n; ++j) {
            v3 d = v3_sub(xyz + 3 * (size_t)i, xyz + 3 * (size_t)j);
            double lhs;
            d.x = minimg(d.x, L[0], invL[0]); d.y = minimg(d.y, L[1], invL[1]); d.z = minimg(d.z, L[2], invL[2]);
            lhs = 10.0 * v3_dot(d, d);
            if (lhs < cutoff) {
                if (pairs && count < cap) { pairs[2 * count] = i; pairs[2 * count + 1] = j; }
                ++count;
            } else if (ties && lhs == (double)cutoff) {
                ties->pair_cutoff_equal++;
            }
            if (ties && fabs(lhs - (double)cutoff) <= 2.8421709430404007e-14) ties->pair_cutoff_near++;
        }
    return count;
}

/* box descriptor for the periodic variants: rectangular (tric == NULL) or triclinic */
typedef struct { const double* L; const double* invL; const double* tric; /* ax,by,cz,bx,cx,cy */ } pbox;

static inline v3 minimg_box(v3 d, const pbox* b)
{
    if (b->tric) {
        const double ax = b->tric[0], by = b->tric[1], cz = b->tric[2], bx = b->tric[3], cx = b->tric[4], cy = b->tric[5];
        const double k = rint(d.z / cz);
        d.x = d.x - k * cx; d.y = d.y - k * cy; d.z = d.z - k * cz;
        {
            const double j = rint(d.y / by);
            d.x = d.x - j * bx; d.y = d.y - j * by;
        }
        {
            const double i = rint(d.x / ax);
            d.x = d.x - i * ax;
        }
    } else {
        d.x = minimg(d.x, b->L[0], b->invL[0]); d.y = minimg(d.y, b->L[1], b->invL[1]); d.z = minimg(d.z, b->L[2], b->invL[2]);
    }
    return d;
}

/* make_edges (src/waternetwork.cpp:182-263) on minimum-image displacements */
static size_t make_edges_box(const int32_t* pairs, size_t m, const double* site_xyz, int n,
                             const double* h_xyz, const int32_t* h_off, const uint8_t* site_type,
                             const pbox* bx_,
                             wno_edge* out, uint64_t* n_evals, wno_ties* ties)
{
    size_t count = 0, k;
    uint64_t evals = 0;
    int maxh = 0, s;
    float *coss, *energies;
    for (s = 0; s < n; ++s) { int nh = h_off[s + 1] - h_off[s]; if (nh > maxh) maxh = nh; }
    coss = (float*)malloc(sizeof(float) * (size_t)(2 * maxh + 1));
    energies = (float*)malloc(sizeof(float) * (size_t)(2 * maxh + 1));
    for (k = 0; k < m; ++k) {
        const int a = pairs[2 * k], b = pairs[2 * k + 1];
        const double* pa = site_xyz + 3 * (size_t)a;
        const double* pb = site_xyz + 3 * (size_t)b;
        const int nha = h_off[a + 1] - h_off[a], nhb = h_off[b + 1] - h_off[b];
        int isedge = 0;
        v3 OO = minimg_box(v3_sub(pa, pb), bx_);
        float distance;
        distance = (float)sqrt(v3_dot(OO, OO));
        OO = v3_div(OO, (double)distance);
        distance = (float)((double)distance * 10.0);
        if (site_type[a] != site_type[b]) isedge = 1;
        if (site_type[a] == WNO_WATER && site_type[b] == WNO_WATER) isedge = 1;
        if (ties && distance == 6.5f && isedge && !(nha == 0 && nhb == 0)) ties->length_equal_cut++;
        if ((double)distance > 6.5) isedge = 0;
        if (nha == 0 && nhb == 0) isedge = 0;
        if (isedge) {
            int nc = 0, h, pos = -1, i;
            float cosmax = 0.0f;
            ++evals;
            for (h = 0; h < nha + nhb; ++h) {
                const int first = h < nha;
                const double* hp = h_xyz + 3 * (size_t)(first ? h_off[a] + h : h_off[b] + h - nha);
                v3 DH = minimg_box(v3_sub(hp, first ? pa : pb), bx_);
                float proj;
                DH = v3_div(DH, sqrt(v3_dot(DH, DH)));
                proj = first ? (float)v3_dot(DH, OO) : (float)v3_dot(DH, v3_scale(-1.0, OO));
                energies[nc] = (float)wno_compute_energy((double)distance, (double)proj);
                coss[nc] = proj;
                ++nc;
            }
            for (i = 0; i < nc; ++i) {
                if (coss[i] > cosmax) { cosmax = coss[i]; pos = i; }
                else if (ties && cosmax > 0.0f && coss[i] == cosmax) ties->cosine_argmax_equal++;
            }
            if (ties && pos == 0) ties->pos_zero++;
            if (pos > 0) {
                out[count].i = a; out[count].j = b;
                out[count].energy = energies[pos]; out[count].length = distance; out[count].cosine = coss[pos];
                ++count;
            }
        }
    }
    free(coss); free(energies);
    if (n_evals) *n_evals = evals;
    return count;
}

size_t wno_frame_hbonds_pbc(const float* frame_xyz, int natoms,
                            const int32_t* site_atom, const int32_t* h_atom, int hmax,
                            const uint8_t* site_type, int n, int first_limit, const float box[3],
                            wno_edge* edges, size_t cap, double* stats,
                            uint64_t* n_pairs, uint64_t* n_evals, wno_ties* ties)
{
    float* wrapped = (float*)malloc(sizeof(float) * 3 * (size_t)(natoms > 0 ? natoms : 1));
    double* site_xyz = (double*)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
    double* h_xyz = (double*)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1) * (size_t)(hmax > 0 ? hmax : 1));
    int32_t* h_off = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n + 1));
    int32_t* pairs;
    wno_edge* tmp;
    size_t m, ne, k;
    double sum_e = 0.0, L[3], invL[3];
    int c;
    for (c = 0; c < 3; ++c) { L[c] = (double)box[c]; invL[c] = 1.0 / L[c]; }
    wno_wrap_frame(frame_xyz, natoms, box, wrapped);
    wno_gather_sites(wrapped, natoms, site_atom, h_atom, hmax, n, site_xyz, h_xyz, h_off);
    m = wno_brute_find_pairs_pbc(site_xyz, n, first_limit, box, NULL, 0, NULL);
    pairs = (int32_t*)malloc(sizeof(int32_t) * 2 * (m ? m : 1));
    wno_brute_find_pairs_pbc(site_xyz, n, first_limit, box, pairs, m, ties);
    tmp = (wno_edge*)malloc(sizeof(wno_edge) * (m ? m : 1));
    {
        pbox pb_;
        pb_.L = L; pb_.invL = invL; pb_.tric = NULL;
        ne = make_edges_box(pairs, m, site_xyz, n, h_xyz, h_off, site_type, &pb_, tmp, n_evals, ties);
    }
    if (n_pairs) *n_pairs = m;
    for (k = 0; k < ne; ++k) sum_e += (double)tmp[k].energy;
    if (stats) {
        stats[0] = (double)n; stats[1] = (double)ne;
        stats[2] = n > 0 ? 1.0 * (double)ne / (double)n : 0.0; stats[3] = sum_e;
    }
    if (ne <= cap && edges) memcpy(edges, tmp, sizeof(wno_edge) * ne);
    free(tmp); free(pairs); free(h_off); free(h_xyz); free(site_xyz); free(wrapped);
    return ne <= cap ? ne : (size_t)-1;
}

/* ---- scattered rows of one frame (test infrastructure for the large configurations) ----
 * The pairs (i, j > i) of the listed rows only, in the order of the list (ascending rows give the
 * reference's lexicographic order restricted to those rows), through the same pair test
 * (src/brute-find-pairs.cpp:19-30) and the same make_edges (src/waternetwork.cpp:182-263).
 * box == NULL: open boundary; else the rectangular periodic definition above. */
size_t wno_frame_hbonds_rows(const float* frame_xyz, int natoms,
                             const int32_t* site_atom, const int32_t* h_atom, int hmax,
                             const uint8_t* site_type, int n, const int32_t* rows, int n_rows,
                             const float* box, wno_edge* edges, size_t cap, uint64_t* n_evals)
{
    float* wrapped = NULL;
    double* site_xyz = (double*)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
    double* h_xyz = (double*)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1) * (size_t)(hmax > 0 ? hmax : 1));
    int32_t* h_off = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n + 1));
    int32_t* pairs = NULL;
    wno_edge* tmp;
    size_t m = 0, pcap = 0, ne;
    double L[3] = {0, 0, 0}, invL[3] = {0, 0, 0};
    float cutoff = 6.5f;
    int r, j, c;
    cutoff *= cutoff;
    if (box) {
        for (c = 0; c < 3; ++c) { L[c] = (double)box[c]; invL[c] = 1.0 / L[c]; }
        wrapped = (float*)malloc(sizeof(float) * 3 * (size_t)(natoms > 0 ? natoms : 1));
        wno_wrap_frame(frame_xyz, natoms, box, wrapped);
        frame_xyz = wrapped;
    }
    wno_gather_sites(frame_xyz, natoms, site_atom, h_atom, hmax, n, site_xyz, h_xyz, h_off);
    for (r = 0; r < n_rows; ++r) {
        const int i = rows[r];
        if (i < 0 || i >= n) continue;
        for (j = i + 1; j < n; ++j) {
            v3 d = v3_sub(site_xyz + 3 * (size_t)i, site_xyz + 3 * (size_t)j);
            if (box) { d.x = minimg(d.x, L[0], invL[0]); d.y = minimg(d.y, L[1], invL[1]); d.z = minimg(d.z, L[2], invL[2]); }
            if (10.0 * v3_dot(d, d) < cutoff) {
                if (m == pcap) {
                    pcap = pcap ? 2 * pcap : 4096;
                    pairs = (int32_t*)realloc(pairs, sizeof(int32_t) * 2 * pcap);
                }
                pairs[2 * m] = i; pairs[2 * m + 1] = j; ++m;
            }
        }
    }
    tmp = (wno_edge*)malloc(sizeof(wno_edge) * (m ? m : 1));
    if (box) {
        pbox pb_;
        pb_.L = L; pb_.invL = invL; pb_.tric = NULL;
        ne = make_edges_box(pairs, m, site_xyz, n, h_xyz, h_off, site_type, &pb_, tmp, n_evals, NULL);
    } else {
        ne = wno_make_edges(pairs, m, site_xyz, n, h_xyz, h_off, site_type, tmp, n_evals, NULL);
    }
    if (ne <= cap && edges) memcpy(edges, tmp, sizeof(wno_edge) * ne);
    free(tmp); free(pairs); free(h_off); free(h_xyz); free(site_xyz); free(wrapped);
    return ne <= cap ? ne : (size_t)-1;
}

/* ======================= triclinic boxes (definition: include/wn_b200.h) ======================= */

static void tric_unpack(const float box9[9], double t[6])
{
    t[0] = (double)box9[0]; t[1] = (double)box9[4]; t[2] = (double)box9[8];   /* ax, by, cz */
    t[3] = (double)box9[3]; t[4] = (double)box9[6]; t[5] = (double)box9[7];   /* bx, cx, cy */
}

void wno_wrap_frame_tric(const float* frame_xyz, int natoms, const float box9[9], float* out_xyz)
{
    double t[6];
    int a;
    tric_unpack(box9, t);
    for (a = 0; a < natoms; ++a) {
        double x = (double)frame_xyz[3 * (size_t)a], y = (double)frame_xyz[3 * (size_t)a + 1], z = (double)frame_xyz[3 * (size_t)a + 2];
        const double fz = floor(z / t[2]);
        double fy, fx;
        x = x - fz * t[4]; y = y - fz * t[5]; z = z - fz * t[2];
        fy = floor(y / t[1]);
        x = x - fy * t[3]; y = y - fy * t[1];
        fx = floor(x / t[0]);
        x = x - fx * t[0];
        out_xyz[3 * (size_t)a] = (float)x; out_xyz[3 * (size_t)a + 1] = (float)y; out_xyz[3 * (size_t)a + 2] = (float)z;
    }
}

size_t wno_brute_find_pairs_tric(const double* xyz, int n, int first_limit, const float box9[9],
                                 int32_t* pairs, size_t cap, wno_ties* ties)
{
    float cutoff = 6.5f;
    size_t count = 0;
    int i, j;
    double t[6];
    pbox pb_;
    cutoff *= cutoff;
    tric_unpack(box9, t);
    pb_.L = NULL; pb_.invL = NULL; pb_.tric = t;
    if (first_limit > n) first_limit = n;
    for (i = 0; i < first_limit; ++i)
        for (j = i + 1; j < n; ++j) {
            const v3 d = minimg_box(v3_sub(xyz + 3 * (size_t)i, xyz + 3 * (size_t)j), &pb_);
            const double lhs = 10.0 * v3_dot(d, d);
            if (lhs < cutoff) {
                if (pairs && count < cap) { pairs[2 * count] = i; pairs[2 * count + 1] = j; }
                ++count;
            } else if (ties && lhs == (double)cutoff) {
                ties->pair_cutoff_equal++;
            }
            if (ties && fabs(lhs - (double)cutoff) <= 2.8421709430404007e-14) ties->pair_cutoff_near++;
        }
    return count;
}

size_t wno_frame_hbonds_tric(const float* frame_xyz, int natoms,
                             const int32_t* site_atom, const int32_t* h_atom, int hmax,
                             const uint8_t* site_type, int n, int first_limit, const float box9[9],
                             wno_edge* edges, size_t cap, double* stats,
                             uint64_t* n_pairs, uint64_t* n_evals, wno_ties* ties)
{
    float* wrapped = (float*)malloc(sizeof(float) * 3 * (size_t)(natoms > 0 ? natoms : 1));
    double* site_xyz = (double*)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
    double* h_xyz = (double*)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1) * (size_t)(hmax > 0 ? hmax : 1));
    int32_t* h_off = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n + 1));
    int32_t* pairs;
    wno_edge* tmp;
    size_t m, ne, k;
    double sum_e = 0.0, t[6];
    pbox pb_;
    tric_unpack(box9, t);
    pb_.L = NULL; pb_.invL = NULL; pb_.tric = t;
    wno_wrap_frame_tric(frame_xyz, natoms, box9, wrapped);
    wno_gather_sites(wrapped, natoms, site_atom, h_atom, hmax, n, site_xyz, h_xyz, h_off);
    m = wno_brute_find_pairs_tric(site_xyz, n, first_limit, box9, NULL, 0, NULL);
    pairs = (int32_t*)malloc(sizeof(int32_t) * 2 * (m ? m : 1));
    wno_brute_find_pairs_tric(site_xyz, n, first_limit, box9, pairs, m, ties);
    tmp = (wno_edge*)malloc(sizeof(wno_edge) * (m ? m : 1));
    ne = make_edges_box(pairs, m, site_xyz, n, h_xyz, h_off, site_type, &pb_, tmp, n_evals, ties);
    if (n_pairs) *n_pairs = m;
    for (k = 0; k < ne; ++k) sum_e += (double)tmp[k].energy;
    if (stats) {
        stats[0] = (double)n; stats[1] = (double)ne;
        stats[2] = n > 0 ? 1.0 * (double)ne / (double)n : 0.0; stats[3] = sum_e;
    }
    if (ne <= cap && edges) memcpy(edges, tmp, sizeof(wno_edge) * ne);
    free(tmp); free(pairs); free(h_off); free(h_xyz); free(site_xyz); free(wrapped);
    return ne <= cap ? ne : (size_t)-1;
}


/* ---- connected components (python/network-analysis.py:16-35, 141-171) ---- */
static int cc_root(int32_t* parent, int v)
{
    int r = v;
    while (parent[r] != r) r = parent[r];
    while (parent[v] != r) { const int nx = parent[v]; parent[v] = r; v = nx; }
    return r;
}

void wno_components(const wno_edge* edges, size_t n_edges, int n_sites, float threshold,
                    int32_t* labels, int32_t stats[4])
{
    int32_t* parent = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_sites > 0 ? n_sites : 1));
    int32_t* count = (int32_t*)calloc((size_t)(n_sites > 0 ? n_sites : 1), sizeof(int32_t));
    unsigned char* has = (unsigned char*)calloc((size_t)(n_sites > 0 ? n_sites : 1), 1);
    int32_t kept = 0, comps = 0, largest = 0, nodes = 0;
    for (int v = 0; v < n_sites; ++v) parent[v] = v;
    for (size_t e = 0; e < n_edges; ++e) {
        const float w = fabsf(edges[e].energy);            /* weight=abs(float(line[2]))  :169 */
        if (w <= threshold) continue;                      /* removed iff weight <= threshold  :22 */
        ++kept;
        has[edges[e].i] = has[edges[e].j] = 1;
        const int a = cc_root(parent, edges[e].i), b = cc_root(parent, edges[e].j);
        if (a < b) parent[b] = a; else if (b < a) parent[a] = b;   /* root = smallest member */
    }
    for (int v = 0; v < n_sites; ++v) {
        if (!has[v]) { labels[v] = -1; continue; }         /* remove_node_degree  :31-34 */
        labels[v] = cc_root(parent, v);
        ++count[labels[v]];
        ++nodes;
    }
    for (int v = 0; v < n_sites; ++v)
        if (labels[v] == v) { ++comps; if (count[v] > largest) largest = count[v]; }
    stats[0] = comps; stats[1] = largest; stats[2] = nodes; stats[3] = kept;
    free(parent); free(count); free(has);
}
