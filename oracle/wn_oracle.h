/*
 * wn_oracle.h -- CPU oracle for the gmx-waternetwork hydrogen-bond hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing shipped under gmx-acqueduct_b200/ or
 * include/ may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * It is a fresh plain-C restatement of the reference algorithm (SURVEY.md
 * Appendix A); every function cites the reference file:line it follows
 * (paths relative to /root/reference).  All point arithmetic that the
 * reference delegates to CGAL's Cartesian kernel (absent from the reference
 * tree; "CGAL 4.X", version unpinned, README.md:12) is restated here from
 * CGAL's published semantics: component-wise subtraction, left-to-right
 * scalar product, true division, std::sqrt.
 *
 * Parity status: the reference holds NO golden vectors, known-answer tests or
 * fixtures for this path (test/tests.cpp is `int main(){return 0;}`).  The
 * oracle is pinned instead against oracle/_ref: the reference's own
 * brute-find-pairs.cpp, make_edges and computeEnergy source text compiled
 * in place against a CGAL-free arithmetic shim (oracle/Makefile, oracle/shim/).
 * See DESIGN.md section "Oracle".
 *
 * Build: gcc -O2 -ffp-contract=off (reference flags: -O2, x86-64 baseline,
 * no fast-math, CMakeLists.txt:59).
 */
#ifndef WN_ORACLE_H
#define WN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* include/utils.hpp:7  enum SiteType {DONOR, ACCEPTOR, WATER} */
enum { WNO_DONOR = 0, WNO_ACCEPTOR = 1, WNO_WATER = 2 };

/* Compact form of include/utils.hpp:25-32 HydrogenBond: the three doubles of the
 * reference each hold a float-rounded value (src/waternetwork.cpp:195-197,
 * 211-212,221-223), so float storage is lossless. */
typedef struct wno_edge {
    int32_t i, j;     /* pair.first, pair.second (0-based site indices) */
    float   energy;   /* energies.at(pos) */
    float   length;   /* distance, Angstrom */
    float   cosine;   /* coss.at(pos) */
} wno_edge;

/* Counters of exact comparison ties, reported next to every parity claim. */
typedef struct wno_ties {
    uint64_t pair_cutoff_equal;   /* 10*d2 == 42.25 in the brute test          */
    uint64_t length_equal_cut;    /* (float)dist == 6.5f in make_edges, for pairs that pass the type and hydrogen rules */
    uint64_t cosine_argmax_equal; /* c[k] == cosmax > 0 for a later k          */
    uint64_t pos_zero;            /* arg-max landed on list index 0 (dropped)  */
    uint64_t pair_cutoff_near;    /* |10*d2 - 42.25| <= 4 ulp(42.25), incl. == */
} wno_ties;

/* src/waternetwork.cpp:12-28 */
double wno_switch_function_radius(double r, double r_on, double r_off);
/* src/waternetwork.cpp:30-47 */
double wno_switch_function_angle(double a, double a_on, double a_off);
/* src/waternetwork.cpp:49-66 with explicit on/off parameters */
double wno_compute_energy_ex(double r, double cosine, double r_on, double r_off,
                             double theta_on, double theta_off);
/* src/waternetwork.cpp:49-66 with the defaults every live call uses (:222,:234) */
double wno_compute_energy(double r, double cosine);

/* src/brute-find-pairs.cpp:7-32.  xyz is AoS double[n][3].  Writes up to `cap`
 * pairs as (i,j) int32 couples in loop order; returns the full count. */
size_t wno_brute_find_pairs(const double* xyz, int n, int32_t* pairs, size_t cap,
                            wno_ties* ties);

/* Same test restricted to rows i < first_limit (north-star "asymmetric search";
 * not in the reference).  first_limit >= n gives the reference behaviour. */
size_t wno_brute_find_pairs_limited(const double* xyz, int n, int first_limit,
                                    int32_t* pairs, size_t cap, wno_ties* ties);

/* src/waternetwork.cpp:182-263.  Hydrogens of site s are
 * h_xyz[3*h_off[s] .. 3*h_off[s+1]).  Returns number of edges written to out
 * (capacity must be >= m).  n_evals counts executions of the `if (isedge)` body. */
size_t wno_make_edges(const int32_t* pairs, size_t m,
                      const double* site_xyz, int n,
                      const double* h_xyz, const int32_t* h_off,
                      const uint8_t* site_type,
                      wno_edge* out, uint64_t* n_evals, wno_ties* ties);

/* src/waternetwork.cpp:68-81 (float->double widening) and :538-577 (gather).
 * site_atom[n]: atom index of each site.  h_atom[n*hmax]: atom indices of its
 * hydrogens, -1 padded.  Outputs: site_xyz[n*3], h_xyz[(sum nH)*3], h_off[n+1]. */
void wno_gather_sites(const float* frame_xyz, int natoms,
                      const int32_t* site_atom, const int32_t* h_atom, int hmax, int n,
                      double* site_xyz, double* h_xyz, int32_t* h_off);

/* Reference water-site table (src/waternetwork.cpp:564-575): water w is site w,
 * site atom 3w, hydrogen list [3w+0, 3w+1] (the oxygen itself, then H1). */
void wno_water_site_table(int n_waters, int32_t* site_atom, int32_t* h_atom /*n*2*/,
                          uint8_t* site_type);

/* One frame end to end as WaterNetwork::analyzeFrame does it with a
 * BruteFindPairs finder (src/waternetwork.cpp:516-523,538-577,583,590-633,
 * 654-666): gather, brute pairs, make_edges, stats.
 * stats[4] = {num_sites, num_edges, 1.0*num_edges/num_sites, sum of (double)energy}.
 * Returns number of edges; edges capacity `cap` (returns SIZE_MAX on overflow). */
size_t wno_frame_hbonds(const float* frame_xyz, int natoms,
                        const int32_t* site_atom, const int32_t* h_atom, int hmax,
                        const uint8_t* site_type, int n, int first_limit,
                        wno_edge* edges, size_t cap, double* stats,
                        uint64_t* n_pairs, uint64_t* n_evals, wno_ties* ties);

/* ---- periodic boundaries (north-star extension; the reference has no PBC).
 * There is no reference oracle for this mode; it is defined in include/wn_b200.h
 * ("Periodic boundaries", after SURVEY Appendix A.4) and restated here independently.  Rectangular box L = (Lx,Ly,Lz), float32 as in a GROMACS frame:
 *   1. every coordinate is first put in the box in fp64 and rounded back to float:
 *        pw = (float)(x - floor(x * (1.0/L)) * L)          (identity for 0 <= x < L)
 *   2. every displacement the path forms (site - site, hydrogen - site) is the minimum
 *      image of the fp64 difference of those floats:
 *        d = (double)pa - (double)pb;  d = d - rint(d * (1.0/L)) * L
 *      (round-half-even: a displacement of exactly +-L/2 is kept; d(b,a) == -d(a,b))
 *   3. everything else is Appendix A.1-A.3 unchanged on those displacements. */
void wno_wrap_frame(const float* frame_xyz, int natoms, const float box[3], float* out_xyz);
size_t wno_brute_find_pairs_pbc(const double* xyz, int n, int first_limit, const float box[3],
                                int32_t* pairs, size_t cap, wno_ties* ties);
size_t wno_frame_hbonds_pbc(const float* frame_xyz, int natoms,
                            const int32_t* site_atom, const int32_t* h_atom, int hmax,
                            const uint8_t* site_type, int n, int first_limit, const float box[3],
                            wno_edge* edges, size_t cap, double* stats,
                            uint64_t* n_pairs, uint64_t* n_evals, wno_ties* ties);

/* Scattered rows of one frame: the edges (i, j > i) of the listed rows only, in list order
 * (test infrastructure for the 100k- and 1M-atom configurations, where a whole frame costs the
 * oracle seconds to minutes).  box == NULL: open boundary (the reference); else rectangular PBC. */
size_t wno_frame_hbonds_rows(const float* frame_xyz, int natoms,
                             const int32_t* site_atom, const int32_t* h_atom, int hmax,
                             const uint8_t* site_type, int n, const int32_t* rows, int n_rows,
                             const float* box, wno_edge* edges, size_t cap, uint64_t* n_evals);

/* ---- triclinic boxes (GROMACS convention: box9 = rows a=(ax,0,0), b=(bx,by,0), c=(cx,cy,cz)).
 * Definition in include/wn_b200.h ("Periodic boundaries", triclinic part), restated here:
 *   wrap:     fz = floor(z/cz): (x,y,z) -= fz*c;  fy = floor(y/by): (x,y) -= fy*(bx,by);
 *             fx = floor(x/ax): x -= fx*ax;   all in fp64, result rounded to float
 *   min image of d = pa - pb:  k = rint(d.z/cz): d -= k*c;  j = rint(d.y/by): d -= j*b;
 *             i = rint(d.x/ax): d.x -= i*ax */
void wno_wrap_frame_tric(const float* frame_xyz, int natoms, const float box9[9], float* out_xyz);
size_t wno_brute_find_pairs_tric(const double* xyz, int n, int first_limit, const float box9[9],
                                 int32_t* pairs, size_t cap, wno_ties* ties);
size_t wno_frame_hbonds_tric(const float* frame_xyz, int natoms,
                             const int32_t* site_atom, const int32_t* h_atom, int hmax,
                             const uint8_t* site_type, int n, int first_limit, const float box9[9],
                             wno_edge* edges, size_t cap, double* stats,
                             uint64_t* n_pairs, uint64_t* n_evals, wno_ties* ties);

/* ---- connected components of one frame's network -----------------------------------------
 * Restates what python/network-analysis.py does with the edge rows of a frame before any graph
 * metric: weight = |energy| (:169-171); filter_edge_by_weight removes an edge iff weight <= threshold
 * (:16-24; NaN compares false and stays); remove_node_degree removes the sites without an edge
 * (:26-35); nx.node_connected_component / connected_components split the rest (:155).
 * labels[s] = smallest site index of the component of s, -1 if s kept no edge.
 * stats = {components, largest component (sites), sites in components, kept edges}. */
void wno_components(const wno_edge* edges, size_t n_edges, int n_sites, float threshold,
                    int32_t* labels, int32_t stats[4]);

#ifdef __cplusplus
}
#endif
#endif
