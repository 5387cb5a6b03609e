/*
 * wn_energy.h -- the hydrogen-bond energy of an edge on the host, on demand.
 *
 * Results in the WN_LAYOUT_CSR12 layout carry {j, length, cosine} only: the energy make_edges stores
 * (reference src/waternetwork.cpp:222,234,257) is a pure function of those two floats,
 *     energy = (float) computeEnergy((double)length, (double)cosine)        (:49-66)
 * so a consumer that needs it for some edges (a weight threshold, the edge file) recomputes it here
 * with the reference's own arithmetic -- pow() from libm, the same association order, the argument
 * swap of the angle switch (:59-60) -- and gets the value the reference binary would have stored.
 * Header-only, plain C, no CUDA: it is host convenience, not part of the GPU path.
 */
#ifndef WN_ENERGY_H
#define WN_ENERGY_H

#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wn_energy_params {
    double r_on, r_off;          /* Angstrom; defaults 5.5 / 6.5   (src/waternetwork.cpp:51-52) */
    double theta_on, theta_off;  /* defaults 0.25 / 0.0            (:53-54)                      */
} wn_energy_params;

static inline wn_energy_params wn_energy_defaults(void)
{
    wn_energy_params p;
    p.r_on = 5.5; p.r_off = 6.5; p.theta_on = 0.25; p.theta_off = 0.0;
    return p;
}

/* radial switch on squared radii (:12-28): smooth between on and off, 1 below, 0 otherwise */
static inline double wn_energy_switch_radius(double q, double q_on, double q_off)
{
    if (q_off > q && q > q_on) {
        double s = pow(pow(q, 2) - pow(q_off, 2), 2);
        s *= pow(q_off, 2) + 2 * pow(q, 2) - 3 * pow(q_on, 2);
        s /= pow(pow(q_off, 2) - pow(q_on, 2), 3);
        return s;
    }
    return q < q_on ? 1.0 : 0.0;
}

/* angular switch (:30-47), called by computeEnergy with (cos^2, theta_off, theta_on) (:59-60) */
static inline double wn_energy_switch_angle(double a, double a_on, double a_off)
{
    if (a_off < a && a < a_on) {
        double s = pow(pow(a, 2) - pow(a_off, 2), 2);
        s *= pow(a_off, 2) + 2 * pow(a, 2) - 3 * pow(a_on, 2);
        s /= pow(pow(a_off, 2) - pow(a_on, 2), 3);
        return -1.0 * s;
    }
    return a > a_on ? 1.0 : 0.0;
}

/* computeEnergy (:49-66) with explicit parameters; C = 3855 and D = 738 are floats in the reference */
static inline double wn_compute_energy_host(double r, double cosine, const wn_energy_params* p)
{
    const float C = 3855.0f, D = 738.0f;
    const double sr = wn_energy_switch_radius(r * r, p->r_on * p->r_on, p->r_off * p->r_off);
    const double sa = wn_energy_switch_angle(cosine * cosine, p->theta_off, p->theta_on);
    double e = (C / pow(r, 6.0)) - (D / pow(r, 4.0));
    e *= pow(cosine, 4.0);
    e *= sr;
    e *= sa;
    return e;
}

/* the value make_edges stores for an edge: float energy of float (length, cosine) */
static inline float wn_edge_energy(float length, float cosine, const wn_energy_params* p)
{
    const wn_energy_params d = wn_energy_defaults();
    return (float)wn_compute_energy_host((double)length, (double)cosine, p ? p : &d);
}

#ifdef __cplusplus
}
#endif
#endif /* WN_ENERGY_H */
