/*
 * wn_geometry.cuh -- the exact arithmetic of make_edges and computeEnergy as device functions,
 * shared by the fused kernel (wn_kernels_edges.cu), the ordered copy (wn_kernels_finish.cu) and the
 * pair-list kernel (wn_kernels_pairlist.cu).  Reference: src/waternetwork.cpp:12-66 (switches, energy),
 * :182-263 (make_edges).  Bit-exactness rules: DESIGN.md section 4.
 */
#ifndef WN_GEOMETRY_CUH
#define WN_GEOMETRY_CUH

#include "wn_internal.cuh"

namespace {

__device__ __forceinline__ float wn_dot_f32(const double* __restrict__ d, double ux, double uy, double uz)
{
    /* float proj = DH*OO : (x*x' + y*y') + z*z' in double, then one rounding to float */
    return __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(d[0], ux), __dmul_rn(d[1], uy)), __dmul_rn(d[2], uz)));
}

/* minimum image of a displacement component (periodic definition, include/wn_b200.h, "Periodic boundaries"):
 * d - rint(d * (1/L)) * L, every operation rounded, round-half-even in rint */
__device__ __forceinline__ double wn_minimg(double d, double L, double invL)
{
    return __dsub_rn(d, __dmul_rn(rint(__dmul_rn(d, invL)), L));
}

/* OO /= distance (src/waternetwork.cpp:196): three IEEE divisions by the same divisor.
 * The divisor is a float widened to double (<= 24 significant bits).  With y = RN(1/b),
 * q0 = RN(a*y), r = a - q0*b (exact in one FMA), the value q0 + r*y differs from a/b by
 * <= 2^-105 relative, while a/b (a with <= 53, b with <= 24 significant bits) is never
 * closer than 2^-78 relative to a rounding midpoint and never on one -- so RN(q0 + r*y)
 * IS the correctly rounded quotient, at 3 instructions per component after one shared
 * reciprocal (vs ~15 each for the generic IEEE division).  Outside a safe exponent range
 * (or for non-finite input) the generic division is used.  wn_selftest_div compares the
 * two bit for bit on the device. */
__device__ __forceinline__ void
wn_div3_by_float(double ox, double oy, double oz, double dd, double& ux, double& uy, double& uz)
{
    const double big = 1.152921504606846976e18;        /* 2^60 */
    const bool safe = (dd >= 1.0 / big) && (dd <= big) && (fabs(ox) <= big) && (fabs(oy) <= big) && (fabs(oz) <= big);
    if (safe) {
        const double y = __drcp_rn(dd);
        double q;
        q = __dmul_rn(ox, y); ux = __fma_rn(__fma_rn(-q, dd, ox), y, q);
        q = __dmul_rn(oy, y); uy = __fma_rn(__fma_rn(-q, dd, oy), y, q);
        q = __dmul_rn(oz, y); uz = __fma_rn(__fma_rn(-q, dd, oz), y, q);
    } else {
        ux = __ddiv_rn(ox, dd); uy = __ddiv_rn(oy, dd); uz = __ddiv_rn(oz, dd);
    }
}

struct PairOut {
    bool counted, edge, t_len, t_cos, t_pos0;
    float len_f, cosmax;
};

/* Exact geometry of make_edges (src/waternetwork.cpp:192-259) for one pair.  a4/b4: float
 * positions + site index bits of the two sites in ANY order (see the role symmetry in the
 * file header); dha/dhb: their unit D-H vectors (stride 4 doubles); ma/mb: meta words. */
template <bool SIMPLE, int PBC>
__device__ __forceinline__ PairOut
wn_pair_geometry(const float4 a4, const float4 b4, const double* __restrict__ dha,
                 const double* __restrict__ dhb, uint32_t ma, uint32_t mb, const WnGrid& g)
{
    PairOut r;
    r.counted = r.edge = r.t_len = r.t_cos = r.t_pos0 = false;
    r.cosmax = 0.f;
    /* Vector_3 OO = sitePoints.at(first) - sitePoints.at(second)      (:194)
     * computed here as S_a - S_b; the other role is its exact negation */
    double ox = __dsub_rn((double)a4.x, (double)b4.x);
    double oy = __dsub_rn((double)a4.y, (double)b4.y);
    double oz = __dsub_rn((double)a4.z, (double)b4.z);
    if (PBC == 1) {
        ox = wn_minimg(ox, g.L[0], g.invL[0]);
        oy = wn_minimg(oy, g.L[1], g.invL[1]);
        oz = wn_minimg(oz, g.L[2], g.invL[2]);
    } else if (PBC == 2) {
        wn_minimg_tric(ox, oy, oz, g.L, g.tri);
    }
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ox, ox), __dmul_rn(oy, oy)), __dmul_rn(oz, oz));
    /* float distance = CGAL::sqrt(OO*OO)                               (:195) */
    const float dist_f = __double2float_rn(__dsqrt_rn(d2));
    const double dd = (double)dist_f;
    /* distance *= 10.0                                                 (:197) */
    r.len_f = __double2float_rn(__dmul_rn(dd, 10.0));
    /* if (distance > 6.5) isedge = false                               (:204) */
    r.counted = r.len_f <= WN_LEN_CUT_F;
    if (!r.counted) return r;
    r.t_len = (r.len_f == WN_LEN_CUT_F);
    /* OO /= distance                                                   (:196) */
    double ux, uy, uz;
    wn_div3_by_float(ox, oy, oz, dd, ux, uy, uz);
    if (SIMPLE) {
        /* every site: list [site atom itself (NaN, never wins), H] -> one cosine each;
         * arg-max from 0 with strict '>' (:240-249) is max(c_first, c_second) when > 0 */
        const float ca = wn_dot_f32(dha, ux, uy, uz);
        const float cb = -wn_dot_f32(dhb, ux, uy, uz);
        if (ca > 0.f) r.cosmax = ca;
        r.t_cos = (ca > 0.f) && (cb == ca);
        if (cb > r.cosmax) r.cosmax = cb;
        r.edge = r.cosmax > 0.f;
    } else {
        const bool a_first = __float_as_int(a4.w) < __float_as_int(b4.w);
        const uint32_t m_first = a_first ? ma : mb;
        bool win = false, win0 = false;
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            const bool use_a = (s == 0) == a_first;
            const uint32_t m = use_a ? ma : mb;
            const int nv = (int)((m >> WN_META_NV_SHIFT) & WN_META_NV_MASK);
            /* list index 0 is the first site's position-0 hydrogen, or the second
             * site's when the first site has no hydrogen list at all (:251) */
            const bool zero_here = (m & WN_META_FIRST0) && (s == 0 || !(m_first & WN_META_HASH));
            for (int v = 0; v < nv; ++v) {
                const float cv = use_a ? wn_dot_f32(dha + 4 * v, ux, uy, uz)        /* :219-221 */
                                       : -wn_dot_f32(dhb + 4 * v, ux, uy, uz);      /* :231-233 */
                if (cv > r.cosmax) {                                                /* :242-247 */
                    r.cosmax = cv; win = true; win0 = zero_here && (v == 0);
                } else if (r.cosmax > 0.f && cv == r.cosmax) {
                    r.t_cos = true;
                }
            }
        }
        r.t_pos0 = win && win0;
        r.edge = win && !win0;                                                      /* if (pos > 0) */
    }
    return r;
}

/* switch_function_radius (src/waternetwork.cpp:12-28); r, r_on, r_off are the
 * squared values the caller passes (:57).  pow(x,2) is restated as x*x. */
__device__ __forceinline__ double wn_switch_radius(double r, double r_on, double r_off, double inv_denom)
{
    double sw = 0.0;
    if (r_off > r && r > r_on) {
        const double r2 = r * r, off2 = r_off * r_off, on2 = r_on * r_on;
        const double a = r2 - off2;
        sw = a * a;
        sw *= off2 + 2 * r2 - 3 * on2;
        sw *= inv_denom;          /* 1 / (r_off^2 - r_on^2)^3, precomputed in fp64 (<= 1 ulp) */
    } else if (r < r_on) {
        sw = 1.0;
    }
    return sw;
}

/* switch_function_angle (src/waternetwork.cpp:30-47) */
__device__ __forceinline__ double wn_switch_angle(double a, double a_on, double a_off)
{
    double sw = 0.0;
    if (a_off < a && a < a_on) {
        const double a2 = a * a, off2 = a_off * a_off, on2 = a_on * a_on;
        const double t = a2 - off2;
        sw = t * t;
        sw *= off2 + 2 * a2 - 3 * on2;
        const double d = off2 - on2;
        sw /= d * d * d;
        sw *= -1.0;
    } else if (a > a_on) {
        sw = 1.0;
    }
    return sw;
}

/* computeEnergy (src/waternetwork.cpp:49-66): C=3855, D=738 (static float),
 * pow(r,6), pow(r,4), pow(cosine,4) restated as products, the two quotients merged
 * into one and the constant switch denominator inverted once (all <= a few ulp from
 * the glibc-pow evaluation; the contract is 1e-6 relative, the result is stored as
 * float). */
__device__ __forceinline__ double wn_compute_energy(double r, double cosine, const WnEnergyParams& ep)
{
    const double r2 = r * r, r4 = r2 * r2, r6 = r4 * r2;
    const double radius_switch = wn_switch_radius(r2, ep.on2, ep.off2, ep.inv_denom_r);
    const double c2 = cosine * cosine;
    const double angle_switch = wn_switch_angle(c2, ep.a_on, ep.a_off);
    /* C/r^6 - D/r^4 as one division (C - D r^2)/r^6: same value to ~2 ulp except within
     * ~1e-9 of the zero crossing r = sqrt(C/D), which no float32 length reaches */
    /* 1/r^6 from the hardware seed (MUFU.RCP64H, ~2^-20) and two Newton steps: <= 2 ulp, ~9 instructions instead of
     * the ~25 of the IEEE division; r^6 of an edge (0 < r <= 6.5 A, float) is a normal double far from the flush range */
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r6));
    y = fma(y, fma(-r6, y, 1.0), y);
    y = fma(y, fma(-r6, y, 1.0), y);
    double energy = (3855.0 - 738.0 * r2) * y;
    energy *= c2 * c2;
    energy *= radius_switch;
    energy *= angle_switch;
    return energy;
}

}  // namespace

#endif
