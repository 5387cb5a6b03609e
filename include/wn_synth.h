/*
 * wn_synth.h -- bit-reproducible synthetic water-box frames.
 *
 * The reference ships no trajectory, so every benchmark and parity test runs on
 * synthetic TIP3P-like (or SPC) rigid-water boxes.  This header is the single
 * definition of that generator.  It is plain C99 / CUDA C++ and uses only
 * integer hashing plus IEEE-754 double  + - * / sqrt  (no FMA, no libm), so the
 * host build (gcc -ffp-contract=off) and the device build (nvcc -fmad=false,
 * see csrc/wn_synth.cu) produce bit-identical float32 coordinates.
 *
 * Layout produced: GROMACS-like AoS float32, atom order OW,HW1,HW2 per water,
 * i.e. frame[(3*w + a)*3 + c]  (what gmx::Selection::coordinates() hands to
 * WaterNetwork::analyzeFrame, reference src/waternetwork.cpp:516-523).
 */
#ifndef WN_SYNTH_H
#define WN_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define WN_HD __host__ __device__ __forceinline__
#else
#define WN_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define WN_SQRT(x) __dsqrt_rn(x)
#else
#include <math.h>
#define WN_SQRT(x) sqrt(x)
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Water model geometry (nm, and sin/cos of half the H-O-H angle). */
typedef struct wn_water_model {
    double r_oh;      /* O-H bond length, nm                         */
    double sin_half;  /* sin(theta_HOH / 2)                          */
    double cos_half;  /* cos(theta_HOH / 2)                          */
} wn_water_model;

/* TIP3P: r=0.09572 nm, theta=104.52 deg.  SPC: r=0.1 nm, theta=109.47 deg. */
#define WN_MODEL_TIP3P 0
#define WN_MODEL_SPC   1

WN_HD wn_water_model wn_synth_model(int which)
{
    wn_water_model m;
    if (which == WN_MODEL_SPC) {
        m.r_oh = 0.1;
        m.sin_half = 0.81649043092495239;  /* sin(54.735 deg) */
        m.cos_half = 0.57735764305523601;  /* cos(54.735 deg) */
    } else {
        m.r_oh = 0.09572;
        m.sin_half = 0.79081030116067216;  /* sin(52.26 deg) */
        m.cos_half = 0.61205789274439245;  /* cos(52.26 deg) */
    }
    return m;
}

/* Box description shared by every frame of a synthetic trajectory. */
typedef struct wn_synth_box {
    int32_t  n_waters;
    int32_t  m;        /* lattice points per axis, m^3 >= n_waters      */
    double   box;      /* cubic box edge L, nm                          */
    double   a;        /* lattice spacing L/m, nm                       */
    uint64_t seed;
    int32_t  model;    /* WN_MODEL_*                                    */
} wn_synth_box;

/* Number density 33.4 nm^-3  =>  L = (N/33.4)^(1/3).  Integer cube root by
 * search and L by Newton iteration in exact IEEE ops, so that host and device
 * agree without cbrt(). */
WN_HD wn_synth_box wn_synth_make_box(int32_t n_waters, uint64_t seed, int32_t model)
{
    wn_synth_box b;
    int32_t m = 1;
    double v, x;
    int it;
    while ((int64_t)m * m * m < (int64_t)n_waters) ++m;
    v = (double)n_waters / 33.4;
    x = (double)m * 0.32;                 /* start near the root       */
    for (it = 0; it < 60; ++it)
        x = x - (x * x * x - v) / (3.0 * x * x);
    b.n_waters = n_waters;
    b.m = m;
    b.box = x;
    b.a = x / (double)m;
    b.seed = seed;
    b.model = model;
    return b;
}

WN_HD uint64_t wn_splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* Counter-based uniform in [0,1): keyed (seed, frame, water, component). */
WN_HD double wn_synth_u01(uint64_t seed, uint64_t frame, uint64_t water, uint64_t comp)
{
    uint64_t h = wn_splitmix64(seed ^ 0xD1B54A32D192ED03ull);
    h = wn_splitmix64(h ^ (frame * 0x9E3779B97F4A7C15ull + 1ull));
    h = wn_splitmix64(h ^ (water * 0xC2B2AE3D27D4EB4Full + 2ull));
    h = wn_splitmix64(h ^ (comp + 3ull));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

/* Write the 9 float32 coordinates (O, H1, H2) of one water of one frame.
 * Oxygen: jittered simple-cubic lattice (+-0.15 a per axis) on lattice slot
 * (w * P) mod m^3 -- a fixed scramble so that site index order carries no
 * spatial locality and the m^3 - N vacancies are spread over the box.
 * Orientation: normalised random quaternion. */
WN_HD void wn_synth_water(const wn_synth_box* b, uint64_t frame, int32_t w, float* out9)
{
    const wn_water_model mdl = wn_synth_model(b->model);
    const uint64_t M = (uint64_t)b->m * (uint64_t)b->m * (uint64_t)b->m;
    const uint64_t slot = ((uint64_t)w * 2654435761ull) % M;
    const uint64_t ix = slot % (uint64_t)b->m;
    const uint64_t iy = (slot / (uint64_t)b->m) % (uint64_t)b->m;
    const uint64_t iz = slot / ((uint64_t)b->m * (uint64_t)b->m);
    double o[3], q[4], n2, inv, R[3][3], h1[3], h2[3];
    int c;

    o[0] = ((double)ix + 0.5 + 0.3 * (wn_synth_u01(b->seed, frame, (uint64_t)w, 0) - 0.5)) * b->a;
    o[1] = ((double)iy + 0.5 + 0.3 * (wn_synth_u01(b->seed, frame, (uint64_t)w, 1) - 0.5)) * b->a;
    o[2] = ((double)iz + 0.5 + 0.3 * (wn_synth_u01(b->seed, frame, (uint64_t)w, 2) - 0.5)) * b->a;

    n2 = 0.0;
    for (c = 0; c < 4; ++c) {
        q[c] = 2.0 * wn_synth_u01(b->seed, frame, (uint64_t)w, 3 + (uint64_t)c) - 1.0;
        n2 = n2 + q[c] * q[c];
    }
    if (n2 < 1e-12) { q[0] = 1.0; q[1] = q[2] = q[3] = 0.0; n2 = 1.0; }
    inv = 1.0 / WN_SQRT(n2);
    for (c = 0; c < 4; ++c) q[c] = q[c] * inv;

    /* rotation matrix of unit quaternion (w,x,y,z) = q[0..3] */
    R[0][0] = 1.0 - 2.0 * (q[2] * q[2] + q[3] * q[3]);
    R[0][1] = 2.0 * (q[1] * q[2] - q[0] * q[3]);
    R[0][2] = 2.0 * (q[1] * q[3] + q[0] * q[2]);
    R[1][0] = 2.0 * (q[1] * q[2] + q[0] * q[3]);
    R[1][1] = 1.0 - 2.0 * (q[1] * q[1] + q[3] * q[3]);
    R[1][2] = 2.0 * (q[2] * q[3] - q[0] * q[1]);
    R[2][0] = 2.0 * (q[1] * q[3] - q[0] * q[2]);
    R[2][1] = 2.0 * (q[2] * q[3] + q[0] * q[1]);
    R[2][2] = 1.0 - 2.0 * (q[1] * q[1] + q[2] * q[2]);

    /* molecule frame: H1 = r( sin, cos, 0), H2 = r(-sin, cos, 0) */
    {
        const double sx = mdl.r_oh * mdl.sin_half;
        const double cy = mdl.r_oh * mdl.cos_half;
        for (c = 0; c < 3; ++c) {
            h1[c] = o[c] + (R[c][0] * sx + R[c][1] * cy);
            h2[c] = o[c] + (R[c][1] * cy - R[c][0] * sx);
        }
    }
    for (c = 0; c < 3; ++c) {
        out9[c]     = (float)o[c];
        out9[3 + c] = (float)h1[c];
        out9[6 + c] = (float)h2[c];
    }
}

#ifdef __cplusplus
}
#endif
#endif /* WN_SYNTH_H */
