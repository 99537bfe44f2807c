/*
 * oracle/det_math.h -- TEST INFRASTRUCTURE (CPU oracle), not part of the product path.
 *
 * "Deterministic math" specification used by the MPPI oracle in MATH_DET mode.
 *
 * Why it exists: the reference (NVIDIA Warp kernels, thesis_master/warp_implementation/
 * projection_warp.py:236-237 `wp.cos/wp.sin`, critics_warp.py:347 `wp.exp`, and `wp.randn`
 * sampling_warp.py:73) calls libdevice sinf/cosf/expf/logf whose bit patterns cannot be
 * reproduced on a CPU.  To make "bit-exact cell indices and argmin" a checkable property
 * between a CPU oracle and a GPU kernel, both sides evaluate the SAME specified sequence of
 * IEEE-754 binary32 operations (+, *, fma, rint, bit moves).  Every operation below is
 * correctly rounded on x86-64 (SSE/FMA, FLT_EVAL_METHOD==0, -ffp-contract=off) and on
 * sm_100a (-fmad=false, explicit fmaf), so the results are bit-identical by construction.
 * Accuracy versus the true functions (<= ~1.5 ulp on the ranges used) is checked in
 * tests/test_oracle_detmath.py against float64 libm, i.e. these are as close to the true
 * function as libdevice's own implementations are.
 *
 * Polynomials: Cephes single-precision sinf/cosf/logf/expf minimax coefficients
 * (public-domain numerical recipes by S. Moshier), evaluated in Horner form with fmaf.
 */
#ifndef ORACLE_DET_MATH_H
#define ORACLE_DET_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#define DM_S1 -0x1.9943f2p-13f
#define DM_S2 0x1.11073cp-7f
#define DM_S3 -0x1.555546p-3f
#define DM_C1 0x1.99eb9cp-16f
#define DM_C2 -0x1.6c0c34p-10f
#define DM_C3 0x1.55554ap-5f
#define DM_L0 0x1.204376p-4f
#define DM_L1 -0x1.d7a37p-4f
#define DM_L2 0x1.de4a34p-4f
#define DM_L3 -0x1.fcba9ep-4f
#define DM_L4 0x1.23d37ep-3f
#define DM_L5 -0x1.555cap-3f
#define DM_L6 0x1.999d58p-3f
#define DM_L7 -0x1.fffff8p-3f
#define DM_L8 0x1.555554p-2f
#define DM_E0 0x1.a0d2cep-13f
#define DM_E1 0x1.6e879cp-10f
#define DM_E2 0x1.111210p-7f
#define DM_E3 0x1.555382p-5f
#define DM_E4 0x1.555554p-3f
#define DM_E5 0x1.0p-1f
#define DM_SQRTHF 0x1.6a09e6p-1f
#define DM_LN2_HI 0x1.63p-1f
#define DM_LN2_LO -0x1.bd0106p-13f
#define DM_LOG2E 0x1.715476p+0f
#define DM_TWO_OVER_PI 0x1.45f306p-1f
#define DM_PIO2_HI 0x1.921fb6p+0f
#define DM_PIO2_MID -0x1.777a5cp-25f
#define DM_PIO2_LO -0x1.ee59dap-50f

static inline float dm_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t dm_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* sin/cos of r for |r| <= pi/4 (reduced argument). */
static inline void dm_sincos_reduced(float r, float *s, float *c)
{
    float z = r * r;
    float sp = fmaf(DM_S1, z, DM_S2);
    sp = fmaf(sp, z, DM_S3);
    *s = fmaf(sp * z, r, r);
    float cp = fmaf(DM_C1, z, DM_C2);
    cp = fmaf(cp, z, DM_C3);
    *c = fmaf(cp * z, z, fmaf(-0.5f, z, 1.0f));
}

static inline void dm_quadrant(int q, float sr, float cr, float *s, float *c)
{
    switch (q & 3) {
    case 0: *s = sr;  *c = cr;  break;
    case 1: *s = cr;  *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
    }
}

/* sin(x), cos(x); intended for |x| < ~1e4 (3-term Cody-Waite reduction with fma). */
static inline void dm_sincosf(float x, float *s, float *c)
{
    float kf = rintf(x * DM_TWO_OVER_PI);
    float r = fmaf(kf, -DM_PIO2_HI, x);
    r = fmaf(kf, -DM_PIO2_MID, r);
    r = fmaf(kf, -DM_PIO2_LO, r);
    float sr, cr;
    dm_sincos_reduced(r, &sr, &cr);
    dm_quadrant((int)kf, sr, cr, s, c);
}

/* sin(2*pi*u), cos(2*pi*u) for u in [0,1): exact quadrant reduction (4u and 4u-k are exact). */
static inline void dm_sincos2pif(float u, float *s, float *c)
{
    float a = 4.0f * u;
    float kf = rintf(a);
    float r = (a - kf) * DM_PIO2_HI;
    float sr, cr;
    dm_sincos_reduced(r, &sr, &cr);
    dm_quadrant((int)kf, sr, cr, s, c);
}

/* natural log for positive normal x. */
static inline float dm_logf(float x)
{
    uint32_t ix = dm_as_uint(x);
    int e = (int)(ix >> 23) - 126;
    float m = dm_as_float((ix & 0x007fffffu) | 0x3f000000u); /* [0.5, 1) */
    if (m < DM_SQRTHF) { e -= 1; m = (m + m) - 1.0f; } else { m = m - 1.0f; }
    float z = m * m;
    float p = fmaf(DM_L0, m, DM_L1);
    p = fmaf(p, m, DM_L2);
    p = fmaf(p, m, DM_L3);
    p = fmaf(p, m, DM_L4);
    p = fmaf(p, m, DM_L5);
    p = fmaf(p, m, DM_L6);
    p = fmaf(p, m, DM_L7);
    p = fmaf(p, m, DM_L8);
    float fe = (float)e;
    float y = (p * m) * z;
    y = fmaf(DM_LN2_LO, fe, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    return fmaf(DM_LN2_HI, fe, r);
}

/* exp(x); returns 0 for x < -87 (below the normal range), inf for x > 88. */
static inline float dm_expf(float x)
{
    if (!(x >= -87.0f)) return (x != x) ? x : 0.0f;
    if (x > 88.0f) return INFINITY;
    float kf = rintf(x * DM_LOG2E);
    float r = fmaf(kf, -DM_LN2_HI, x);
    r = fmaf(kf, -DM_LN2_LO, r);
    float z = r * r;
    float p = fmaf(DM_E0, r, DM_E1);
    p = fmaf(p, r, DM_E2);
    p = fmaf(p, r, DM_E3);
    p = fmaf(p, r, DM_E4);
    p = fmaf(p, r, DM_E5);
    float res = fmaf(p, z, r) + 1.0f;
    int k = (int)kf;
    /* k in [-126, 127]; split the scale so that 2^k never overflows the exponent field */
    int k1 = k / 2, k2 = k - k1;
    float s1 = dm_as_float((uint32_t)(k1 + 127) << 23);
    float s2 = dm_as_float((uint32_t)(k2 + 127) << 23);
    return (res * s1) * s2;
}

/* atan(x), any x (Cephes atanf: two range reductions + degree-4 odd polynomial in x^2); atan(+-inf) = +-pi/2,
 * atan(NaN) = NaN.  Used by the goal-angle critic (critics_warp.py:37). */
#define DM_A0 0x1.49e1a2p-4f
#define DM_A1 -0x1.1c370ap-3f
#define DM_A2 0x1.9924bep-3f
#define DM_A3 -0x1.555454p-2f
#define DM_TAN3PIO8 0x1.3504f4p+1f
#define DM_TANPIO8 0x1.a8279ap-2f
#define DM_PIO4 0x1.921fb6p-1f
static inline float dm_atanf(float xx)
{
    float x = fabsf(xx), y = 0.0f;
    if (x > DM_TAN3PIO8) { y = DM_PIO2_HI; x = -(1.0f / x); }
    else if (x > DM_TANPIO8) { y = DM_PIO4; x = (x - 1.0f) / (x + 1.0f); }
    float z = x * x;
    float p = fmaf(DM_A0, z, DM_A1);
    p = fmaf(p, z, DM_A2);
    p = fmaf(p, z, DM_A3);
    float r = y + fmaf(p * z, x, x);
    return copysignf(r, xx);
}

#endif /* ORACLE_DET_MATH_H */
