// det_math.cuh -- specified ("det") fp32 transcendental functions of the STRICT flavour.
//
// The reference's Warp kernels call libdevice sinf/cosf (projection_warp.py:236-237), expf
// (critics_warp.py:347) and, inside wp.randn, logf/cosf (sampling_warp.py:73).  Their bit patterns
// cannot be reproduced on a CPU, so the STRICT flavour evaluates a *specified* sequence of
// correctly-rounded IEEE-754 binary32 operations instead (Cephes minimax polynomials, Horner form,
// explicit fmaf; Cody-Waite reductions).  The CPU oracle restates the same specification
// independently (oracle/det_math.h), which makes whole rollouts bit-comparable.  Accuracy vs the true
// functions is <= 1.6 ulp on the ranges used (tests/test_oracle_detmath.py).  A side benefit on
// sm_100a: no slow-path branches, ~12 FP32 instructions per sincos.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mppi {
namespace dm {

__device__ __forceinline__ void sincos_reduced(float r, float& s, float& c)
{
    const float z = r * r;
    float sp = fmaf(-0x1.9943f2p-13f, z, 0x1.11073cp-7f);
    sp = fmaf(sp, z, -0x1.555546p-3f);
    s = fmaf(sp * z, r, r);
    float cp = fmaf(0x1.99eb9cp-16f, z, -0x1.6c0c34p-10f);
    cp = fmaf(cp, z, 0x1.55554ap-5f);
    c = fmaf(cp * z, z, fmaf(-0.5f, z, 1.0f));
}

__device__ __forceinline__ void quadrant(int q, float sr, float cr, float& s, float& c)
{
    q &= 3;
    const float a = (q & 1) ? cr : sr;      // |sin| source
    const float b = (q & 1) ? sr : cr;      // |cos| source
    s = (q & 2) ? -a : a;                   // q: 0 s | 1 c | 2 -s | 3 -c
    c = (q == 1 || q == 2) ? -b : b;        // q: 0 c | 1 -s | 2 -c | 3 s
}

// sin(x), cos(x) for |x| < ~1e4.
__device__ __forceinline__ void sincosf_det(float x, float& s, float& c)
{
    const float kf = rintf(x * 0x1.45f306p-1f);
    float r = fmaf(kf, -0x1.921fb6p+0f, x);
    r = fmaf(kf, 0x1.777a5cp-25f, r);
    r = fmaf(kf, 0x1.ee59dap-50f, r);
    float sr, cr;
    sincos_reduced(r, sr, cr);
    quadrant((int)kf, sr, cr, s, c);
}

// sin(2*pi*u), cos(2*pi*u) for u in [0, 1).
__device__ __forceinline__ void sincos2pif_det(float u, float& s, float& c)
{
    const float a = 4.0f * u;
    const float kf = rintf(a);
    const float r = (a - kf) * 0x1.921fb6p+0f;
    float sr, cr;
    sincos_reduced(r, sr, cr);
    quadrant((int)kf, sr, cr, s, c);
}

// natural log, positive normal x.
__device__ __forceinline__ float logf_det(float x)
{
    const uint32_t ix = __float_as_uint(x);
    int e = (int)(ix >> 23) - 126;
    float m = __uint_as_float((ix & 0x007fffffu) | 0x3f000000u);
    if (m < 0x1.6a09e6p-1f) { e -= 1; m = (m + m) - 1.0f; } else { m = m - 1.0f; }
    const float z = m * m;
    float p = fmaf(0x1.204376p-4f, m, -0x1.d7a37p-4f);
    p = fmaf(p, m, 0x1.de4a34p-4f);
    p = fmaf(p, m, -0x1.fcba9ep-4f);
    p = fmaf(p, m, 0x1.23d37ep-3f);
    p = fmaf(p, m, -0x1.555cap-3f);
    p = fmaf(p, m, 0x1.999d58p-3f);
    p = fmaf(p, m, -0x1.fffff8p-3f);
    p = fmaf(p, m, 0x1.555554p-2f);
    const float fe = (float)e;
    float y = (p * m) * z;
    y = fmaf(-0x1.bd0106p-13f, fe, y);
    y = fmaf(-0.5f, z, y);
    const float r = m + y;
    return fmaf(0x1.63p-1f, fe, r);
}

// exp(x): 0 below -87 (sub-normal results are flushed by specification), +inf above 88.
__device__ __forceinline__ float expf_det(float x)
{
    if (!(x >= -87.0f)) return (x != x) ? x : 0.0f;
    if (x > 88.0f) return __int_as_float(0x7f800000);
    const float kf = rintf(x * 0x1.715476p+0f);
    float r = fmaf(kf, -0x1.63p-1f, x);
    r = fmaf(kf, 0x1.bd0106p-13f, r);
    const float z = r * r;
    float p = fmaf(0x1.a0d2cep-13f, r, 0x1.6e879cp-10f);
    p = fmaf(p, r, 0x1.111210p-7f);
    p = fmaf(p, r, 0x1.555382p-5f);
    p = fmaf(p, r, 0x1.555554p-3f);
    p = fmaf(p, r, 0x1.0p-1f);
    const float res = fmaf(p, z, r) + 1.0f;
    const int k = (int)kf;
    const int k1 = k / 2, k2 = k - k1;
    const float s1 = __uint_as_float((uint32_t)(k1 + 127) << 23);
    const float s2 = __uint_as_float((uint32_t)(k2 + 127) << 23);
    return (res * s1) * s2;
}

// atan(x), any x: atan(+-inf) = +-pi/2, atan(NaN) = NaN (goal-angle critic, critics_warp.py:37).  Evaluated once per
// sample at most, so the two IEEE divisions are spelled __fdiv_rn whatever the flavour's division is.
__device__ __forceinline__ float atanf_det(float xx)
{
    float x = fabsf(xx), y = 0.0f;
    if (x > 0x1.3504f4p+1f) { y = 0x1.921fb6p+0f; x = -__fdiv_rn(1.0f, x); }
    else if (x > 0x1.a8279ap-2f) { y = 0x1.921fb6p-1f; x = __fdiv_rn(x - 1.0f, x + 1.0f); }
    const float z = x * x;
    float p = fmaf(0x1.49e1a2p-4f, z, -0x1.1c370ap-3f);
    p = fmaf(p, z, 0x1.9924bep-3f);
    p = fmaf(p, z, -0x1.555454p-2f);
    const float r = y + fmaf(p * z, x, x);
    return copysignf(r, xx);
}

}  // namespace dm
}  // namespace mppi
