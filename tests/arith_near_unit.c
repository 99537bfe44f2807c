/* CPU proof (by exhaustive enumeration) of the MUFU-free normalisation shortcut of the STRICT kernels
 * (csrc/mppi_device.cuh, normalize3).  For a vector whose squared norm d is within 2^-14 of 1:
 *     s  = fma(fma(-d, d, d), 0.5, d)                       == RN(sqrt(d))   for EVERY float d in the interval
 *     r0 = 2 - s ; r = fma(r0, fma(-s, r0, 1), r0) ; r = 1 + 2^-23 if s == 1 - 2^-24
 *     q  = fma(r, fma(-s, v*r, v), v*r)                     == RN(v / s)     for EVERY such s and EVERY mantissa of v
 * (division rounding depends on the mantissas only, so one binade of v covers all normal numerators).
 * Prints "bad_sqrt bad_div tested_div"; exit status 1 on any mismatch.
 * Build: gcc -O2 -ffp-contract=off -mfma  (fmaf = one correctly-rounded instruction, as FFMA on the GPU). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static uint32_t asu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float asf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int main(void)
{
    long bad_s = 0, bad_q = 0, n = 0;
    const float thr = 0x1.0p-14f;
    float s_lo = 2.0f, s_hi = 0.0f;
    for (uint32_t u = asu(1.0f - thr); u <= asu(1.0f + thr); ++u) {
        const float d = asf(u);
        const float s = fmaf(fmaf(-d, d, d), 0.5f, d);
        if (asu(s) != asu(sqrtf(d))) bad_s++;
        if (s < s_lo) s_lo = s;
        if (s > s_hi) s_hi = s;
    }
    for (uint32_t us = asu(s_lo); us <= asu(s_hi); ++us) {
        const float s = asf(us);
        const float r0 = 2.0f - s;
        float r = fmaf(r0, fmaf(-s, r0, 1.0f), r0);
        if (s == 0x1.fffffep-1f) r = 0x1.000002p+0f;
        for (uint32_t uv = asu(1.0f); uv < asu(2.0f); ++uv) {
            const float v = asf(uv), q0 = v * r, q = fmaf(r, fmaf(-s, q0, v), q0);
            bad_q += (q != v / s);
        }
        n += 1L << 23;
    }
    printf("%ld %ld %ld\n", bad_s, bad_q, n);
    return (bad_s || bad_q) ? 1 : 0;
}
