/* CPU proof (by exhaustive enumeration) of the MUFU-free normalisation shortcut of the STRICT kernels
 * (csrc/mppi_device.cuh, near_unit / renormalize3).  For a vector whose squared norm d is within 2^-15 of 1:
 *     s  = fma(fma(-d, d, d), 0.5, d)                       == RN(sqrt(d))   for EVERY float d in the window
 *     g  = fma(-0.5, d, 1.5)
 *     r  = fma(g, fma(-s, g, 1), g) ; r = 1 + 2^-23 if s == 1 - 2^-24   == RN(1 / s)
 *     q  = fma(r, fma(-s, v*g, v), v*g)                     == RN(v / s)     for EVERY such d and EVERY mantissa of v
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
    const float thr = 0x1.0p-15f;
    for (uint32_t u = asu(1.0f - thr); u <= asu(1.0f + thr); ++u) {
        const float d = asf(u);
        const float s = fmaf(fmaf(-d, d, d), 0.5f, d);
        if (asu(s) != asu(sqrtf(d))) bad_s++;
        const float g = fmaf(-0.5f, d, 1.5f);
        float r = fmaf(g, fmaf(-s, g, 1.0f), g);
        if (s == 0x1.fffffep-1f) r = 0x1.000002p+0f;
        if (r != 1.0f / s) bad_q++;
        for (uint32_t uv = asu(1.0f); uv < asu(2.0f); ++uv) {
            const float v = asf(uv), q0 = v * g, q = fmaf(r, fmaf(-s, q0, v), q0);
            bad_q += (q != v / s);
        }
        n += 1L << 23;
    }
    printf("%ld %ld %ld\n", bad_s, bad_q, n);
    return (bad_s || bad_q) ? 1 : 0;
}
