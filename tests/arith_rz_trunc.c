/* Exhaustive proof of the tile-index shortcut of csrc/mppi_device.cuh (tile_rel): for every binary32 value f with
 * 0 <= f <= 32768 (the largest grid the header allows),
 *     bits(RZ(f + 2^23)) - 0x4B000000 == (int)f          and, for the non-positive operand g = -f,
 *     bits(RZ(g - 2^23)) - 0xCB000000 == -(int)g
 * i.e. one round-toward-zero FADD puts the truncated integer into the mantissa, which is what the kernel uses instead
 * of F2I.  Prints "<mismatches> <values checked>".  Build: gcc -O2 -frounding-math (the rounding mode is dynamic). */
#include <fenv.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static inline uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float fl(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int main(void)
{
    unsigned long long bad = 0, n = 0;
    volatile float big = 8388608.0f;
    fesetround(FE_TOWARDZERO);
    for (uint32_t u = 0; u <= 0x47000000u; ++u) {
        const float f = fl(u);
        const float rp = f + big;
        const float rn = -f - big;
        const int32_t t = (int32_t)f;
        bad += (int32_t)(bits(rp) - 0x4B000000u) != t;
        bad += (int32_t)(bits(rn) - 0xCB000000u) != t;
        ++n;
    }
    fesetround(FE_TONEAREST);
    printf("%llu %llu\n", bad, n);
    return bad != 0;
}
