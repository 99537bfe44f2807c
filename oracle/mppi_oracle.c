/*
 * oracle/mppi_oracle.c -- TEST INFRASTRUCTURE. Scalar fp32 C restatement of ONE MPPI step of the
 * reference (thesis_master/warp_implementation; all file:line cites are relative to that dir
 * unless stated).  Written from the Warp kernels in reference operation order; every fp32
 * expression is evaluated left-to-right exactly as the Python source reads, with no FMA
 * contraction (build with -ffp-contract=off).
 *
 * Not the product: only tests/, smoke() and bench.py's CPU-baseline legs load this.
 *
 * Parity status: the reference has no tests for this path; this file is pinned against outputs of the
 * reference's own controller + kernel sources executed in the build container under oracle/warp_shim.py
 * (tests/golden/reference_mppi_steps.npz, tests/test_reference_kernels_golden.py) -- u, v, omega bit-identical,
 * everything else to ~1e-7.  warp-lang's own arithmetic (wp.randn, libdevice bit patterns) stays unpinned; see
 * oracle/README.md.
 *
 * Indexing note: the reference computes per-sample offsets in float32 (critics_warp.py:325-329,
 * `wp.float(tid)*iterations`), exact only while K*T < 2^24.  This restatement uses integers,
 * which is identical below that bound and the only well-defined reading above it.
 */
#include "mppi_oracle.h"
#include "det_math.h"

#include <stdlib.h>
#include <stdio.h>
#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------ math mode switch */
typedef struct { int det; } MathMode;

static inline void m_sincos(const MathMode *mm, float x, float *s, float *c)
{
    if (mm->det) dm_sincosf(x, s, c);
    else { *s = sinf(x); *c = cosf(x); }
}
static inline float m_exp(const MathMode *mm, float x) { return mm->det ? dm_expf(x) : expf(x); }

/* ------------------------------------------------------------------ Philox4x32-10
 * Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11); Random123 philox4x32_R(10).
 * The reference draws noise with warp-lang's wp.randn (sampling_warp.py:73,78,84,89), a third-party
 * PCG+Box-Muller that is absent from /root/reference (unpinned warp-lang, pyproject.toml:19); the
 * product replaces it with this counter-based stream and parity is anchored on injected noise. */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Box-Muller pair from two 32-bit words: log argument in (0,1], angle fraction in [0,1). */
static inline void box_muller(const MathMode *mm, uint32_t ra, uint32_t rb, float *n0, float *n1)
{
    float ua = (float)(ra >> 8) * 0x1.0p-24f + 0x1.0p-25f;
    float ub = (float)(rb >> 8) * 0x1.0p-24f;
    float s, c, lg;
    if (mm->det) { lg = dm_logf(ua); dm_sincos2pif(ub, &s, &c); }
    else { lg = logf(ua); float a = 6.283185307179586f * ub; s = sinf(a); c = cosf(a); }
    float rad = sqrtf(-2.0f * lg);
    *n0 = rad * c;
    *n1 = rad * s;
}

void oracle_philox_normals(uint64_t seed, uint64_t offset, uint32_t rover, uint32_t k0,
                           int32_t K, int32_t T, float *eps1, float *eps2, int32_t math)
{
    MathMode mm = { math == ORACLE_MATH_DET };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32) };
    for (int32_t k = 0; k < K; ++k) {
        for (int32_t pr = 0; 2 * pr < T; ++pr) {
            uint32_t ctr[4] = { k0 + (uint32_t)k, (uint32_t)pr, rover, (uint32_t)offset };
            uint32_t r[4];
            oracle_philox4x32_10(ctr, key, r);
            float a0, a1, b0, b1;
            box_muller(&mm, r[0], r[1], &a0, &a1);
            box_muller(&mm, r[2], r[3], &b0, &b1);
            int t = 2 * pr;
            eps1[(size_t)k * T + t] = a0; eps2[(size_t)k * T + t] = b0;
            if (t + 1 < T) { eps1[(size_t)k * T + t + 1] = a1; eps2[(size_t)k * T + t + 1] = b1; }
        }
    }
}

void oracle_detmath_eval(int32_t fn, const float *x, float *y0, float *y1, int32_t n)
{
    for (int32_t i = 0; i < n; ++i) {
        switch (fn) {
        case 0: dm_sincosf(x[i], &y0[i], &y1[i]); break;
        case 1: dm_sincos2pif(x[i], &y0[i], &y1[i]); break;
        case 2: y0[i] = dm_logf(x[i]); break;
        case 4: y0[i] = dm_atanf(x[i]); break;
        default: y0[i] = dm_expf(x[i]); break;
        }
    }
}

/* ------------------------------------------------------------------ small vector helpers */
typedef struct { float x, y, z; } V3;

static inline float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross3(V3 a, V3 b)
{
    V3 r = { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
    return r;
}
static inline V3 divs(V3 a, float s) { V3 r = { a.x / s, a.y / s, a.z / s }; return r; }
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

typedef struct { const OrTerrain *t; int32_t oob; } Lookup;

static inline int32_t clampi(int32_t v, int32_t lo, int32_t hi, int32_t *oob)
{
    if (v < lo) { (*oob)++; return lo; }
    if (v > hi) { (*oob)++; return hi; }
    return v;
}

/* wp.int(x) = C cast, truncation toward zero.  The cast is undefined in C for NaN and for values outside int32 (a
 * rollout that touched a no-data cell of the DEM); the GPU conversion is defined (NaN -> 0, saturation), and that
 * definition is restated here so that such runs are comparable instead of crashing the checker. */
static inline int32_t f2i(float v)
{
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}
static inline int32_t negi(int32_t v) { return (int32_t)(0u - (uint32_t)v); }   /* two's-complement wrap, as the GPU's integer negate */

/* projection_warp.py:39-40 -- cell index of (x, y).  x_min = y_min = -half_width (MPPI_isaac.py:584-585). */
static inline void dem_index(const OrTerrain *t, float x, float y, int32_t *i, int32_t *j)
{
    float x_min = -t->half_width, y_min = -t->half_width;
    *i = f2i((x - x_min) / t->res);
    *j = negi(f2i((y + y_min) / t->res));
}

/* projection_warp.py:8-48 */
static inline void corners(Lookup *L, float x, float y, float q[2][2], int32_t *io, int32_t *jo)
{
    int32_t i, j;
    dem_index(L->t, x, y, &i, &j);
    *io = i; *jo = j;
    int32_t gs = L->t->gs;
    i = clampi(i, 0, gs - 2, &L->oob);
    j = clampi(j, 0, gs - 2, &L->oob);
    const float *Z = L->t->dem;
    q[0][0] = Z[(size_t)j * gs + i];
    q[0][1] = Z[(size_t)j * gs + i + 1];
    q[1][0] = Z[(size_t)(j + 1) * gs + i];
    q[1][1] = Z[(size_t)(j + 1) * gs + i + 1];
}

/* projection_warp.py:70-100 */
static inline float bilinear(float x, float y, float q[2][2], float res)
{
    float xn = x / res, yn = y / res;
    float x2 = xn - truncf(xn);
    float y2 = yn - truncf(yn);
    return (1.0f - x2) * (1.0f - y2) * q[0][0] + x2 * (1.0f - y2) * q[1][0]
         + (1.0f - x2) * y2 * q[0][1] + x2 * y2 * q[1][1];
}

/* projection_warp.py:129-151 */
static inline V3 normal_on_grid(float q[2][2], float res)
{
    float vx = -res / 2.0f * (q[0][1] - q[0][0] - q[1][0] + q[1][1]);
    float vy = -res / 2.0f * (q[1][0] - q[0][0] - q[0][1] + q[1][1]);
    float vz = res * res;
    float norm = sqrtf(vx * vx + vy * vy + vz * vz);
    V3 v = { vx, vy, vz };
    return divs(v, norm);
}

/* projection_warp.py:168-190 */
static inline V3 tangent(V3 n, V3 prev)
{
    float d = dot3(prev, n);
    V3 p = { prev.x - d * n.x, prev.y - d * n.y, prev.z - d * n.z };
    float norm = sqrtf(dot3(p, p));
    return divs(p, norm);
}

/* projection_warp.py:207-223 */
static inline void update_position(float *x, float *y, V3 h, float v, float dt)
{
    h = divs(h, sqrtf(dot3(h, h)));
    float dx = h.x * v * dt, dy = h.y * v * dt;
    *x = *x + dx;
    *y = *y + dy;
}

/* projection_warp.py:225-248 */
static inline V3 update_orientation(const MathMode *mm, V3 h, float w, V3 n, float dt)
{
    h = divs(h, sqrtf(dot3(h, h)));
    float angle = w * dt;
    float s, c;
    m_sincos(mm, angle, &s, &c);
    V3 cr = cross3(n, h);
    float d = dot3(n, h);
    float omc = 1.0f - c;
    V3 r = { h.x * c + cr.x * s + n.x * d * omc,
             h.y * c + cr.y * s + n.y * d * omc,
             h.z * c + cr.z * s + n.z * d * omc };
    return divs(r, sqrtf(dot3(r, r)));
}

/* projection_warp.py:251-275 */
static inline V3 update_orientation_2d(const MathMode *mm, V3 h, float w, float dt)
{
    float th = w * dt, s, c;
    m_sincos(mm, th, &s, &c);
    float nx = c * h.x - s * h.y;
    float ny = s * h.x + c * h.y;
    float norm = sqrtf(nx * nx + ny * ny);
    if (norm > 0.0f) { nx /= norm; ny /= norm; }
    V3 r = { nx, ny, 0.0f };
    return r;
}

/* sampling_warp.py:96-138 -- first-order wheel lag + differential drive, sequential in t. */
static void wheel_filter(const OrParams *p, const OrState *st, float k, float a, int T,
                         const float *u1, const float *u2, float *v, float *w)
{
    float l = st->wheel_l, r = st->wheel_r;
    for (int t = 0; t < T; ++t) {
        l = l * a + u1[t] * k * (1.0f - a);
        r = r * a + u2[t] * k * (1.0f - a);
        v[t] = clampf((l + r) / 2.0f, p->v_min, p->v_max);
        w[t] = clampf((-l + r) / p->r_wheels, p->w_min, p->w_max);
    }
}

/* projection_warp.py:284-350 (3-D) and :353-382 (2-D) for one sample.  traj/heading/lw/rw are [T*3]. */
static void rollout(const OrParams *p, Lookup *L, const OrState *st, const MathMode *mm,
                    const float *v, const float *w,
                    float *traj, float *heading, float *lw, float *rw,
                    int32_t *dem_ij, int32_t *lw_ij, int32_t *rw_ij)
{
    const OrTerrain *t = L->t;
    int T = p->T;
    float x = st->x, y = st->y;
    V3 prev0 = { st->hx, st->hy, st->hz };
    float q[2][2];
    int32_t ii, jj;
    if (p->proj == 3) {
        corners(L, x, y, q, &ii, &jj);
        float height = bilinear(x, y, q, t->res);
        (void)height;
        V3 n = normal_on_grid(q, t->res);
        V3 prev = tangent(n, prev0);
        for (int k = 0; k < T; ++k) {
            update_position(&x, &y, prev, v[k], p->dt);
            corners(L, x, y, q, &ii, &jj);
            height = bilinear(x, y, q, t->res);
            n = normal_on_grid(q, t->res);
            prev = tangent(n, prev);
            V3 cur = update_orientation(mm, prev, w[k], n, p->dt);
            heading[3 * k] = cur.x; heading[3 * k + 1] = cur.y; heading[3 * k + 2] = cur.z;
            traj[3 * k] = x; traj[3 * k + 1] = y; traj[3 * k + 2] = height;
            if (dem_ij) { dem_ij[2 * k] = ii; dem_ij[2 * k + 1] = jj; }
            /* wheel points, projection_warp.py:332-348 (nearest cell, no interpolation) */
            V3 cr = cross3(n, cur);
            float rx = p->wheel_offset * cr.x, ry = p->wheel_offset * cr.y;
            int32_t gs = t->gs, wi, wj, ci, cj;
            float xw = x + rx, yw = y + ry;
            dem_index(t, xw, yw, &wi, &wj);
            if (lw_ij) { lw_ij[2 * k] = wi; lw_ij[2 * k + 1] = wj; }
            ci = clampi(wi, 0, gs - 1, &L->oob); cj = clampi(wj, 0, gs - 1, &L->oob);
            lw[3 * k] = xw; lw[3 * k + 1] = yw; lw[3 * k + 2] = t->dem[(size_t)cj * gs + ci];
            xw = x - rx; yw = y - ry;
            dem_index(t, xw, yw, &wi, &wj);
            if (rw_ij) { rw_ij[2 * k] = wi; rw_ij[2 * k + 1] = wj; }
            ci = clampi(wi, 0, gs - 1, &L->oob); cj = clampi(wj, 0, gs - 1, &L->oob);
            rw[3 * k] = xw; rw[3 * k + 1] = yw; rw[3 * k + 2] = t->dem[(size_t)cj * gs + ci];
            prev = cur;
        }
    } else {
        V3 prev = prev0;
        for (int k = 0; k < T; ++k) {
            update_position(&x, &y, prev, v[k], p->dt);
            V3 cur = update_orientation_2d(mm, prev, w[k], p->dt);
            heading[3 * k] = cur.x; heading[3 * k + 1] = cur.y; heading[3 * k + 2] = cur.z;
            corners(L, x, y, q, &ii, &jj);
            float height = bilinear(x, y, q, t->res);
            traj[3 * k] = x; traj[3 * k + 1] = y; traj[3 * k + 2] = height;
            if (dem_ij) { dem_ij[2 * k] = ii; dem_ij[2 * k + 1] = jj; }
            /* lw / rw are never written by the 2-D kernel: they stay at their zero initial value
             * (MPPI_isaac.py:482-483), so the slope critic sees all-zero wheel points. */
            lw[3 * k] = lw[3 * k + 1] = lw[3 * k + 2] = 0.0f;
            rw[3 * k] = rw[3 * k + 1] = rw[3 * k + 2] = 0.0f;
            if (lw_ij) { lw_ij[2 * k] = lw_ij[2 * k + 1] = 0; }
            if (rw_ij) { rw_ij[2 * k] = rw_ij[2 * k + 1] = 0; }
            prev = cur;
        }
    }
}

/* critics_warp.py:85-127 */
static float path_follow(const OrParams *p, const OrState *st, const float *traj)
{
    int T = p->T;
    float x_diff = st->goal_x - st->x, y_diff = st->goal_y - st->y;
    float dist = sqrtf(x_diff * x_diff + y_diff * y_diff);
    const float *last = traj + 3 * (T - 1);
    float cost = 0.0f;
    if (dist > p->horizon) {
        float igx = st->x + x_diff * p->horizon / (dist + p->pf_eps);
        float igy = st->y + y_diff * p->horizon / (dist + p->pf_eps);
        /* wp.pow(., 1.0) is the identity */
        cost = (last[0] - igx) * (last[0] - igx) + (last[1] - igy) * (last[1] - igy);
        return cost * (1.0f + 2.0f * p->horizon / dist);
    }
    for (int i = 0; i < T - 1; ++i)
        cost += p->pf_near_gain * (fabsf(traj[3 * i] - st->goal_x) + fabsf(traj[3 * i + 1] - st->goal_y));
    return cost;
}

/* critics_warp.py:168-218 */
static float avoid_slope_wheels(const OrParams *p, const float *lw, const float *rw)
{
    int T = p->T;
    float total = 0.0f;
    for (int i = 0; i < T - 3; i += 2) {
        const float *cl = lw + 3 * (i + 2), *pl = lw + 3 * i;
        const float *cr = rw + 3 * (i + 2), *pr = rw + 3 * i;
        float dz_l = cl[2] - pl[2];
        float d_l = sqrtf((cl[0] - pl[0]) * (cl[0] - pl[0]) + (cl[1] - pl[1]) * (cl[1] - pl[1]));
        float dz_r = cr[2] - pr[2];
        float d_r = sqrtf((cr[0] - pr[0]) * (cr[0] - pr[0]) + (cr[1] - pr[1]) * (cr[1] - pr[1]));
        float ratio_l = fabsf(dz_l / (d_l + p->slope_eps));
        float ratio_r = fabsf(dz_r / (d_r + p->slope_eps));
        float ls = (1.0f + p->slope_gain * ratio_l) * (1.0f + p->slope_gain * ratio_l);
        float rs = (1.0f + p->slope_gain * ratio_r) * (1.0f + p->slope_gain * ratio_r);
        if (ls > rs) total += ls; else total += rs;
    }
    return total;
}

/* critics_warp.py:269-300 */
static float maximise_speed(const OrParams *p, const OrState *st, const float *v)
{
    float x_diff = st->goal_x - st->x, y_diff = st->goal_y - st->y;
    float dist = sqrtf(x_diff * x_diff + y_diff * y_diff);
    if (dist < p->near_goal_cut) return 0.0f;
    float acc = 0.0f;
    for (int i = 0; i < p->T; ++i) acc += (p->target_speed - v[i]) / (v[i] + p->speed_eps);
    return acc;
}

/* critics_warp.py:220-267 */
static float avoid_obstacle(const OrParams *p, Lookup *L, const float *traj, int32_t *cm_ij)
{
    const OrTerrain *t = L->t;
    float acc = 0.0f;
    for (int i = 0; i < p->T; ++i) {
        float idx_x = (traj[3 * i] + t->half_width) / t->cres;
        float idx_y = (-traj[3 * i + 1] + t->half_width) / t->cres;
        int32_t ix = f2i(idx_x), iy = f2i(idx_y);
        if (cm_ij) { cm_ij[2 * i] = ix; cm_ij[2 * i + 1] = iy; }
        ix = clampi(ix, 0, t->cms - 1, &L->oob);
        iy = clampi(iy, 0, t->cms - 1, &L->oob);
        float c = t->costmap[(size_t)ix + (size_t)t->cms * iy];
        if (c > p->lethal_thresh) acc += p->lethal_penalty;
        acc += c;
    }
    return acc;
}

/* ---- optional critics (weight 0 by default) ---- */

/* critics_warp.py:44-83: penalise a last segment that points away from the goal. */
static float path_orientation(const OrParams *p, const OrState *st, const float *traj)
{
    int T = p->T;
    if (T < 2) return 0.0f;                       /* the reference would read the previous sample's last point */
    float x_diff = st->goal_x - st->x, y_diff = st->goal_y - st->y;
    const float *pen = traj + 3 * (T - 2), *last = traj + 3 * (T - 1);
    float x_diff2 = last[0] - pen[0], y_diff2 = last[1] - pen[1];
    float sp = x_diff * x_diff2 + y_diff * y_diff2;
    if (sp <= 0.0f) return -sp / (fabsf(x_diff) + fabsf(y_diff));
    return 0.0f;
}

/* critics_warp.py:131-166: stride-2 slope of the body path (same form as the wheel critic, on trajectory[...].z) */
static float avoid_slope_path(const OrParams *p, const float *traj)
{
    int T = p->T;
    float total = 0.0f;
    for (int i = 0; i < T - 3; i += 2) {
        const float *c = traj + 3 * (i + 2), *q = traj + 3 * i;
        float dz = c[2] - q[2];
        float d = sqrtf((c[0] - q[0]) * (c[0] - q[0]) + (c[1] - q[1]) * (c[1] - q[1]));
        float ratio = fabsf(dz / (d + p->slope_eps));
        total += (1.0f + p->slope_gain * ratio) * (1.0f + p->slope_gain * ratio);
    }
    return total;
}

/* critics_warp.py:5-41: near the goal, |atan(dy/dx of the last segment) - goal_orientation| */
static float goal_angle(const OrParams *p, const OrState *st, const MathMode *mm, const float *traj)
{
    int T = p->T;
    if (T < 2) return 0.0f;
    float dist = sqrtf((st->x - st->goal_x) * (st->x - st->goal_x) + (st->y - st->goal_y) * (st->y - st->goal_y));
    if (dist < p->goal_angle_radius) {
        const float *pen = traj + 3 * (T - 2), *last = traj + 3 * (T - 1);
        float q = (last[1] - pen[1]) / (last[0] - pen[0]);
        float a = mm->det ? dm_atanf(q) : atanf(q);
        return fabsf(a - st->goal_theta);
    }
    return 0.0f;
}

/* extensions (no reference counterpart; BASELINE configuration 5 names roll / pitch critics) */
static float roll_critic(const OrParams *p, const float *lw, const float *rw)
{
    float acc = 0.0f, track = 2.0f * p->wheel_offset;
    for (int i = 0; i < p->T; i += 2) {
        float r = (lw[3 * i + 2] - rw[3 * i + 2]) / track;
        acc += r * r;
    }
    return acc;
}

static float pitch_critic(const OrParams *p, const float *heading)
{
    float acc = 0.0f;
    for (int i = 0; i < p->T; i += 2) acc += heading[3 * i + 2] * heading[3 * i + 2];
    return acc;
}

static float effort_critic(const OrParams *p, const float *u1, const float *u2)
{
    float acc = 0.0f;
    for (int i = 0; i < p->T; ++i) acc += u1[i] * u1[i] + u2[i] * u2[i];
    return acc;
}

int oracle_num_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* One worker = one contiguous slice of samples (static partition; results do not depend on it). */
typedef struct {
    const OrParams *p; const OrTerrain *ter; const OrState *st;
    const float *nom1, *nom2, *eps1, *eps2;
    OrDump *dump; float *cost, *u1, *u2;
    int k_begin, k_end; int32_t oob;
} Worker;

static void *worker_main(void *arg)
{
    Worker *W = (Worker *)arg;
    const OrParams *p = W->p; const OrState *st = W->st; OrDump *dump = W->dump;
    float *u1 = W->u1, *u2 = W->u2, *cost = W->cost;
    const float *nom1 = W->nom1, *nom2 = W->nom2, *eps1 = W->eps1, *eps2 = W->eps2;
    const int T = p->T;
    MathMode mm = { p->math == ORACLE_MATH_DET };
    float *vloc = (float *)malloc(sizeof(float) * T * 2);
    float *tloc = (float *)malloc(sizeof(float) * T * 12);
    Lookup L = { W->ter, 0 };
    for (int k = W->k_begin; k < W->k_end; ++k) {
        size_t o = (size_t)k * T;
        /* A.1 sampling_warp.py:54-92 (receding-horizon shift folded in) */
        for (int t = 0; t < T; ++t) {
            int src = (t != T - 1) ? t + 1 : t;
            if (p->input_model == 1) {   /* velocity space, sampling_warp.py:10-48 */
                u1[o + t] = clampf(nom1[src] + st->sigma1 * eps1[o + t], p->v_min, p->v_max);
                u2[o + t] = clampf(nom2[src] + st->sigma2 * eps2[o + t], p->w_min, p->w_max);
            } else {
                u1[o + t] = clampf(nom1[src] + st->sigma1 * eps1[o + t], p->u1_min, p->u1_max);
                u2[o + t] = clampf(nom2[src] + st->sigma2 * eps2[o + t], p->u2_min, p->u2_max);
            }
        }
        float *v = dump->v ? dump->v + o : vloc;
        float *w = dump->w ? dump->w + o : vloc + T;
        if (p->input_model == 1) {
            for (int t = 0; t < T; ++t) { v[t] = u1[o + t]; w[t] = u2[o + t]; }
        } else {
            wheel_filter(p, st, p->filt_k, p->filt_a, T, u1 + o, u2 + o, v, w);
        }
        float *traj = dump->traj ? dump->traj + 3 * o : tloc;
        float *hd = dump->heading ? dump->heading + 3 * o : tloc + 3 * T;
        float *lw = dump->lw ? dump->lw + 3 * o : tloc + 6 * T;
        float *rw = dump->rw ? dump->rw + 3 * o : tloc + 9 * T;
        rollout(p, &L, st, &mm, v, w, traj, hd, lw, rw,
                dump->dem_ij ? dump->dem_ij + 2 * o : NULL,
                dump->lw_ij ? dump->lw_ij + 2 * o : NULL,
                dump->rw_ij ? dump->rw_ij + 2 * o : NULL);
        /* critics_warp.py:302-329: four `costs[tid] +=` on a zeroed accumulator, in this order */
        float c_path = path_follow(p, st, traj);
        float c_slope = avoid_slope_wheels(p, lw, rw);
        float c_speed = maximise_speed(p, st, v);
        float c_obs = avoid_obstacle(p, &L, traj, dump->cm_ij ? dump->cm_ij + 2 * o : NULL);
        const int want_ext = dump->critics_ext != NULL;
        float x_orient = (want_ext || p->cw_orient != 0.0f) ? path_orientation(p, st, traj) : 0.0f;
        float x_slope = (want_ext || p->cw_slope_path != 0.0f) ? avoid_slope_path(p, traj) : 0.0f;
        float x_angle = (want_ext || p->cw_goal_angle != 0.0f) ? goal_angle(p, st, &mm, traj) : 0.0f;
        float x_roll = (want_ext || p->cw_roll != 0.0f) ? roll_critic(p, lw, rw) : 0.0f;
        float x_pitch = (want_ext || p->cw_pitch != 0.0f) ? pitch_critic(p, hd) : 0.0f;
        float x_effort = (want_ext || p->cw_effort != 0.0f) ? effort_critic(p, u1 + o, u2 + o) : 0.0f;
        float c = 0.0f;
        if (p->cw_orient != 0.0f) c += p->cw_orient * x_orient;             /* :324 */
        c += p->cw_path * c_path;
        if (p->cw_slope_path != 0.0f) c += p->cw_slope_path * x_slope;      /* :326 */
        c += p->cw_slope * c_slope;
        c += p->cw_speed * c_speed;
        c += p->cw_obs * c_obs;
        if (p->cw_goal_angle != 0.0f) c += p->cw_goal_angle * x_angle;
        if (p->cw_roll != 0.0f) c += p->cw_roll * x_roll;
        if (p->cw_pitch != 0.0f) c += p->cw_pitch * x_pitch;
        if (p->cw_effort != 0.0f) c += p->cw_effort * x_effort;
        cost[k] = c;
        if (want_ext) {
            float *e = dump->critics_ext + 6 * (size_t)k;
            e[0] = x_orient; e[1] = x_slope; e[2] = x_angle; e[3] = x_roll; e[4] = x_pitch; e[5] = x_effort;
        }
        if (dump->critics) {
            dump->critics[4 * k] = c_path; dump->critics[4 * k + 1] = c_slope;
            dump->critics[4 * k + 2] = c_speed; dump->critics[4 * k + 3] = c_obs;
        }
    }
    W->oob = L.oob;
    free(vloc); free(tloc);
    return NULL;
}

int oracle_mppi_step(const OrParams *p, const OrTerrain *ter, const OrState *st,
                     const float *nom1, const float *nom2,
                     const float *eps1, const float *eps2,
                     OrDump *dump, OrOut *out, int32_t nthreads)
{
    const int K = p->K, T = p->T;
    if (K <= 0 || T < 2 || !ter->dem || !ter->costmap || !eps1 || !eps2) return -1;
    MathMode mm = { p->math == ORACLE_MATH_DET };
    OrDump nodump = { 0 };
    if (!dump) dump = &nodump;

    float *cost = dump->cost ? dump->cost : (float *)malloc(sizeof(float) * K);
    float *u1 = dump->u1 ? dump->u1 : (float *)malloc(sizeof(float) * (size_t)K * T);
    float *u2 = dump->u2 ? dump->u2 : (float *)malloc(sizeof(float) * (size_t)K * T);
    int32_t oob_total = 0;
    if (nthreads <= 0) nthreads = oracle_num_threads();
    if (nthreads > K) nthreads = K;
    if (nthreads > 256) nthreads = 256;
    {
        Worker W[256]; pthread_t th[256];
        for (int i = 0; i < nthreads; ++i) {
            Worker w0 = { p, ter, st, nom1, nom2, eps1, eps2, dump, cost, u1, u2,
                          (int)((long long)K * i / nthreads), (int)((long long)K * (i + 1) / nthreads), 0 };
            W[i] = w0;
        }
        if (nthreads == 1) worker_main(&W[0]);
        else {
            for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, worker_main, &W[i]);
            for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
        }
        for (int i = 0; i < nthreads; ++i) oob_total += W[i].oob;
    }

    /* A.8 update (critics_warp.py:338-376), race-free intent per old_files/run_mppi.py:222-226 */
    float m = INFINITY; int32_t arg = 0;
    for (int k = 0; k < K; ++k) if (cost[k] < m) { m = cost[k]; arg = k; }
    float S = 0.0f; double S64 = 0.0;
    float *wts = dump->weights ? dump->weights : (float *)malloc(sizeof(float) * K);
    for (int k = 0; k < K; ++k) {
        /* a NaN cost (possible only with the optional goal-angle / orientation critics: atan(0 / 0) for a sample that
         * stands still) would poison every sum in the reference; the product gives such a sample zero weight and
         * counts it (stats[4]) -- restated here so that the rule is checkable */
        float nc = ((cost[k] != cost[k]) ? INFINITY : cost[k]) - m;
        wts[k] = m_exp(&mm, -nc / p->lambda);
        S += wts[k]; S64 += (double)wts[k];
    }
    for (int t = 0; t < T; ++t) {
        float a1 = 0.0f, a2 = 0.0f; double d1 = 0.0, d2 = 0.0;
        for (int k = 0; k < K; ++k) {
            if (wts[k] == 0.0f) continue; /* adds exactly +0 */
            a1 += wts[k] * u1[(size_t)k * T + t] / S;
            a2 += wts[k] * u2[(size_t)k * T + t] / S;
            d1 += (double)wts[k] * (double)u1[(size_t)k * T + t];
            d2 += (double)wts[k] * (double)u2[(size_t)k * T + t];
        }
        /* no valid sample (every cost NaN / +inf): the product keeps the previous nominal (the reference would
         * store 0 / 0) */
        if (!(S > 0.0f)) { a1 = nom1[t]; a2 = nom2[t]; }
        out->nominal1[t] = a1; out->nominal2[t] = a2;
        if (out->nominal1_f64) out->nominal1_f64[t] = d1 / S64;
        if (out->nominal2_f64) out->nominal2_f64[t] = d2 / S64;
    }
    out->min_cost = m; out->argmin = arg; out->weights_sum = S;

    /* A.9: optimal sequence -> (v*, w*) with (k, a) = (opt_k, opt_a)  MPPI_isaac.py:672-692 */
    if (p->input_model == 1) {   /* the weighted (v, w) sequence IS the optimal velocity sequence (old_files/run_mppi.py:228-245) */
        for (int t = 0; t < T; ++t) { out->opt_v[t] = out->nominal1[t]; out->opt_w[t] = out->nominal2[t]; }
    } else {
        wheel_filter(p, st, p->opt_k, p->opt_a, T, out->nominal1, out->nominal2, out->opt_v, out->opt_w);
    }
    /* optimal-trajectory rollout, dim = 1  MPPI_isaac.py:696-720 (always the 3-D kernel) */
    if (out->sim_traj && out->sim_heading) {
        OrParams p3 = *p; p3.proj = 3;
        float *tmp = (float *)malloc(sizeof(float) * T * 6);
        Lookup L = { ter, 0 };
        rollout(&p3, &L, st, &mm, out->opt_v, out->opt_w, out->sim_traj, out->sim_heading, tmp, tmp + 3 * T,
                NULL, NULL, NULL);
        oob_total += L.oob;
        free(tmp);
    }
    out->oob_clamps = oob_total;

    if (!dump->cost) free(cost);
    if (!dump->u1) free(u1);
    if (!dump->u2) free(u2);
    if (!dump->weights) free(wts);
    return 0;
}

void oracle_combine_partials(const float *parts, int32_t G, int32_t T, float lambda, int32_t math,
                             float *nominal1, float *nominal2, float *min_cost, int32_t *argmin, float *wsum)
{
    MathMode mm = { math == ORACLE_MATH_DET };
    const int stride = 3 + 2 * T;
    float M = INFINITY; int32_t arg = -1;
    for (int g = 0; g < G; ++g) {
        const float *pg = parts + (size_t)g * stride;
        if (pg[0] < M) { M = pg[0]; int32_t a; memcpy(&a, &pg[2], 4); arg = a; }
    }
    float S = 0.0f;
    for (int g = 0; g < G; ++g) {
        const float *pg = parts + (size_t)g * stride;
        float sc = m_exp(&mm, -(pg[0] - M) / lambda);
        S += pg[1] * sc;
    }
    for (int t = 0; t < T; ++t) {
        float a1 = 0.0f, a2 = 0.0f;
        for (int g = 0; g < G; ++g) {
            const float *pg = parts + (size_t)g * stride;
            float sc = m_exp(&mm, -(pg[0] - M) / lambda);
            a1 += pg[3 + t] * sc;
            a2 += pg[3 + T + t] * sc;
        }
        nominal1[t] = a1 / S; nominal2[t] = a2 / S;
    }
    *min_cost = M; *argmin = arg; *wsum = S;
}
