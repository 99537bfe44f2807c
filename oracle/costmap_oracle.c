/*
 * oracle/costmap_oracle.c -- TEST INFRASTRUCTURE.  CPU restatement of the reference's obstacle-costmap builder
 * Surface.create_obstacles_costmap (thesis_master/warp_implementation/MPPI_isaac.py:361-378):
 *
 *     obs = 255 everywhere; for every rock: cells with (Xc - x_local)^2 + (Yc - y_local)^2 <= R^2 become 0,
 *           x_local = y_global - y0, y_local = x_global - x0, R = r/2 + r_robot + 0.1        (:365-372, float64)
 *     d   = cv2.distanceTransform(obs, cv2.DIST_L2, 5)                                        (:374)
 *     d   = cv2.normalize(d, None, 0, 1.0, cv2.NORM_MINMAX)                                   (:375)
 *     c   = (1 - d) ** 20                                                                     (:376)
 *
 * cv2.distanceTransform is third-party arithmetic (OpenCV, 4.13 in this image; not under /root/reference).  Its
 * published algorithm for DIST_L2 with a 5x5 mask is Borgefors' two-pass chamfer transform with the weights a = 1,
 * b = 1.4, c = 2.1969 (OpenCV imgproc/src/distransform.cpp, distanceTransform_5x5; 4.x accumulates the path length in
 * float32: cv2 returns exactly 1.4f and 2.1969f for the diagonal and knight neighbours, which 16.16 fixed point would
 * not); restated here and pinned against cv2 itself in tests/test_costmap.py (bit-identical distance maps).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define INIT_DIST0 3.402823466e+38f   /* FLT_MAX: x + 2.1969f stays FLT_MAX */

/* src: [n][n] uint8 (0 = obstacle), dist: [n][n] float.  Returns 0. */
int oracle_chamfer5x5(const uint8_t *src, int rows, int cols, float *dist)
{
    const int B = 2, step = cols + 2 * B;
    const float HV = 1.0f, DG = 1.4f, LG = 2.1969f;
    float *temp = (float *)malloc(sizeof(float) * (size_t)(rows + 2 * B) * step);
    if (!temp) return -1;
    for (size_t i = 0; i < (size_t)(rows + 2 * B) * step; ++i) temp[i] = INIT_DIST0;
    /* forward pass: row-major, neighbours above and to the left */
    for (int i = 0; i < rows; ++i) {
        const uint8_t *s = src + (size_t)i * cols;
        float *t = temp + (size_t)(i + B) * step + B;
        for (int j = 0; j < cols; ++j) {
            if (!s[j]) { t[j] = 0; continue; }
            float t0 = t[j - step * 2 - 1] + LG, v;
            v = t[j - step * 2 + 1] + LG; if (t0 > v) t0 = v;
            v = t[j - step - 2] + LG; if (t0 > v) t0 = v;
            v = t[j - step - 1] + DG; if (t0 > v) t0 = v;
            v = t[j - step] + HV; if (t0 > v) t0 = v;
            v = t[j - step + 1] + DG; if (t0 > v) t0 = v;
            v = t[j - step + 2] + LG; if (t0 > v) t0 = v;
            v = t[j - 1] + HV; if (t0 > v) t0 = v;
            t[j] = t0;
        }
    }
    /* backward pass: neighbours below and to the right; fixed point -> float */
    for (int i = rows - 1; i >= 0; --i) {
        float *d = dist + (size_t)i * cols;
        float *t = temp + (size_t)(i + B) * step + B;
        for (int j = cols - 1; j >= 0; --j) {
            float t0 = t[j], v;
            if (t0 > HV) {
                v = t[j + step * 2 + 1] + LG; if (t0 > v) t0 = v;
                v = t[j + step * 2 - 1] + LG; if (t0 > v) t0 = v;
                v = t[j + step + 2] + LG; if (t0 > v) t0 = v;
                v = t[j + step + 1] + DG; if (t0 > v) t0 = v;
                v = t[j + step] + HV; if (t0 > v) t0 = v;
                v = t[j + step - 1] + DG; if (t0 > v) t0 = v;
                v = t[j + step - 2] + LG; if (t0 > v) t0 = v;
                v = t[j + 1] + HV; if (t0 > v) t0 = v;
                t[j] = t0;
            }
            d[j] = t0;
        }
    }
    free(temp);
    return 0;
}

/* Rasterisation of the rocks, MPPI_isaac.py:361-372 (all float64, numpy.linspace grid).  obstacles: [n][3]
 * (x_global, y_global, r_obs).  mask: [cms][cms] uint8, row = Y index, column = X index. */
int oracle_rasterize_obstacles(const double *obstacles, int n_obs, double x0, double y0, int cms, double half_width,
                               double r_robot, double radius_scale, double inflate, uint8_t *mask)
{
    memset(mask, 255, (size_t)cms * cms);
    const double start = -half_width, stepd = (half_width - start) / (double)(cms - 1);
    double *xc = (double *)malloc(sizeof(double) * cms);
    if (!xc) return -1;
    for (int i = 0; i < cms; ++i) xc[i] = (double)i * stepd + start;       /* numpy.linspace */
    xc[cms - 1] = half_width;
    for (int o = 0; o < n_obs; ++o) {
        const double xl = obstacles[3 * o + 1] - y0, yl = obstacles[3 * o] - x0;
        const double R = obstacles[3 * o + 2] * radius_scale + r_robot + inflate, R2 = R * R;
        for (int r = 0; r < cms; ++r) {
            const double dy = xc[r] - yl, dy2 = dy * dy;
            if (dy2 > R2) continue;
            for (int c = 0; c < cms; ++c) {
                const double dx = xc[c] - xl;
                if (dx * dx + dy2 <= R2) mask[(size_t)r * cms + c] = 0;
            }
        }
    }
    free(xc);
    return 0;
}
