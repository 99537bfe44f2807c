/*
 * oracle/mppi_oracle.h -- TEST INFRASTRUCTURE. CPU restatement of the reference MPPI step.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product path (husky-rover-mppi-isaacsim_b200/) never does.
 *
 * Parity status: the reference holds no golden vectors for this path.  The oracle
 * is pinned against outputs of the reference's own controller + kernel sources run in the build container under
 * oracle/warp_shim.py (tests/golden/reference_mppi_steps.npz) plus the pins listed in oracle/README.md
 * (function-level goldens from the reference's CPU script, trajectory_2D.csv, Random123 Philox KATs, a second
 * independent NumPy restatement).  warp-lang's own arithmetic (wp.randn, libdevice) stays unpinned.
 */
#ifndef MPPI_ORACLE_H
#define MPPI_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_MATH_LIBM = 0, ORACLE_MATH_DET = 1 };

typedef struct OrParams {
    int32_t K;            /* number_of_trajectories      config.yaml:7  */
    int32_t T;            /* number_of_iterations        config.yaml:5  */
    int32_t proj;         /* 2 = "2d", 3 = "3d"          MPPI_isaac.py:554,578 */
    int32_t math;         /* ORACLE_MATH_*                                */
    float dt;             /* config.yaml:6 */
    float u1_min, u1_max, u2_min, u2_max;   /* config.yaml:21-24 */
    float v_min, v_max, w_min, w_max;       /* config.yaml:11-12,15-16 */
    float lambda;         /* temperature config.yaml:28 */
    float r_wheels;       /* robot.radius MPPI_isaac.py:537 */
    float filt_k, filt_a; /* 3.5, 0.96 MPPI_isaac.py:548-549 */
    float opt_k, opt_a;   /* 3.0, 0.92 MPPI_isaac.py:688-689 */
    float wheel_offset;   /* 0.2 projection_warp.py:333 */
    float cw_path, cw_slope, cw_speed, cw_obs; /* 100.5 50.5 0.5 25 critics_warp.py:325-329 */
    float lethal_thresh, lethal_penalty;       /* 0.99, 1e5 critics_warp.py:251-252 */
    float near_goal_cut;  /* 2.0 critics_warp.py:285 */
    float speed_eps;      /* 1e-4 critics_warp.py:297 */
    float pf_eps;         /* 1e-6 critics_warp.py:111 */
    float pf_near_gain;   /* 10.0 critics_warp.py:126 */
    float slope_eps;      /* 1e-6 critics_warp.py:188 */
    float slope_gain;     /* 5.0 critics_warp.py:209-210 */
    float horizon;        /* dt*v_max*T MPPI_isaac.py:440 (host float64 -> fp32) */
    float target_speed;   /* v_max_linear MPPI_isaac.py:619 */
    int32_t input_model;  /* 0: wheel inputs u1,u2 + first-order filter (_generate_inputs_kernel sampling_warp.py:54-92,
                             _convert_inputs_to_velocities :96-138); 1: velocity space (v, w) sampled directly
                             (_generate_velocities_kernel sampling_warp.py:10-48), no filter */
    /* optional critics, weight 0 = not evaluated, not added.  Cost order: orient, path, slope_path, slope (wheels),
     * speed, obstacle, goal_angle, roll, pitch, effort. */
    float cw_orient;        /* _path_orientation_critic critics_warp.py:44-83 (commented `+=` at :324) */
    float cw_slope_path;    /* _avoid_slope critics_warp.py:131-166 (commented `+=` at :326) */
    float cw_goal_angle;    /* _goal_angle_critic critics_warp.py:5-41 (never called by the kernel) */
    float goal_angle_radius;/* 0.5 critics_warp.py:33 */
    float cw_roll;          /* extension (no reference): sum over even t of ((lw_z - rw_z) / (2 wheel_offset))^2 */
    float cw_pitch;         /* extension: sum over even t of heading_z^2 */
    float cw_effort;        /* extension: sum over t of u1^2 + u2^2 */
} OrParams;

typedef struct OrTerrain {
    const float *dem;     /* Z, row-major gs*gs  MPPI_isaac.py:463 */
    int32_t gs;
    float half_width;
    float res;            /* 2*hw/gs (host float64 -> fp32) MPPI_isaac.py:265 */
    const float *costmap; /* row-major cms*cms MPPI_isaac.py:464 */
    int32_t cms;
    float cres;           /* 2*hw/cms MPPI_isaac.py:272 */
} OrTerrain;

typedef struct OrState {
    float x, y;           /* robot.x[-1], robot.y[-1] */
    float hx, hy, hz;     /* robot.heading_vector / norm (MPPI_isaac.py:493) */
    float wheel_l, wheel_r;
    float sigma1, sigma2;
    float goal_x, goal_y, goal_theta;
} OrState;

/* Optional per-(k,t) dumps; any pointer may be NULL. */
typedef struct OrDump {
    float *u1, *u2, *v, *w;               /* [K*T]   */
    float *traj, *heading, *lw, *rw;      /* [K*T*3] */
    int32_t *dem_ij, *lw_ij, *rw_ij;      /* [K*T*2] (i, j) of projection_warp.py:39-40 / 338-339 / 345-346 */
    int32_t *cm_ij;                       /* [K*T*2] (ix, iy) of critics_warp.py:245-248 */
    float *critics;                       /* [K*4] path, slope, speed, obstacle (unweighted) */
    float *critics_ext;                   /* [K*6] orient, slope_path, goal_angle, roll, pitch, effort (unweighted, always evaluated) */
    float *cost;                          /* [K] */
    float *weights;                       /* [K] */
} OrDump;

typedef struct OrOut {
    float *nominal1, *nominal2;           /* [T] updated optimal_u1/u2 (critics_warp.py:363-376) */
    float *opt_v, *opt_w;                 /* [T] MPPI_isaac.py:672-692 */
    float *sim_traj, *sim_heading;        /* [T*3] MPPI_isaac.py:696-720 (may be NULL) */
    double *nominal1_f64, *nominal2_f64;  /* [T] float64-accumulated shadow (may be NULL) */
    float min_cost;
    int32_t argmin;
    float weights_sum;
    int32_t oob_clamps;                   /* lookups that needed index clamping (reference: UB) */
} OrOut;

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Noise stream of the production mode (spec in DESIGN.md "Noise"): eps{1,2}[k*T+t], k0 = first
 * global sample id. math selects libm or det Box-Muller. */
void oracle_philox_normals(uint64_t seed, uint64_t offset, uint32_t rover, uint32_t k0,
                           int32_t K, int32_t T, float *eps1, float *eps2, int32_t math);

/* fn: 0 sincos, 1 sincos2pi, 2 log, 3 exp, 4 atan */
void oracle_detmath_eval(int32_t fn, const float *x, float *y0, float *y1, int32_t n);

int oracle_mppi_step(const OrParams *p, const OrTerrain *ter, const OrState *st,
                     const float *nom1, const float *nom2,
                     const float *eps1, const float *eps2,
                     OrDump *dump, OrOut *out, int32_t nthreads);

/* Rank-partial combine of the sample-sharded mode: parts[g] = {M, S, argmin(as float bits),
 * A1[T], A2[T]} ; writes nominal (A/S) */
void oracle_combine_partials(const float *parts, int32_t G, int32_t T, float lambda, int32_t math,
                             float *nominal1, float *nominal2, float *min_cost, int32_t *argmin, float *wsum);

int oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
