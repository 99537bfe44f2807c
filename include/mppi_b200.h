/*
 * mppi_b200.h -- C ABI of libmppi_b200.so: the B200-native (sm_100a) MPPI controller core.
 *
 * Drop-in boundary.  The reference has no FFI: its boundary is the Python class
 * `MPPI_Controller` (thesis_master/warp_implementation/MPPI_isaac.py:402-805) whose methods issue
 * nine `wp.launch` calls per control iteration (MPPI_isaac.py:505-720).  Each entry point below
 * names the reference interface it replaces.  All pointers are plain device or host pointers as
 * documented; no torch / Warp types cross this boundary.  Every function returns an int status
 * (0 = MPPI_OK, negative = error; see mppi_strerror).
 *
 * Threading: a handle is thread-compatible (one thread at a time).  All device work is issued
 * asynchronously on the `stream` argument (a cudaStream_t passed as void*; NULL = legacy default
 * stream) unless stated otherwise.
 */
#ifndef MPPI_B200_H
#define MPPI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPPI_B200_ABI_VERSION 4

enum {
    MPPI_OK = 0,
    MPPI_ERR_INVALID_ARG = -1,
    MPPI_ERR_CUDA = -2,
    MPPI_ERR_NO_TERRAIN = -3,
    MPPI_ERR_ALLOC = -4,
    MPPI_ERR_UNSUPPORTED = -5
};

enum { MPPI_PROJ_2D = 2, MPPI_PROJ_3D = 3 };   /* MPPI_step(proj="2d"|"3d")  MPPI_isaac.py:554,578 */

/* Arithmetic flavour of the kernels.
 *  STRICT: every fp32 operation in reference order, no FMA contraction, IEEE div/sqrt, specified
 *          ("det") sin/cos/exp/log -> bit-identical to oracle/mppi_oracle.c in MATH_DET mode.
 *  FAST:   FMA contraction + approximate reciprocal/rsqrt/MUFU intrinsics (throughput mode; matches the
 *          oracle within tolerance, not bit-for-bit). */
enum { MPPI_MATH_STRICT = 0, MPPI_MATH_FAST = 1 };

/* Fused-kernel variant.  Both compute the same bits; they differ in how a sample's step is mapped to warps.
 *  MONO: one thread per sample does everything (throughput regime, large K).
 *  PIPE: six specialised warps per 32 samples (2 x noise, filter -> dependent chain -> wheel/slope and obstacle
 *        critics) handing stages over through shared-memory rings (latency regime, K of a few thousand). */
enum { MPPI_VARIANT_AUTO = 0, MPPI_VARIANT_MONO = 1, MPPI_VARIANT_PIPE = 2 };

/* What is sampled.
 *  SKID_STEER: wheel inputs (u1, u2) perturbed around the nominal, then the first-order wheel filter -> (v, w):
 *              _generate_inputs_kernel + _convert_inputs_to_velocities (sampling_warp.py:54-138), what MPPI_isaac.py runs.
 *  UNICYCLE:   (v, w) perturbed directly around the previous optimal velocity sequence, clamped to the velocity
 *              limits, no filter: _generate_velocities_kernel (sampling_warp.py:10-48; driven by old_files/run_mppi.py).
 *              The nominal / optimal_u1,u2 buffers then hold (v, w) and optimal_v/w are copies of them;
 *              MppiState.sigma1 / sigma2 are std_dev_linear / std_dev_angular. */
enum { MPPI_INPUT_SKID_STEER = 0, MPPI_INPUT_UNICYCLE = 1 };

/* Every tunable / literal of the reference hot path. Defaults via mppi_default_params. */
typedef struct MppiParams {
    int32_t K;              /* number_of_trajectories                       config.yaml:7   */
    int32_t T;              /* number_of_iterations (horizon steps), 2..512  config.yaml:5   */
    int32_t math;           /* MPPI_MATH_*                                                   */
    int32_t variant;        /* MPPI_VARIANT_*: which fused kernel runs the step (0 = choose by K)             */
    float dt;               /* config.yaml:6 */
    float u1_min, u1_max, u2_min, u2_max;       /* config.yaml:21-24 */
    float v_min, v_max, w_min, w_max;           /* config.yaml:11-12,15-16 */
    float lambda;           /* temperature config.yaml:28 */
    float r_wheels;         /* robot.radius used as track width  MPPI_isaac.py:537 */
    float filt_k, filt_a;   /* sample wheel filter  (3.5, 0.96)  MPPI_isaac.py:548-549 */
    float opt_k, opt_a;     /* optimal-sequence filter (3.0, 0.92) MPPI_isaac.py:688-689 */
    float wheel_offset;     /* 0.2 m  projection_warp.py:333 */
    float cw_path, cw_slope, cw_speed, cw_obs;  /* 100.5, 50.5, 0.5, 25  critics_warp.py:325-329 */
    float lethal_thresh, lethal_penalty;        /* 0.99, 1e5  critics_warp.py:251-252 */
    float near_goal_cut;    /* 2.0   critics_warp.py:285 */
    float speed_eps;        /* 1e-4  critics_warp.py:297 */
    float pf_eps;           /* 1e-6  critics_warp.py:111 */
    float pf_near_gain;     /* 10.0  critics_warp.py:126 */
    float slope_eps;        /* 1e-6  critics_warp.py:188 */
    float slope_gain;       /* 5.0   critics_warp.py:209-210 */
    float horizon;          /* dt*v_max*T  MPPI_isaac.py:440 (host double -> float) */
    float target_speed;     /* v_max_linear MPPI_isaac.py:619 */
    int32_t input_model;    /* MPPI_INPUT_* */
    /* ---- optional critics: weight 0 (the default) = off, the term is not evaluated and not added.  The first three
     * are the reference's dormant critics (defined in critics_warp.py, their `costs[tid] +=` lines commented out or
     * absent in _evaluate_trajectories_kernel :324,326); the last three are extensions with no reference counterpart
     * (BASELINE configuration 5 names roll / pitch critics).  Cost order, all fp32 `+=` on the zeroed accumulator:
     * orient, path, slope_path, slope(wheels), speed, obstacle, goal_angle, roll, pitch, effort. */
    float cw_orient;        /* _path_orientation_critic  critics_warp.py:44-83 (commented coefficient 1.0, :324); T >= 2 */
    float cw_slope_path;    /* _avoid_slope: stride-2 slope of the BODY path  critics_warp.py:131-166 (commented 50.5, :326) */
    float cw_goal_angle;    /* _goal_angle_critic  critics_warp.py:5-41 (defined, never called); uses MppiState.goal_theta */
    float goal_angle_radius;/* 0.5   critics_warp.py:33 */
    float cw_roll;          /* extension: sum over even t of ((lw_z - rw_z) / (2 wheel_offset))^2   (tan of body roll) */
    float cw_pitch;         /* extension: sum over even t of heading_z^2                            (sin of body pitch) */
    float cw_effort;        /* extension: sum over t of u1^2 + u2^2 of the clamped sampled inputs */
    int32_t reserved;       /* must be 0 (keeps sizeof(MppiParams) = 172 = 12 mod 16, the layout the kernels' parameter
                               loads were tuned with) */
} MppiParams;

/* Terrain = DEM `Z_wp` + obstacle `costmap_wp` (MPPI_isaac.py:463-464), borrowed device pointers. */
typedef struct MppiTerrain {
    const float *dem;       /* device, row-major [grid_size*grid_size]; row 0 at y=+half_width (projection_warp.py:40) */
    int32_t grid_size;
    float half_width;       /* x_min = y_min = -half_width  MPPI_isaac.py:584-585 */
    float resolution;       /* 2*half_width/grid_size  MPPI_isaac.py:265 */
    const float *costmap;   /* device, row-major [costmap_size*costmap_size]  critics_warp.py:248 */
    int32_t costmap_size;
    float costmap_resolution; /* 2*half_width/costmap_size MPPI_isaac.py:272 */
} MppiTerrain;

/* Per-iteration inputs the caller mutates between steps (visual_terrain_stack_full_terrain.py:494-515,
 * 574-576; MPPI_isaac.py:769-784). */
typedef struct MppiState {
    float x, y;             /* robot.x[-1], robot.y[-1] */
    float hx, hy, hz;       /* robot.heading_vector / |.|  (MPPI_isaac.py:493) */
    float wheel_l, wheel_r; /* robot.left/right_wheel_speed */
    float sigma1, sigma2;   /* std_dev_u1 / std_dev_u2 */
    float goal_x, goal_y, goal_theta;
} MppiState;

/* Device-resident results of the last step (pointers owned by the handle, valid until mppi_destroy). */
typedef struct MppiOutputs {
    const float *optimal_u1, *optimal_u2;   /* [T]  optimal_u1_wp / optimal_u2_wp (= next nominal) */
    const float *optimal_v, *optimal_w;     /* [T]  optimal_lin_vel_wp / optimal_ang_vel_wp; [0] is the command */
    const float *costs;                     /* [n_rovers*K] costs_wp */
    const float *stats;                     /* [n_rovers*8] {min_cost, argmin(int bits), weights_sum, oob_count(int bits),
                                                nan_count(int bits), ess, 0, 0} */
    const float *sim_traj, *sim_heading;    /* [T*3] trajectories_sim / heading_vectors_sim (after mppi_sim_rollout) */
} MppiOutputs;

/* Optional dump of the K x T intermediates (validation / visualiser only; the fused step never
 * materialises them).  Any pointer may be NULL.  Layouts match the reference arrays. */
typedef struct MppiDebugDump {
    float *u1, *u2, *v, *w;                 /* device [K*T]   u1,u2 (MPPI_isaac.py:448-449), linear/angular_velocities */
    float *traj, *heading, *lw, *rw;        /* device [K*T*3] trajectories, heading_vectors, left/right_wheel_pos */
    int32_t *dem_ij, *lw_ij, *rw_ij;        /* device [K*T*2] DEM cell (i, j) of body / left / right wheel */
    int32_t *cm_ij;                         /* device [K*T*2] costmap cell (ix, iy) */
    float *critics;                         /* device [K*4]   path, slope, speed, obstacle (unweighted) */
    float *weights;                         /* device [K]     exp(-(c-min)/lambda) with the global min */
    float *critics_ext;                     /* device [K*6]   orient, slope_path, goal_angle, roll, pitch, effort (unweighted;
                                                              evaluated in the dump whatever their weights) */
} MppiDebugDump;

typedef struct MppiHandle MppiHandle;

const char *mppi_strerror(int status);
int mppi_abi_version(void);

/* Fills *p with the reference defaults (config.yaml + kernel literals) for the given K, T. */
int mppi_default_params(MppiParams *p, int32_t K, int32_t T);

/* Replaces MPPI_Controller.__init__ + warp_setup (MPPI_isaac.py:404-487): allocates all device scratch
 * for up to `max_rovers` independent controllers (1 for the reference use). Nominal sequences start at 0
 * (MPPI_isaac.py:446-447). */
int mppi_create(const MppiParams *params, int32_t device, int32_t max_rovers, MppiHandle **out);
int mppi_destroy(MppiHandle *h);
int mppi_set_params(MppiHandle *h, const MppiParams *params);   /* K, T must not exceed the created sizes */

/* Replaces `Z_wp = ...` / `costmap_wp.assign(...)` (MPPI_isaac.py:463-464, driver :561-567).
 * Pointers are borrowed (zero-copy) and must stay valid while steps run. */
int mppi_set_terrain(MppiHandle *h, const MppiTerrain *terrain);
/* Batched mode: terrains_dev = device array [n_rovers] of MppiTerrain (per-rover maps). */
int mppi_set_terrain_batched(MppiHandle *h, const MppiTerrain *terrains_dev, int32_t n_rovers);

/* optimal_u1_wp / optimal_u2_wp accessors (host pointers, [n_rovers*T]); synchronous on `stream`. */
int mppi_set_nominal(MppiHandle *h, const float *u1_host, const float *u2_host, int32_t n_rovers, void *stream);
int mppi_get_nominal(MppiHandle *h, float *u1_host, float *u2_host, int32_t n_rovers, void *stream);

/* Replaces reset("controller") + MPPI_step(proj) launches 1-8 (MPPI_isaac.py:489-692): ONE fused kernel
 * (sample -> wheel filter -> rollout on the DEM -> critics -> online softmax -> update -> (v*, w*)).
 *  noise_dev: NULL = production mode, counter-based Philox4x32-10 keyed by (seed, offset, sample id);
 *             else device [2][K][T] standard-normal eps injected in validation mode (shared-noise parity).
 *  Asynchronous on `stream`. */
int mppi_step(MppiHandle *h, const MppiState *state, int32_t proj, const float *noise_dev,
              uint64_t seed, uint64_t offset, void *stream);

/* Same, then copies the command (v*[0], w*[0]) to cmd_host[2] and synchronises: the call the reference
 * driver makes as MPPI_step + 2 x `.numpy()[0]` (visual_terrain_stack_full_terrain.py:468-472). */
int mppi_step_host(MppiHandle *h, const MppiState *state, int32_t proj, uint64_t seed, uint64_t offset,
                   float *cmd_host, void *stream);

/* Multi-rover batch (BASELINE config 4): n_rovers independent controllers, states_dev = device [n_rovers]
 * MppiState, terrains from mppi_set_terrain_batched. Rover r uses Philox stream (seed, offset, rover=r). */
int mppi_step_batched(MppiHandle *h, const MppiState *states_dev, int32_t n_rovers, int32_t proj,
                      uint64_t seed, uint64_t offset, void *stream);

/* Sample-sharded multi-GPU mode (BASELINE config 3): this rank rolls out global samples
 * [k_begin, k_begin + params.K) and writes its softmax partial {M, S, argmin(int bits), A1[T], A2[T]}
 * (4 + 2T floats) to partial_dev instead of updating the nominal.  After the ranks exchange partials
 * (one all-gather), mppi_combine_partials folds `n_parts` partials in rank order -- identically on every
 * rank -- and finishes the update (nominal, v*, w*). */
int mppi_step_partial(MppiHandle *h, const MppiState *state, int32_t proj, const float *noise_dev,
                      uint64_t seed, uint64_t offset, uint32_t k_begin, float *partial_dev, void *stream);
int mppi_combine_partials(MppiHandle *h, const MppiState *state, const float *partials_dev, int32_t n_parts,
                          void *stream);
int mppi_partial_floats(int32_t T);   /* 4 + 2T */

/* The same sharded step as ONE launch per rank, exchanging the rank partials over NVLink peer memory inside the
 * fused kernel (no collective-library call, no second kernel).  One process per GPU on one NVSwitch node:
 *   1. every rank: mppi_comm_export(h, world, handle)        -> 64-byte CUDA IPC handle of its exchange buffer
 *   2. the ranks all-gather the handles (any host transport; sharding.py uses torch.distributed)
 *   3. every rank: mppi_comm_connect(h, rank, world, handles)   with handles = world x 64 bytes, in rank order
 *   4. per iteration, on every rank: mppi_step_sharded(...).  Latency regime (pipelined kernel): every worker block
 *      stores its softmax partial as flag-in-data lines {value, sequence} into every rank's buffer (a partial that is
 *      provably weightless: 48 bytes per peer); every rank's updater block polls the world x nblocks slots in its own
 *      memory and folds them in global block order -- no ticket, no fence, no flag round, bitwise what one GPU computes
 *      over the same blocks.  Throughput regime (monolithic kernel): the rank folds its own partials, the rank partials
 *      are exchanged with one flag per peer and folded in rank order.  A rank that never arrives makes the others trap
 *      after MPPI_SPIN_LIMIT_MS (environment, default 10000 ms of wall time) instead of hanging the GPUs.
 * All ranks must call mppi_step_sharded the same number of times. */
int mppi_comm_export(MppiHandle *h, int32_t world, unsigned char *ipc_handle_out /* [64] */);
int mppi_comm_connect(MppiHandle *h, int32_t rank, int32_t world, const unsigned char *ipc_handles);
int mppi_step_sharded(MppiHandle *h, const MppiState *state, int32_t proj, const float *noise_dev,
                      uint64_t seed, uint64_t offset, uint32_t k_begin, void *stream);
/* Same + the command (v*[0], w*[0]) delivered to cmd_host[2] as in mppi_step_host (zero-copy store, host poll). */
int mppi_step_sharded_host(MppiHandle *h, const MppiState *state, int32_t proj, uint64_t seed, uint64_t offset,
                           uint32_t k_begin, float *cmd_host, void *stream);

/* Replaces launch 9 (MPPI_isaac.py:696-720): rollout of the optimal sequence (dim = 1) from `state`,
 * filling sim_traj / sim_heading. Lazy: only run() consumes element [0] (MPPI_isaac.py:769-772). */
int mppi_sim_rollout(MppiHandle *h, const MppiState *state, void *stream);

/* Replaces the offline closed loop MPPI_Controller.run (MPPI_isaac.py:755-805), whose plant is the controller's own
 * model, with a device-resident loop: one fused launch per control iteration, the robot state lives in device memory.
 * After (v*, w*) are known, the launch's last block advances the robot by the first step of the optimal-trajectory
 * rollout (launch 9, MPPI_isaac.py:696-720, :769-772) and applies run()'s host logic -- sigma1/2 = max(b, b -/+ g w^2)
 * (:777-778, b = sigma_base = 0.4, g = sigma_gain = 1), wheel speeds v -/+ w r/2 (:783-784), goal test
 * abs(x - gx) <= goal_tol and abs(y - gy) <= goal_tol (:763, 0.5) -- so no host round trip happens until the loop ends.
 *  state_inout: host; in: the initial state, out: the state after the last executed iteration.
 *  noise_dev:   NULL (Philox, offset = offset0 + iteration) or device [max_iters][2][K][T] injected noise.
 *  log_host:    optional host [max_iters][8] rows {x, y, z, hx, hy, hz, v*, w*} per executed iteration (what run()
 *               appends to robot.x / y / z, heading_vector, lin_vel, ang_vel).
 * Synchronous: returns when the loop has ended (goal reached or max_iters). */
int mppi_run_closed_loop(MppiHandle *h, MppiState *state_inout, int32_t proj, const float *noise_dev,
                         uint64_t seed, uint64_t offset0, int32_t max_iters, float goal_tol, float sigma_base,
                         float sigma_gain, float *log_host, int32_t *iters_done, int32_t *goal_reached, void *stream);

/* The state the LAST executed iteration of mppi_run_closed_loop sampled from (its pose, sigmas and wheel speeds BEFORE
 * that iteration's plant step): together with (seed, offset0 + iterations - 1) it replays that iteration's fan of
 * rollouts through mppi_debug_dump / mppi_export_trajectories, as MPPI_step's own bookkeeping does for a single step. */
int mppi_closed_loop_last_input(MppiHandle *h, MppiState *state_out);

/* Validation / visualiser path: re-runs sampling + rollout + critics for rover 0 and writes the requested
 * K x T intermediates (what the unfused reference keeps in `trajectories`, `left_wheel_pos`, ...). Uses the
 * nominal sequence as it was BEFORE the last step when `use_previous_nominal` != 0. */
int mppi_debug_dump(MppiHandle *h, const MppiState *state, int32_t proj, const float *noise_dev,
                    uint64_t seed, uint64_t offset, int32_t use_previous_nominal,
                    const MppiDebugDump *dump, void *stream);

/* Replaces Surface.create_obstacles_costmap (MPPI_isaac.py:361-378; called at start-up and at every terrain-block
 * change, visual_terrain_stack_full_terrain.py:449,561-563): rocks -> inflated discs (float64 test on the
 * numpy.linspace grid, identical mask) -> two-pass 5x5 chamfer distance (what cv2.distanceTransform(DIST_L2, 5)
 * computes: weights 1, 1.4, 2.1969, float32 path sums) -> min-max normalisation (cv2.normalize NORM_MINMAX) ->
 * (1 - d)^power, written straight into a device costmap (no host round trip, no H2D).
 *  obstacles_host: [n_obs][3] doubles (x_global, y_global, r_obs); a rock covers the cells within
 *                  r_obs * radius_scale + r_robot + inflate of (y_global - origin_y, x_global - origin_x)  (:365-372;
 *                  radius_scale 0.5, inflate 0.1, power 20 in MPPI_isaac.py; 1, 0.2, 10 in create_costmap.py:14-28).
 *  costmap_dev:    device [costmap_size^2] float out (e.g. the buffer passed to mppi_set_terrain).
 *  distance_dev / mask_dev: optional device outputs of the intermediate distance map (float) and mask (uint8).
 * costmap_size <= 1024 (MPPI_ERR_UNSUPPORTED above).  Asynchronous on `stream`; not re-entrant (one workspace). */
int mppi_build_costmap(int32_t device, const double *obstacles_host, int32_t n_obs, double origin_x, double origin_y,
                       int32_t costmap_size, double half_width, double r_robot, double radius_scale, double inflate,
                       double power, float *costmap_dev, float *distance_dev, unsigned char *mask_dev, void *stream);

/* Visualiser feed (visual_terrain_stack_full_terrain.py:252-261, 520-528; consumer
 * src/terrain_management/large_scale_terrain/mppi_instancer.py:65-90): the driver shows every 50th sampled trajectory
 * at every 10th step -- `trajectories.numpy().reshape(-1, T, 3)[::50]` then `[::10]` -- which in the reference costs a
 * D2H copy of the whole K x T x 3 tensor.  This re-rolls only the requested samples of the LAST step and writes only
 * the requested points: points_dev = device [ceil(K / k_stride)][ceil(T / t_stride)][3] (x, y, height).  Arguments as
 * mppi_debug_dump. */
int mppi_export_trajectories(MppiHandle *h, const MppiState *state, int32_t proj, const float *noise_dev, uint64_t seed,
                             uint64_t offset, int32_t use_previous_nominal, int32_t k_stride, int32_t t_stride,
                             float *points_dev, void *stream);

int mppi_get_outputs(MppiHandle *h, MppiOutputs *out);

/* Last measured device time of mppi_step* in microseconds (CUDA events on the step's stream); optional
 * profiling aid, enabled with mppi_enable_timing(h, 1). Synchronises the stream. */
int mppi_enable_timing(MppiHandle *h, int32_t on);
int mppi_last_step_us(MppiHandle *h, float *us);
/* p50 / p99 / max device time (microseconds) over the last up-to-1024 steps issued since mppi_enable_timing(h, 1) that
 * have completed (n_out of them): the control-update latency record SURVEY.md 8d asks for, kept inside the library so
 * that a deployed loop can read it without a profiler.  Does not synchronise (steps still in flight are left out). */
int mppi_latency_stats(MppiHandle *h, float *p50_us, float *p99_us, float *max_us, int32_t *n_out);

/* Profiling aid: when trace_dev != NULL every block of rover 0 stores 32 64-bit words per step: slots 0-6 and 8-15
 * are %globaltimer nanoseconds of the kernel's phases (entry, set-up done, rollout start / end, roles joined, partial
 * published, update finished, cost ready, block softmax, rows published, and the update phases), slot 7 the SM id,
 * slots 16-28 SM-clock stamps inside the update (tools/timeline.py decodes them).  trace_dev: device [rows * 32] uint64;
 * rows is returned through nblocks_out (the pipelined launch has one more block than partials: its last block is the
 * updater); NULL switches the stamps off (default). */
int mppi_set_trace(MppiHandle *h, uint64_t *trace_dev, int32_t *nblocks_out);

/* Measurement aid (no reference counterpart): two roofline denominators of the device the bench runs on that the
 * driver-written MEASURED_PEAKS.json does not carry -- FP32 FMA throughput (TFLOP/s, independent FFMA chains on every
 * sub-partition) and the L2 gather rate (random 4-byte gathers from an L2-resident window of l2_window_bytes >= 1 MiB:
 * 1e9 32-byte sectors per second, and the same in GB/s).  Synchronous; allocates and frees its own scratch. */
int mppi_measure_peaks(int32_t device, uint64_t l2_window_bytes, float *fp32_tflops, float *l2_gather_gsectors,
                       float *l2_gather_gbs);

/* Measurement aid: a one-thread kernel that stores %globaltimer (the clock of the mppi_set_trace stamps) into *out_dev. */
int mppi_test_timestamp(uint64_t *out_dev, void *stream);

/* Test hook: evaluates the specified ("det") math on the device. fn: 0 sincos, 1 sincos(2*pi*u), 2 log, 3 exp, 4 atan.
 * x, y0, y1 are device pointers [n]. */
int mppi_test_detmath(int32_t fn, const float *x_dev, float *y0_dev, float *y1_dev, int32_t n, void *stream);
/* Test hook: Philox normals exactly as the production kernel draws them -> eps1/eps2 device [K*T]. */
int mppi_test_noise(uint64_t seed, uint64_t offset, uint32_t rover, uint32_t k_begin, int32_t K, int32_t T,
                    int32_t math, float *eps1_dev, float *eps2_dev, void *stream);

/* Test hooks: the STRICT flavour's branch-free normalize3 / fdiv / fsqrt next to the IEEE intrinsics
 * (__fsqrt_rn, __fdiv_rn).  v_dev [3n] -> out_dev / ref_dev [3n];  a_dev, b_dev [n] -> out_dev / ref_dev [2n]
 * = {a/b, sqrt|a|}. */
int mppi_test_normalize(const float *v_dev, float *out_dev, float *ref_dev, int32_t n, void *stream);
int mppi_test_divsqrt(const float *a_dev, const float *b_dev, float *out_dev, float *ref_dev, int32_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MPPI_B200_H */
