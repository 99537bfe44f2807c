// mppi_capi.cu -- C ABI of libmppi_b200.so (include/mppi_b200.h): handle management and launch plumbing.
// Host side only; all arithmetic lives in mppi_kernels.cu.  No torch types, no CPU fallback: every entry
// that computes needs a CUDA device and fails with MPPI_ERR_CUDA otherwise.
#include "mppi_kernels.cuh"

#include <cuda.h>
#include <nvtx3/nvToolsExt.h>

#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include <vector>

using namespace mppi;

struct MppiHandle {
    MppiParams p;
    int device;
    int max_rovers;
    int K_cap, T_cap;
    int block, nblocks;
    bool pipe;              // warp-specialised variant selected
    bool has_terrain;
    MppiTerrain terrain;
    const MppiTerrain* terrains_dev;
    int n_terrains;
    // device buffers
    float *nominal1, *nominal2, *prev1, *prev2, *opt_v, *opt_w, *costs, *partials, *stats, *sim_traj, *sim_heading;
    float* dbg_costs;
    unsigned int* counters;
    // LL protocol of the pipelined kernel (mppi_kernels.cuh): per-block partial lines, polled by the updater block
    uint4* ll;                   // device [ll_slots][ll_lines(T_cap)] lines, grown on demand
    size_t ll_slots;
    unsigned long long* minkey;  // device [max_rovers]
    uint32_t ll_seq;             // sequence number of the last LL launch (1 .. 2^31 - 1)
    uint32_t spin_limit_ms;      // MPPI_SPIN_LIMIT_MS: how long a kernel may wait for another block / rank before it traps
    float* cmd_pinned;      // host pinned + mapped [4]: {v*, w*, sequence, 0}, written by the kernel itself
    float* cmd_pinned_dev;  // device view of cmd_pinned
    uint32_t host_seq;
    cudaEvent_t ev0, ev1;
    bool timing;
    bool timed_valid;
    // latency ring (mppi_latency_stats): the event pairs of the last kLatRing timed steps
    cudaEvent_t* lat_ev;         // [2 * kLatRing], created on the first timed step
    uint32_t lat_count;          // timed steps so far
    unsigned long long* trace;   // optional device buffer for kernel timeline stamps (mppi_set_trace)
    // sample-sharded multi-GPU exchange over peer memory (mppi_comm_*)
    void* comm_local;            // this rank's exchange allocation: [LL lines (flat variant)] + [2][world][stride] floats + [2][world] flags
    size_t comm_ll_bytes;        // size of the LL region (0: laid out for the two-level variant only)
    size_t comm_bytes;
    void* comm_peer[kMaxRanks];  // peers' allocations opened with CUDA IPC (nullptr for this rank)
    PeerComm peers;
    int comm_nblocks;            // blocks per rank the exchange buffers were laid out for
    // device-resident closed loop (mppi_run_closed_loop)
    MppiState* loop_state;       // device: [0] the loop's state, [1] the state the last executed iteration sampled from
    MppiState loop_last_input;   // host copy of [1] after mppi_run_closed_loop
    bool loop_last_input_valid;
    int32_t* loop_ctl;           // device {iterations done, goal reached}
    float* loop_log;             // device [loop_log_cap][8]
    int32_t loop_log_cap;
    // TMA descriptor of the DEM for the pipelined kernel's shared-memory tile, re-encoded when its key changes
    TmaDesc dem_desc;
    const float* desc_dem; int desc_gs, desc_w, desc_h; bool desc_ok;
};

static constexpr int kLatRing = 1024;
static thread_local char g_cuda_err[256];

static int cuda_fail(cudaError_t e, const char* where)
{
    snprintf(g_cuda_err, sizeof(g_cuda_err), "CUDA error at %s: %s", where, cudaGetErrorString(e));
    return MPPI_ERR_CUDA;
}
#define CK(call)                                         \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

extern "C" const char* mppi_strerror(int status)
{
    switch (status) {
    case MPPI_OK: return "ok";
    case MPPI_ERR_INVALID_ARG: return "invalid argument";
    case MPPI_ERR_CUDA: return g_cuda_err[0] ? g_cuda_err : "CUDA error";
    case MPPI_ERR_NO_TERRAIN: return "terrain not set (call mppi_set_terrain first)";
    case MPPI_ERR_ALLOC: return "allocation failed";
    case MPPI_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
    }
}

extern "C" int mppi_abi_version(void) { return MPPI_B200_ABI_VERSION; }

extern "C" int mppi_default_params(MppiParams* p, int32_t K, int32_t T)
{
    if (!p || K <= 0 || T < 2) return MPPI_ERR_INVALID_ARG;
    memset(p, 0, sizeof(*p));
    p->K = K; p->T = T; p->math = MPPI_MATH_STRICT; p->variant = MPPI_VARIANT_AUTO;
    p->dt = 0.045f;
    p->u1_min = -1.f; p->u1_max = 1.f; p->u2_min = -1.f; p->u2_max = 1.f;
    p->v_min = 0.f; p->v_max = 2.f; p->w_min = -1.f; p->w_max = 1.f;
    p->lambda = 0.3f; p->r_wheels = 1.2f;
    p->filt_k = 3.5f; p->filt_a = 0.96f; p->opt_k = 3.0f; p->opt_a = 0.92f;
    p->wheel_offset = 0.2f;
    p->cw_path = 100.5f; p->cw_slope = 50.5f; p->cw_speed = 0.5f; p->cw_obs = 25.0f;
    p->lethal_thresh = 0.99f; p->lethal_penalty = 100000.0f;
    p->near_goal_cut = 2.0f; p->speed_eps = 0.0001f; p->pf_eps = 1e-6f; p->pf_near_gain = 10.0f;
    p->slope_eps = 1e-6f; p->slope_gain = 5.0f;
    p->horizon = (float)(0.045 * 2.0 * (double)T);
    p->target_speed = 2.0f;
    p->goal_angle_radius = 0.5f;                 // critics_warp.py:33; the optional critics' weights stay 0 (off)
    return MPPI_OK;
}

// Variant + block size.  Small K is latency-bound: the warp-specialised kernel (32 samples per six-warp CTA) wins
// while the grid fits the machine a few times over; large K wants the monolithic kernel with full CTAs.
// What counts is the number of samples in flight in the launch: K per rover x rovers.
static void pick_launch(const MppiParams& p, int n_rovers, int* block, int* nblocks, bool* pipe)
{
    const int K = p.K;
    const long long total = (long long)K * (n_rovers > 0 ? n_rovers : 1);
    *pipe = (p.variant == MPPI_VARIANT_PIPE || (p.variant == MPPI_VARIANT_AUTO && total <= 148 * 32 * 2));
    if (*pipe) { *block = 192; *nblocks = (K + 31) / 32; return; }
    int b;
    if (total <= 148 * 32 * 2) b = 32;
    else if (total <= 148 * 64 * 8) b = 64;
    else b = 128;
    while (b > 32 && b / 2 >= K) b /= 2;                    // no block wider than one rover's samples need
    while ((K + b - 1) / b > 8192 && b < kMaxBlock) b *= 2;
    *block = b;
    *nblocks = (K + b - 1) / b;
}

// Kernel namespace of a parameter set: arithmetic flavour x (optional critics compiled in or not).  The XC builds are
// used only when one of the optional critic weights is non-zero, so the default hot path is untouched by them.
static bool wants_xc(const MppiParams& p)
{
    return p.cw_orient != 0.f || p.cw_slope_path != 0.f || p.cw_goal_angle != 0.f || p.cw_roll != 0.f ||
           p.cw_pitch != 0.f || p.cw_effort != 0.f;
}
#define MPPI_BY_NS(P, CALL)                                                                             \
    (wants_xc(P) ? (((P).math == MPPI_MATH_FAST) ? fast_xc::CALL : strict_xc::CALL)                       \
                 : (((P).math == MPPI_MATH_FAST) ? fast::CALL : strict::CALL))

static bool params_ok(const MppiParams* p)
{
    return p && p->K > 0 && p->T >= 2 && p->T <= 512 && p->lambda > 0.f && p->dt > 0.f &&
           (p->math == MPPI_MATH_STRICT || p->math == MPPI_MATH_FAST) && p->variant >= 0 && p->variant <= 2 &&
           (p->input_model == MPPI_INPUT_SKID_STEER || p->input_model == MPPI_INPUT_UNICYCLE);
}

extern "C" int mppi_create(const MppiParams* params, int32_t device, int32_t max_rovers, MppiHandle** out)
{
    if (!out || !params_ok(params) || max_rovers < 1) return MPPI_ERR_INVALID_ARG;
    CK(cudaSetDevice(device));
    MppiHandle* h = new (std::nothrow) MppiHandle();
    if (!h) return MPPI_ERR_ALLOC;
    memset(h, 0, sizeof(*h));
    h->p = *params; h->device = device; h->max_rovers = max_rovers;
    h->K_cap = params->K; h->T_cap = params->T;
    pick_launch(*params, 1, &h->block, &h->nblocks, &h->pipe);
    if ((size_t)h->nblocks > 8192) { delete h; return MPPI_ERR_UNSUPPORTED; }
    const size_t R = (size_t)max_rovers, T = (size_t)params->T, K = (size_t)params->K;
    const size_t stride = (size_t)partial_stride(params->T);
    // worst-case block count over the launch configs we may pick (block >= 32)
    const size_t nb_cap = (K + 31) / 32;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](float** p, size_t n) { if (e == cudaSuccess) { e = cudaMalloc((void**)p, n * sizeof(float)); if (e == cudaSuccess) e = cudaMemset(*p, 0, n * sizeof(float)); } };
    alloc(&h->nominal1, R * T); alloc(&h->nominal2, R * T);
    alloc(&h->prev1, R * T); alloc(&h->prev2, R * T);
    alloc(&h->opt_v, R * T); alloc(&h->opt_w, R * T);
    alloc(&h->costs, R * K);
    alloc(&h->partials, R * nb_cap * stride);
    alloc(&h->stats, R * kStatsStride);
    alloc(&h->sim_traj, 3 * T); alloc(&h->sim_heading, 3 * T);
    alloc(&h->dbg_costs, K);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->counters, R * kCounterStride * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(h->counters, 0, R * kCounterStride * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->minkey, R * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(h->minkey, 0, R * sizeof(unsigned long long));
    if (const char* lim = getenv("MPPI_SPIN_LIMIT_MS")) h->spin_limit_ms = (uint32_t)strtoul(lim, nullptr, 10);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&h->cmd_pinned, 4 * sizeof(float), cudaHostAllocMapped);
    if (e == cudaSuccess) { memset(h->cmd_pinned, 0, 4 * sizeof(float)); e = cudaHostGetDevicePointer((void**)&h->cmd_pinned_dev, h->cmd_pinned, 0); }
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e != cudaSuccess) { mppi_destroy(h); return cuda_fail(e, "mppi_create"); }
    *out = h;
    return MPPI_OK;
}

extern "C" int mppi_destroy(MppiHandle* h)
{
    if (!h) return MPPI_OK;
    cudaSetDevice(h->device);
    float* bufs[] = { h->nominal1, h->nominal2, h->prev1, h->prev2, h->opt_v, h->opt_w, h->costs, h->partials,
                      h->stats, h->sim_traj, h->sim_heading, h->dbg_costs };
    for (float* b : bufs) if (b) cudaFree(b);
    for (int r = 0; r < kMaxRanks; ++r) if (h->comm_peer[r]) cudaIpcCloseMemHandle(h->comm_peer[r]);
    if (h->comm_local) cudaFree(h->comm_local);
    if (h->loop_state) cudaFree(h->loop_state);
    if (h->loop_ctl) cudaFree(h->loop_ctl);
    if (h->loop_log) cudaFree(h->loop_log);
    if (h->counters) cudaFree(h->counters);
    if (h->ll) cudaFree(h->ll);
    if (h->minkey) cudaFree(h->minkey);
    if (h->cmd_pinned) cudaFreeHost(h->cmd_pinned);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->lat_ev) { for (int i = 0; i < 2 * kLatRing; ++i) if (h->lat_ev[i]) cudaEventDestroy(h->lat_ev[i]); delete[] h->lat_ev; }
    delete h;
    return MPPI_OK;
}

extern "C" int mppi_set_params(MppiHandle* h, const MppiParams* params)
{
    if (!h || !params_ok(params) || params->K > h->K_cap || params->T > h->T_cap) return MPPI_ERR_INVALID_ARG;
    h->p = *params;
    pick_launch(*params, 1, &h->block, &h->nblocks, &h->pipe);
    return MPPI_OK;
}

static bool terrain_ok(const MppiTerrain* t)
{
    return t && t->dem && t->costmap && t->grid_size >= 2 && t->costmap_size >= 1 && t->resolution > 0.f &&
           t->costmap_resolution > 0.f;
}

extern "C" int mppi_set_terrain(MppiHandle* h, const MppiTerrain* terrain)
{
    if (!h || !terrain_ok(terrain)) return MPPI_ERR_INVALID_ARG;
    h->terrain = *terrain;
    h->has_terrain = true;
    return MPPI_OK;
}

extern "C" int mppi_set_terrain_batched(MppiHandle* h, const MppiTerrain* terrains_dev, int32_t n_rovers)
{
    if (!h || !terrains_dev || n_rovers < 1 || n_rovers > h->max_rovers) return MPPI_ERR_INVALID_ARG;
    h->terrains_dev = terrains_dev;
    h->n_terrains = n_rovers;
    return MPPI_OK;
}

extern "C" int mppi_set_nominal(MppiHandle* h, const float* u1, const float* u2, int32_t n_rovers, void* stream)
{
    if (!h || !u1 || !u2 || n_rovers < 1 || n_rovers > h->max_rovers) return MPPI_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)n_rovers * h->p.T * sizeof(float);
    CK(cudaMemcpyAsync(h->nominal1, u1, n, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->nominal2, u2, n, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    return MPPI_OK;
}

extern "C" int mppi_get_nominal(MppiHandle* h, float* u1, float* u2, int32_t n_rovers, void* stream)
{
    if (!h || !u1 || !u2 || n_rovers < 1 || n_rovers > h->max_rovers) return MPPI_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)n_rovers * h->p.T * sizeof(float);
    CK(cudaMemcpyAsync(u1, h->nominal1, n, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(u2, h->nominal2, n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MPPI_OK;
}

// Box of the shared-memory DEM tile of the pipelined kernel (DemTile in mppi_kernels.cuh): the cells the body can reach
// in T steps plus a margin, rows a multiple of 16 bytes for the TMA tensor copy.  The box depends on the parameters and
// the map only; the kernel places its corner around the robot.  Returns w = 0 (no tile) when the tile would not fit
// beside the pipeline's rings or the rows cannot be aligned.
static DemTile plan_dem_tile(const MppiParams& p, const MppiTerrain& t, size_t other_smem_bytes)
{
    DemTile g = { 0, 0, 0, 0 };
    if (getenv("MPPI_NO_DEM_TILE")) return g;
    if ((t.grid_size & 3) != 0 || (reinterpret_cast<uintptr_t>(t.dem) & 15) != 0) return g;
    // body reach + the lateral wheel offset: the wheel role reads its two nearest-cell heights from the tile too
    const float reach = p.dt * fmaxf(fabsf(p.v_max), fabsf(p.v_min)) * (float)p.T + fabsf(p.wheel_offset);
    if (!(reach > 0.f) || !(t.resolution > 0.f)) return g;
    const int hc = (int)(reach / t.resolution) + 4;                         // half extent in cells, 3+ cells of margin
    // columns [i0, i0 + w) with i0 = (ic - hc) rounded down to a multiple of 4: w = 2 hc + 8 covers ic - hc .. ic + hc + 1
    const int w = (2 * hc + 2 + 3 + 3) & ~3, hgt = 2 * hc + 2;
    if (w > 256 || hgt > 256 || w > t.grid_size || hgt > t.grid_size) return g;
    if ((size_t)w * (size_t)hgt * sizeof(float) + other_smem_bytes > (size_t)227 * 1024) return g;
    g.w = w; g.h = hgt;
    return g;
}

// cuTensorMapEncodeTiled through the runtime's driver-entry-point lookup (no link against libcuda).
static bool encode_dem_desc(MppiHandle* h, const MppiTerrain& t, int w, int hgt)
{
    if (h->desc_ok && h->desc_dem == t.dem && h->desc_gs == t.grid_size && h->desc_w == w && h->desc_h == hgt) return true;
    h->desc_ok = false;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
        fn = reinterpret_cast<EncodeFn>(p);
    }
    static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc), "CUtensorMap is 128 bytes");
    const cuuint64_t gdim[2] = { (cuuint64_t)t.grid_size, (cuuint64_t)t.grid_size };
    const cuuint64_t gstride[1] = { (cuuint64_t)t.grid_size * sizeof(float) };
    const cuuint32_t box[2] = { (cuuint32_t)w, (cuuint32_t)hgt };
    const cuuint32_t estr[2] = { 1, 1 };
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(&h->dem_desc), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                    const_cast<float*>(t.dem), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    h->desc_dem = t.dem; h->desc_gs = t.grid_size; h->desc_w = w; h->desc_h = hgt; h->desc_ok = true;
    return true;
}

static int do_step(MppiHandle* h, const MppiState* state, const MppiState* states_dev, int n_rovers, int proj,
                   const float* noise, uint64_t seed, uint64_t offset, uint32_t k_begin, float* rank_partial,
                   cudaStream_t s, bool to_host = false, bool sharded = false, const LoopCtl* loop = nullptr)
{
    if (!h || (proj != MPPI_PROJ_2D && proj != MPPI_PROJ_3D)) return MPPI_ERR_INVALID_ARG;
    if (!state && !states_dev && !loop) return MPPI_ERR_INVALID_ARG;
    if (states_dev) {
        if (!h->terrains_dev || n_rovers > h->n_terrains) return MPPI_ERR_NO_TERRAIN;
    } else if (!h->has_terrain) {
        return MPPI_ERR_NO_TERRAIN;
    }
    CK(cudaSetDevice(h->device));
    pick_launch(h->p, n_rovers, &h->block, &h->nblocks, &h->pipe);
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    a.p = h->p;
    if (state) a.state = *state;
    a.terrain = h->terrain;
    a.states = states_dev;
    a.terrains = states_dev ? h->terrains_dev : nullptr;
    a.noise = noise;
    a.nominal1 = h->nominal1; a.nominal2 = h->nominal2; a.prev1 = h->prev1; a.prev2 = h->prev2;
    a.opt_v = h->opt_v; a.opt_w = h->opt_w; a.costs = h->costs; a.partials = h->partials; a.stats = h->stats;
    a.counters = h->counters;
    a.rank_partial = rank_partial;
    a.seed = seed; a.offset = offset; a.k_begin = k_begin; a.nblocks = h->nblocks;
    a.trace = h->trace;
    if (to_host) { a.host_cmd = h->cmd_pinned_dev; a.host_seq = ++h->host_seq; }
    if (sharded) { a.peers = h->peers; a.peers.seq = ++h->peers.seq; }
    if (loop) a.loop = *loop;
    a.spin_limit_ms = h->spin_limit_ms;
    if (h->pipe) {
        // LL protocol: one slot of lines per worker block (grown on demand; first step only), a launch sequence number
        // that is never 0.  A flat sharded step takes its sequence number from the exchange (identical on all ranks)
        // and tags its running-minimum entries in the upper half of the tag space.
        if (sharded && !h->peers.ll[h->peers.rank]) return MPPI_ERR_UNSUPPORTED;      // buffers laid out for two-level
        const size_t need = (size_t)n_rovers * h->nblocks;
        if (h->ll_slots < need) {
            CK(cudaStreamSynchronize(s));
            if (h->ll) CK(cudaFree(h->ll));
            h->ll = nullptr; h->ll_slots = 0;
            const size_t bytes = need * ll_lines(h->T_cap) * sizeof(uint4);
            CK(cudaMalloc((void**)&h->ll, bytes));
            CK(cudaMemset(h->ll, 0, bytes));
            h->ll_slots = need;
        }
        if (h->ll_seq >= 0x7fffffffu) {                      // wrap: stale lines / keys must not look current again
            CK(cudaMemsetAsync(h->ll, 0, h->ll_slots * ll_lines(h->T_cap) * sizeof(uint4), s));
            CK(cudaMemsetAsync(h->minkey, 0, (size_t)h->max_rovers * sizeof(unsigned long long), s));
            h->ll_seq = 0;
        }
        ++h->ll_seq;
        a.ll = h->ll; a.minkey = h->minkey;
        a.ll_seq = sharded ? a.peers.seq : h->ll_seq;
        a.mk_tag = sharded ? (a.peers.seq | 0x80000000u) : h->ll_seq;
    }
    if (h->pipe && !states_dev && n_rovers == 1 && proj == MPPI_PROJ_3D && h->nblocks <= 148)
    {
        a.tile = plan_dem_tile(h->p, h->terrain,
                               MPPI_BY_NS(h->p, pipe_smem_bytes_no_tile(h->p.T, h->nblocks * (a.peers.world > 0 ? a.peers.world : 1))));
        if (a.tile.w > 0) {
            if (encode_dem_desc(h, h->terrain, a.tile.w, a.tile.h)) a.dem_desc = h->dem_desc;
            else a.tile.w = a.tile.h = 0;
        }
    }
    cudaEvent_t ring0 = nullptr, ring1 = nullptr;
    if (h->timing) {
        if (!h->lat_ev) {
            h->lat_ev = new (std::nothrow) cudaEvent_t[2 * kLatRing]();
            if (!h->lat_ev) return MPPI_ERR_ALLOC;
            for (int i = 0; i < 2 * kLatRing; ++i) CK(cudaEventCreate(&h->lat_ev[i]));
        }
        ring0 = h->lat_ev[2 * (h->lat_count % kLatRing)];
        ring1 = h->lat_ev[2 * (h->lat_count % kLatRing) + 1];
        CK(cudaEventRecord(h->ev0, s));
        CK(cudaEventRecord(ring0, s));
    }
    // NVTX range around the launch (a no-op without a profiler attached): mppi_step [pipe | mono]
    nvtxRangePushA(h->pipe ? "mppi_step [pipelined kernel]" : "mppi_step [monolithic kernel]");
    cudaError_t e;
    if (h->pipe)
        e = MPPI_BY_NS(h->p, launch_fused_pipe(a, proj, n_rovers, s));
    else
        e = MPPI_BY_NS(h->p, launch_fused(a, proj, n_rovers, h->block, s));
    nvtxRangePop();
    if (e != cudaSuccess) return cuda_fail(e, "launch_fused");
    if (h->timing) { CK(cudaEventRecord(ring1, s)); CK(cudaEventRecord(h->ev1, s)); h->timed_valid = true; ++h->lat_count; }
    return MPPI_OK;
}

extern "C" int mppi_step(MppiHandle* h, const MppiState* state, int32_t proj, const float* noise_dev,
                         uint64_t seed, uint64_t offset, void* stream)
{
    if (!state) return MPPI_ERR_INVALID_ARG;
    return do_step(h, state, nullptr, 1, proj, noise_dev, seed, offset, 0u, nullptr, (cudaStream_t)stream);
}

// The kernel's last block stores {v*, w*, sequence} straight into mapped pinned host memory (one 16-byte store):
// poll the sequence word instead of paying a D2H copy plus a stream synchronisation.  The stream is queried now
// and then so that a faulted launch is reported instead of spinning forever.
static int wait_host_command(MppiHandle* h, float* cmd_host, cudaStream_t s)
{
    volatile uint32_t* seq = reinterpret_cast<volatile uint32_t*>(h->cmd_pinned + 2);
    const uint32_t want = h->host_seq;
    for (unsigned spin = 1; *seq != want; ++spin) {
        __builtin_ia32_pause();
        if ((spin & 0x3fffu) == 0) {
            cudaError_t q = cudaStreamQuery(s);
            if (q == cudaSuccess) break;                       // finished: the store is visible by now
            if (q != cudaErrorNotReady) return cuda_fail(q, "mppi_step_host");
        }
    }
    if (*seq != want) CK(cudaStreamSynchronize(s));
    cmd_host[0] = h->cmd_pinned[0];
    cmd_host[1] = h->cmd_pinned[1];
    return MPPI_OK;
}

extern "C" int mppi_step_host(MppiHandle* h, const MppiState* state, int32_t proj, uint64_t seed, uint64_t offset,
                              float* cmd_host, void* stream)
{
    if (!state || !cmd_host) return MPPI_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = do_step(h, state, nullptr, 1, proj, nullptr, seed, offset, 0u, nullptr, s, true);
    if (rc != MPPI_OK) return rc;
    return wait_host_command(h, cmd_host, s);
}

extern "C" int mppi_step_batched(MppiHandle* h, const MppiState* states_dev, int32_t n_rovers, int32_t proj,
                                 uint64_t seed, uint64_t offset, void* stream)
{
    if (!h || !states_dev || n_rovers < 1 || n_rovers > h->max_rovers) return MPPI_ERR_INVALID_ARG;
    return do_step(h, nullptr, states_dev, n_rovers, proj, nullptr, seed, offset, 0u, nullptr, (cudaStream_t)stream);
}

extern "C" int mppi_partial_floats(int32_t T) { return partial_stride(T); }

extern "C" int mppi_step_partial(MppiHandle* h, const MppiState* state, int32_t proj, const float* noise_dev,
                                 uint64_t seed, uint64_t offset, uint32_t k_begin, float* partial_dev, void* stream)
{
    if (!state || !partial_dev) return MPPI_ERR_INVALID_ARG;
    return do_step(h, state, nullptr, 1, proj, noise_dev, seed, offset, k_begin, partial_dev, (cudaStream_t)stream);
}

extern "C" int mppi_combine_partials(MppiHandle* h, const MppiState* state, const float* partials_dev,
                                     int32_t n_parts, void* stream)
{
    if (!h || !state || !partials_dev || n_parts < 1 || n_parts > 8192) return MPPI_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    CombineArgs a;
    memset(&a, 0, sizeof(a));
    a.p = h->p; a.state = *state; a.parts = partials_dev; a.n_parts = n_parts;
    a.nominal1 = h->nominal1; a.nominal2 = h->nominal2; a.prev1 = h->prev1; a.prev2 = h->prev2;
    a.opt_v = h->opt_v; a.opt_w = h->opt_w; a.stats = h->stats;
    cudaError_t e = (h->p.math == MPPI_MATH_FAST) ? fast::launch_combine(a, (cudaStream_t)stream)
                                                  : strict::launch_combine(a, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch_combine");
    return MPPI_OK;
}

// ---------------------------------------------------------------- sample-sharded step over peer memory
// Exchange allocation of one rank: [LL region: 2 parities x world x nblocks slots of ll_lines(T) lines -- flat variant,
// pipelined kernel only] [2 x world rank partials of partial_stride(T) floats] [2 x world arrival flags].
static size_t comm_floats(int world, int T) { return (size_t)2 * world * partial_stride(T); }

extern "C" int mppi_comm_export(MppiHandle* h, int32_t world, unsigned char* ipc_handle_out)
{
    if (!h || world < 1 || world > kMaxRanks || !ipc_handle_out) return MPPI_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    if (h->comm_local) return MPPI_ERR_INVALID_ARG;             // already exported
    pick_launch(h->p, 1, &h->block, &h->nblocks, &h->pipe);
    h->comm_nblocks = h->nblocks;
    h->comm_ll_bytes = h->pipe ? (size_t)2 * world * h->comm_nblocks * ll_lines(h->T_cap) * sizeof(uint4) : 0;
    h->comm_bytes = h->comm_ll_bytes + comm_floats(world, h->T_cap) * sizeof(float) + (size_t)2 * world * sizeof(unsigned);
    CK(cudaMalloc(&h->comm_local, h->comm_bytes));
    CK(cudaMemset(h->comm_local, 0, h->comm_bytes));
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t mh;
    CK(cudaIpcGetMemHandle(&mh, h->comm_local));
    static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(ipc_handle_out, &mh, sizeof(mh));
    return MPPI_OK;
}

extern "C" int mppi_comm_connect(MppiHandle* h, int32_t rank, int32_t world, const unsigned char* ipc_handles)
{
    if (!h || !h->comm_local || world < 1 || world > kMaxRanks || rank < 0 || rank >= world || !ipc_handles)
        return MPPI_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    for (int r = 0; r < kMaxRanks; ++r)
        if (h->comm_peer[r]) { cudaIpcCloseMemHandle(h->comm_peer[r]); h->comm_peer[r] = nullptr; }
    memset(&h->peers, 0, sizeof(h->peers));
    // the sequence numbers restart at 1: lines / flags of an earlier connection must not look current
    CK(cudaMemset(h->comm_local, 0, h->comm_bytes));
    CK(cudaDeviceSynchronize());
    for (int r = 0; r < world; ++r) {
        void* base = h->comm_local;
        if (r != rank) {
            cudaIpcMemHandle_t mh;
            memcpy(&mh, ipc_handles + (size_t)r * sizeof(mh), sizeof(mh));
            CK(cudaIpcOpenMemHandle(&base, mh, cudaIpcMemLazyEnablePeerAccess));
            h->comm_peer[r] = base;
        }
        char* b = static_cast<char*>(base);
        h->peers.ll[r] = h->comm_ll_bytes ? reinterpret_cast<uint4*>(b) : nullptr;
        h->peers.x[r] = reinterpret_cast<float*>(b + h->comm_ll_bytes);
        h->peers.f[r] = reinterpret_cast<unsigned*>(h->peers.x[r] + comm_floats(world, h->T_cap));
    }
    h->peers.rank = rank;
    h->peers.world = world;
    h->peers.nblocks = h->comm_nblocks;
    h->peers.seq = 0;
    return MPPI_OK;
}

extern "C" int mppi_step_sharded(MppiHandle* h, const MppiState* state, int32_t proj, const float* noise_dev,
                                 uint64_t seed, uint64_t offset, uint32_t k_begin, void* stream)
{
    if (!h || !state || h->peers.world < 1) return MPPI_ERR_INVALID_ARG;
    pick_launch(h->p, 1, &h->block, &h->nblocks, &h->pipe);
    if (h->p.T != h->T_cap || h->nblocks != h->comm_nblocks) return MPPI_ERR_UNSUPPORTED;   // slots laid out at export
    return do_step(h, state, nullptr, 1, proj, noise_dev, seed, offset, k_begin, nullptr, (cudaStream_t)stream,
                   false, true);
}

extern "C" int mppi_step_sharded_host(MppiHandle* h, const MppiState* state, int32_t proj, uint64_t seed,
                                      uint64_t offset, uint32_t k_begin, float* cmd_host, void* stream)
{
    if (!h || !state || !cmd_host || h->peers.world < 1) return MPPI_ERR_INVALID_ARG;
    pick_launch(h->p, 1, &h->block, &h->nblocks, &h->pipe);
    if (h->p.T != h->T_cap || h->nblocks != h->comm_nblocks) return MPPI_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = do_step(h, state, nullptr, 1, proj, nullptr, seed, offset, k_begin, nullptr, s, true, true);
    if (rc != MPPI_OK) return rc;
    return wait_host_command(h, cmd_host, s);
}

// ---------------------------------------------------------------- device-resident closed loop
extern "C" int mppi_run_closed_loop(MppiHandle* h, MppiState* state_inout, int32_t proj, const float* noise_dev,
                                    uint64_t seed, uint64_t offset0, int32_t max_iters, float goal_tol,
                                    float sigma_base, float sigma_gain, float* log_host, int32_t* iters_done,
                                    int32_t* goal_reached, void* stream)
{
    if (!h || !state_inout || max_iters < 1 || (proj != MPPI_PROJ_2D && proj != MPPI_PROJ_3D)) return MPPI_ERR_INVALID_ARG;
    if (!h->has_terrain) return MPPI_ERR_NO_TERRAIN;
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaSetDevice(h->device));
    if (!h->loop_state) {
        CK(cudaMalloc((void**)&h->loop_state, 2 * sizeof(MppiState)));
        CK(cudaMalloc((void**)&h->loop_ctl, 2 * sizeof(int32_t)));
    }
    if (log_host && h->loop_log_cap < max_iters) {
        if (h->loop_log) CK(cudaFree(h->loop_log));
        h->loop_log = nullptr; h->loop_log_cap = 0;
        CK(cudaMalloc((void**)&h->loop_log, (size_t)max_iters * 8 * sizeof(float)));
        h->loop_log_cap = max_iters;
    }
    CK(cudaMemcpyAsync(h->loop_state, state_inout, sizeof(MppiState), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(h->loop_ctl, 0, 2 * sizeof(int32_t), s));
    LoopCtl lc;
    memset(&lc, 0, sizeof(lc));
    lc.state = h->loop_state; lc.prev_state = h->loop_state + 1; lc.ctl = h->loop_ctl; lc.log = log_host ? h->loop_log : nullptr;
    lc.goal_tol = goal_tol; lc.sigma_base = sigma_base; lc.sigma_gain = sigma_gain;
    // Launches are enqueued back to back; a launch that finds the goal flag set returns at once.  The host looks at
    // the flag every `kCheck` iterations only, so the device never waits for it.
    const int kCheck = 64;
    const size_t noise_stride = (size_t)2 * h->p.K * h->p.T;
    int32_t ctl[2] = { 0, 0 };
    for (int it = 0; it < max_iters; ++it) {
        lc.iter = it;
        int rc = do_step(h, nullptr, nullptr, 1, proj, noise_dev ? noise_dev + (size_t)it * noise_stride : nullptr,
                         seed, offset0 + (uint64_t)it, 0u, nullptr, s, false, false, &lc);
        if (rc != MPPI_OK) return rc;
        if ((it + 1) % kCheck == 0 && it + 1 < max_iters) {
            CK(cudaMemcpyAsync(ctl, h->loop_ctl, sizeof(ctl), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            if (ctl[1]) break;
        }
    }
    CK(cudaMemcpyAsync(ctl, h->loop_ctl, sizeof(ctl), cudaMemcpyDeviceToHost, s));
    h->loop_last_input = *state_inout;                           // stands when no iteration runs
    CK(cudaMemcpyAsync(state_inout, h->loop_state, sizeof(MppiState), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (ctl[0] > 0) CK(cudaMemcpy(&h->loop_last_input, h->loop_state + 1, sizeof(MppiState), cudaMemcpyDeviceToHost));
    h->loop_last_input_valid = true;
    if (log_host && ctl[0] > 0)
        CK(cudaMemcpy(log_host, h->loop_log, (size_t)ctl[0] * 8 * sizeof(float), cudaMemcpyDeviceToHost));
    if (iters_done) *iters_done = ctl[0];
    if (goal_reached) *goal_reached = ctl[1];
    return MPPI_OK;
}

extern "C" int mppi_closed_loop_last_input(MppiHandle* h, MppiState* state_out)
{
    if (!h || !state_out || !h->loop_last_input_valid) return MPPI_ERR_INVALID_ARG;
    *state_out = h->loop_last_input;
    return MPPI_OK;
}

extern "C" int mppi_sim_rollout(MppiHandle* h, const MppiState* state, void* stream)
{
    if (!h || !state) return MPPI_ERR_INVALID_ARG;
    if (!h->has_terrain) return MPPI_ERR_NO_TERRAIN;
    CK(cudaSetDevice(h->device));
    SimArgs a;
    memset(&a, 0, sizeof(a));
    a.p = h->p; a.state = *state; a.terrain = h->terrain;
    a.opt_v = h->opt_v; a.opt_w = h->opt_w; a.sim_traj = h->sim_traj; a.sim_heading = h->sim_heading;
    cudaError_t e = (h->p.math == MPPI_MATH_FAST) ? fast::launch_sim(a, (cudaStream_t)stream)
                                                  : strict::launch_sim(a, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch_sim");
    return MPPI_OK;
}

extern "C" int mppi_debug_dump(MppiHandle* h, const MppiState* state, int32_t proj, const float* noise_dev,
                               uint64_t seed, uint64_t offset, int32_t use_previous_nominal,
                               const MppiDebugDump* dump, void* stream)
{
    if (!h || !state || !dump || (proj != MPPI_PROJ_2D && proj != MPPI_PROJ_3D)) return MPPI_ERR_INVALID_ARG;
    if (!h->has_terrain) return MPPI_ERR_NO_TERRAIN;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    DumpArgs a;
    memset(&a, 0, sizeof(a));
    a.p = h->p; a.state = *state; a.terrain = h->terrain; a.noise = noise_dev;
    a.nominal1 = use_previous_nominal ? h->prev1 : h->nominal1;
    a.nominal2 = use_previous_nominal ? h->prev2 : h->nominal2;
    a.d = *dump; a.costs = h->dbg_costs; a.seed = seed; a.offset = offset; a.k_begin = 0u;
    const bool fastm = (h->p.math == MPPI_MATH_FAST);
    cudaError_t e = fastm ? fast::launch_dump(a, proj, s) : strict::launch_dump(a, proj, s);
    if (e != cudaSuccess) return cuda_fail(e, "launch_dump");
    if (dump->weights) {
        e = fastm ? fast::launch_weights(h->dbg_costs, h->p.K, h->p.lambda, dump->weights, s)
                  : strict::launch_weights(h->dbg_costs, h->p.K, h->p.lambda, dump->weights, s);
        if (e != cudaSuccess) return cuda_fail(e, "launch_weights");
    }
    return MPPI_OK;
}

extern "C" int mppi_export_trajectories(MppiHandle* h, const MppiState* state, int32_t proj, const float* noise_dev,
                                        uint64_t seed, uint64_t offset, int32_t use_previous_nominal, int32_t k_stride,
                                        int32_t t_stride, float* points_dev, void* stream)
{
    if (!h || !state || !points_dev || k_stride < 1 || t_stride < 1 || (proj != MPPI_PROJ_2D && proj != MPPI_PROJ_3D))
        return MPPI_ERR_INVALID_ARG;
    if (!h->has_terrain) return MPPI_ERR_NO_TERRAIN;
    CK(cudaSetDevice(h->device));
    ExportArgs a;
    memset(&a, 0, sizeof(a));
    a.p = h->p; a.state = *state; a.terrain = h->terrain; a.noise = noise_dev;
    a.nominal1 = use_previous_nominal ? h->prev1 : h->nominal1;
    a.nominal2 = use_previous_nominal ? h->prev2 : h->nominal2;
    a.seed = seed; a.offset = offset; a.k_stride = k_stride; a.t_stride = t_stride; a.points = points_dev;
    cudaError_t e = (h->p.math == MPPI_MATH_FAST) ? fast::launch_export(a, proj, (cudaStream_t)stream)
                                                  : strict::launch_export(a, proj, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch_export");
    return MPPI_OK;
}

extern "C" int mppi_get_outputs(MppiHandle* h, MppiOutputs* out)
{
    if (!h || !out) return MPPI_ERR_INVALID_ARG;
    out->optimal_u1 = h->nominal1; out->optimal_u2 = h->nominal2;
    out->optimal_v = h->opt_v; out->optimal_w = h->opt_w;
    out->costs = h->costs; out->stats = h->stats;
    out->sim_traj = h->sim_traj; out->sim_heading = h->sim_heading;
    return MPPI_OK;
}

extern "C" int mppi_enable_timing(MppiHandle* h, int32_t on)
{
    if (!h) return MPPI_ERR_INVALID_ARG;
    h->timing = on != 0;
    h->timed_valid = false;
    h->lat_count = 0;
    return MPPI_OK;
}

extern "C" int mppi_last_step_us(MppiHandle* h, float* us)
{
    if (!h || !us || !h->timed_valid) return MPPI_ERR_INVALID_ARG;
    CK(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *us = ms * 1000.f;
    return MPPI_OK;
}

extern "C" int mppi_latency_stats(MppiHandle* h, float* p50_us, float* p99_us, float* max_us, int32_t* n_out)
{
    if (!h || !p50_us || !p99_us || !max_us || !n_out) return MPPI_ERR_INVALID_ARG;
    CK(cudaSetDevice(h->device));
    std::vector<float> us;
    const uint32_t have = std::min<uint32_t>(h->lat_count, (uint32_t)kLatRing);
    for (uint32_t i = 0; i < have && h->lat_ev; ++i) {
        if (cudaEventQuery(h->lat_ev[2 * i + 1]) != cudaSuccess) continue;      // still in flight: not counted
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->lat_ev[2 * i], h->lat_ev[2 * i + 1]) == cudaSuccess) us.push_back(ms * 1000.f);
    }
    cudaGetLastError();
    *n_out = (int32_t)us.size();
    if (us.empty()) { *p50_us = *p99_us = *max_us = 0.f; return MPPI_OK; }
    std::sort(us.begin(), us.end());
    *p50_us = us[us.size() / 2];
    *p99_us = us[std::min(us.size() - 1, (size_t)(0.99 * (double)us.size()))];
    *max_us = us.back();
    return MPPI_OK;
}

extern "C" int mppi_set_trace(MppiHandle* h, uint64_t* trace_dev, int32_t* nblocks_out)
{
    if (!h) return MPPI_ERR_INVALID_ARG;
    h->trace = reinterpret_cast<unsigned long long*>(trace_dev);
    if (nblocks_out) *nblocks_out = h->nblocks + (h->pipe ? 1 : 0);      // the pipelined launch has an updater block
    return MPPI_OK;
}

// ---------------------------------------------------------------- obstacle costmap builder
namespace mppi { namespace costmap {
cudaError_t build(const double* obstacles_dev, int n_obs, double x0, double y0, int cms, double hw, double r_robot,
                  double radius_scale, double inflate, double power, uint8_t* mask, float* tmp, float* dist, float* minmax,
                  float* costmap, float* dist_out, uint8_t* mask_out, cudaStream_t s);
size_t workspace_cells(int cms);      // the builder keeps mask / distances in a wavefront-major layout (costmap_kernels.cu)
} }

struct CostmapWorkspace {          // grown on demand, one per process (the builder is not re-entrant)
    int device = -1;
    size_t cells = 0, obs = 0;
    uint8_t* mask = nullptr; float* tmp = nullptr; float* dist = nullptr; float* minmax = nullptr; double* obstacles = nullptr;
};
static CostmapWorkspace g_cm;

extern "C" int mppi_build_costmap(int32_t device, const double* obstacles_host, int32_t n_obs, double origin_x,
                                  double origin_y, int32_t costmap_size, double half_width, double r_robot,
                                  double radius_scale, double inflate, double power, float* costmap_dev,
                                  float* distance_dev, unsigned char* mask_dev, void* stream)
{
    if (n_obs < 0 || (n_obs > 0 && !obstacles_host) || costmap_size < 2 || costmap_size > 1024 || !(half_width > 0.0) ||
        !costmap_dev)
        return (costmap_size > 1024) ? MPPI_ERR_UNSUPPORTED : MPPI_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaSetDevice(device));
    const size_t cells = mppi::costmap::workspace_cells(costmap_size);
    if (g_cm.device != device || g_cm.cells < cells) {
        if (g_cm.mask) { cudaFree(g_cm.mask); cudaFree(g_cm.tmp); cudaFree(g_cm.dist); cudaFree(g_cm.minmax); }
        g_cm.mask = nullptr; g_cm.cells = 0;
        CK(cudaMalloc((void**)&g_cm.mask, cells));
        CK(cudaMalloc((void**)&g_cm.tmp, cells * sizeof(float)));
        CK(cudaMalloc((void**)&g_cm.dist, cells * sizeof(float)));
        CK(cudaMalloc((void**)&g_cm.minmax, 2 * sizeof(float)));
        g_cm.cells = cells;
        if (g_cm.device != device) { if (g_cm.obstacles) cudaFree(g_cm.obstacles); g_cm.obstacles = nullptr; g_cm.obs = 0; }
        g_cm.device = device;
    }
    if ((size_t)n_obs > g_cm.obs) {
        if (g_cm.obstacles) cudaFree(g_cm.obstacles);
        g_cm.obstacles = nullptr; g_cm.obs = 0;
        CK(cudaMalloc((void**)&g_cm.obstacles, (size_t)n_obs * 3 * sizeof(double)));
        g_cm.obs = (size_t)n_obs;
    }
    if (n_obs > 0)
        CK(cudaMemcpyAsync(g_cm.obstacles, obstacles_host, (size_t)n_obs * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    cudaError_t e = mppi::costmap::build(g_cm.obstacles, n_obs, origin_x, origin_y, costmap_size, half_width, r_robot,
                                         radius_scale, inflate, power, g_cm.mask, g_cm.tmp, g_cm.dist, g_cm.minmax,
                                         costmap_dev, distance_dev, mask_dev, s);
    if (e != cudaSuccess) return cuda_fail(e, "mppi_build_costmap");
    return MPPI_OK;
}

extern "C" int mppi_test_detmath(int32_t fn, const float* x, float* y0, float* y1, int32_t n, void* stream)
{
    if (!x || !y0 || n < 1 || fn < 0 || fn > 4) return MPPI_ERR_INVALID_ARG;
    cudaError_t e = strict::launch_detmath(fn, x, y0, y1, n, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch_detmath");
    return MPPI_OK;
}

extern "C" int mppi_test_noise(uint64_t seed, uint64_t offset, uint32_t rover, uint32_t k_begin, int32_t K, int32_t T,
                               int32_t math, float* e1, float* e2, void* stream)
{
    if (!e1 || !e2 || K < 1 || T < 1) return MPPI_ERR_INVALID_ARG;
    cudaError_t e = (math == MPPI_MATH_FAST) ? fast::launch_noise(seed, offset, rover, k_begin, K, T, e1, e2, (cudaStream_t)stream)
                                             : strict::launch_noise(seed, offset, rover, k_begin, K, T, e1, e2, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch_noise");
    return MPPI_OK;
}

extern "C" int mppi_test_normalize(const float* v_dev, float* out_dev, float* ref_dev, int32_t n, void* stream)
{
    if (!v_dev || !out_dev || !ref_dev || n < 1) return MPPI_ERR_INVALID_ARG;
    cudaError_t e = strict::launch_normalize_test(v_dev, out_dev, ref_dev, n, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch_normalize_test");
    return MPPI_OK;
}

extern "C" int mppi_test_divsqrt(const float* a_dev, const float* b_dev, float* out_dev, float* ref_dev, int32_t n,
                                 void* stream)
{
    if (!a_dev || !b_dev || !out_dev || !ref_dev || n < 1) return MPPI_ERR_INVALID_ARG;
    cudaError_t e = strict::launch_divsqrt_test(a_dev, b_dev, out_dev, ref_dev, n, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch_divsqrt_test");
    return MPPI_OK;
}
