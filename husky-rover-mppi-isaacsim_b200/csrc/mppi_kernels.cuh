// mppi_kernels.cuh -- host-visible launch interface of the kernel translation units.
// mppi_kernels.cu is compiled four times (STRICT and FAST arithmetic flavours, each without and with the optional
// critics -DMPPI_XC, see mppi_device.cuh); each compilation exports one set of launchers in its own namespace.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mppi_b200.h"

namespace mppi {

// Softmax partial: {min cost M, sum w, argmin (int bits), sum w^2, A1[T], A2[T]} with w = exp(-(c - M)/lambda)
// and A = sum_k w_k u[k, t].  Same layout for block partials, rank partials and the all-gather payload.
constexpr int kPartialHeader = 4;
__host__ __device__ inline int partial_stride(int T) { return kPartialHeader + 2 * T; }

constexpr int kMaxBlock = 256;        // threads per block upper bound
constexpr int kStatsStride = 8;       // floats per rover in the stats buffer
constexpr int kCounterStride = 4;     // uints per rover: {ticket, oob, nan, running-minimum key of the block minima}

// ---------------------------------------------------------------- flag-in-data ("LL") softmax partials
// The pipelined kernel (latency regime) hands its block partials to a designated UPDATER block -- block `nblocks` of
// the grid, resident on an otherwise idle SM -- as 16-byte lines {value0, seq, value1, seq}: every 8-byte half carries
// the launch's sequence number, so a line is valid exactly when both halves show the current `seq` (8-byte stores are
// single transactions; the scheme of the low-latency collectives).  No ticket atomic, no release fence, no separate flag
// round and nothing to re-arm between launches: a worker block stores its lines and exits, the updater polls them.
// Slot layout (lines of 16 bytes):
//   0  {min cost m_b, argmin (int bits)}      1  {sum w, sum w^2}      2  {out-of-range count, NaN count}      3  unused
//   4 .. 4+P-1     {A1[2p], A1[2p+1]}         4+P .. 4+2P-1  {A2[2p], A2[2p+1]}        P = ceil(T / 2) step pairs
// A block that can prove its partial will be scaled by exactly 0 (partial_is_dead) publishes lines 0-2 only, with
// sum w = 0; the updater never reads the A lines of a slot whose sum w is 0.
constexpr int kLLHeaderLines = 4;
__host__ __device__ inline int ll_pairs(int T) { return (T + 1) >> 1; }
__host__ __device__ inline int ll_lines(int T) { return kLLHeaderLines + 2 * ll_pairs(T); }

// Peer exchange of the sample-sharded multi-GPU step (one process per GPU; buffers mapped with CUDA IPC).
// FLAT variant (pipelined kernel): rank r owns ll[r]: [2 parities][world][nblocks][ll_lines(T)] lines.  EVERY worker
// block of rank g stores its LL partial into slot (g, block) of every rank's buffer over NVLink as soon as it has it
// (dead partials: 3 lines per peer); every rank's updater block polls all world x nblocks slots in its OWN memory and
// folds them in global block (= global sample) order, exactly as an unsharded run over the same blocks would: compute
// + exchange + update in ONE launch, no collective library call, no second kernel, no system-scope fence.  The parity
// half = seq & 1 keeps a rank that runs one step ahead from overwriting lines its neighbour is still reading.
// TWO-LEVEL variant (monolithic kernel, thousands of blocks): the rank folds its own partials first, only the rank
// partial crosses NVLink (x[r]: [2][world][partial_stride] floats + f[r]: [2][world] arrival flags).
constexpr int kMaxRanks = 8;
struct PeerComm {
    uint4* ll[kMaxRanks];             // flat variant (nullptr when the buffers were laid out for the two-level one)
    float* x[kMaxRanks];
    unsigned int* f[kMaxRanks];
    int32_t rank, world;              // world == 0: no peer exchange
    int32_t nblocks;                  // blocks per rank the buffers were laid out for
    uint32_t seq;                     // step sequence number (parity = seq & 1 selects the buffer half)
};

// Device-resident closed loop (mppi_run_closed_loop): the plant of MPPI_Controller.run (MPPI_isaac.py:755-805) is the
// controller's own model, so the whole loop can stay on the device.  The fused kernel reads its MppiState from
// `state` and, once (v*, w*) are known, thread 0 of the last block advances the robot by the first step of the
// optimal-trajectory rollout (launch 9, MPPI_isaac.py:696-720) and applies run()'s host logic (:769-784) to produce
// the next iteration's state -- one launch per control iteration, no host round trip.
struct LoopCtl {
    MppiState* state;                 // device, in/out; nullptr: loop mode off
    MppiState* prev_state;            // device, out: the state the LAST executed iteration sampled from (for replays)
    float* log;                       // device [max_iters][8] {x, y, z, hx, hy, hz, v*, w*} after each iteration, or nullptr
    int32_t* ctl;                     // device {iterations done, goal reached}
    float goal_tol;                   // stop when |x - goal_x| <= tol and |y - goal_y| <= tol   (0.5, MPPI_isaac.py:763)
    float sigma_base, sigma_gain;     // sigma1/2 = max(base, base -/+ gain * w^2)                (0.4, 1: MPPI_isaac.py:777-778)
    int32_t iter;                     // index of this iteration (row of `log`)
};

// Shared-memory DEM tile of the pipelined kernel (latency regime): the square of cells the body can reach, copied
// once per CTA by ONE 2-D TMA tensor copy while the pipeline fills; the chain warp's four corner gathers
// per step then hit shared memory (4-byte bank granularity) instead of L1 (128-byte line granularity).  Geometry is
// computed on the host from the same state the kernel gets.  w == 0: no tile (gathers go through L1 / L2).
struct DemTile {
    int32_t i0, j0;                   // DEM column / row of tile element (0, 0); i0 and w are multiples of 4 (16-byte rows)
    int32_t w, h;                     // tile width / height in cells (the box of the TMA descriptor)
};

// 128-byte TMA descriptor (CUtensorMap) of the DEM as a 2-D fp32 tensor with a w x h box, encoded on the host
// (cuTensorMapEncodeTiled) whenever the terrain or the box changes; opaque to the device code.
struct alignas(64) TmaDesc { unsigned long long opaque[16]; };

struct FusedArgs {
    MppiParams p;
    MppiState state;                  // used when states == nullptr
    MppiTerrain terrain;              // used when terrains == nullptr
    const MppiState* states;          // device [n_rovers] (batched mode)
    const MppiTerrain* terrains;      // device [n_rovers] (batched mode)
    const float* noise;               // device [2][K][T] injected eps, or nullptr (Philox)
    float* nominal1;                  // device [n_rovers][T]  in: nominal, out: updated nominal
    float* nominal2;
    float* prev1;                     // device [n_rovers][T]  copy of the nominal before the update
    float* prev2;
    float* opt_v;                     // device [n_rovers][T]
    float* opt_w;
    float* costs;                     // device [n_rovers][K]
    float* partials;                  // device [n_rovers][nblocks][stride]
    float* stats;                     // device [n_rovers][kStatsStride]
    unsigned int* counters;           // device [n_rovers][kCounterStride]
    float* rank_partial;              // if non-null: write the rank partial here and skip the update
    uint64_t seed, offset;
    uint32_t k_begin;                 // first global sample id of this rank
    int32_t nblocks;                  // grid.x
    unsigned long long* trace;        // optional device [grid.x][32] %globaltimer stamps (profiling aid), or nullptr
    float* host_cmd;                  // optional mapped pinned host memory [4]: {v*, w*, sequence, 0} (mppi_step_host)
    uint32_t host_seq;                // sequence number stored with the command
    PeerComm peers;                   // sample-sharded multi-GPU exchange (world == 0: off)
    // LL protocol of the pipelined kernel (see above): grid.x = nblocks + 1, block `nblocks` is the updater
    uint4* ll;                        // device [n_rovers][nblocks][ll_lines(T)] lines (single-GPU launches)
    unsigned long long* minkey;       // device [n_rovers]: (seq << 32) | order-reversed key of the lowest block minimum
    uint32_t ll_seq;                  // sequence number of this launch (never 0); 0: legacy ticket protocol
    uint32_t mk_tag;                  // tag of this launch's entries in `minkey` (entries with another tag are stale)
    uint32_t spin_limit_ms;           // a wait for lines / flags that lasts longer traps (a missing rank must fail loudly)
    LoopCtl loop;                     // device-resident closed loop (state == nullptr: off)
    DemTile tile;                     // pipelined kernel only
    int32_t uhist;                    // pipelined kernel: 1 = the block keeps its sampled u in shared memory ([T][2][32]
                                      // floats after the DEM tile) for the A rows of the update; set by the launcher
    TmaDesc dem_desc;                 // valid when tile.w > 0
};

struct CombineArgs {
    MppiParams p;
    MppiState state;
    const float* parts;               // device [n_parts][stride]
    int32_t n_parts;
    float* nominal1; float* nominal2; float* prev1; float* prev2;
    float* opt_v; float* opt_w; float* stats;
};

struct DumpArgs {
    MppiParams p;
    MppiState state;
    MppiTerrain terrain;
    const float* noise;
    const float* nominal1; const float* nominal2;   // the nominal to sample around
    MppiDebugDump d;
    float* costs;                     // device [K] (debug copy)
    uint64_t seed, offset;
    uint32_t k_begin;
};

struct ExportArgs {                  // strided trajectory export for the visualiser (driver transform_trajs :252-261)
    MppiParams p;
    MppiState state;
    MppiTerrain terrain;
    const float* noise;
    const float* nominal1; const float* nominal2;
    uint64_t seed, offset;
    int32_t k_stride, t_stride;       // every k_stride-th sample, every t_stride-th step
    float* points;                    // device [ceil(K / k_stride)][ceil(T / t_stride)][3]
};

struct SimArgs {
    MppiParams p;
    MppiState state;
    MppiTerrain terrain;
    const float* opt_v; const float* opt_w;
    float* sim_traj; float* sim_heading;
};

#define MPPI_DECLARE_LAUNCHERS(NS)                                                                              \
    namespace NS {                                                                                              \
    cudaError_t launch_fused(const FusedArgs& a, int proj, int n_rovers, int block, cudaStream_t s);          \
    cudaError_t launch_fused_pipe(const FusedArgs& a, int proj, int n_rovers, cudaStream_t s);                 \
    cudaError_t launch_combine(const CombineArgs& a, cudaStream_t s);                                          \
    cudaError_t launch_dump(const DumpArgs& a, int proj, cudaStream_t s);                                      \
    cudaError_t launch_export(const ExportArgs& a, int proj, cudaStream_t s);                                  \
    cudaError_t launch_weights(const float* costs, int K, float lambda, float* weights, cudaStream_t s);      \
    cudaError_t launch_sim(const SimArgs& a, cudaStream_t s);                                                  \
    cudaError_t launch_detmath(int fn, const float* x, float* y0, float* y1, int n, cudaStream_t s);          \
    cudaError_t launch_noise(uint64_t seed, uint64_t offset, uint32_t rover, uint32_t k_begin, int K, int T,  \
                             float* e1, float* e2, cudaStream_t s);                                            \
    cudaError_t launch_normalize_test(const float* v, float* out, float* ref, int n, cudaStream_t s);         \
    cudaError_t launch_divsqrt_test(const float* a, const float* b, float* out, float* ref, int n, cudaStream_t s); \
    size_t fused_smem_bytes(int T, int block, int nblocks);                                                    \
    size_t pipe_smem_bytes_no_tile(int T, int nblocks);                                                        \
    }

MPPI_DECLARE_LAUNCHERS(strict)
MPPI_DECLARE_LAUNCHERS(fast)
MPPI_DECLARE_LAUNCHERS(strict_xc)     // the same translation unit compiled with -DMPPI_XC (optional critics)
MPPI_DECLARE_LAUNCHERS(fast_xc)

}  // namespace mppi
