// mppi_kernels.cu -- the fused MPPI iteration for sm_100a.
//
// One launch replaces launches 1-8 of MPPI_Controller.MPPI_step (MPPI_isaac.py:512-692):
//   phase 1  one thread per sample: Philox noise (or injected eps) -> u -> wheel filter -> rollout on the
//            DEM -> streaming critics -> cost[k].  No K x T tensor is written.
//   phase 2  per block: min/argmin, w = exp(-(c - m_b)/lambda), sum w, and A[t] = sum_k w_k u[k,t] where u is
//            REGENERATED from the counter-based noise only for samples with w > 0 (a zero weight adds
//            exactly +0, so skipping is bit-exact).  One softmax partial per block goes to global memory.
//   phase 3  the last block to finish (atomic ticket) folds the partials in block order (deterministic
//            online softmax), writes the updated nominal, runs the (opt_k, opt_a) wheel filter and
//            publishes (v*, w*).  In sample-sharded multi-GPU mode it writes the rank partial instead.
//
//            A block that can prove its partial will be scaled by exactly 0 (running minimum of the block
//            minima, see partial_is_dead) publishes only its header.
//
// Compiled four times: STRICT (-fmad=false ...) and FAST (-DMPPI_FLAVOR_FAST -use_fast_math), each without and with
// the optional critics (-DMPPI_XC), into the namespaces strict / fast / strict_xc / fast_xc.
#include "mppi_kernels.cuh"
#include "mppi_device.cuh"

#include <math_constants.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <utility>

namespace mppi {
namespace MPPI_NS {

// ------------------------------------------------------------------ optional timeline stamps (mppi_set_trace)
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounds a wait on another block / rank by WALL time (%globaltimer), not by an iteration count: ranks of a sharded step
// are launched by different processes and a late one (lazy module load, a garbage collection) must not turn into a
// trap on every rank after a second.  limit_ms == 0: the default of 10 s.
__device__ __forceinline__ void spin_guard(unsigned& spins, unsigned long long& t0, unsigned limit_ms)
{
    if ((++spins & 0x3ffu) != 0u) return;
    const unsigned long long now = globaltimer_ns();
    if (t0 == 0ull) { t0 = now; return; }
    if (now - t0 > (unsigned long long)(limit_ms ? limit_ms : 10000u) * 1000000ull) __trap();
}
// slot: 0 entry, 1 set-up done, 2 rollout start, 3 rollout end, 4 all roles joined, 5 partial published,
//       6 update finished (last block only), 7 SM id, 8 cost ready, 9 block min, 10 block sum + compaction,
//       11 ticket taken, 12..15 last block: global min, ordered fold, nominal written, (v*, w*) written,
//       25 DEM tile landed (pipelined kernel)
constexpr int kTraceSlots = 32;       // 16..24: SM-clock stamps inside the last block's update (cycles)
__device__ __forceinline__ void trace_stamp(const FusedArgs& A, int slot)
{
    if (A.trace != nullptr && blockIdx.y == 0) A.trace[(size_t)blockIdx.x * kTraceSlots + slot] = globaltimer_ns();
}

// ------------------------------------------------------------------ L2 warm-up of the reachable terrain window
// Every sample starts at the robot, so all blocks gather from the same window of the DEM / costmap: the square of
// half-side R = T dt v_max (+ wheel offset + one cell) around the start.  After an L2 flush the first block to touch
// a line pays HBM latency on its dependent chain (that block then finishes last: +3 us of tail at C2).  Each block
// prefetches a 1/nblocks share of the window's 128-byte lines into L2 while its pipeline fills (C2: 8 lines per
// block; C5: 32).  Purely a hint: no data dependence, out-of-range rows / columns are clipped.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void prefetch_window(const float* base, int n, float half_width, float res, float cx, float cy,
                                                float reach, int part, int nparts, int tid, int nthreads)
{
    // columns i = (x + hw) / res, rows j = (hw - y) / res  (projection_warp.py:39-40, critics_warp.py:245-248)
    const float inv = 1.0f / res;
    int i0 = (int)((cx - reach + half_width) * inv) - 1, i1 = (int)((cx + reach + half_width) * inv) + 2;
    int j0 = (int)((half_width - cy - reach) * inv) - 1, j1 = (int)((half_width - cy + reach) * inv) + 2;
    i0 = max(i0, 0); j0 = max(j0, 0); i1 = min(i1, n - 1); j1 = min(j1, n - 1);
    if (i1 < i0 || j1 < j0) return;
    const int lines_per_row = ((i1 - i0) >> 5) + 2;             // 32 floats per 128-byte line, unaligned start
    const int total = (j1 - j0 + 1) * lines_per_row;
    const int per_part = (total + nparts - 1) / nparts;
    const int begin = part * per_part, end = min(begin + per_part, total);
    for (int l = begin + tid; l < end; l += nthreads) {
        const int row = j0 + l / lines_per_row, seg = l - (l / lines_per_row) * lines_per_row;
        const int col = min(i0 + seg * 32, n - 1);
        prefetch_l2(base + (size_t)row * n + col);
    }
}

__device__ __forceinline__ void prefetch_terrain(const MppiParams& p, const MppiState& st, const MppiTerrain& tr,
                                                 int part, int nparts, int tid, int nthreads)
{
    const float reach = p.dt * p.v_max * (float)p.T + p.wheel_offset;
    prefetch_window(tr.dem, tr.grid_size, tr.half_width, tr.resolution, st.x, st.y, reach + tr.resolution, part,
                    nparts, tid, nthreads);
    prefetch_window(tr.costmap, tr.costmap_size, tr.half_width, tr.costmap_resolution, st.x, st.y, reach, part, nparts,
                    tid, nthreads);
}

// ------------------------------------------------------------------ block-level helpers
__device__ __forceinline__ void pair_min(float& c, int& k, float oc, int ok)
{
    if (oc < c || (oc == c && ok < k)) { c = oc; k = ok; }
}

__device__ __forceinline__ void warp_min(float& c, int& k)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float oc = __shfl_xor_sync(0xffffffffu, c, off);
        const int ok = __shfl_xor_sync(0xffffffffu, k, off);
        pair_min(c, k, oc, ok);
    }
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Sample-sharded peer mode comes in two variants.  FLAT (pipelined kernel, LL protocol): every worker block stores
// its partial into every rank and every rank's updater block folds world x nblocks partials -- bitwise the unsharded
// result over the same blocks, no second-level fold on the critical path.  TWO-LEVEL (monolithic kernel, many blocks
// per rank): per-block NVLink stores would stall thousands of blocks, so the rank folds its own partials first and only
// the rank partial crosses NVLink (measured at 8 x 32768 samples: 119 us two-level vs 150 us flat).
__host__ __device__ inline bool peers_flat(const FusedArgs& a)
{
    return a.peers.world > 0 && a.rank_partial == nullptr && a.ll_seq != 0u;
}
// Number of partials the final fold may see: the blocks of this launch, times the ranks in flat peer mode.
__host__ __device__ inline int list_cap(const FusedArgs& a)
{
    return a.nblocks * (peers_flat(a) ? a.peers.world : 1);
}

// Shared-memory carve-up (floats).  `nblocks` only matters for the block that runs phase 3.
struct Smem {
    float* nom1;        // [T]
    float* nom2;        // [T]
    float* red_f;       // [3 * 32] warp partials
    int* red_i;         // [2 * 32]
    int* list_i;        // [max(B, nblocks)]  compacted sample / block indices
    float* list_w;      // [max(B, nblocks)]  their weights / scales
    float* acc;         // [4 * B] group accumulators, later new nominal [2T]
};

__host__ __device__ inline size_t smem_floats(int T, int B, int nblocks)
{
    const int L = (B > nblocks) ? B : nblocks;
    const int accn = (4 * B > 2 * T) ? 4 * B : 2 * T;
    return (size_t)2 * T + 96 + 64 + 2 * (size_t)L + accn;
}

__device__ __forceinline__ Smem carve(float* base, int T, int B, int nblocks)
{
    const int L = (B > nblocks) ? B : nblocks;
    Smem s;
    s.nom1 = base;
    s.nom2 = s.nom1 + T;
    s.red_f = s.nom2 + T;
    s.red_i = reinterpret_cast<int*>(s.red_f + 96);
    s.list_i = s.red_i + 64;
    s.list_w = reinterpret_cast<float*>(s.list_i + L);
    s.acc = s.list_w + L;
    return s;
}

// Block-wide (min, argmin) with ties to the lowest index; result broadcast to all threads.
__device__ __forceinline__ void block_min(float& c, int& k, const Smem& s)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    warp_min(c, k);
    __syncthreads();
    if (lane == 0) { s.red_f[warp] = c; s.red_i[warp] = k; }
    __syncthreads();
    c = s.red_f[0]; k = s.red_i[0];
    for (int w = 1; w < nw; ++w) pair_min(c, k, s.red_f[w], s.red_i[w]);
}

// Block-wide sums of two values in a fixed order; broadcast.
__device__ __forceinline__ void block_sum2(float& a, float& b, const Smem& s)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();
    if (lane == 0) { s.red_f[32 + warp] = a; s.red_f[64 + warp] = b; }
    __syncthreads();
    a = s.red_f[32]; b = s.red_f[64];
    for (int w = 1; w < nw; ++w) { a += s.red_f[32 + w]; b += s.red_f[64 + w]; }
}

// Ordered compaction of the threads with `keep` into (list_i, list_w), appended at `base`.
// Returns the new element count (uniform).
__device__ __forceinline__ int block_compact(bool keep, int idx, float val, int base, const Smem& s)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    __syncthreads();
    if (lane == 0) s.red_i[32 + warp] = __popc(mask);
    __syncthreads();
    int off = base, total = base;
    for (int w = 0; w < nw; ++w) {
        const int c = s.red_i[32 + w];
        if (w < warp) off += c;
        total += c;
    }
    if (keep) {
        const int pos = off + __popc(mask & ((1u << lane) - 1u));
        s.list_i[pos] = idx;
        s.list_w[pos] = val;
    }
    return total;
}

// ------------------------------------------------------------------ phase 3: fold partials, finish the update
// parts: [n][stride] softmax partials (read through L2).  Called by every thread of ONE block; all but warp 0 return
// early (at once when n <= 128, after the block-wide scan of the partial headers otherwise): the rest is a few hundred
// values, where block-wide barriers and reductions would only add latency to the tail of the iteration.
// smem nom1/nom2 hold the old nominal.  `tr`: optional timeline row (see trace_stamp).
__device__ void combine_and_finalize(const MppiParams& p, const MppiState& st, const float* parts, int n,
                                     const Smem& s, float* nominal1, float* nominal2, float* prev1, float* prev2,
                                     float* opt_v, float* opt_w, float* stats, float* rank_partial,
                                     unsigned oob_count, unsigned nan_count, unsigned long long* tr,
                                     float* host_cmd, unsigned host_seq, const PeerComm* pc = nullptr,
                                     unsigned spin_limit_ms = 0u)
{
    const int T = p.T, lane = threadIdx.x & 31, tid = threadIdx.x, B = blockDim.x;
    const int stride = partial_stride(T);
    const unsigned FULL = 0xffffffffu;
    float M = CUDART_INF_F;
    int arg = 0x7fffffff, cnt = 0;
#define MPPI_CLK(slot) do { if (tr != nullptr && (threadIdx.x & 31) == 0 && threadIdx.x < 32) tr[slot] = (unsigned long long)clock64(); } while (0)
    MPPI_CLK(16);

    if (n > 128) {
        // ---- many partials (throughput regime, thousands of blocks): steps 1-2 use the whole block
        for (int b = tid; b < n; b += B) {
            const float mb = __ldcg(parts + (size_t)b * stride);
            const int kb = __float_as_int(__ldcg(parts + (size_t)b * stride + 2));
            pair_min(M, arg, mb, kb);
        }
        block_min(M, arg, s);
        for (int base = 0; base < n; base += B) {
            const int b = base + tid;
            float sc = 0.0f;
            if (b < n) sc = fexp(fdiv(-(__ldcg(parts + (size_t)b * stride) - M), p.lambda));
            cnt = block_compact(sc > 0.0f, b, sc, cnt, s);
        }
        __syncthreads();
        if (tid >= 32) return;
    } else {
        // ---- few partials (latency regime): one warp, no block barrier anywhere
        if (tid >= 32) return;
        // 1. global min / argmin (ties -> lowest sample id); at most four partials per lane.  ALL header loads are
        //    issued before the first comparison: each is an L2 round trip (~700 cycles for a line another SM has just
        //    written), and interleaved with the comparisons they ran one after the other (3700 cycles for this step
        //    in the kernel-internal timeline; one round trip is enough).
        float mb4[4];
        int kb4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float* hp = parts + (size_t)min(lane + 32 * q, n - 1) * stride;      // clamped: no branch around the loads
            mb4[q] = __ldcg(hp);
            kb4[q] = __float_as_int(__ldcg(hp + 2));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (lane + 32 * q < n) pair_min(M, arg, mb4[q], kb4[q]);
            else mb4[q] = CUDART_INF_F;
        }
        warp_min(M, arg);
        MPPI_CLK(17);
        // 2. scale of every partial relative to M; the non-zero ones are kept, in partial order
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int b = lane + 32 * q;
            float sc = 0.0f;
            if (b < n) sc = fexp(fdiv(-(mb4[q] - M), p.lambda));
            const unsigned keep = __ballot_sync(FULL, sc > 0.0f);
            if (sc > 0.0f) {
                const int pos = cnt + __popc(keep & ((1u << lane) - 1u));
                s.list_i[pos] = b;
                s.list_w[pos] = sc;
            }
            cnt += __popc(keep);
        }
        __syncwarp();
        MPPI_CLK(18);
    }
    if (tr != nullptr && lane == 0) tr[12] = globaltimer_ns();

    // 3. ordered fold (one warp from here on).  Every lane accumulates S (identical on all lanes); lane l owns columns
    //    l, l + 32, ...: eight per pass, all loads of an entry issued together (one L2 round trip per kept partial).
    float S = 0.0f, S2 = 0.0f;
    for (int col0 = 0; col0 < 2 * T; col0 += 256) {
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.0f;
        // two kept partials per pass: their loads share one L2 round trip; the additions keep the partial order
        for (int e = 0; e < cnt; e += 2) {
            const bool two = e + 1 < cnt;
            const float* pb0 = parts + (size_t)s.list_i[e] * stride;
            const float* pb1 = two ? parts + (size_t)s.list_i[e + 1] * stride : pb0;
            const float sc0 = s.list_w[e], sc1 = two ? s.list_w[e + 1] : 0.0f;
            float v0[8], v1[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int col = col0 + lane + 32 * q;
                v0[q] = (col < 2 * T) ? __ldcg(pb0 + kPartialHeader + col) : 0.0f;
                v1[q] = (col < 2 * T) ? __ldcg(pb1 + kPartialHeader + col) : 0.0f;
            }
            float s0 = 0.0f, q0 = 0.0f, s1 = 0.0f, q1 = 0.0f;
            if (col0 == 0) { s0 = __ldcg(pb0 + 1); q0 = __ldcg(pb0 + 3); s1 = __ldcg(pb1 + 1); q1 = __ldcg(pb1 + 3); }
            if (col0 == 0) { S += s0 * sc0; S2 += q0 * sc0 * sc0; }
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] += v0[q] * sc0;
            if (two) {
                if (col0 == 0) { S += s1 * sc1; S2 += q1 * sc1 * sc1; }
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] += v1[q] * sc1;
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int col = col0 + lane + 32 * q;
            if (col < 2 * T) s.acc[col] = acc[q];
        }
    }
    __syncwarp();
    MPPI_CLK(19);
    if (tr != nullptr && lane == 0) tr[13] = globaltimer_ns();

    if (rank_partial != nullptr) {           // sample-sharded mode (NCCL transport): publish {M, S, argmin, S2, A1, A2}
        if (lane == 0) {
            rank_partial[0] = M; rank_partial[1] = S; rank_partial[2] = __int_as_float(arg); rank_partial[3] = S2;
        }
        for (int col = lane; col < 2 * T; col += 32) rank_partial[kPartialHeader + col] = s.acc[col];
        return;
    }

    if (pc != nullptr && pc->world > 0) {
        // ---- sample-sharded peer mode, TWO-LEVEL variant (many blocks per rank): this rank's partials have just been
        //      folded into one rank partial; exchange the rank partials and fold them in rank order.  Slot layout:
        //      x[r] = [2 parities][world][stride].
        const int world = pc->world, me = pc->rank;
        const unsigned seq = pc->seq;
        const size_t half = (size_t)world * stride;                        // floats per parity half
        const size_t slot = (size_t)(seq & 1u) * half + (size_t)me * stride;
        for (int r = 0; r < world; ++r) {                  // NVLink stores (plain local stores for r == me)
            float* dst = pc->x[r] + slot;
            if (lane == 0) { dst[0] = M; dst[1] = S; dst[2] = __int_as_float(arg); dst[3] = S2; }
            for (int col = lane; col < 2 * T; col += 32) dst[kPartialHeader + col] = s.acc[col];
        }
        // __syncwarp orders every lane's stores before the system-scope release of the flags below
        __syncwarp();
        if (lane < world) {
            unsigned int* theirs = pc->f[lane] + (seq & 1u) * world + me;
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(seq) : "memory");
            const unsigned int* mine = pc->f[me] + (seq & 1u) * world + lane;
            unsigned got, spins = 0;
            unsigned long long t0 = 0;
            for (;;) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(mine) : "memory");
                if (got == seq) break;
                __nanosleep(64);
                spin_guard(spins, t0, spin_limit_ms);      // a missing rank must fail loudly, never hang the GPU
            }
        }
        __syncwarp();
        combine_and_finalize(p, st, pc->x[me] + (size_t)(seq & 1u) * half, world, s, nominal1, nominal2,
                             prev1, prev2, opt_v, opt_w, stats, nullptr, oob_count, nan_count, nullptr, host_cmd,
                             host_seq, nullptr);
        return;
    }

    // 4. updated nominal = A / S  (critics_warp.py:363-376); keep the previous one for replay.  The smem copy of
    //    the OLD nominal is then overwritten with the filter drive u * k * (1 - a) of step 5.
    {
        // No valid sample at all (every cost NaN / +inf, e.g. the whole fan of rollouts crossed no-data cells): S = 0
        // and A / S would write a NaN nominal that poisons every later step -- the previous nominal is kept instead
        // (the reference has no such case handling; stats[4] = K tells the caller).
        const bool any_valid = S > 0.0f;
        const Recip rS = make_recip(S);
        const float oma = 1.0f - p.opt_a;
        for (int col = lane; col < 2 * T; col += 32) {
            const float old = (col < T) ? s.nom1[col] : s.nom2[col - T];
            const float nv = any_valid ? fdiv(s.acc[col], rS) : old;
            const float drive = nv * p.opt_k * oma;
            s.acc[col] = nv;
            if (col < T) { prev1[col] = s.nom1[col]; nominal1[col] = nv; s.nom1[col] = drive; }
            else { prev2[col - T] = s.nom2[col - T]; nominal2[col - T] = nv; s.nom2[col - T] = drive; }
        }
    }
    __syncwarp();
    MPPI_CLK(20);
    if (tr != nullptr && lane == 0) tr[14] = globaltimer_ns();

    if (p.input_model == MPPI_INPUT_UNICYCLE) {
        // velocity-space model: the weighted (v, w) sequence is the optimal velocity sequence (s.acc holds it)
        for (int t = lane; t < T; t += 32) {
            const float v = s.acc[t], w = s.acc[T + t];
            opt_v[t] = v; opt_w[t] = w;
            if (t == 0) {
                stats[6] = v; stats[7] = w;
                if (host_cmd != nullptr)
                    *reinterpret_cast<float4*>(host_cmd) = make_float4(v, w, __uint_as_float(host_seq), 0.0f);
            }
        }
    } else {
        // 5. optimal sequence -> (v*, w*) with (opt_k, opt_a) (MPPI_isaac.py:672-692).  The two wheel recurrences
        //    l <- l a + drive_l[t], r <- r a + drive_r[t] are the only sequential part: lane 0 runs the left wheel and
        //    lane 1 the right wheel in the same instruction stream, in register batches whose successor is loaded
        //    before the batch is stored in place; then all lanes map (l, r) -> (v, w).  Same operations in the same
        //    order as a one-thread loop.
        if (lane < 2) {
            float* d = (lane == 0) ? s.nom1 : s.nom2;
            float x = (lane == 0) ? st.wheel_l : st.wheel_r;
            constexpr int N = 8;
            float cur[N], nxt[N];
#pragma unroll
            for (int i = 0; i < N; ++i) cur[i] = (i < T) ? d[i] : 0.0f;
            for (int t0 = 0; t0 < T; t0 += N) {
#pragma unroll
                for (int i = 0; i < N; ++i) nxt[i] = (t0 + N + i < T) ? d[t0 + N + i] : 0.0f;
#pragma unroll
                for (int i = 0; i < N; ++i) { x = x * p.opt_a + cur[i]; cur[i] = x; }
#pragma unroll
                for (int i = 0; i < N; ++i) if (t0 + i < T) d[t0 + i] = cur[i];
#pragma unroll
                for (int i = 0; i < N; ++i) cur[i] = nxt[i];
            }
        }
        __syncwarp();
        MPPI_CLK(21);
        const Recip rw = make_recip(p.r_wheels);
#pragma unroll 4
        for (int t = lane; t < T; t += 32) {
            const float l = s.nom1[t], r = s.nom2[t];
            const float v = clampf((l + r) / 2.0f, p.v_min, p.v_max);
            const float w = clampf(fdiv(-l + r, rw), p.w_min, p.w_max);
            opt_v[t] = v; opt_w[t] = w;
            if (t == 0) {
                stats[6] = v; stats[7] = w;                 // the command, contiguous for one 8-byte D2H
                // zero-copy result for mppi_step_host: ONE 16-byte store {v*, w*, sequence number, 0} into mapped
                // pinned host memory; the host polls the sequence word
                if (host_cmd != nullptr)
                    *reinterpret_cast<float4*>(host_cmd) = make_float4(v, w, __uint_as_float(host_seq), 0.0f);
            }
        }
    }
    if (lane == 0) {
        stats[0] = M;
        stats[1] = __int_as_float(arg);
        stats[2] = S;
        stats[3] = __uint_as_float(oob_count);
        stats[4] = __uint_as_float(nan_count);
        stats[5] = (S > 0.0f) ? fdiv(S * S, S2) : 0.0f;           // effective sample size
    }
    __syncwarp();
    MPPI_CLK(22);
    if (tr != nullptr && lane == 0) tr[15] = globaltimer_ns();
#undef MPPI_CLK
}

// ------------------------------------------------------------------ closed loop: the plant step (one thread)
// Launch 9 of the reference restricted to the step run() consumes (MPPI_isaac.py:696-720, :769-772), followed by the
// host logic of run() (:774-784).  All float32 in the reference's order (run() keeps float32 NumPy scalars / arrays).
__device__ void loop_advance(const FusedArgs& A, const MppiState& st, const float* stats)
{
    const MppiParams& p = A.p;
    const float v0 = stats[6], w0 = stats[7];
    const Terr ter = make_terr(A.terrain);
    int oob = 0, i, j;
    float dev = 0.0f;
    float x = st.x, y = st.y;
    Quad q = corners(ter, x, y, i, j, oob);
    float3 n = normal_on_grid(q, ter.res);
    float3 prev = tangent(n, make_float3(st.hx, st.hy, st.hz));
    update_position(x, y, prev, v0, p.dt, dev);
    q = corners(ter, x, y, i, j, oob);
    const float z = bilinear(x, y, q, ter.rres);
    n = normal_on_grid(q, ter.res);
    prev = tangent(n, prev);
    const float3 cur = update_orientation(prev, w0, n, p.dt, dev);

    MppiState ns = st;
    ns.x = x; ns.y = y;
    // reset("controller"): heading / np.linalg.norm(heading) on the float32 array run() stored (MPPI_isaac.py:493)
    const Recip rn = make_recip(fsqrt(cur.x * cur.x + cur.y * cur.y + cur.z * cur.z));
    ns.hx = fdiv(cur.x, rn); ns.hy = fdiv(cur.y, rn); ns.hz = fdiv(cur.z, rn);
    const float w2 = w0 * w0;
    ns.sigma1 = fmaxf(A.loop.sigma_base, A.loop.sigma_base - A.loop.sigma_gain * w2);
    ns.sigma2 = fmaxf(A.loop.sigma_base, A.loop.sigma_base + A.loop.sigma_gain * w2);
    ns.wheel_l = v0 - w0 * p.r_wheels / 2.0f;
    ns.wheel_r = v0 + w0 * p.r_wheels / 2.0f;
    if (A.loop.prev_state != nullptr) *A.loop.prev_state = st;
    *A.loop.state = ns;
    if (A.loop.log != nullptr) {
        float* row = A.loop.log + (size_t)A.loop.iter * 8;
        row[0] = x; row[1] = y; row[2] = z; row[3] = cur.x; row[4] = cur.y; row[5] = cur.z; row[6] = v0; row[7] = w0;
    }
    A.loop.ctl[0] = A.loop.iter + 1;
    const bool far = fabsf(x - st.goal_x) > A.loop.goal_tol || fabsf(y - st.goal_y) > A.loop.goal_tol;
    if (!far) A.loop.ctl[1] = 1;
}

// Running minimum of the block minima published so far in this launch (counters[3] of the rover), as an order-reversed
// unsigned key so that atomicMax keeps the SMALLEST cost and 0 (the re-armed value) means "nothing published yet".
// A block whose own minimum is more than 90 lambda above ANY already published minimum will get the scale
// exp(-(m_b - M) / lambda) == 0 in the final fold whatever the other blocks do (M can only be lower still; the exp is
// flushed to zero below -87), i.e. its partial is never read: it skips the exponentials, the regeneration of its best
// sample's inputs and the store of its A rows -- with lambda = 0.3 that is all but one or two blocks, including,
// almost always, the block that finishes last and sits on the critical path.  The result does not depend on which
// blocks skipped: a skipped partial is exactly one that the fold drops.
__device__ __forceinline__ unsigned min_key(float c)
{
    const unsigned u = __float_as_uint(c);
    return ~((u & 0x80000000u) ? ~u : (u | 0x80000000u));
}
__device__ __forceinline__ float min_key_cost(unsigned key)
{
    const unsigned o = ~key;
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ bool partial_is_dead(const MppiParams& p, float m_b, unsigned snapshot_key)
{
    return snapshot_key != 0u && (m_b - min_key_cost(snapshot_key)) > 90.0f * p.lambda;
}

// ------------------------------------------------------------------ A[t] = sum_k w_k u[k, t] of one block
// Every thread of the block calls this.  (list_i, list_w)[0 .. n_e) are the block's samples with w > 0 in sample order;
// u is REGENERATED from the counter-based noise (or re-read from the injected noise) for those samples only.  One
// thread per step pair; when the block has at least twice as many threads as pairs, G groups of threads split the
// entries and the group sums are folded in group order.  `put(pair, a1[2p], a1[2p+1], a2[2p], a2[2p+1])` stores a
// finished pair.
// `uhist` != nullptr (pipelined kernel, when it fits): the block kept every u it sampled in shared memory as
// uhist[t][channel][lane]; the rows are then plain reads instead of a second Philox + Box-Muller pass.  Same values,
// same summation order as the regenerating path.
template <bool INJECT, typename Put>
__device__ __forceinline__ void accumulate_rows(const FusedArgs& A, const MppiState& st, const NoiseKey& nk,
                                                const Smem& s, int rover, int spb, int n_e, Put put,
                                                const float* uhist = nullptr)
{
    const MppiParams& p = A.p;
    const UBounds ub = make_ubounds(p);
    const int T = p.T, K = p.K, B = blockDim.x, tid = threadIdx.x;
    const int P = (T + 1) >> 1;                       // step pairs
    const int G = (B >= 2 * P && n_e > 1) ? B / P : 1;   // entry groups working in parallel
    const int g = tid / P, pr0 = tid - g * P;
    for (int pbase = 0; pbase < P; pbase += B) {      // one pass unless P > B
        const int pr = (G > 1) ? pr0 : pbase + tid;
        const bool active = (G > 1) ? (g < G) : (pr < P);
        float a1a = 0.f, a1b = 0.f, a2a = 0.f, a2b = 0.f;
        if (active) {
            const int t = 2 * pr;
            for (int e = (G > 1) ? g : 0; e < n_e; e += G) {
                const int kl = blockIdx.x * spb + s.list_i[e];
                const float we = s.list_w[e];
                if (uhist != nullptr) {
                    const float* u = uhist + (size_t)t * 64 + s.list_i[e];
                    a1a += we * u[0];
                    a2a += we * u[32];
                    if (t + 1 < T) { a1b += we * u[64]; a2b += we * u[96]; }
                    continue;
                }
                float e1a, e1b, e2a, e2b;
                if (INJECT) {
                    const float* q1 = A.noise + ((size_t)rover * 2 * K + kl) * T;
                    const float* q2 = q1 + (size_t)K * T;
                    e1a = q1[t]; e2a = q2[t];
                    e1b = (t + 1 < T) ? q1[t + 1] : 0.f;
                    e2b = (t + 1 < T) ? q2[t + 1] : 0.f;
                } else {
                    noise_pair(nk, A.k_begin + (uint32_t)kl, (uint32_t)pr, e1a, e1b, e2a, e2b);
                }
                a1a += we * sample_u(s.nom1, t, T, st.sigma1, e1a, ub.lo1, ub.hi1);
                a2a += we * sample_u(s.nom2, t, T, st.sigma2, e2a, ub.lo2, ub.hi2);
                if (t + 1 < T) {
                    a1b += we * sample_u(s.nom1, t + 1, T, st.sigma1, e1b, ub.lo1, ub.hi1);
                    a2b += we * sample_u(s.nom2, t + 1, T, st.sigma2, e2b, ub.lo2, ub.hi2);
                }
            }
        }
        if (G > 1) {                                  // fold the groups in order
            __syncthreads();
            if (active) {
                float* q = s.acc + 4 * (g * P + pr);
                q[0] = a1a; q[1] = a1b; q[2] = a2a; q[3] = a2b;
            }
            __syncthreads();
            if (tid < P) {
                a1a = a1b = a2a = a2b = 0.f;
                for (int gg = 0; gg < G; ++gg) {
                    const float* q = s.acc + 4 * (gg * P + tid);
                    a1a += q[0]; a1b += q[1]; a2a += q[2]; a2b += q[3];
                }
            }
        }
        const bool writer = (G > 1) ? (tid < P) : active;
        if (writer) put((G > 1) ? tid : pr, a1a, a1b, a2a, a2b);
        if (G > 1) break;
    }
}

// ------------------------------------------------------------------ phases 2 + 3, ticket protocol (monolithic kernel)
// Every thread of the block calls this.  `valid` threads own one sample each (local index `k_in_block`, cost
// `cost`); `spb` = samples per block.  INJECT: u is re-read from the injected noise instead of regenerated.
template <bool INJECT>
__device__ __forceinline__ void block_update(const FusedArgs& A, const MppiState& st, const NoiseKey& nk, const Smem& s,
                                             int rover, int spb, bool valid, int k_in_block, float cost,
                                             unsigned my_oob, unsigned my_nan, float* nominal1, float* nominal2,
                                             unsigned snapshot_key /* thread 0: counters[3] read a little earlier */)
{
    const MppiParams& p = A.p;
    const UBounds ub = make_ubounds(p);
    const int T = p.T, K = p.K, B = blockDim.x, tid = threadIdx.x;
    const uint32_t kg = A.k_begin + (uint32_t)(blockIdx.x * spb + k_in_block);
    if (my_oob) atomicAdd(&A.counters[rover * kCounterStride + 1], my_oob);
    if (my_nan) atomicAdd(&A.counters[rover * kCounterStride + 2], my_nan);

    // ---------------- phase 2: block softmax partial
    float m_b = cost;
    int arg_b = valid ? (int)kg : 0x7fffffff;
    float w = 0.0f, s_b, s2_b;
    int n_e;
    bool dead_partial;
    if (tid == 0) trace_stamp(A, 8);
    if (spb == 32) {
        // all samples of the block live in warp 0: warp-level reductions, one barrier to publish the list
        if (tid < 32) {
            warp_min(m_b, arg_b);
            if (tid == 0) atomicMax(&A.counters[rover * kCounterStride + 3], min_key(m_b));
            const bool dead = partial_is_dead(p, m_b, __shfl_sync(0xffffffffu, snapshot_key, 0));
            unsigned mask = 0u;
            s_b = 0.0f; s2_b = 0.0f;
            if (!dead) {
                if (valid && cost < CUDART_INF_F) w = fexp(fdiv(-(cost - m_b), p.lambda));   // critics_warp.py:346-347
                s_b = warp_sum(w);
                s2_b = warp_sum(w * w);
                mask = __ballot_sync(0xffffffffu, w > 0.0f);
                if (w > 0.0f) {
                    const int pos = __popc(mask & ((1u << tid) - 1u));
                    s.list_i[pos] = k_in_block;
                    s.list_w[pos] = w;
                }
            }
            if (tid == 0) {
                s.red_i[62] = __popc(mask); s.red_f[0] = m_b; s.red_f[1] = s_b; s.red_f[2] = s2_b; s.red_i[0] = arg_b;
                s.red_i[61] = dead;
            }
        }
        __syncthreads();
        n_e = s.red_i[62]; m_b = s.red_f[0]; s_b = s.red_f[1]; s2_b = s.red_f[2]; arg_b = s.red_i[0];
        dead_partial = s.red_i[61] != 0;
    } else {
        if (tid == 0) s.red_i[60] = (int)snapshot_key;          // visible to the block after block_min's barriers
        block_min(m_b, arg_b, s);
        if (tid == 0) atomicMax(&A.counters[rover * kCounterStride + 3], min_key(m_b));
        dead_partial = partial_is_dead(p, m_b, (unsigned)s.red_i[60]);      // block-uniform
        s_b = 0.0f; s2_b = 0.0f; n_e = 0;
        if (!dead_partial) {
            if (valid && cost < CUDART_INF_F) w = fexp(fdiv(-(cost - m_b), p.lambda));   // critics_warp.py:346-347
            s_b = w; s2_b = w * w;
            block_sum2(s_b, s2_b, s);
            n_e = block_compact(w > 0.0f, k_in_block, w, 0, s);
        }
        __syncthreads();
    }
    if (tid == 0) trace_stamp(A, 10);

    const int stride = partial_stride(T);
    float* part_local = A.partials + ((size_t)rover * A.nblocks + blockIdx.x) * stride;
    auto put = [&](int idx, float v) { part_local[idx] = v; };
    if (!dead_partial)
        accumulate_rows<INJECT>(A, st, nk, s, rover, spb, n_e, [&](int pr, float a1a, float a1b, float a2a, float a2b) {
            const int t = 2 * pr;
            put(kPartialHeader + t, a1a);
            put(kPartialHeader + T + t, a2a);
            if (t + 1 < T) { put(kPartialHeader + t + 1, a1b); put(kPartialHeader + T + t + 1, a2b); }
        });
    if (tid == 0) { put(0, m_b); put(1, s_b); put(2, __int_as_float(arg_b)); put(3, s2_b); trace_stamp(A, 5); }

    // ---------------- phase 3: last block folds everything
    // publish: the barrier orders every thread's partial stores before thread 0's gpu-scope release
    __syncthreads();
    if (tid == 0) {
        if (A.trace != nullptr && rover == 0) A.trace[(size_t)blockIdx.x * kTraceSlots + 24] = (unsigned long long)clock64();
        unsigned ticket;
        // The partials of this block are ordered before the ticket by the release; an acquire on EVERY block's ticket
        // would add an L1 invalidation to the tail of every block, so only the block that wins the ticket executes an
        // acquire fence (below) before it reads the other blocks' partials.
        asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;"
                     : "=r"(ticket) : "l"(&A.counters[rover * kCounterStride + 0]) : "memory");
        s.red_i[63] = (ticket == (unsigned)(A.nblocks - 1));
        if (s.red_i[63]) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        trace_stamp(A, 11);
    }
    __syncthreads();
    if (!s.red_i[63]) return;
    // every partial was published before its block's release + ticket; the fence above synchronises with them

    if (A.trace != nullptr && rover == 0 && tid == 0) A.trace[(size_t)blockIdx.x * kTraceSlots + 23] = (unsigned long long)clock64();
    const unsigned oob_count = __ldcg(&A.counters[rover * kCounterStride + 1]);
    const unsigned nan_count = __ldcg(&A.counters[rover * kCounterStride + 2]);
    const float* all_parts = A.partials + (size_t)rover * A.nblocks * stride;
    const int n_parts = A.nblocks;
    combine_and_finalize(p, st, all_parts, n_parts, s,
                         nominal1, nominal2, A.prev1 + (size_t)rover * T, A.prev2 + (size_t)rover * T,
                         A.opt_v + (size_t)rover * T, A.opt_w + (size_t)rover * T,
                         A.stats + (size_t)rover * kStatsStride,
                         A.rank_partial ? A.rank_partial + (size_t)rover * stride : nullptr, oob_count, nan_count,
                         (A.trace != nullptr && rover == 0) ? A.trace + (size_t)blockIdx.x * kTraceSlots : nullptr,
                         (rover == 0) ? A.host_cmd : nullptr, A.host_seq,
                         (rover == 0 && A.peers.world > 0 && A.rank_partial == nullptr) ? &A.peers : nullptr, A.spin_limit_ms);
    if (tid == 0 && A.loop.state != nullptr && A.rank_partial == nullptr)
        loop_advance(A, st, A.stats + (size_t)rover * kStatsStride);
    if (tid == 0) {                                       // re-arm for the next launch
        trace_stamp(A, 6);
        A.counters[rover * kCounterStride + 0] = 0u;
        A.counters[rover * kCounterStride + 1] = 0u;
        A.counters[rover * kCounterStride + 2] = 0u;
        A.counters[rover * kCounterStride + 3] = 0u;
    }
}

// ------------------------------------------------------------------ LL protocol (pipelined kernel): worker side
// 16-byte line {v0, seq, v1, seq}; volatile accesses go straight to L2 (or over NVLink into the peer's L2).
__device__ __forceinline__ void st_ll(uint4* line, float v0, float v1, uint32_t seq)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};"
                 ::"l"(line), "r"(__float_as_uint(v0)), "r"(seq), "r"(__float_as_uint(v1)), "r"(seq) : "memory");
}
__device__ __forceinline__ uint4 ld_ll(const uint4* line)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(line) : "memory");
    return v;
}
// (relaxed.gpu instead of volatile accesses, and polling without the 20 ns back-off, were measured at N = 1: no
// difference -- profiles/r2_ab_recurrence_and_poll_scope.txt)
#define MPPI_POLL_SLEEP() __nanosleep(20)
__device__ __forceinline__ bool ll_ok(const uint4& l, uint32_t seq) { return l.y == seq && l.w == seq; }

// Worker tail, part 1 (ONE warp): total cost of the warp's 32 samples from the critic sums the roles left in shared
// memory (critics_warp.py:325-329), block minimum / argmin, dead test, weights, and the three header lines, which leave
// the SM as soon as they exist.  `dry`: warm-up pass.  The tail is the first use of ~300 instructions and of a dozen
// kernel-parameter lines; after an L2 flush each of those is a DRAM round trip, taken one after the other by the
// ONE warp every other block is waiting for (measured: 2.2 us between "cost ready" and "header stored" on all 128
// blocks at once).  The obstacle warp, idle while the pipeline fills, therefore executes this same code once at the
// start of the kernel with every global side effect predicated off: instruction and constant caches are warm when it
// matters.  The function is inlined at ONE call site inside a two-trip loop so that both passes run the same addresses.
template <typename PS>
__device__ __forceinline__ void pipe_header(const FusedArgs& A, const MppiState& st, const SampleConsts& sc,
                                            const Smem& s, PS& ps, int rover, int lane, bool valid, unsigned snapshot_key,
                                            bool dry, const float* uhist)
{
    const MppiParams& p = A.p;
    const int T = p.T, K = p.K;
    const uint32_t seq = A.ll_seq;
    const unsigned FULL = 0xffffffffu;
    const bool flat = peers_flat(A) && (rover == 0);
    const int ndst = flat ? A.peers.world : 1;
    const int L = ll_lines(T);
    const size_t flat_off = flat ? (((size_t)(seq & 1u) * A.peers.world + A.peers.rank) * A.nblocks + blockIdx.x) * L : 0;
    uint4* local_slot = flat ? A.peers.ll[A.peers.rank] + flat_off
                             : A.ll + ((size_t)rover * A.nblocks + blockIdx.x) * L;

    // ---- cost (critics_warp.py:325-329)
    SampleAcc a;
    a.speed = ps.crit[0][lane]; a.slope = ps.crit[1][lane]; a.obs = ps.crit[2][lane];
    a.pf_near = ps.crit[3][lane]; a.last_x = ps.crit[4][lane]; a.last_y = ps.crit[5][lane];
    if (kXC) {
        a.effort = ps.crit[6][lane]; a.roll = ps.crit[7][lane]; a.pitch = ps.crit[8][lane];
        a.slope_c = ps.crit[9][lane]; a.pen_x = ps.crit[10][lane]; a.pen_y = ps.crit[11][lane];
    }
    float cost = sample_cost(p, st, sc, a, nullptr);
    unsigned my_oob = 0u, my_nan = 0u;
    const int k_local = blockIdx.x * 32 + lane;
    if (valid) {
        if (!dry) A.costs[(size_t)rover * K + k_local] = cost;
        for (int r = 0; r < 6; ++r) my_oob += (unsigned)ps.oob[r][lane];
        if (cost != cost) { my_nan = 1u; cost = CUDART_INF_F; }       // a NaN rollout gets zero weight
    } else {
        cost = CUDART_INF_F;
    }
    if (lane == 0 && !dry) trace_stamp(A, 8);

    // ---- block softmax partial
    float m_b = cost, w = 0.0f, s_b = 0.0f, s2_b = 0.0f;
    int arg_b = valid ? (int)(A.k_begin + (uint32_t)k_local) : 0x7fffffff;
    warp_min(m_b, arg_b);
    const unsigned oob_b = __reduce_add_sync(FULL, my_oob), nan_b = __reduce_add_sync(FULL, my_nan);
    if (lane == 0 && !dry)
        atomicMax(&A.minkey[rover], ((unsigned long long)A.mk_tag << 32) | (unsigned long long)min_key(m_b));
    const bool dead = partial_is_dead(p, m_b, __shfl_sync(FULL, snapshot_key, 0));
    unsigned mask = 0u;
    if (!dead) {
        if (valid && cost < CUDART_INF_F) w = fexp(fdiv(-(cost - m_b), p.lambda));   // critics_warp.py:346-347
        s_b = warp_sum(w);
        s2_b = warp_sum(w * w);
        mask = __ballot_sync(FULL, w > 0.0f);
        if (w > 0.0f && !dry) {
            const int pos = __popc(mask & ((1u << lane) - 1u));
            s.list_i[pos] = lane;
            s.list_w[pos] = w;
        }
    }
    // header lines 0..2 of every destination: lane 3 r + j stores line j of rank r (one lane per line)
    if (lane < 3 * ndst && !dry) {
        const int r = lane / 3, j = lane - 3 * r;
        const float v0 = (j == 0) ? m_b : (j == 1) ? s_b : __uint_as_float(oob_b);
        // bit 31 of the argmin word: "this block has out-of-range / NaN counts in line 2" (sample ids are < 2^31)
        const int arg_w = arg_b | (((oob_b | nan_b) != 0u) ? (int)0x80000000 : 0);
        const float v1 = (j == 0) ? __int_as_float(arg_w) : (j == 1) ? s2_b : __uint_as_float(nan_b);
        st_ll((flat ? A.peers.ll[r] + flat_off : local_slot) + j, v0, v1, seq);
    }
    int n_e = dead ? 0 : __popc(mask);
    if (lane == 0 && !dry) trace_stamp(A, 5);

    // ---- the A rows of a live partial, by this same warp when the block kept its sampled u in shared memory: plain
    //      reads, no block barrier, and the code is covered by the warm-up trip -- the rows leave ~0.2 us after the header
    //      instead of ~1.1 us (they are what the updater's fold waits for).  Summation order = accumulate_rows': G
    //      groups of entries (e = g, g + G, ...) summed separately, then the group sums folded in order.
    if (uhist != nullptr && n_e > 0) {
        __syncwarp();
        const int P = ll_pairs(T), B = blockDim.x;
        const int G = (B >= 2 * P && n_e > 1) ? B / P : 1;
        for (int pr = lane; pr < P; pr += 32) {
            const int t = 2 * pr;
            float a1a = 0.f, a1b = 0.f, a2a = 0.f, a2b = 0.f;
            for (int g = 0; g < G && g < n_e; ++g) {
                float g1a = 0.f, g1b = 0.f, g2a = 0.f, g2b = 0.f;
                for (int e = g; e < n_e; e += G) {
                    const float* u = uhist + (size_t)t * 64 + (s.list_i[e] & 31);      // (& 31: stale list in the dry trip)
                    const float we = s.list_w[e];
                    g1a += we * u[0];
                    g2a += we * u[32];
                    if (t + 1 < T) { g1b += we * u[64]; g2b += we * u[96]; }
                }
                a1a += g1a; a1b += g1b; a2a += g2a; a2b += g2b;
            }
            if (!dry) {
                for (int r = 0; r < ndst; ++r) {
                    uint4* d = (flat ? A.peers.ll[r] + flat_off : local_slot) + kLLHeaderLines + pr;
                    st_ll(d, a1a, a1b, seq);
                    st_ll(d + P, a2a, a2b, seq);
                }
            }
        }
        if (lane == 0 && !dry) trace_stamp(A, 11);
        n_e = 0;                                    // nothing left for pipe_rows
    }
    if (lane == 0 && !dry) s.red_i[62] = n_e;
}

// Worker tail, part 2 (every thread, after a barrier): the A rows of a live partial.
template <bool INJECT>
__device__ __forceinline__ void pipe_rows(const FusedArgs& A, const MppiState& st, const NoiseKey& nk, const Smem& s,
                                          int rover, const float* uhist)
{
    const int T = A.p.T, tid = threadIdx.x;
    const uint32_t seq = A.ll_seq;
    const bool flat = peers_flat(A) && (rover == 0);
    const int ndst = flat ? A.peers.world : 1;
    const int L = ll_lines(T), P = ll_pairs(T);
    const size_t flat_off = flat ? (((size_t)(seq & 1u) * A.peers.world + A.peers.rank) * A.nblocks + blockIdx.x) * L : 0;
    uint4* local_slot = flat ? A.peers.ll[A.peers.rank] + flat_off
                             : A.ll + ((size_t)rover * A.nblocks + blockIdx.x) * L;
    const int n_e = s.red_i[62];
    if (tid == 0) trace_stamp(A, 10);
    if (n_e == 0) return;               // dead, or no finite cost at all (sum w = 0: the updater skips the slot)
    accumulate_rows<INJECT>(A, st, nk, s, rover, 32, n_e, [&](int pr, float a1a, float a1b, float a2a, float a2b) {
        for (int r = 0; r < ndst; ++r) {
            uint4* d = (flat ? A.peers.ll[r] + flat_off : local_slot) + kLLHeaderLines + pr;
            st_ll(d, a1a, a1b, seq);
            st_ll(d + P, a2a, a2b, seq);
        }
    }, uhist);
    if (tid == 0) trace_stamp(A, 11);
}

// ------------------------------------------------------------------ LL protocol: the updater block
// Block `nblocks` of the pipelined launch.  It sits on an SM no worker uses (C2: 128 workers on 148 SMs), has loaded
// the old nominal and is polling the header lines long before the first worker finishes, so the tail of the
// iteration is: slowest worker's header store -> L2 -> one poll -> block-wide min / scales / fold over the whole
// block (one column pair per thread) -> command.  The command (v*, w*)[0] needs only step 0 of the final wheel filter
// and is stored to the host BEFORE the T-step recurrence and the (v, w) arrays are finished.
// `hdr`: shared memory for 4 x n header values.  Arithmetic and summation order are those of combine_and_finalize.
// `dry`: warm-up pass, run once while the workers roll out -- the same code on the previous launch's lines with every
// wait and every global store predicated off, so that instructions and kernel parameters are cached when the real
// pass starts (see pipe_header).
// (As a __noinline__ function the updater would leave the workers' code where it is, but its parameter loads turn into
// generic-address loads: +2.5 us measured, profiles/r2_ab_unroll_and_noinline.txt.)
__device__ __forceinline__ void pipe_updater(const FusedArgs& A, const MppiState& st, const Smem& s, float* hdr,
                                             int rover, float* nominal1, float* nominal2, bool dry)
{
    const MppiParams& p = A.p;
    const int T = p.T, tid = threadIdx.x, B = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t seq = A.ll_seq;
    const bool flat = peers_flat(A) && (rover == 0);
    const int n = A.nblocks * (flat ? A.peers.world : 1);
    const int L = ll_lines(T), P = ll_pairs(T);
    const size_t half_off = flat ? (size_t)(seq & 1u) * n * L : 0;
    const uint4* slots = flat ? A.peers.ll[A.peers.rank] + half_off : A.ll + (size_t)rover * A.nblocks * L;
    auto slot_of = [&](int b) -> const uint4* { return slots + (size_t)b * L; };
    unsigned long long* tr = (A.trace != nullptr && rover == 0 && !dry) ? A.trace + (size_t)blockIdx.x * kTraceSlots : nullptr;
    float* hm = hdr;
    float* hs = hdr + n;
    float* hs2 = hdr + 2 * n;
    int* harg = reinterpret_cast<int*>(hdr + 3 * n);
    const unsigned FULL = 0xffffffffu;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    if (tr != nullptr && tid == 0) tr[1] = globaltimer_ns();
#define UPD_CLK(slot) do { if (tr != nullptr && tid == 0) tr[slot] = (unsigned long long)clock64(); } while (0)
    UPD_CLK(16);

    // 1. headers.  Thread t owns the contiguous slots [t per, (t + 1) per): it polls lines 0 and 1 of all of them
    //    together (up to eight in flight per round trip; line 2, the out-of-range / NaN counts, is not needed before
    //    the command and is collected at the very end).
    float M = CUDART_INF_F;
    int arg = 0x7fffffff;
    unsigned any_counts = 0u;
    const int per = (n + B - 1) / B;
    const int b_lo = min(tid * per, n), b_hi = min(b_lo + per, n);
    for (int g0 = b_lo; g0 < b_hi; g0 += 8) {
        unsigned pending = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (g0 + j < b_hi) pending |= 1u << j;
        while (pending) {
            uint4 l[8][2];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (pending & (1u << j)) {
                    const uint4* lp = slot_of(g0 + j);
                    l[j][0] = ld_ll(lp); l[j][1] = ld_ll(lp + 1);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if ((pending & (1u << j)) && (dry || (ll_ok(l[j][0], seq) && ll_ok(l[j][1], seq)))) {
                    const int b = g0 + j;
                    const float mb = __uint_as_float(l[j][0].x);
                    const int kb = (int)(l[j][0].z & 0x7fffffffu);
                    any_counts |= l[j][0].z >> 31;
                    hm[b] = mb; harg[b] = (int)l[j][0].z;
                    hs[b] = __uint_as_float(l[j][1].x); hs2[b] = __uint_as_float(l[j][1].z);
                    pair_min(M, arg, mb, kb);
                    pending &= ~(1u << j);
                }
            }
            if (pending) { spin_guard(spins, t0, A.spin_limit_ms); MPPI_POLL_SLEEP(); }
        }
    }
    if (tr != nullptr && tid == 0) tr[2] = globaltimer_ns();
    UPD_CLK(17);
    block_min(M, arg, s);                    // two barriers
    if (tr != nullptr && tid == 0) tr[12] = globaltimer_ns();
    UPD_CLK(18);

    // 2. scale of every partial relative to M; the ones that carry weight (scale > 0 and sum w > 0: a dead partial
    //    published sum w = 0) are kept, in slot order -- with lambda = 0.3 one to three of them.  The slots of a thread
    //    are contiguous, so ONE exclusive scan of the per-thread counts orders the kept list (two barriers whatever
    //    the number of ranks).  The scale overwrites the slot's minimum, which is no longer needed.
    int mine = 0;
    for (int b = b_lo; b < b_hi; ++b) {
        float sc = 0.0f;
        if (hs[b] > 0.0f) sc = fexp(fdiv(-(hm[b] - M), p.lambda));
        hm[b] = sc;
        mine += (sc > 0.0f);
    }
    int incl;
    if (per == 1) {                                  // at most one slot per thread (n <= 192): a ballot is the scan
        const unsigned kept = __ballot_sync(FULL, mine != 0);
        incl = __popc(kept & (0xffffffffu >> (31 - lane)));
    } else {
        incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int up = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += up;
        }
    }
    if (lane == 31) s.red_i[32 + warp] = incl;
    __syncthreads();
    int pos = incl - mine, cnt = 0;
    for (int wv = 0; wv < (B >> 5); ++wv) {
        const int c = s.red_i[32 + wv];
        if (wv < warp) pos += c;
        cnt += c;
    }
    for (int b = b_lo; b < b_hi; ++b) {
        const float sc = hm[b];
        if (sc > 0.0f) { s.list_i[pos] = b; s.list_w[pos] = sc; ++pos; }
    }
    __syncthreads();
    UPD_CLK(19);
    // the A lines of the first kept partials are requested at once: the L2 round trip overlaps the sums below
    uint4 v0[4];
    if (tid < 2 * P) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < cnt) v0[j] = ld_ll(slot_of(s.list_i[j]) + kLLHeaderLines + tid);
    }

    // 3. S, S2 in slot order (every thread, identically) and the ordered fold: one 16-byte line = one step pair of one
    //    channel per thread; up to four kept partials in flight per round trip
    float S = 0.0f, S2 = 0.0f;
    for (int e = 0; e < cnt; ++e) {
        const int b = s.list_i[e];
        const float sc = s.list_w[e];
        S += hs[b] * sc;
        S2 += hs2[b] * sc * sc;
    }
    UPD_CLK(20);
    for (int l = tid; l < 2 * P; l += B) {
        float acc0 = 0.0f, acc1 = 0.0f;
        for (int e0 = 0; e0 < cnt; e0 += 4) {
            uint4 v[4];
            unsigned pending = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (e0 + j < cnt) pending |= 1u << j;
            const unsigned all = pending;
            bool have = (l == tid && e0 == 0);            // the first group was requested before the exponentials
            while (pending) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (pending & (1u << j))
                        v[j] = have ? v0[j] : ld_ll(slot_of(s.list_i[e0 + j]) + kLLHeaderLines + l);
                }
                have = false;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((pending & (1u << j)) && (dry || ll_ok(v[j], seq))) pending &= ~(1u << j);
                if (pending) { spin_guard(spins, t0, A.spin_limit_ms); MPPI_POLL_SLEEP(); }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (all & (1u << j)) {
                    const float sc = s.list_w[e0 + j];
                    acc0 += __uint_as_float(v[j].x) * sc;
                    acc1 += __uint_as_float(v[j].z) * sc;
                }
            }
        }
        const int c0 = (l < P) ? 2 * l : T + 2 * (l - P);
        s.acc[c0] = acc0;
        if (2 * ((l < P) ? l : l - P) + 1 < T) s.acc[c0 + 1] = acc1;
    }
    UPD_CLK(21);
    __syncthreads();
    UPD_CLK(22);
    if (tr != nullptr && tid == 0) tr[13] = globaltimer_ns();

    if (A.rank_partial != nullptr) {         // NCCL transport: publish the rank partial {M, S, argmin, S2, A1, A2}
        float* rp = A.rank_partial + (size_t)rover * partial_stride(T);
        if (tid == 0 && !dry) { rp[0] = M; rp[1] = S; rp[2] = __int_as_float(arg); rp[3] = S2; }
        for (int col = tid; col < 2 * T; col += B) if (!dry) rp[kPartialHeader + col] = s.acc[col];
        return;
    }

    // 4. updated nominal = A / S (critics_warp.py:363-376); the previous one is kept for replay and when no sample is
    //    valid (see combine_and_finalize).  nom1 / nom2 then hold the filter drive u k (1 - a) of step 5.
    float* prev1 = A.prev1 + (size_t)rover * T;
    float* prev2 = A.prev2 + (size_t)rover * T;
    float* opt_v = A.opt_v + (size_t)rover * T;
    float* opt_w = A.opt_w + (size_t)rover * T;
    float* stats = A.stats + (size_t)rover * kStatsStride;
    {
        const bool any_valid = S > 0.0f;
        const Recip rS = make_recip(S);
        const float oma = 1.0f - p.opt_a;
        for (int col = tid; col < 2 * T; col += B) {
            const float old = (col < T) ? s.nom1[col] : s.nom2[col - T];
            const float nv = any_valid ? fdiv(s.acc[col], rS) : old;
            const float drive = nv * p.opt_k * oma;
            s.acc[col] = nv;
            if (col < T) { if (!dry) { prev1[col] = old; nominal1[col] = nv; } s.nom1[col] = drive; }
            else { if (!dry) { prev2[col - T] = old; nominal2[col - T] = nv; } s.nom2[col - T] = drive; }
        }
    }
    __syncthreads();
    UPD_CLK(23);
    if (tr != nullptr && tid == 0) tr[14] = globaltimer_ns();

    // 5. the command first: step 0 of the (opt_k, opt_a) wheel filter (MPPI_isaac.py:672-692) is all (v*, w*)[0] needs
    const bool unicycle = (p.input_model == MPPI_INPUT_UNICYCLE);
    const Recip rw = make_recip(p.r_wheels);
    if (tid == 0) {
        float v, w;
        if (unicycle) {
            v = s.acc[0]; w = s.acc[T];
        } else {
            const float l = st.wheel_l * p.opt_a + s.nom1[0], r = st.wheel_r * p.opt_a + s.nom2[0];
            v = clampf((l + r) / 2.0f, p.v_min, p.v_max);
            w = clampf(fdiv(-l + r, rw), p.w_min, p.w_max);
        }
        const float ess = (S > 0.0f) ? fdiv(S * S, S2) : 0.0f;    // effective sample size
        if (!dry) {
            stats[6] = v; stats[7] = w;                    // the command, contiguous for one 8-byte D2H
            // zero-copy result for mppi_step_host: ONE 16-byte store {v*, w*, sequence number, 0} into mapped pinned
            // host memory; the host polls the sequence word
            if (rover == 0 && A.host_cmd != nullptr)
                *reinterpret_cast<float4*>(A.host_cmd) = make_float4(v, w, __uint_as_float(A.host_seq), 0.0f);
            if (tr != nullptr) tr[15] = globaltimer_ns();
            stats[0] = M;
            stats[1] = __int_as_float(arg);
            stats[2] = S;
            stats[3] = 0.0f;                               // out-of-range / NaN counts: accumulated in step 6
            stats[4] = 0.0f;
            stats[5] = ess;
            if (A.loop.state != nullptr) loop_advance(A, st, stats);   // closed loop: the plant step needs only the command
        }
        UPD_CLK(24);
    }
    if (unicycle) {
        for (int t = tid; t < T; t += B) if (!dry) { opt_v[t] = s.acc[t]; opt_w[t] = s.acc[T + t]; }
    } else {
        // the two T-step recurrences l <- l a + drive_l[t], r <- r a + drive_r[t] run on one thread each (warps 1 and 2:
        // warp 0 is busy with the command), 16 steps per trip in registers; same operations in the same order as a
        // one-thread loop (same-node A/B against 8-step scalar batches: 1900 vs 2900 cycles, -0.6 us per launch)
        //  -- results go to s.acc (the nominal copy is no longer needed): thread 0 is still reading the drives
        if (lane == 0 && (warp == 1 || warp == 2)) {
            const float* d = (warp == 1) ? s.nom1 : s.nom2;
            float* out = (warp == 1) ? s.acc : s.acc + T;
            float x = (warp == 1) ? st.wheel_l : st.wheel_r;
            // batches of 16 steps held in registers: 4 vector loads, 32 dependent operations, 4 vector stores (the drive
            // and result arrays are 16-byte aligned when T is a multiple of 4; otherwise the scalar loop below)
            int t = 0;
            if ((T & 3) == 0) {
                const float a = p.opt_a;
                for (; t + 16 <= T; t += 16) {
                    float4 c0 = *reinterpret_cast<const float4*>(d + t), c1 = *reinterpret_cast<const float4*>(d + t + 4);
                    float4 c2 = *reinterpret_cast<const float4*>(d + t + 8), c3 = *reinterpret_cast<const float4*>(d + t + 12);
                    x = x * a + c0.x; c0.x = x; x = x * a + c0.y; c0.y = x; x = x * a + c0.z; c0.z = x; x = x * a + c0.w; c0.w = x;
                    x = x * a + c1.x; c1.x = x; x = x * a + c1.y; c1.y = x; x = x * a + c1.z; c1.z = x; x = x * a + c1.w; c1.w = x;
                    x = x * a + c2.x; c2.x = x; x = x * a + c2.y; c2.y = x; x = x * a + c2.z; c2.z = x; x = x * a + c2.w; c2.w = x;
                    x = x * a + c3.x; c3.x = x; x = x * a + c3.y; c3.y = x; x = x * a + c3.z; c3.z = x; x = x * a + c3.w; c3.w = x;
                    *reinterpret_cast<float4*>(out + t) = c0; *reinterpret_cast<float4*>(out + t + 4) = c1;
                    *reinterpret_cast<float4*>(out + t + 8) = c2; *reinterpret_cast<float4*>(out + t + 12) = c3;
                }
            }
            for (; t < T; ++t) { x = x * p.opt_a + d[t]; out[t] = x; }
            if (tr != nullptr && warp == 1) tr[26] = (unsigned long long)clock64();
        }
        __syncthreads();
        UPD_CLK(27);
        for (int t = tid; t < T; t += B) {
            const float l = s.acc[t], r = s.acc[T + t];
            const float v = clampf((l + r) / 2.0f, p.v_min, p.v_max);
            const float w = clampf(fdiv(-l + r, rw), p.w_min, p.w_max);
            if (!dry) { opt_v[t] = v; opt_w[t] = w; }
        }
    }
    // 6. out-of-range / NaN counts: diagnostics, collected after everything the caller is waiting for, and only when
    //    some block flagged that it has any (bit 31 of its argmin word) -- the barrier is all this costs otherwise
    if (__syncthreads_or((int)any_counts)) {
        unsigned oob = 0u, nan = 0u;
        for (int b = b_lo; b < b_hi; ++b) {
            if (harg[b] >= 0) continue;
            uint4 l2 = ld_ll(slot_of(b) + 2);
            while (!(dry || ll_ok(l2, seq))) { spin_guard(spins, t0, A.spin_limit_ms); l2 = ld_ll(slot_of(b) + 2); }
            oob += l2.x; nan += l2.z;
        }
        oob = __reduce_add_sync(FULL, oob);
        nan = __reduce_add_sync(FULL, nan);
        if (lane == 0 && !dry) {
            if (oob) atomicAdd(reinterpret_cast<unsigned*>(stats) + 3, oob);
            if (nan) atomicAdd(reinterpret_cast<unsigned*>(stats) + 4, nan);
        }
    }
    UPD_CLK(28);
    if (tr != nullptr && tid == 0) tr[6] = globaltimer_ns();
#undef UPD_CLK
}

// ------------------------------------------------------------------ the fused kernel
// All T steps of one sample (monolithic kernel).  Steps come in (even, odd) pairs: one Philox call feeds both, and
// the odd step skips the wheel points (dead: the slope critic reads even steps only).
// SWP = true: fully software-pipelined rollout -- noise one pair ahead AND critics one step behind (critic_step(t - 1)
// next to chain_step(t), conditions as selects).  Bit-identical (whole GPU suite), measured on one B200
// (profiles/r2_ab_mono_swp.txt, r2_ab_mono_regs.txt): under the 128-register cap that four resident blocks per SM need
// the scheduler has no room to interleave (C3 / C4 / C5 lose 2-3 % to the 3.6 % more instructions; K = 16384 .. 32768
// gain 1-4 %), but with the 166 registers it asks for K = 8192 / 16384 / 32768 gain 20 / 19 / 8 % and K >= 65536 loses
// 5 %.  It is therefore a SECOND instantiation of the kernel (LOWOCC, own launch bounds -- the throughput instantiation's
// code and registers are untouched), launched for single-rover grids of at most kLowOccMaxSamples samples.
// SWP = false: noise one pair ahead only (K = 16384 / 32768: -3 % / -2 %, nothing lost at full occupancy;
// profiles/r2_ab_mono_noise_ahead.txt).
constexpr int kLowOccMaxSamples = 40960;
template <int PROJ, bool INJECT, bool CLAMP, bool SWP>
__device__ __forceinline__ void rollout_sample(const MppiParams& p, const MppiState& st, const Terr& ter,
                                               const SampleConsts& sc, const NoiseKey& nk, const Smem& s, SampleAcc& a,
                                               uint32_t kg, const float* eps1, const float* eps2)
{
    const int T = p.T;
    const UBounds ub = make_ubounds(p);
    const DumpPtrs nod = {};
    if (SWP) {
    // One thread's step is a long DEPENDENT chain (position -> cell -> normal -> tangent -> Rodrigues) followed by
    // critics that only read its outputs: a warp issues one instruction every ~4.6 cycles, so below four warps per
    // scheduler the launch time is that latency, not the issue rate (K = 16384 .. 65536 per GPU: the strong-scaling
    // end of C3, and C5).  The loop is therefore software-pipelined by hand: the iteration that runs chain_step(t)
    // also generates the noise of the NEXT pair of steps (counter-based, depends on nothing here; the pair after the
    // last one is computed and never used) and runs critic_step(t - 1) on the saved outputs of the previous step --
    // independent instruction streams inside one basic block (conditions are selects, see critic_step), which the
    // scheduler interleaves with the chain.  Every accumulator still receives its terms in the order t = 0, 1, 2, ...
    float e1a, e1b, e2a, e2b;
    if (INJECT) {
        e1a = eps1[0]; e2a = eps2[0];
        e1b = (1 < T) ? eps1[1] : 0.f;
        e2b = (1 < T) ? eps2[1] : 0.f;
    } else {
        noise_pair(nk, kg, 0u, e1a, e1b, e2a, e2b);
    }
    StepOut se, so;                                    // outputs of the pending even / odd step
    chain_step<PROJ, false, CLAMP, true>(p, ter, sc, a,
                                         sample_u(s.nom1, 0, T, st.sigma1, e1a, ub.lo1, ub.hi1),
                                         sample_u(s.nom2, 0, T, st.sigma2, e2a, ub.lo2, ub.hi2), se);
    int t = 0;
    for (; t + 2 < T; t += 2) {                        // steps t (pending), t + 1 and t + 2 exist
        float n1a, n1b, n2a, n2b;
        if (INJECT) {
            const int tb = min(t + 3, T - 1);
            n1a = eps1[t + 2]; n2a = eps2[t + 2]; n1b = eps1[tb]; n2b = eps2[tb];
        } else {
            noise_pair(nk, kg, (uint32_t)(t >> 1) + 1u, n1a, n1b, n2a, n2b);
        }
        chain_step<PROJ, false, CLAMP, true>(p, ter, sc, a,
                                             sample_u(s.nom1, t + 1, T, st.sigma1, e1b, ub.lo1, ub.hi1),
                                             sample_u(s.nom2, t + 1, T, st.sigma2, e2b, ub.lo2, ub.hi2), so);
        critic_step<PROJ, false, CLAMP, true, true>(p, st, ter, sc, a, t, se, 0.0f, nod, 0);
        chain_step<PROJ, false, CLAMP, true>(p, ter, sc, a,
                                             sample_u(s.nom1, t + 2, T, st.sigma1, n1a, ub.lo1, ub.hi1),
                                             sample_u(s.nom2, t + 2, T, st.sigma2, n2a, ub.lo2, ub.hi2), se);
        critic_step<PROJ, false, CLAMP, false, true>(p, st, ter, sc, a, t + 1, so, 0.0f, nod, 0);
        e1b = n1b; e2b = n2b;
    }
    critic_step<PROJ, false, CLAMP, true, true>(p, st, ter, sc, a, t, se, 0.0f, nod, 0);
    if (t + 1 < T) {
        chain_step<PROJ, false, CLAMP, true>(p, ter, sc, a,
                                             sample_u(s.nom1, t + 1, T, st.sigma1, e1b, ub.lo1, ub.hi1),
                                             sample_u(s.nom2, t + 1, T, st.sigma2, e2b, ub.lo2, ub.hi2), so);
        critic_step<PROJ, false, CLAMP, false, true>(p, st, ter, sc, a, t + 1, so, 0.0f, nod, 0);
    }
    } else {
    // Philox + Box-Muller of the NEXT pair of steps depends on nothing in this one; generated here, unconditionally
    // (the stream is counter-based: the pair after the last one is computed and never used), it shares a basic block
    // with the head of the chain and fills the latency of its four corner gathers instead of preceding them.
    float e1a, e1b, e2a, e2b;
    if (INJECT) {
        e1a = eps1[0]; e2a = eps2[0];
        e1b = (1 < T) ? eps1[1] : 0.f;
        e2b = (1 < T) ? eps2[1] : 0.f;
    } else {
        noise_pair(nk, kg, 0u, e1a, e1b, e2a, e2b);
    }
    for (int t = 0; t < T; t += 2) {
        float n1a, n1b, n2a, n2b;
        if (INJECT) {
            const int ta = min(t + 2, T - 1), tb = min(t + 3, T - 1);
            n1a = eps1[ta]; n2a = eps2[ta]; n1b = eps1[tb]; n2b = eps2[tb];
        } else {
            noise_pair(nk, kg, (uint32_t)(t >> 1) + 1u, n1a, n1b, n2a, n2b);
        }
        {
            const float u1 = sample_u(s.nom1, t, T, st.sigma1, e1a, ub.lo1, ub.hi1);
            const float u2 = sample_u(s.nom2, t, T, st.sigma2, e2a, ub.lo2, ub.hi2);
            sample_step<PROJ, false, CLAMP, true>(p, st, ter, sc, a, t, u1, u2, nod, 0);
        }
        if (t + 1 < T) {
            const float u1 = sample_u(s.nom1, t + 1, T, st.sigma1, e1b, ub.lo1, ub.hi1);
            const float u2 = sample_u(s.nom2, t + 1, T, st.sigma2, e2b, ub.lo2, ub.hi2);
            sample_step<PROJ, false, CLAMP, false>(p, st, ter, sc, a, t + 1, u1, u2, nod, 0);
        }
        e1a = n1a; e1b = n1b; e2a = n2a; e2b = n2b;
    }
    }
}

#ifndef MPPI_MONO_MINBLOCKS
#define MPPI_MONO_MINBLOCKS 2       // A/B knob: resident 256-thread blocks per SM the register allocation must allow
                                    // (the -DMPPI_XC build fits the same 128 registers without spilling; left alone it
                                    // takes 156 and C5 no longer fits one wave: 402 us instead of 300)
#endif
template <int PROJ, bool INJECT, bool LOWOCC = false>
__global__ void __launch_bounds__(kMaxBlock, LOWOCC ? 1 : MPPI_MONO_MINBLOCKS)
mppi_fused_kernel(const __grid_constant__ FusedArgs A)
{
    extern __shared__ float smem_raw[];
    const MppiParams& p = A.p;
    const int T = p.T, K = p.K, B = blockDim.x, tid = threadIdx.x;
    const int rover = blockIdx.y;
    const Smem s = carve(smem_raw, T, B, list_cap(A));
    if (tid == 0) {
        trace_stamp(A, 0);
        if (A.trace != nullptr && rover == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            A.trace[(size_t)blockIdx.x * kTraceSlots + 7] = smid;
        }
    }

    if (A.loop.state != nullptr && *reinterpret_cast<volatile const int32_t*>(A.loop.ctl + 1) != 0) return;   // goal reached
    const MppiState st = (A.loop.state != nullptr) ? *A.loop.state : ((A.states != nullptr) ? A.states[rover] : A.state);
    const MppiTerrain tr = (A.terrains != nullptr) ? A.terrains[rover] : A.terrain;
    const Terr ter = make_terr(tr);
    const SampleConsts sc = make_consts(p, st);
    const NoiseKey nk = make_noise_key(A.seed, A.offset, (uint32_t)rover);

    float* nominal1 = A.nominal1 + (size_t)rover * T;
    float* nominal2 = A.nominal2 + (size_t)rover * T;
    for (int t = tid; t < T; t += B) { s.nom1[t] = nominal1[t]; s.nom2[t] = nominal2[t]; }
    prefetch_terrain(p, st, tr, blockIdx.x, A.nblocks, tid, B);
    __syncthreads();

    const int k_local = blockIdx.x * B + tid;
    const bool valid = k_local < K;
    const uint32_t kg = A.k_begin + (uint32_t)k_local;
    const float* eps1 = INJECT ? A.noise + ((size_t)rover * 2 * K + k_local) * T : nullptr;
    const float* eps2 = INJECT ? eps1 + (size_t)K * T : nullptr;

    // ---------------- phase 1: rollout + critics
    float cost = CUDART_INF_F;
    unsigned my_oob = 0, my_nan = 0;
    if (tid == 0) { trace_stamp(A, 1); trace_stamp(A, 2); }
    if (valid) {
        SampleAcc a;
        sample_init<PROJ>(st, ter, a);
        if (terrain_window_safe(p, st, tr))
            rollout_sample<PROJ, INJECT, false, LOWOCC>(p, st, ter, sc, nk, s, a, kg, eps1, eps2);
        else
            rollout_sample<PROJ, INJECT, true, LOWOCC>(p, st, ter, sc, nk, s, a, kg, eps1, eps2);
        cost = sample_cost(p, st, sc, a, nullptr);
        A.costs[(size_t)rover * K + k_local] = cost;
        my_oob = (unsigned)(a.oob + unit_violation(a.dev));
        if (cost != cost) { my_nan = 1; cost = CUDART_INF_F; }   // a NaN rollout gets zero weight
    }
    if (tid == 0) { trace_stamp(A, 3); trace_stamp(A, 4); }
    // snapshot of the running minimum of the blocks that already finished (see partial_is_dead); any value is safe
    const unsigned snap = (tid == 0) ? __ldcg(&A.counters[rover * kCounterStride + 3]) : 0u;
    block_update<INJECT>(A, st, nk, s, rover, B, valid, tid, cost, my_oob, my_nan, nominal1, nominal2, snap);
}

// ------------------------------------------------------------------ warp-specialised fused kernel (latency regime)
// With K of a few thousand there is less than one warp of samples per SM and the step time is the latency of
// ONE warp issuing ~560 mostly dependent instructions per horizon step (profiles/r1_ncu_fused_v1.md).  This
// variant gives every group of 32 samples a CTA of six warps and cuts the step along its data dependences:
//   warps 0,1  noise     Philox + Box-Muller -> u (alternate chunks; the stream is counter-based)   [ring U]
//   warp  4    filter    wheel filter u -> (v, w); speed critic                                     [ring A]
//   warp  2    chain     position / DEM corners / normal / tangent / Rodrigues: the only recurrence  [ring B]
//   warp  3    wheels    wheel points + 2 DEM gathers + stride-2 slope critic
//   warp  5    obstacle  costmap gather + lethal penalty, near-goal path critic, last point
// Warp w runs on SM sub-partition w % 4, so the chain and the wheel warps own a scheduler each and the light
// roles share the other two.  Rings live in shared memory; chunks of kPipeChunk steps are handed over with
// mbarriers (noise -> filter, and every "slot free" edge that a producer with slack waits on) and, for the two
// conditions the CHAIN warp waits on, with release / acquire counters (see PipeSmem), so the chain warp runs its
// dependent chain without issuing the other roles' instructions or paying an mbarrier wait.  Arithmetic per sample
// is unchanged (same device functions => same bits).
// Unroll factors of the roles' per-chunk loops (A/B knobs).  The instruction caches are small (L0 ~6 KB per
// sub-partition, L1.5 32 KB per SM, /opt/skills/guides/B300_MICROARCH.md "I-cache") and the six roles run FIVE different
// loops at once.  Fully unrolled (chain 13.8 KB, filter 9 KB, noise 7.4 KB, wheels 3.7 KB, obstacle 3.2 KB) the hot
// set is ~37 KB: more than the L1.5 holds, and the chain warp -- whose latency IS the kernel's -- waits on instruction
// fetches from L2 whenever the layout of the kernel image shifts (same-node A/B, ns per horizon step: chain unroll
// 4 / 2 / 1 -> 287 / 279 / 271; noise + filter unroll 1 on top: 289 -> 279 after the updater code had moved the loops).
// The producers (noise, filter) have slack and run un-unrolled; the two critic roles keep their unrolling because
// they need the overlap between steps to stay ahead of the chain (un-unrolled they back-pressure it through ring B).
// Hot set now: chain 4.1 + noise 3.7 + filter 2.3 + wheels 3.7 + obstacle 3.2 = 17 KB.
#ifndef MPPI_CHAIN_UNROLL
#define MPPI_CHAIN_UNROLL 1
#endif
#ifndef MPPI_NOISE_UNROLL
#define MPPI_NOISE_UNROLL 1         // step pairs per trip (a chunk has kPipeChunk / 2)
#endif
#ifndef MPPI_FILTER_UNROLL
#define MPPI_FILTER_UNROLL 1
#endif
#ifndef MPPI_WHEELS_UNROLL
#define MPPI_WHEELS_UNROLL 2
#endif
#ifndef MPPI_OBST_UNROLL
#define MPPI_OBST_UNROLL 4
#endif
constexpr int kChainUnroll = MPPI_CHAIN_UNROLL;
constexpr int kNoiseUnroll = MPPI_NOISE_UNROLL, kFilterUnroll = MPPI_FILTER_UNROLL;
constexpr int kWheelsUnroll = MPPI_WHEELS_UNROLL, kObstUnroll = MPPI_OBST_UNROLL;
constexpr int kPipeStages = 4;      // ring depth in chunks
constexpr int kPipeChunk = 4;       // steps per chunk (even: noise comes in pairs of steps)
constexpr int kNoiseWarps = 2;
constexpr int kPipeThreads = 192;
template <bool B> struct FastTag { static constexpr bool value = B; };
enum { ROLE_NOISE0 = 0, ROLE_NOISE1 = 1, ROLE_CHAIN = 2, ROLE_WHEELS = 3, ROLE_FILTER = 4, ROLE_OBST = 5 };

struct PipeSmem {
    unsigned long long full_u[kPipeStages], empty_u[kPipeStages];
    unsigned long long empty_a[kPipeStages];
    unsigned long long full_b[kPipeStages];
    // What the CHAIN warp waits for is published as plain monotonic counters (st.release / ld.acquire at CTA scope)
    // instead of mbarrier phases: an mbarrier try_wait costs ~90 cycles even when the phase completed long ago, twice
    // per chunk on the one warp whose latency is the kernel's latency (14 % of its stall samples,
    // profiles/r1_ncu_c2_strict_pipe.md).  A counter is read with one LDS, issued a whole chunk before it is needed,
    // and one read usually covers several chunks because the filter runs up to kPipeStages chunks ahead.
    int a_ready;                                      // chunks of ring A published by the filter warp
    int b_done[2];                                    // chunks of ring B consumed by the wheel / obstacle warps
    float ring_u[kPipeStages][kPipeChunk][2][32];     // u1, u2
    float ring_a[kPipeStages][kPipeChunk][3][32];     // v, sin(w dt), cos(w dt)
    float ring_b[kPipeStages][kPipeChunk][8][32];     // x, y, n.xyz, cur.xyz
    float crit[kXC ? 12 : 6][32];                     // speed, slope, obs, pf_near, last_x, last_y
                                                      // (+ effort, roll, pitch, slope_c, pen_x, pen_y with the optional critics)
    int oob[6][32];
    unsigned long long tile_bar;                      // completes when the TMA copies of the DEM tile have landed
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void st_release_cta(int* p, int v)
{
    asm volatile("st.release.cta.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta(const int* p)
{
    int v;
    asm volatile("ld.acquire.cta.shared::cta.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
// one warp publishes "chunk count = v": every lane's ring stores happen-before lane 0's release through the warp sync
__device__ __forceinline__ void warp_publish(int* counter, int v, int lane)
{
    __syncwarp();
#ifdef MPPI_AB_PUBLISH_RELAXED
    // A/B knob, MEASUREMENT ONLY (make variant EXTRA=-DMPPI_AB_PUBLISH_RELAXED): a plain volatile store instead of
    // st.release.cta (= MEMBAR.ALL.CTA + STS).  Not a release in the PTX memory model; it exists to measure what the
    // three per-chunk MEMBARs of the sibling warps cost the chain warp.
    if (lane == 0) *reinterpret_cast<volatile int*>(counter) = v;
#else
    if (lane == 0) st_release_cta(counter, v);
#endif
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok = 0;
    for (unsigned spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (spin > (1u << 22)) __trap();              // a broken pipeline must fail loudly, never hang the GPU
    }
}

// TMA tensor copy global -> shared: ONE instruction moves the whole w x h box of the 2-D DEM tensor described by
// `desc` whose corner is (column c0, row c1); completion (bytes) is counted on `bar`.  (Row-by-row 1-D bulk copies were
// tried first: the TMA unit spends ~46 cycles per request, 186 rows cost ~4 us.)
__device__ __forceinline__ void tma_load_tile_2d(void* dst_smem, const TmaDesc* desc, int c0, int c1, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"(desc), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// projection_warp.py:8-48 on the shared-memory tile (indices proven in range by terrain_window_safe + the tile margins)
__device__ __forceinline__ Quad corners_tile(const Terr& t, const float* tile, const TileIdx& ti, float x, float y)
{
    const unsigned rel = tile_rel(t, ti, x, y);
    Quad q;
    q.q00 = tile_at(tile, rel);
    q.q01 = tile_at(tile, rel + 4u);
    q.q10 = tile_at(tile, rel + ti.w4);
    q.q11 = tile_at(tile, rel + ti.w4 + 4u);
    return q;
}

// chain role on the shared-memory DEM tile (3-D projection only)
template <int PROJ>
__device__ __forceinline__ void role_chain_tile(const MppiParams& p, const Terr& ter, const float* tile, const TileIdx& g,
                                                float& x, float& y, float3& prev, float v, float sn, float cs, float3& n,
                                                float& dev)
{
    update_position(x, y, prev, v, p.dt, dev);
    const Quad q = corners_tile(ter, tile, g, x, y);
    n = normal_on_grid(q, ter.res);
    const float3 tg = tangent(n, prev);
    prev = update_orientation_sc(tg, sn, cs, n, dev);
}

// Shared-memory layout of the pipelined kernel: [nominal, reduction scratch, sample lists, accumulators (carve)]
// [PipeSmem: barriers + rings] [DEM tile] [u history].  The UPDATER block has no rings, tile or history: its per-partial
// arrays (kept list 2 n + header values 4 n floats, n = partials it folds) overlay that region, so the workers' layout
// does not grow with the number of ranks.
__host__ __device__ inline size_t pipe_smem_offset_floats(int T)
{
    return (smem_floats(T, kPipeThreads, kPipeThreads) + 31) & ~(size_t)31;    // 128-byte aligned (TMA destination follows)
}
__host__ __device__ inline size_t pipe_updater_floats(int n) { return 6 * (size_t)n; }

template <int PROJ, bool INJECT>
__global__ void __launch_bounds__(kPipeThreads, 1)
mppi_fused_pipe_kernel(const __grid_constant__ FusedArgs A)
{
    extern __shared__ __align__(128) float smem_raw[];
    const MppiParams& p = A.p;
    const int T = p.T, K = p.K, tid = threadIdx.x;
    const int lane = tid & 31;
    // Warp w issues on SM sub-partition w % 4.  When two CTAs share an SM (grids beyond one CTA per SM: the second
    // wave of the block rasteriser lands on the same SMs), every second CTA swaps the chain and wheel warps so that the
    // two dependent chains do not compete for the same scheduler.
    int role = tid >> 5;
    if (((blockIdx.x / 148) & 1) && (role == 2 || role == 3)) role ^= 1;
    const int rover = blockIdx.y;
    const Smem s = carve(smem_raw, T, kPipeThreads, kPipeThreads);
    PipeSmem& ps = *reinterpret_cast<PipeSmem*>(smem_raw + pipe_smem_offset_floats(T));
    const bool is_updater = ((int)blockIdx.x == A.nblocks);                   // LL protocol: grid.x = nblocks + 1
    float* tile = reinterpret_cast<float*>(reinterpret_cast<char*>(&ps) + ((sizeof(PipeSmem) + 127) & ~(size_t)127));
    // every sampled u of the block, [t][channel][lane], kept for the A rows of the update when it fits beside the tile
    float* uhist = A.uhist ? tile + (size_t)A.tile.w * A.tile.h : nullptr;
    if (tid == 0) {
        trace_stamp(A, 0);
        if (A.trace != nullptr && rover == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            A.trace[(size_t)blockIdx.x * kTraceSlots + 7] = smid;
        }
    }

    // Programmatic dependent launch (back-to-back steps: closed loop, un-flushed streams).  Every block lets the NEXT
    // launch in the stream be scheduled at once: its blocks take the SMs this launch's workers leave, initialise their
    // barriers, and then stop at griddepcontrol.wait -- which returns when THIS grid has completed and its writes
    // (nominal, loop state, running minimum) are visible.  Nothing above the wait reads or writes global memory that
    // a launch produces; with an ordinary launch on either side both instructions do nothing.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!is_updater && tid == 0) {
        for (int i = 0; i < kPipeStages; ++i) {
            mbar_init(&ps.full_u[i], 32); mbar_init(&ps.empty_u[i], 32);
            mbar_init(&ps.empty_a[i], 32);
            mbar_init(&ps.full_b[i], 32);
        }
        ps.a_ready = 0; ps.b_done[0] = 0; ps.b_done[1] = 0;
        mbar_init(&ps.tile_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (A.loop.state != nullptr && *reinterpret_cast<volatile const int32_t*>(A.loop.ctl + 1) != 0) return;   // goal reached
    const MppiState st = (A.loop.state != nullptr) ? *A.loop.state : ((A.states != nullptr) ? A.states[rover] : A.state);
    const MppiTerrain tr = (A.terrains != nullptr) ? A.terrains[rover] : A.terrain;
    const Terr ter = make_terr(tr);
    const SampleConsts sc = make_consts(p, st);
    const NoiseKey nk = make_noise_key(A.seed, A.offset, (uint32_t)rover);
    // DEM tile: the host fixes the box (w x h, part of the TMA descriptor); its corner follows the robot, which may
    // live in device memory (closed loop), so it is placed here with the same index formula as the gathers
    DemTile tg = A.tile;
    if (tg.w > 0) {
        const int hc = (tg.h - 2) / 2;
        int ic, jc;
        dem_index(ter, st.x, st.y, ic, jc);
        tg.i0 = (ic - hc) & ~3;
        tg.j0 = jc - hc;
        if (!terrain_window_safe(p, st, tr) || tg.i0 < 0 || tg.j0 < 0 || tg.i0 + tg.w > tr.grid_size ||
            tg.j0 + tg.h > tr.grid_size)
            tg.w = 0;
    }

    const TileIdx tidx = make_tile_idx(tg.i0, tg.j0, tg.w, tg.h);

    float* nominal1 = A.nominal1 + (size_t)rover * T;
    float* nominal2 = A.nominal2 + (size_t)rover * T;
    // (the updater block's code is placed AFTER the workers' -- see the end of the kernel)
    if (__builtin_expect(!is_updater, 1)) {
    if (tid == 0) {
        // The TMA copy of the DEM tile starts before the block even synchronises, so that the tile lands while the
        // nominal is loaded and the noise / filter stages fill the pipeline.
        if (tg.w > 0) {
            mbar_expect_tx(&ps.tile_bar, (unsigned)(tg.w * tg.h) * 4u);
            tma_load_tile_2d(tile, &A.dem_desc, tg.i0, tg.j0, &ps.tile_bar);
        }
    }
    for (int t = tid; t < T; t += kPipeThreads) { s.nom1[t] = nominal1[t]; s.nom2[t] = nominal2[t]; }
    __syncthreads();
    if (tid == 0) trace_stamp(A, 1);

    const int k_local = blockIdx.x * 32 + lane;
    const bool valid = k_local < K;
    const int k_read = valid ? k_local : K - 1;                 // idle lanes shadow the last sample (results unused)
    const uint32_t kg = A.k_begin + (uint32_t)k_read;
    const int nchunks = (T + kPipeChunk - 1) / kPipeChunk;
    // chunks that take the check-free path: full chunks of a rollout that provably stays inside the maps
    const int nfast = terrain_window_safe(p, st, tr) ? T / kPipeChunk : 0;
    int oob = 0;
    unsigned snap = 0u;

    // Two trips: trip 0 is the warm-up of the tail (pipe_header, dry) by the obstacle warp while the pipeline fills,
    // trip 1 runs the roles and then the real tail on warp 0 -- ONE copy of the tail's code, the same addresses.
#ifdef MPPI_AB_NO_WARM
    constexpr int kFirstPhase = 1;                        // A/B knob: no warm-up trip
#else
    constexpr int kFirstPhase = 0;
#endif
#pragma unroll 1
    for (int phase = kFirstPhase; phase < 2; ++phase) {
    const bool dry = (phase == 0);
    if (!dry) {
    if (role < kNoiseWarps) {
        // ---- noise: eps -> u for chunks c = role, role + kNoiseWarps, ...
        const UBounds ub = make_ubounds(p);
        const float* eps1 = INJECT ? A.noise + ((size_t)rover * 2 * K + k_read) * T : nullptr;
        const float* eps2 = INJECT ? eps1 + (size_t)K * T : nullptr;
        for (int c = role; c < nchunks; c += kNoiseWarps) {
            const int sg = c % kPipeStages;
            mbar_wait(&ps.empty_u[sg], ((c / kPipeStages) & 1) ^ 1);
#pragma unroll kNoiseUnroll
            for (int i = 0; i < kPipeChunk; i += 2) {
                const int t = c * kPipeChunk + i;
                if (t < T) {
                    float e1a, e1b, e2a, e2b;
                    if (INJECT) {
                        e1a = eps1[t]; e2a = eps2[t];
                        e1b = (t + 1 < T) ? eps1[t + 1] : 0.f;
                        e2b = (t + 1 < T) ? eps2[t + 1] : 0.f;
                    } else {
                        noise_pair(nk, kg, (uint32_t)(t >> 1), e1a, e1b, e2a, e2b);
                    }
                    const float u1a = sample_u(s.nom1, t, T, st.sigma1, e1a, ub.lo1, ub.hi1);
                    const float u2a = sample_u(s.nom2, t, T, st.sigma2, e2a, ub.lo2, ub.hi2);
                    ps.ring_u[sg][i][0][lane] = u1a;
                    ps.ring_u[sg][i][1][lane] = u2a;
                    if (uhist != nullptr) { uhist[t * 64 + lane] = u1a; uhist[t * 64 + 32 + lane] = u2a; }
                    if (t + 1 < T) {
                        const float u1b = sample_u(s.nom1, t + 1, T, st.sigma1, e1b, ub.lo1, ub.hi1);
                        const float u2b = sample_u(s.nom2, t + 1, T, st.sigma2, e2b, ub.lo2, ub.hi2);
                        ps.ring_u[sg][i + 1][0][lane] = u1b;
                        ps.ring_u[sg][i + 1][1][lane] = u2b;
                        if (uhist != nullptr) { uhist[t * 64 + 64 + lane] = u1b; uhist[t * 64 + 96 + lane] = u2b; }
                    }
                }
            }
            mbar_arrive(&ps.full_u[sg]);
        }
    } else if (role == ROLE_FILTER) {
        // ---- wheel filter u -> (v, w) (sequential in t) + speed critic
        float wl = st.wheel_l, wr = st.wheel_r, speed = 0.0f, effort = 0.0f;
        for (int c = 0; c < nchunks; ++c) {
            const int sg = c % kPipeStages;
            const unsigned ph = (c / kPipeStages) & 1;
            mbar_wait(&ps.full_u[sg], ph);
            mbar_wait(&ps.empty_a[sg], ph ^ 1);
#pragma unroll kFilterUnroll
            for (int i = 0; i < kPipeChunk; ++i) {
                const int t = c * kPipeChunk + i;
                if (t < T) {
                    float v, sn, cs;
                    const float u1 = ps.ring_u[sg][i][0][lane], u2 = ps.ring_u[sg][i][1][lane];
                    role_filter(p, sc, wl, wr, u1, u2, v, sn, cs, speed);
                    if (kXC) effort += u1 * u1 + u2 * u2;
                    ps.ring_a[sg][i][0][lane] = v; ps.ring_a[sg][i][1][lane] = sn; ps.ring_a[sg][i][2][lane] = cs;
                }
            }
            mbar_arrive(&ps.empty_u[sg]);
            warp_publish(&ps.a_ready, c + 1, lane);
        }
        ps.crit[0][lane] = speed;
        if (kXC) ps.crit[6][lane] = effort;
    } else if (role == ROLE_CHAIN) {
        // ---- chain: the recurrence
        float x = st.x, y = st.y;
        float3 prev = make_float3(st.hx, st.hy, st.hz), n;
        float dev = 0.0f;
        if (PROJ == MPPI_PROJ_3D) {
            int i0, j0;
            const Quad q = corners(ter, x, y, i0, j0, oob);
            prev = tangent(normal_on_grid(q, ter.res), prev);
        }
        if (lane == 0) trace_stamp(A, 2);
        const bool use_tile = (tg.w > 0) && (PROJ == MPPI_PROJ_3D) && (nfast > 0);
        if (use_tile) mbar_wait(&ps.tile_bar, 0);
        if (lane == 0) trace_stamp(A, 25);                   // DEM tile landed
        // FAST: a full chunk whose cell indices need no clamping (terrain_window_safe) -> no per-step checks at all.
        // TILE: additionally the four corner gathers read the shared-memory DEM tile.
        int a_seen = 0, b_seen = 0;          // last counter values this warp has read (monotonic, so stale is safe)
        auto chunk = [&](int c, auto fast_tag, auto tile_tag) {
            constexpr bool FAST = decltype(fast_tag)::value;
            constexpr bool TILE = decltype(tile_tag)::value;
            const int sg = c % kPipeStages;
            // chunk c of ring A published, and the slot of ring B released by both critic warps (they finished chunk
            // c - kPipeStages); rarely taken after the pipeline has filled: the values were prefetched a chunk ago
            for (unsigned spin = 0; a_seen <= c || b_seen < c + 1 - kPipeStages; ++spin) {
                a_seen = ld_acquire_cta(&ps.a_ready);
                b_seen = min(ld_acquire_cta(&ps.b_done[0]), ld_acquire_cta(&ps.b_done[1]));
                if (spin > (1u << 24)) __trap();          // a broken pipeline must fail loudly, never hang the GPU
            }
            const int a_next = ld_acquire_cta(&ps.a_ready);          // for chunk c + 1: in flight during this chunk
            const int b_next0 = ld_acquire_cta(&ps.b_done[0]), b_next1 = ld_acquire_cta(&ps.b_done[1]);
#pragma unroll kChainUnroll
            for (int i = 0; i < kPipeChunk; ++i) {
                const int t = c * kPipeChunk + i;
                if (FAST || t < T) {
                    const float v = ps.ring_a[sg][i][0][lane];
                    const float sn = ps.ring_a[sg][i][1][lane], cs = ps.ring_a[sg][i][2][lane];
                    if (TILE) role_chain_tile<PROJ>(p, ter, tile, tidx, x, y, prev, v, sn, cs, n, dev);
                    else role_chain<PROJ, !FAST>(p, ter, x, y, prev, v, sn, cs, n, oob, dev);
                    float* o = &ps.ring_b[sg][i][0][lane];
                    o[0] = x; o[32] = y;
                    if ((i & 1) == 0) {                  // only even steps feed the wheel / slope role
                        o[64] = n.x; o[96] = n.y; o[128] = n.z;
                        o[160] = prev.x; o[192] = prev.y; o[224] = prev.z;
                    }
                }
            }
            mbar_arrive(&ps.empty_a[sg]);
            mbar_arrive(&ps.full_b[sg]);
            a_seen = a_next; b_seen = min(b_next0, b_next1);
        };
        int c = 0;
        if (use_tile) { for (; c < nfast; ++c) chunk(c, FastTag<true>{}, FastTag<true>{}); }
        else { for (; c < nfast; ++c) chunk(c, FastTag<true>{}, FastTag<false>{}); }
        for (; c < nchunks; ++c) chunk(c, FastTag<false>{}, FastTag<false>{});
        oob += unit_violation(dev);
        if (lane == 0) trace_stamp(A, 3);
    } else if (role == ROLE_WHEELS) {
        // ---- wheels + slope critic
        float3 lw_e = make_float3(0.f, 0.f, 0.f), rw_e = lw_e, ctr_e = lw_e;
        float slope = 0.0f, roll = 0.0f, pitch = 0.0f, slope_c = 0.0f;
        const bool use_tile = (tg.w > 0) && (PROJ == MPPI_PROJ_3D) && (nfast > 0);
        if (use_tile) mbar_wait(&ps.tile_bar, 0);
        auto chunk = [&](int c, auto fast_tag, auto tile_tag) {
            constexpr bool FAST = decltype(fast_tag)::value;
            constexpr bool TILE = decltype(tile_tag)::value;
            const int sg = c % kPipeStages;
            mbar_wait(&ps.full_b[sg], (c / kPipeStages) & 1);
#pragma unroll kWheelsUnroll
            for (int i = 0; i < kPipeChunk; i += 2) {                 // even steps only feed the critic
                const int t = c * kPipeChunk + i;
                if (FAST || t < T) {
                    const float* o = &ps.ring_b[sg][i][0][lane];
                    role_wheels<PROJ, !FAST, TILE>(p, ter, t, o[0], o[32], make_float3(o[64], o[96], o[128]),
                                                   make_float3(o[160], o[192], o[224]), lw_e, rw_e, slope, oob,
                                                   tile, &tidx);
                    if (kXC) {
                        // optional critics of the even steps: lw_e / rw_e now hold this step's wheel points; the
                        // body height is re-interpolated here (the chain role does not need it)
                        int bi, bj, dummy = 0;
                        const Quad q = TILE ? corners_tile(ter, tile, tidx, o[0], o[32])
                                            : corners<!FAST>(ter, o[0], o[32], bi, bj, dummy);
                        extras_even(p, t, o[0], o[32], bilinear(o[0], o[32], q, ter.rres), lw_e.z, rw_e.z, o[224],
                                    roll, pitch, slope_c, ctr_e);
                    }
                }
            }
            warp_publish(&ps.b_done[0], c + 1, lane);
        };
        int c = 0;
        if (use_tile) { for (; c < nfast; ++c) chunk(c, FastTag<true>{}, FastTag<true>{}); }
        else { for (; c < nfast; ++c) chunk(c, FastTag<true>{}, FastTag<false>{}); }
        for (; c < nchunks; ++c) chunk(c, FastTag<false>{}, FastTag<false>{});
        ps.crit[1][lane] = slope;
        if (kXC) { ps.crit[7][lane] = roll; ps.crit[8][lane] = pitch; ps.crit[9][lane] = slope_c; }
    } else {
        // ---- obstacle + near-goal path critic + last point.  This warp idles while the pipeline fills: it first
        //      warms L2 with this block's share of the reachable terrain window.
        prefetch_terrain(p, st, tr, blockIdx.x, A.nblocks, lane, 32);
        float obs = 0.0f, pf_near = 0.0f, lx = st.x, ly = st.y, px = st.x, py = st.y;
        auto chunk = [&](int c, auto fast_tag) {
            constexpr bool FAST = decltype(fast_tag)::value;
            const int sg = c % kPipeStages;
            mbar_wait(&ps.full_b[sg], (c / kPipeStages) & 1);
#pragma unroll kObstUnroll
            for (int i = 0; i < kPipeChunk; ++i) {
                const int t = c * kPipeChunk + i;
                if (FAST || t < T) {
                    if (kXC) { px = lx; py = ly; }
                    lx = ps.ring_b[sg][i][0][lane]; ly = ps.ring_b[sg][i][1][lane];
                    role_obstacle<!FAST>(p, st, ter, sc, t, lx, ly, pf_near, obs, oob);
                }
            }
            warp_publish(&ps.b_done[1], c + 1, lane);
        };
        int c = 0;
        for (; c < nfast; ++c) chunk(c, FastTag<true>{});
        for (; c < nchunks; ++c) chunk(c, FastTag<false>{});
        ps.crit[2][lane] = obs; ps.crit[3][lane] = pf_near; ps.crit[4][lane] = lx; ps.crit[5][lane] = ly;
        if (kXC) { ps.crit[10][lane] = px; ps.crit[11][lane] = py; }
    }
    ps.oob[role][lane] = oob;
    // thread 0 (a noise warp, done long before the chain) reads the running minimum of the blocks that have already
    // finished while it waits for the other roles: the L2 round trip is off the critical path (see partial_is_dead)
    if (tid == 0) {
        // ... but not too early: thread 0 idles until the critic warps are within one chunk of the end, so that the
        // snapshot covers the blocks that finished up to ~1 us before this one
        for (unsigned spin = 0; ld_acquire_cta(&ps.b_done[1]) < nchunks - 1 && spin < (1u << 16); ++spin) __nanosleep(128);
        const unsigned long long mk = __ldcg(&A.minkey[rover]);
        snap = ((uint32_t)(mk >> 32) == A.mk_tag) ? (uint32_t)mk : 0u;       // entries of earlier launches do not count
    }
    __syncthreads();
    if (tid == 0) trace_stamp(A, 4);
    }   // !dry: roles

    // ---- tail part 1: cost, block partial, header lines (one warp; see pipe_header for the dry trip)
    // (which idle warp runs the warm-up trip -- obstacle, wheels, filter, chain -- makes no difference at the end of the
    // launch: profiles/r2_ab_dry_role.txt)
    if (role == (dry ? ROLE_OBST : ROLE_NOISE0)) pipe_header(A, st, sc, s, ps, rover, lane, valid, snap, dry, uhist);
    }   // phase
    __syncthreads();
    // ---- tail part 2: the A rows of a live partial
    pipe_rows<INJECT>(A, st, nk, s, rover, uhist);
    return;
    }   // !is_updater

    // ---- the updater block.  Its code sits after the workers' on purpose: the position of the role loops in the kernel
    //      image (instruction-cache footprint / alignment) moves the rollout by several percent (profiles/r2_timeline.md).
    {
#pragma unroll 1
#ifdef MPPI_AB_NO_WARM
        for (int pass = 1; pass < 2; ++pass) {
#else
        for (int pass = 0; pass < 2; ++pass) {            // pass 0: warm-up on the previous launch's lines (dry)
#endif
            for (int t = tid; t < T; t += kPipeThreads) { s.nom1[t] = nominal1[t]; s.nom2[t] = nominal2[t]; }
            __syncthreads();
            // the updater's kept list and header values overlay the rings / tile region it does not use
            Smem su = s;
            float* ubase = smem_raw + pipe_smem_offset_floats(T);
            su.list_i = reinterpret_cast<int*>(ubase);
            su.list_w = ubase + list_cap(A);
            const MppiState st_u = st;
            pipe_updater(A, st_u, su, ubase + 2 * (size_t)list_cap(A), rover, nominal1, nominal2, pass == 0);
            __syncthreads();
        }
        return;
    }
}

// ------------------------------------------------------------------ rank-partial combine (multi-GPU epilogue)
__global__ void __launch_bounds__(kMaxBlock) mppi_combine_kernel(const __grid_constant__ CombineArgs A)
{
    extern __shared__ float smem_raw[];
    const int T = A.p.T, B = blockDim.x;
    const Smem s = carve(smem_raw, T, B, A.n_parts);
    for (int t = threadIdx.x; t < T; t += B) { s.nom1[t] = A.nominal1[t]; s.nom2[t] = A.nominal2[t]; }
    __syncthreads();
    combine_and_finalize(A.p, A.state, A.parts, A.n_parts, s, A.nominal1, A.nominal2, A.prev1, A.prev2,
                         A.opt_v, A.opt_w, A.stats, nullptr, 0u, 0u, nullptr, nullptr, 0u);
}

// ------------------------------------------------------------------ validation / visualiser dump (unfused view)
template <int PROJ, bool INJECT>
__global__ void __launch_bounds__(128) mppi_dump_kernel(const __grid_constant__ DumpArgs A)
{
    const MppiParams& p = A.p;
    const int T = p.T, K = p.K;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const Terr ter = make_terr(A.terrain);
    const SampleConsts sc = make_consts(p, A.state);
    const NoiseKey nk = make_noise_key(A.seed, A.offset, 0u);
    DumpPtrs d;
    d.u1 = A.d.u1; d.u2 = A.d.u2; d.v = A.d.v; d.w = A.d.w;
    d.traj = A.d.traj; d.heading = A.d.heading; d.lw = A.d.lw; d.rw = A.d.rw;
    d.dem_ij = A.d.dem_ij; d.lw_ij = A.d.lw_ij; d.rw_ij = A.d.rw_ij; d.cm_ij = A.d.cm_ij;
    SampleAcc a;
    sample_init<PROJ>(A.state, ter, a);
    const UBounds ub = make_ubounds(p);
    const float* eps1 = INJECT ? A.noise + (size_t)k * T : nullptr;
    const float* eps2 = INJECT ? eps1 + (size_t)K * T : nullptr;
    for (int t = 0; t < T; t += 2) {
        float e1a, e1b, e2a, e2b;
        if (INJECT) {
            e1a = eps1[t]; e2a = eps2[t];
            e1b = (t + 1 < T) ? eps1[t + 1] : 0.f;
            e2b = (t + 1 < T) ? eps2[t + 1] : 0.f;
        } else {
            noise_pair(nk, A.k_begin + (uint32_t)k, (uint32_t)(t >> 1), e1a, e1b, e2a, e2b);
        }
        {
            const float u1 = sample_u(A.nominal1, t, T, A.state.sigma1, e1a, ub.lo1, ub.hi1);
            const float u2 = sample_u(A.nominal2, t, T, A.state.sigma2, e2a, ub.lo2, ub.hi2);
            sample_step<PROJ, true>(p, A.state, ter, sc, a, t, u1, u2, d, (size_t)k * T + t);
        }
        if (t + 1 < T) {
            const float u1 = sample_u(A.nominal1, t + 1, T, A.state.sigma1, e1b, ub.lo1, ub.hi1);
            const float u2 = sample_u(A.nominal2, t + 1, T, A.state.sigma2, e2b, ub.lo2, ub.hi2);
            sample_step<PROJ, true>(p, A.state, ter, sc, a, t + 1, u1, u2, d, (size_t)k * T + t + 1);
        }
    }
    float cr[4], cx[6];
    const float cost = sample_cost<true>(p, A.state, sc, a, cr, cx);
    if (A.costs) A.costs[k] = cost;
    if (A.d.critics) { for (int i = 0; i < 4; ++i) A.d.critics[4 * k + i] = cr[i]; }
    if (A.d.critics_ext) { for (int i = 0; i < 6; ++i) A.d.critics_ext[6 * k + i] = cx[i]; }
}

// Strided export for the visualiser: the driver shows every 50th sampled trajectory at every 10th step
// (visual_terrain_stack_full_terrain.py:252-261, 520-528), i.e. 2 % of the samples and 0.2 % of the K x T points.
// One thread per exported sample re-rolls it (same device functions, same bits as the step) and stores only the
// requested points.
template <int PROJ, bool INJECT>
__global__ void __launch_bounds__(128) mppi_export_kernel(const __grid_constant__ ExportArgs A)
{
    const MppiParams& p = A.p;
    const int T = p.T, K = p.K;
    const int ke = blockIdx.x * blockDim.x + threadIdx.x;          // exported sample index
    const int k = ke * A.k_stride;
    if (k >= K) return;
    const int nt = (T + A.t_stride - 1) / A.t_stride;
    const Terr ter = make_terr(A.terrain);
    const SampleConsts sc = make_consts(p, A.state);
    const NoiseKey nk = make_noise_key(A.seed, A.offset, 0u);
    const UBounds ub = make_ubounds(p);
    DumpPtrs none = {}, pts = {};
    pts.traj = A.points;
    SampleAcc a;
    sample_init<PROJ>(A.state, ter, a);
    const float* eps1 = INJECT ? A.noise + (size_t)k * T : nullptr;
    const float* eps2 = INJECT ? eps1 + (size_t)K * T : nullptr;
    for (int t = 0; t < T; t += 2) {
        float e1a, e1b, e2a, e2b;
        if (INJECT) {
            e1a = eps1[t]; e2a = eps2[t];
            e1b = (t + 1 < T) ? eps1[t + 1] : 0.f;
            e2b = (t + 1 < T) ? eps2[t + 1] : 0.f;
        } else {
            noise_pair(nk, (uint32_t)k, (uint32_t)(t >> 1), e1a, e1b, e2a, e2b);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int tt = t + h;
            if (tt >= T) break;
            const float u1 = sample_u(A.nominal1, tt, T, A.state.sigma1, h ? e1b : e1a, ub.lo1, ub.hi1);
            const float u2 = sample_u(A.nominal2, tt, T, A.state.sigma2, h ? e2b : e2a, ub.lo2, ub.hi2);
            const bool keep = (tt % A.t_stride) == 0;
            sample_step<PROJ, true>(p, A.state, ter, sc, a, tt, u1, u2, keep ? pts : none,
                                    (size_t)ke * nt + (size_t)(tt / A.t_stride));
        }
    }
}

// weights with the global minimum, as the reference intends (critics_warp.py:338-347, run_mppi.py:222-226)
__global__ void __launch_bounds__(1024) mppi_weights_kernel(const float* costs, int K, float lambda, float* weights)
{
    __shared__ float red[32];
    float m = CUDART_INF_F;
    for (int k = threadIdx.x; k < K; k += blockDim.x) m = fminf(m, costs[k]);
    for (int off = 16; off > 0; off >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < (int)((blockDim.x + 31) >> 5); ++w) m = fminf(m, red[w]);
    for (int k = threadIdx.x; k < K; k += blockDim.x) weights[k] = fexp(fdiv(-(costs[k] - m), lambda));
}

// optimal-trajectory rollout, dim = 1 (MPPI_isaac.py:696-720): always the 3-D kernel, driven by (v*, w*)
__global__ void mppi_sim_kernel(const __grid_constant__ SimArgs A)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const MppiParams& p = A.p;
    const Terr ter = make_terr(A.terrain);
    int oob = 0, i, j;
    float dev = 0.0f;
    float x = A.state.x, y = A.state.y;
    Quad q = corners(ter, x, y, i, j, oob);
    float3 n = normal_on_grid(q, ter.res);
    float3 prev = tangent(n, make_float3(A.state.hx, A.state.hy, A.state.hz));
    for (int t = 0; t < p.T; ++t) {
        update_position(x, y, prev, A.opt_v[t], p.dt, dev);
        q = corners(ter, x, y, i, j, oob);
        const float h = bilinear(x, y, q, ter.rres);
        n = normal_on_grid(q, ter.res);
        prev = tangent(n, prev);
        const float3 cur = update_orientation(prev, A.opt_w[t], n, p.dt, dev);
        A.sim_traj[3 * t] = x; A.sim_traj[3 * t + 1] = y; A.sim_traj[3 * t + 2] = h;
        A.sim_heading[3 * t] = cur.x; A.sim_heading[3 * t + 1] = cur.y; A.sim_heading[3 * t + 2] = cur.z;
        prev = cur;
    }
}

// ------------------------------------------------------------------ test hooks
__global__ void mppi_detmath_kernel(int fn, const float* x, float* y0, float* y1, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = 0.f, b = 0.f;
    switch (fn) {
    case 0: dm::sincosf_det(x[i], a, b); break;
    case 1: dm::sincos2pif_det(x[i], a, b); break;
    case 2: a = dm::logf_det(x[i]); break;
    case 4: a = dm::atanf_det(x[i]); break;
    default: a = dm::expf_det(x[i]); break;
    }
    y0[i] = a;
    if (y1) y1[i] = b;
}

// product normalize3 / fdiv / fsqrt against the IEEE intrinsics (only meaningful in the STRICT flavour)
__global__ void mppi_normalize_test_kernel(const float* v, float* out, float* ref, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 a = make_float3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    const float3 r = normalize3_maybe_unit(a);
    out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    const float s = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)), __fmul_rn(a.z, a.z)));
    ref[3 * i] = __fdiv_rn(a.x, s); ref[3 * i + 1] = __fdiv_rn(a.y, s); ref[3 * i + 2] = __fdiv_rn(a.z, s);
}

// scalar fdiv(a, b) and fsqrt(|a|) against the IEEE intrinsics
__global__ void mppi_divsqrt_test_kernel(const float* a, const float* b, float* out, float* ref, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[2 * i] = fdiv(a[i], b[i]);
    ref[2 * i] = __fdiv_rn(a[i], b[i]);
    out[2 * i + 1] = fsqrt(fabsf(a[i]));
    ref[2 * i + 1] = __fsqrt_rn(fabsf(a[i]));
}

__global__ void mppi_noise_kernel(uint64_t seed, uint64_t offset, uint32_t rover, uint32_t k_begin, int K, int T,
                                  float* e1, float* e2)
{
    const int P = (T + 1) >> 1;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)K * P) return;
    const int k = (int)(idx / P), pr = (int)(idx - (size_t)k * P);
    const NoiseKey nk = make_noise_key(seed, offset, rover);
    float a0, a1, b0, b1;
    noise_pair(nk, k_begin + (uint32_t)k, (uint32_t)pr, a0, a1, b0, b1);
    const int t = 2 * pr;
    e1[(size_t)k * T + t] = a0; e2[(size_t)k * T + t] = b0;
    if (t + 1 < T) { e1[(size_t)k * T + t + 1] = a1; e2[(size_t)k * T + t + 1] = b1; }
}

// ------------------------------------------------------------------ launchers
size_t fused_smem_bytes(int T, int block, int nblocks) { return smem_floats(T, block, nblocks) * sizeof(float); }

// Raises the dynamic shared-memory limit of a kernel (to the 227 KB a CTA may have) once per (device, kernel), not on
// every launch: cudaFuncSetAttribute applies to the current device only, and one process may hold handles on several.
template <typename Kern>
static cudaError_t ensure_smem(Kern k, size_t bytes)
{
    if (bytes <= 48 * 1024) return cudaSuccess;
    static std::mutex mu;
    static std::set<std::pair<int, const void*>> raised;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    const std::pair<int, const void*> key(dev, reinterpret_cast<const void*>(k));
    if (raised.count(key)) return cudaSuccess;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) raised.insert(key);
    return e;
}

cudaError_t launch_fused(const FusedArgs& a, int proj, int n_rovers, int block, cudaStream_t s)
{
    const dim3 grid(a.nblocks, n_rovers);
    const size_t smem = fused_smem_bytes(a.p.T, block, list_cap(a));
    // partial occupancy (one rover, at most kLowOccMaxSamples samples): the software-pipelined instantiation
    const bool no_lowocc = getenv("MPPI_NO_LOWOCC") != nullptr;      // A/B knob (read at every launch: tests toggle it)
    const bool lowocc = !no_lowocc && n_rovers == 1 && (long long)a.nblocks * block <= kLowOccMaxSamples;
    cudaError_t e;
#define MPPI_LAUNCH_FUSED(PROJ, INJ)                                              \
    do {                                                                          \
        if (lowocc) {                                                             \
            e = ensure_smem(mppi_fused_kernel<PROJ, INJ, true>, smem);            \
            if (e != cudaSuccess) return e;                                       \
            mppi_fused_kernel<PROJ, INJ, true><<<grid, block, smem, s>>>(a);      \
        } else {                                                                  \
            e = ensure_smem(mppi_fused_kernel<PROJ, INJ, false>, smem);           \
            if (e != cudaSuccess) return e;                                       \
            mppi_fused_kernel<PROJ, INJ, false><<<grid, block, smem, s>>>(a);     \
        }                                                                         \
    } while (0)
    if (proj == MPPI_PROJ_3D) {
        if (a.noise) MPPI_LAUNCH_FUSED(MPPI_PROJ_3D, true); else MPPI_LAUNCH_FUSED(MPPI_PROJ_3D, false);
    } else {
        if (a.noise) MPPI_LAUNCH_FUSED(MPPI_PROJ_2D, true); else MPPI_LAUNCH_FUSED(MPPI_PROJ_2D, false);
    }
#undef MPPI_LAUNCH_FUSED
    return cudaGetLastError();
}

size_t pipe_smem_bytes_no_tile(int T, int /*nblocks*/)
{
    return pipe_smem_offset_floats(T) * sizeof(float) + ((sizeof(PipeSmem) + 127) & ~(size_t)127);
}

cudaError_t launch_fused_pipe(const FusedArgs& a, int proj, int n_rovers, cudaStream_t s)
{
    if (a.ll_seq == 0u || a.ll == nullptr || a.minkey == nullptr) return cudaErrorInvalidValue;
    const dim3 grid(a.nblocks + 1, n_rovers);             // + the updater block (LL protocol)
    size_t smem = pipe_smem_bytes_no_tile(a.p.T, list_cap(a)) + (size_t)a.tile.w * a.tile.h * sizeof(float);
    // u history ([T][2][32] floats) when it fits beside the tile: the A rows of the update become shared-memory reads
    FusedArgs b = a;
    static const bool no_uhist = getenv("MPPI_NO_UHIST") != nullptr;          // A/B knob
    const size_t hist = (size_t)a.p.T * 64 * sizeof(float);
    b.uhist = (!no_uhist && smem + hist <= (size_t)227 * 1024) ? 1 : 0;
    if (b.uhist) smem += hist;
    // the updater block's overlay (see pipe_smem_offset_floats) must fit too
    const size_t upd = (pipe_smem_offset_floats(a.p.T) + pipe_updater_floats(list_cap(a))) * sizeof(float);
    if (upd > smem) smem = upd;
    if (smem > (size_t)227 * 1024) return cudaErrorInvalidValue;
    // Programmatic stream serialisation: this launch may be SCHEDULED while the previous pipelined launch of the stream
    // is still in its update (see griddepcontrol in the kernel); after any other kind of work it starts as usual.
    // Off for the sharded step (its completion order against the peers' launches is what tests/multi_gpu_check.py
    // and the N = 8 timeline were measured with) and with MPPI_NO_PDL set (A/B knob).
    static const bool no_pdl = getenv("MPPI_NO_PDL") != nullptr;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = dim3(kPipeThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cfg.attrs = attr; cfg.numAttrs = (no_pdl || a.peers.world > 0) ? 0 : 1;
    cudaError_t e;
#define MPPI_LAUNCH_PIPE(PROJ, INJ)                                                    \
    do {                                                                               \
        e = ensure_smem(mppi_fused_pipe_kernel<PROJ, INJ>, smem);                      \
        if (e != cudaSuccess) return e;                                                \
        e = cudaLaunchKernelEx(&cfg, mppi_fused_pipe_kernel<PROJ, INJ>, b);            \
        if (e != cudaSuccess) return e;                                                \
    } while (0)
    if (proj == MPPI_PROJ_3D) {
        if (a.noise) MPPI_LAUNCH_PIPE(MPPI_PROJ_3D, true); else MPPI_LAUNCH_PIPE(MPPI_PROJ_3D, false);
    } else {
        if (a.noise) MPPI_LAUNCH_PIPE(MPPI_PROJ_2D, true); else MPPI_LAUNCH_PIPE(MPPI_PROJ_2D, false);
    }
#undef MPPI_LAUNCH_PIPE
    return cudaGetLastError();
}

cudaError_t launch_combine(const CombineArgs& a, cudaStream_t s)
{
    const int block = 256;
    const size_t smem = fused_smem_bytes(a.p.T, block, a.n_parts);
    cudaError_t e = ensure_smem(mppi_combine_kernel, smem);
    if (e != cudaSuccess) return e;
    mppi_combine_kernel<<<1, block, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_dump(const DumpArgs& a, int proj, cudaStream_t s)
{
    const int block = 128, grid = (a.p.K + block - 1) / block;
    if (proj == MPPI_PROJ_3D) {
        if (a.noise) mppi_dump_kernel<MPPI_PROJ_3D, true><<<grid, block, 0, s>>>(a);
        else mppi_dump_kernel<MPPI_PROJ_3D, false><<<grid, block, 0, s>>>(a);
    } else {
        if (a.noise) mppi_dump_kernel<MPPI_PROJ_2D, true><<<grid, block, 0, s>>>(a);
        else mppi_dump_kernel<MPPI_PROJ_2D, false><<<grid, block, 0, s>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_export(const ExportArgs& a, int proj, cudaStream_t s)
{
    const int ne = (a.p.K + a.k_stride - 1) / a.k_stride;
    const int block = 128, grid = (ne + block - 1) / block;
    if (proj == MPPI_PROJ_3D) {
        if (a.noise) mppi_export_kernel<MPPI_PROJ_3D, true><<<grid, block, 0, s>>>(a);
        else mppi_export_kernel<MPPI_PROJ_3D, false><<<grid, block, 0, s>>>(a);
    } else {
        if (a.noise) mppi_export_kernel<MPPI_PROJ_2D, true><<<grid, block, 0, s>>>(a);
        else mppi_export_kernel<MPPI_PROJ_2D, false><<<grid, block, 0, s>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_weights(const float* costs, int K, float lambda, float* weights, cudaStream_t s)
{
    mppi_weights_kernel<<<1, 1024, 0, s>>>(costs, K, lambda, weights);
    return cudaGetLastError();
}

cudaError_t launch_sim(const SimArgs& a, cudaStream_t s)
{
    mppi_sim_kernel<<<1, 32, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_detmath(int fn, const float* x, float* y0, float* y1, int n, cudaStream_t s)
{
    mppi_detmath_kernel<<<(n + 255) / 256, 256, 0, s>>>(fn, x, y0, y1, n);
    return cudaGetLastError();
}

cudaError_t launch_normalize_test(const float* v, float* out, float* ref, int n, cudaStream_t s)
{
    mppi_normalize_test_kernel<<<(n + 255) / 256, 256, 0, s>>>(v, out, ref, n);
    return cudaGetLastError();
}

cudaError_t launch_divsqrt_test(const float* a, const float* b, float* out, float* ref, int n, cudaStream_t s)
{
    mppi_divsqrt_test_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, b, out, ref, n);
    return cudaGetLastError();
}

cudaError_t launch_noise(uint64_t seed, uint64_t offset, uint32_t rover, uint32_t k_begin, int K, int T,
                         float* e1, float* e2, cudaStream_t s)
{
    const size_t n = (size_t)K * ((T + 1) / 2);
    mppi_noise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(seed, offset, rover, k_begin, K, T, e1, e2);
    return cudaGetLastError();
}

}  // namespace MPPI_NS
}  // namespace mppi
