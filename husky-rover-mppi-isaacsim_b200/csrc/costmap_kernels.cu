// costmap_kernels.cu -- GPU obstacle-costmap builder, sm_100a.
//
// Replaces Surface.create_obstacles_costmap (thesis_master/warp_implementation/MPPI_isaac.py:361-378), which the
// Isaac driver calls at start-up and again at every terrain-block change (visual_terrain_stack_full_terrain.py:449,
// 561-563) and which stalls the control loop there: a full-grid NumPy mask per rock (seconds for 750 rocks on an
// 875^2 map), cv2.distanceTransform(DIST_L2, 5), cv2.normalize(NORM_MINMAX), (1 - d)^20, then an H2D copy.
//
//   k1  costmap_rasterize_kernel   one block per rock stamps the rock's bounding box; the disc test is the
//                                   reference's float64 expression on the numpy.linspace grid -> identical mask.
//   k2  costmap_chamfer_kernel     Borgefors' two-pass 5x5 chamfer transform (weights 1, 1.4, 2.1969, float32 path
//                                   sums: what cv2's DIST_L2 / 5 computes).  The raster scan is sequential in
//                                   (row, column), but cells with equal j + 3 i are independent, so ONE block walks
//                                   the wavefronts with one thread per row: the left neighbour stays in a register,
//                                   the two rows above are read from 16-deep shared-memory rings written by the
//                                   neighbouring threads.  Same candidates, same float operations as the sequential
//                                   scan -> bit-identical to the CPU restatement used by the tests.  One barrier per wavefront: 2 x (cols + 3 rows).
//   k3  costmap_finish_kernel      min / max come out of k2; d * scale + shift (cv2.normalize), (1 - d)^p.
//
// Compiled with the STRICT flags (no FMA contraction: NumPy does not contract either).
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace mppi {
namespace costmap {

constexpr int kMaxSize = 1024;      // one thread per row
constexpr int kRing = 16;

// WAVEFRONT-MAJOR ("skewed") layout of the mask and of the two distance buffers: cell (row i, column j) lives at
// [j + 3 i][i], i.e. all cells of one wavefront of the chamfer scan are contiguous.  The chamfer kernel has one thread
// per ROW; in the natural row-major layout every warp-wide access of a wavefront touched 32 different 128-byte lines
// (~290 active rows x 2 lines per wavefront through one SM's L1: that, not the barrier or the arithmetic, was what a
// wavefront cost).  Skewed, a warp reads 32 consecutive bytes of mask and writes 128 consecutive bytes of distance.
__host__ __device__ inline int skew_pitch(int rows) { return (rows + 31) & ~31; }
__host__ __device__ inline size_t skew_cells(int rows, int cols) { return (size_t)(cols + 3 * (rows - 1)) * skew_pitch(rows); }
__host__ __device__ inline size_t skew_at(int i, int j, int pitch) { return (size_t)(j + 3 * i) * pitch + i; }

__global__ void costmap_fill_kernel(uint8_t* mask, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mask[i] = 255;
}

// numpy.linspace(-hw, hw, n)[i]
__device__ __forceinline__ double grid_coord(int i, int n, double hw, double step)
{
    return (i == n - 1) ? hw : (double)i * step + (-hw);
}

// obstacles: [n][3] (x_global, y_global, r_obs) float64.  MPPI_isaac.py:365-372.
__global__ void costmap_rasterize_kernel(const double* __restrict__ obstacles, int n_obs, double x0, double y0, int cms,
                                         double hw, double r_robot, double radius_scale, double inflate, uint8_t* mask)
{
    const int o = blockIdx.x;
    if (o >= n_obs) return;
    const double step = (hw - (-hw)) / (double)(cms - 1);
    const double xl = obstacles[3 * o + 1] - y0;          // x_local = y_global - y0
    const double yl = obstacles[3 * o] - x0;              // y_local = x_global - x0
    const double R = obstacles[3 * o + 2] * radius_scale + r_robot + inflate;
    const double R2 = R * R;
    // conservative bounding box in cells (exact test below)
    int c0 = (int)floor((xl - R + hw) / step) - 1, c1 = (int)ceil((xl + R + hw) / step) + 1;
    int r0 = (int)floor((yl - R + hw) / step) - 1, r1 = (int)ceil((yl + R + hw) / step) + 1;
    c0 = max(c0, 0); r0 = max(r0, 0); c1 = min(c1, cms - 1); r1 = min(r1, cms - 1);
    if (c1 < c0 || r1 < r0) return;
    const int w = c1 - c0 + 1, total = w * (r1 - r0 + 1);
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int r = r0 + e / w, c = c0 + e % w;
        const double dx = grid_coord(c, cms, hw, step) - xl, dy = grid_coord(r, cms, hw, step) - yl;
        if (dx * dx + dy * dy <= R2) mask[skew_at(r, c, skew_pitch(cms))] = 0;
    }
}

// One block, blockDim.x >= rows.  mask, tmp (scratch), dist (out): skewed layout; minmax: [2] out.
__global__ void __launch_bounds__(kMaxSize) costmap_chamfer_kernel(const uint8_t* __restrict__ mask, int rows, int cols,
                                                                   float* tmp, float* dist, float* minmax)
{
    extern __shared__ float ring[];                       // [rows + 4][kRing], two guard rows at either end
    const float HV = 1.0f, DG = 1.4f, LG = 2.1969f, INIT = FLT_MAX;
    const int i = threadIdx.x;
    const bool active = i < rows;
    const int pitch = skew_pitch(rows);
    for (int e = threadIdx.x; e < (rows + 4) * kRing; e += blockDim.x) ring[e] = INIT;
    __syncthreads();
    float* mine = ring + (size_t)(i + 2) * kRing;
    const float* up1 = mine - kRing;
    const float* up2 = mine - 2 * kRing;

    // ---------------- forward pass: neighbours above and to the left
    {
        float left = INIT;
        const int nw = cols + 3 * (rows - 1);
        // Cost of a wavefront, measured in round 2 on one B200 (875^2, both passes = 6994 wavefronts; every variant
        // bit-identical to the oracle): round 1 (row-major buffers, two barriers per wavefront) 2.90 ms; one barrier
        // 2.83 ms; wavefront-major buffers 3.29 ms (coalesced, but every wavefront is a fresh line: an L2 round trip on the
        // dependent chain) and 2.26 ms once the next wavefront's input is requested one wavefront ahead (this version);
        // letting the ~18 warps without a cell in the current wavefront go straight to the barrier: no gain (2.56 ms);
        // a warp-synchronous variant without barriers or rings (32 rows per warp, neighbours by shuffle out of register
        // histories, flag-in-data hand-over between warps): 5.2 ms.  ~630 cycles per wavefront remain against a dependency
        // depth of ~8 (FADD + FMNMX): open.
        // the mask byte of the NEXT wavefront is requested one wavefront ahead: in the skewed layout every wavefront is a
        // fresh line (an L2 round trip), which must not sit on the dependent chain of the step
        uint8_t m_next = (active && i == 0 && cols > 0) ? mask[skew_at(0, 0, pitch)] : (uint8_t)255;
        for (int w = 0; w < nw; ++w) {
            const int j = w - 3 * i;
            float t = INIT;
            const bool on = active && j >= 0 && j < cols;
            const uint8_t m_cur = m_next;
            {
                const int jn = j + 1;
                if (active && jn >= 0 && jn < cols) m_next = mask[skew_at(i, jn, pitch)];
            }
            if (on) {
                if (m_cur == 0) {
                    t = 0.0f;
                } else {
                    // rows above hold INIT outside the image: guard rows and the ring slots of columns < 0 / >= cols
                    const float a_m1 = (j >= 1) ? up2[(j - 1) & (kRing - 1)] : INIT;
                    const float a_p1 = (j + 1 < cols) ? up2[(j + 1) & (kRing - 1)] : INIT;
                    const float b_m2 = (j >= 2) ? up1[(j - 2) & (kRing - 1)] : INIT;
                    const float b_m1 = (j >= 1) ? up1[(j - 1) & (kRing - 1)] : INIT;
                    const float b_0 = up1[j & (kRing - 1)];
                    const float b_p1 = (j + 1 < cols) ? up1[(j + 1) & (kRing - 1)] : INIT;
                    const float b_p2 = (j + 2 < cols) ? up1[(j + 2) & (kRing - 1)] : INIT;
                    t = a_m1 + LG;
                    t = fminf(t, a_p1 + LG);
                    t = fminf(t, b_m2 + LG);
                    t = fminf(t, b_m1 + DG);
                    t = fminf(t, b_0 + HV);
                    t = fminf(t, b_p1 + DG);
                    t = fminf(t, b_p2 + LG);
                    t = fminf(t, left + HV);
                }
                tmp[skew_at(i, j, pitch)] = t;
                left = t;
                // No barrier between this wavefront's reads and the store: in wavefront w row i writes the slot of column
                // j = w - 3 i while row i + 1 reads this row's columns j - 5 .. j - 1 and row i + 2 its columns j - 7 and
                // j - 5 -- never slot j mod 16, and the column it overwrites (j - 16) was last read nine wavefronts ago.
                mine[j & (kRing - 1)] = t;
            }
            __syncthreads();                 // ONE barrier per wavefront: its stores are visible to the next one's reads
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < (rows + 4) * kRing; e += blockDim.x) ring[e] = INIT;
    __syncthreads();

    // ---------------- backward pass: neighbours below and to the right (mirror image of the forward pass)
    float lo = FLT_MAX, hi = -FLT_MAX;
    {
        const float* dn1 = mine + kRing;
        const float* dn2 = mine + 2 * kRing;
        float right = INIT;
        const int nw = cols + 3 * (rows - 1);
        const int ii = rows - 1 - i;         // mirrored row index
        float t_next = (active && ii == 0 && cols > 0) ? tmp[skew_at(i, cols - 1, pitch)] : INIT;
        for (int w = 0; w < nw; ++w) {
            const int jj = w - 3 * ii;       // mirrored column index
            const int j = cols - 1 - jj;
            float t = INIT;
            const bool on = active && jj >= 0 && jj < cols;
            const float t_cur = t_next;
            {
                const int jjn = jj + 1;      // the forward result of the next wavefront's cell, one wavefront ahead
                if (active && jjn >= 0 && jjn < cols) t_next = tmp[skew_at(i, cols - 1 - jjn, pitch)];
            }
            if (on) {
                t = t_cur;
                if (t > HV) {
                    const float a_p1 = (j + 1 < cols) ? dn2[(j + 1) & (kRing - 1)] : INIT;
                    const float a_m1 = (j >= 1) ? dn2[(j - 1) & (kRing - 1)] : INIT;
                    const float b_p2 = (j + 2 < cols) ? dn1[(j + 2) & (kRing - 1)] : INIT;
                    const float b_p1 = (j + 1 < cols) ? dn1[(j + 1) & (kRing - 1)] : INIT;
                    const float b_0 = dn1[j & (kRing - 1)];
                    const float b_m1 = (j >= 1) ? dn1[(j - 1) & (kRing - 1)] : INIT;
                    const float b_m2 = (j >= 2) ? dn1[(j - 2) & (kRing - 1)] : INIT;
                    t = fminf(t, a_p1 + LG);
                    t = fminf(t, a_m1 + LG);
                    t = fminf(t, b_p2 + LG);
                    t = fminf(t, b_p1 + DG);
                    t = fminf(t, b_0 + HV);
                    t = fminf(t, b_m1 + DG);
                    t = fminf(t, b_m2 + LG);
                    t = fminf(t, right + HV);
                }
                dist[skew_at(i, j, pitch)] = t;
                right = t;
                lo = fminf(lo, t); hi = fmaxf(hi, t);
                mine[j & (kRing - 1)] = t;       // mirrored: the readers are at columns j + 1 .. j + 5 and j + 5, j + 7
            }
            __syncthreads();
        }
    }
    // ---------------- min / max of the distance map (cv2.normalize NORM_MINMAX needs both)
    __shared__ float red_lo[32], red_hi[32];
    for (int off = 16; off > 0; off >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if ((threadIdx.x & 31) == 0) { red_lo[threadIdx.x >> 5] = lo; red_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)((blockDim.x + 31) >> 5); ++k) { lo = fminf(lo, red_lo[k]); hi = fmaxf(hi, red_hi[k]); }
        minmax[0] = lo; minmax[1] = hi;
    }
}

// cv2.normalize(d, None, 0, 1, NORM_MINMAX): scale = 1 / (max - min) (double; 0 when max == min), shift = -min scale,
// d' = float(d * (float)scale + (float)shift); then (1 - d')^power (MPPI_isaac.py:375-376).  The power is evaluated in
// double and rounded once.
// dist: skewed; costmap (and the optional row-major copies of the distance map / mask): [cms][cms].
__global__ void costmap_finish_kernel(const float* __restrict__ dist, const uint8_t* __restrict__ mask,
                                      const float* __restrict__ minmax, int cms, double power, float* costmap,
                                      float* dist_rowmajor, uint8_t* mask_rowmajor)
{
    const size_t n = (size_t)cms * cms;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int r = (int)(e / cms), c = (int)(e - (size_t)r * cms);
    const size_t sk = skew_at(r, c, skew_pitch(cms));
    const double smin = (double)minmax[0], smax = (double)minmax[1];
    const double scale = (smax - smin > DBL_EPSILON) ? 1.0 / (smax - smin) : 0.0;
    const float a = (float)scale, b = (float)(0.0 - smin * scale);
    const float d = dist[sk];
    const float dn = d * a + b;
    costmap[e] = (float)pow((double)(1.0f - dn), power);
    if (dist_rowmajor != nullptr) dist_rowmajor[e] = d;
    if (mask_rowmajor != nullptr) mask_rowmajor[e] = mask[sk];
}

// mask / tmp / dist: workspaces of skew_cells(cms, cms) elements (skewed layout).  dist_out / mask_out: optional
// row-major [cms][cms] copies of the distance map and of the rasterised mask.
cudaError_t build(const double* obstacles_dev, int n_obs, double x0, double y0, int cms, double hw, double r_robot,
                  double radius_scale, double inflate, double power, uint8_t* mask, float* tmp, float* dist, float* minmax,
                  float* costmap, float* dist_out, uint8_t* mask_out, cudaStream_t s)
{
    const size_t n = (size_t)cms * cms, nsk = skew_cells(cms, cms);
    costmap_fill_kernel<<<(unsigned)((nsk + 255) / 256), 256, 0, s>>>(mask, nsk);
    if (n_obs > 0)
        costmap_rasterize_kernel<<<n_obs, 128, 0, s>>>(obstacles_dev, n_obs, x0, y0, cms, hw, r_robot, radius_scale,
                                                        inflate, mask);
    const int threads = ((cms + 31) / 32) * 32;
    const size_t smem = (size_t)(cms + 4) * kRing * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(costmap_chamfer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    costmap_chamfer_kernel<<<1, threads, smem, s>>>(mask, cms, cms, tmp, dist, minmax);
    costmap_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dist, mask, minmax, cms, power, costmap, dist_out,
                                                                       mask_out);
    return cudaGetLastError();
}

size_t workspace_cells(int cms) { return skew_cells(cms, cms); }

}  // namespace costmap
}  // namespace mppi
