// peaks.cu -- two roofline denominators the driver's MEASURED_PEAKS.json does not carry, measured on the device the
// bench runs on (bench.py reports them beside its numbers; tools/measure_peaks.py writes them to profiles/):
//   * FP32 FMA throughput: 148 SMs x 4 sub-partitions issuing independent FFMA chains (the denominator of
//     roofline.fp32; the spec-sheet product 148 x 128 x 2 x clock is what it replaces);
//   * L2 gather rate: random 4-byte gathers from a window that is resident in L2 but far larger than L1
//     (the denominator for the un-staged terrain lookups of the throughput kernel: 32-byte sectors per second).
// Measurement aids only; nothing on the MPPI path calls them.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mppi_b200.h"

namespace {

constexpr int kFmaChains = 8;       // independent accumulators per thread: covers the 4-cycle FFMA latency twice over
constexpr int kFmaIters = 4096;

__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, float a, float b)
{
    float acc[kFmaChains];
#pragma unroll
    for (int i = 0; i < kFmaChains; ++i) acc[i] = (float)(threadIdx.x + i);
#pragma unroll 4
    for (int it = 0; it < kFmaIters; ++it) {
#pragma unroll
        for (int i = 0; i < kFmaChains; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kFmaChains; ++i) s += acc[i];
    if (s == 12345.678f) out[0] = s;                  // keeps the chains alive; never true for the inputs used
}

// Every thread walks its own pseudo-random sequence of word indices inside [0, words): one independent 4-byte gather
// per step and per lane, i.e. 32 distinct sectors per warp-instruction -- the access pattern of late-horizon rollouts.
__global__ void __launch_bounds__(256) l2_gather_kernel(const float* __restrict__ buf, uint32_t words, int iters, float* out)
{
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int it = 0; it < iters; it += 4) {
        uint32_t i0, i1, i2, i3;
        x = x * 1664525u + 1013904223u; i0 = (uint32_t)(((uint64_t)x * words) >> 32);
        x = x * 1664525u + 1013904223u; i1 = (uint32_t)(((uint64_t)x * words) >> 32);
        x = x * 1664525u + 1013904223u; i2 = (uint32_t)(((uint64_t)x * words) >> 32);
        x = x * 1664525u + 1013904223u; i3 = (uint32_t)(((uint64_t)x * words) >> 32);
        s0 += __ldcg(buf + i0); s1 += __ldcg(buf + i1); s2 += __ldcg(buf + i2); s3 += __ldcg(buf + i3);   // .cg: L2 only
    }
    const float s = s0 + s1 + s2 + s3;
    if (s == 12345.678f) out[0] = s;
}

float time_ms(cudaEvent_t e0, cudaEvent_t e1)
{
    float ms = 0.f;
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

}  // namespace

// One thread stores %globaltimer: a timestamp in the SAME clock as the kernel-internal stamps of mppi_set_trace, for
// splitting the time between a launch's neighbours in the stream and its first / last instruction (tools/launch_gap.py).
__global__ void timestamp_kernel(unsigned long long* out)
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *out = t;
}

extern "C" int mppi_test_timestamp(uint64_t* out_dev, void* stream)
{
    if (!out_dev) return MPPI_ERR_INVALID_ARG;
    timestamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(out_dev));
    return cudaGetLastError() == cudaSuccess ? MPPI_OK : MPPI_ERR_CUDA;
}

extern "C" int mppi_measure_peaks(int32_t device, uint64_t l2_window_bytes, float* fp32_tflops, float* l2_gather_gsectors,
                                  float* l2_gather_gbs)
{
    if (!fp32_tflops || !l2_gather_gsectors || !l2_gather_gbs || l2_window_bytes < (1u << 20)) return MPPI_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return MPPI_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MPPI_ERR_CUDA;
    const int sms = prop.multiProcessorCount;
    float* out = nullptr;
    float* buf = nullptr;
    cudaEvent_t e0, e1;
    if (cudaMalloc(&out, 256) != cudaSuccess) return MPPI_ERR_ALLOC;
    if (cudaMalloc(&buf, l2_window_bytes) != cudaSuccess) { cudaFree(out); return MPPI_ERR_ALLOC; }
    cudaMemset(buf, 0, l2_window_bytes);
    cudaEventCreate(&e0); cudaEventCreate(&e1);

    // ---- FP32: 8 resident blocks of 256 threads per SM, best of 5
    const int fgrid = sms * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        fma_peak_kernel<<<fgrid, 256>>>(out, 1.000001f, 1e-7f);
        cudaEventRecord(e1);
        const float ms = time_ms(e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    *fp32_tflops = (float)(2.0 * kFmaChains * (double)kFmaIters * 256.0 * fgrid / (best * 1e-3) / 1e12);

    // ---- L2 gathers: window resident in L2 (touched once by the first repetition), best of 5
    const uint32_t words = (uint32_t)(l2_window_bytes / 4);
    const int ggrid = sms * 8, iters = 2048;
    best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        l2_gather_kernel<<<ggrid, 256>>>(buf, words, iters, out);
        cudaEventRecord(e1);
        const float ms = time_ms(e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double gathers = (double)ggrid * 256.0 * iters;
    *l2_gather_gsectors = (float)(gathers / (best * 1e-3) / 1e9);
    *l2_gather_gbs = (float)(gathers * 32.0 / (best * 1e-3) / 1e9);        // one 32-byte sector per gather

    const cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf); cudaFree(out);
    return e == cudaSuccess ? MPPI_OK : MPPI_ERR_CUDA;
}
