// microbench.cu -- measured peaks of the SM resources that bound the spread / interp kernels
// ("shared-memory / FP32 / FP64 peaks must be MEASURED with a micro-benchmark").
// bench.py calls cufinufft_b200_microbench() once per run and uses the numbers as the denominators of
// `roofline.binding`.  Each probe is a small persistent kernel (one block of 1024 threads per SM x 2)
// timed with CUDA events; rates are per whole GPU.
//   0  shared-memory read bandwidth, conflict-free LDS.128          -> bytes / s
//   1  packed FP32 FMA (FFMA2: fma.rn.f32x2), 8 independent chains  -> FMA / s (2 per lane and instruction)
//   2  scalar FP32 FMA (FFMA)                                       -> FMA / s
//   3  FP64 FMA (DFMA)                                              -> FMA / s
//   4  shared-memory atomicAdd(float), conflict-free spread (compiles to a CAS loop: ATOMS.CAST.SPIN,
//      there is no native shared-memory float add) -> atomics / s
//   5  shared-memory wavefronts through LDS.64 (half-warp granules)  -> bytes / s
#include <cuda_runtime.h>
#include "../../include/cufinufft_b200.h"

namespace {

constexpr int MB_THREADS = 1024;

// the loads are volatile asm: the buffer never changes, an ordinary load would be hoisted out of the loop
__global__ void __launch_bounds__(MB_THREADS) mb_lds128(int iters, float *sink)
{
    __shared__ float4 buf[MB_THREADS];
    buf[threadIdx.x] = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
    __syncthreads();
    float acc = 0;
    unsigned base = (unsigned)__cvta_generic_to_shared(buf);
    unsigned idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float x, y, z, w;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w)
                         : "r"(base + (((idx + u * 32) & (MB_THREADS - 1)) << 4)));
            acc += (x + y) + (z + w);            // all four words used: ptxas must keep the 128-bit load
        }
        idx += 256;
    }
    if (acc == 12345.678f) sink[0] = acc;
}

__global__ void __launch_bounds__(MB_THREADS) mb_lds64(int iters, float *sink)
{
    __shared__ float2 buf[MB_THREADS];
    buf[threadIdx.x] = make_float2(threadIdx.x, 1.f);
    __syncthreads();
    float acc = 0;
    unsigned base = (unsigned)__cvta_generic_to_shared(buf);
    unsigned idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float x, y;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(base + (((idx + u * 32) & (MB_THREADS - 1)) << 3)));
            acc += x + y;
        }
        idx += 256;
    }
    if (acc == 12345.678f) sink[0] = acc;
}

__global__ void __launch_bounds__(MB_THREADS) mb_ffma2(int iters, float *sink)
{
    float2 a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = make_float2(threadIdx.x * 1e-3f + j, j * 0.5f);
    const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __ffma2_rn(a[j], m, c);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j].x + a[j].y;
    if (s == 12345.678f) sink[0] = s;
}

__global__ void __launch_bounds__(MB_THREADS) mb_ffma(int iters, float *sink)
{
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3f + j;
    const float m = 1.0000001f, c = 1e-7f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], m, c);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 12345.678f) sink[0] = s;
}

__global__ void __launch_bounds__(MB_THREADS) mb_dfma(int iters, float *sink)
{
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3 + j;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fma(a[j], m, c);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 12345.678) sink[0] = (float)s;
}

__global__ void __launch_bounds__(MB_THREADS) mb_atoms(int iters, float *sink)
{
    __shared__ float buf[MB_THREADS];
    buf[threadIdx.x] = 0.0f;
    __syncthreads();
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) atomicAdd(&buf[(idx + u * 32) & (MB_THREADS - 1)], 1.0f);
        idx = (idx + 256) & (MB_THREADS - 1);
    }
    __syncthreads();
    if (buf[threadIdx.x] == 12345.678f) sink[0] = 1.0f;
}

template <typename K>
int time_probe(K kernel, int blocks, int iters, double work_per_thread_iter, double *out)
{
    float *sink = nullptr;
    if (cudaMalloc(&sink, sizeof(float)) != cudaSuccess) return CFB_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kernel<<<blocks, MB_THREADS>>>(iters / 8 + 1, sink);          // warm-up (clocks, instruction cache)
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kernel<<<blocks, MB_THREADS>>>(iters, sink);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(sink); return CFB_ERR_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double rate = work_per_thread_iter * (double)iters * MB_THREADS * blocks / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *out = best;
    return cudaGetLastError() == cudaSuccess ? 0 : CFB_ERR_CUDA;
}

}  // namespace

extern "C" int cufinufft_b200_microbench(int what, int device, double *out)
{
    if (!out) return CFB_ERR_BAD_ARG;
    int ndev = 0, prev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return CFB_ERR_CUDA;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int blocks = prop.multiProcessorCount * 2;
    int ier = CFB_ERR_BAD_ARG;
    switch (what) {
        case 0: ier = time_probe(mb_lds128, blocks, 4000, 8 * 16.0, out); break;
        case 1: ier = time_probe(mb_ffma2, blocks, 4000, 4 * 8 * 2.0, out); break;
        case 2: ier = time_probe(mb_ffma, blocks, 4000, 4 * 8 * 1.0, out); break;
        case 3: ier = time_probe(mb_dfma, blocks, 2000, 4 * 8 * 1.0, out); break;
        case 4: ier = time_probe(mb_atoms, blocks, 500, 8 * 1.0, out); break;
        case 5: ier = time_probe(mb_lds64, blocks, 4000, 8 * 8.0, out); break;
        default: break;
    }
    cudaSetDevice(prev);
    return ier;
}
