// njode_diag.cu -- measurement helpers: fp32 FMA-pipe peak microbenchmark (the roofline denominator
// of the narrow NJ-ODE configurations) and an L2 flush.
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void __launch_bounds__(256) nj_fma_peak_kernel(float* out, int iters) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1e-3f * (float)(threadIdx.x + i);
    const float b = 1.0000001f, c = 1e-9f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chain alive
}

// launches the FMA chain on every SM; returns the number of FMAs issued (lane-level) in *fmas
extern "C" int njode_fma_peak_launch(float* scratch, int iters, double* fmas, void* stream) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -2;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256;
    nj_fma_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(scratch, iters);
    if (fmas) *fmas = (double)blocks * threads * (double)iters * 64.0;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

__global__ void nj_l2_flush_kernel(float4* buf, size_t n, float v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        buf[i] = make_float4(v, v, v, v);
}

// overwrites `bytes` of `buf` (caller sizes it larger than L2) so the next kernel starts cold
extern "C" int njode_l2_flush(void* buf, int64_t bytes, void* stream) {
    nj_l2_flush_kernel<<<1184, 256, 0, (cudaStream_t)stream>>>((float4*)buf, (size_t)bytes / 16, 1.f);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
