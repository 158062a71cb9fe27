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

// legacy tensor path microbenchmark: mma.sync.m16n8k8 tf32 (fp32 accumulate), 8 independent accumulator tiles per warp.
// Answers one design question: would a 3xTF32 error-compensated mma.sync dW phase beat the FFMA pipe?
__global__ void __launch_bounds__(256) nj_mma_tf32_peak_kernel(float* out, int iters) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = 0x3f800000u + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 0x3f000000u + threadIdx.x, b1 = b0 + 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// *macs = multiply-accumulates issued (16 x 8 x 8 per warp-level mma)
extern "C" int njode_mma_tf32_peak_launch(float* scratch, int iters, double* macs, void* stream) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -2;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256;
    nj_mma_tf32_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(scratch, iters);
    if (macs) *macs = (double)blocks * (threads / 32) * (double)iters * 8.0 * 1024.0;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
