// njode_api_seg.cu -- the segment kernels (njode_seg.cuh) with their launch function; a translation unit of its own so that
// it compiles concurrently with njode_api.cu (C ABI entry points, generic kernels) and njode_api_path.cu
#include <cuda_runtime.h>
#include "njode_plan.h"

extern __shared__ __align__(16) float nj_smem[];

__global__ void __launch_bounds__(384) nj_seg_fwd_kernel(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjSeg seg,
                                                         const __grid_constant__ NjArgs args) {
    nj_seg_cta_forward(cfg, seg, args, nj_smem);
}

__global__ void __launch_bounds__(384) nj_seg_bwd_kernel(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjSeg seg,
                                                         const __grid_constant__ NjArgs args) {
    nj_seg_cta_backward<false>(cfg, seg, args, nj_smem, blockIdx.x);
}

// ... and for launches whose dW tiles all fit the register slots (no out-of-line overflow code: see nj_seg_dw)
__global__ void __launch_bounds__(384) nj_seg_bwd_kernel_r(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjSeg seg,
                                                           const __grid_constant__ NjArgs args) {
    nj_seg_cta_backward<false, false>(cfg, seg, args, nj_smem, blockIdx.x);
}

// the same kernel for launches with dW helper warps (seg.nt_b > 32 * seg.nw_b), see nj_seg_cta_backward
__global__ void __launch_bounds__(384) nj_seg_bwd_kernel_h(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjSeg seg,
                                                           const __grid_constant__ NjArgs args) {
    nj_seg_cta_backward<true>(cfg, seg, args, nj_smem, blockIdx.x);
}


cudaError_t nj_launch_seg(const NjPlanOut& pl, const NjArgs& a, bool bwd, cudaStream_t st, const char** name) {
    const NjSeg& s = pl.seg;
    if (!bwd) {
        cudaError_t e = cudaFuncSetAttribute(nj_seg_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.seg_smem_f_bytes);
        if (e != cudaSuccess) return e;
        nj_seg_fwd_kernel<<<pl.seg_grid_f, s.nw_f * 32, pl.seg_smem_f_bytes, st>>>(pl.fwd, s, a);
        *name = "nj_seg_fwd_kernel";
        return cudaGetLastError();
    }
    const bool in_regs = s.tiles_total <= s.nt_slots * s.nt_b, helpers = s.nt_b > 32 * s.nw_b;
    auto kern = helpers ? nj_seg_bwd_kernel_h : (in_regs ? nj_seg_bwd_kernel_r : nj_seg_bwd_kernel);
    *name = helpers ? "nj_seg_bwd_kernel_h" : (in_regs ? "nj_seg_bwd_kernel_r" : "nj_seg_bwd_kernel");
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.seg_smem_b_bytes);
    if (e != cudaSuccess) return e;
    kern<<<pl.seg_grid_b, s.nt_b, pl.seg_smem_b_bytes, st>>>(pl.bwd, s, a);
    return cudaGetLastError();
}
