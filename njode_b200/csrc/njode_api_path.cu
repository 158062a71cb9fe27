// njode_api_path.cu -- kernels of whole-path units (njode_path.cuh: warp GEMMs, pipelined backward, K-split stationary)
// with their launch function; a translation unit of its own so that it compiles concurrently with njode_api.cu (the C ABI entry points)
#include <cuda_runtime.h>
#include "njode_plan.h"

extern __shared__ __align__(16) float nj_smem[];

// whole-path units on the warp GEMMs (njode_path.cuh); one kernel per (row groups, rows per group) tile shape
template <int RG, int TR>
__global__ void __launch_bounds__(384) nj_path_fwd_kernel(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjPath path,
                                                          const __grid_constant__ NjArgs args) {
    nj_path_cta_forward<RG, TR>(cfg, path, args, nj_smem);
}
template <int RG, int TR>
__global__ void __launch_bounds__(384) nj_path_bwd_kernel(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjPath path,
                                                          const __grid_constant__ NjArgs args) {
    nj_path_cta_backward<RG, TR>(cfg, path, args, nj_smem, blockIdx.x);
}
// the same units with weight-stationary Euler steps (small batches): all warps of a CTA on one tile
// (13 warps: warps are allocated in groups of 4, so the register file gives a 416-thread CTA 128 registers per thread)
template <int RG, int TR>
__global__ void __launch_bounds__(416) nj_stat_fwd_kernel(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjPath path,
                                                          const __grid_constant__ NjArgs args) {
    nj_stat_cta_forward<RG, TR>(cfg, path, args, nj_smem);
}
template <int RG, int TR>
__global__ void __launch_bounds__(416) nj_stat_bwd_kernel(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjPath path,
                                                          const __grid_constant__ NjArgs args) {
    nj_stat_cta_backward<RG, TR>(cfg, path, args, nj_smem, blockIdx.x);
}
// pipelined backward: dW of the ODE network on helper warps, concurrent with the row warps' next step
template <int RG, int TR>
__global__ void __launch_bounds__(384) nj_path_bwd_pipe_kernel(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjPath path,
                                                               const __grid_constant__ NjArgs args) {
    nj_path_cta_backward_pipe<RG, TR>(cfg, path, args, nj_smem, blockIdx.x);
}
typedef void (*nj_path_kern_t)(const NjCfg, const NjPath, const NjArgs);
static nj_path_kern_t nj_pipe_pick(int rg, int tr, const char** name) {
    if (rg == 1) { *name = "nj_path_bwd_pipe_kernel<1,1>"; return nj_path_bwd_pipe_kernel<1, 1>; }
    if (rg == 2) { *name = "nj_path_bwd_pipe_kernel<2,1>"; return nj_path_bwd_pipe_kernel<2, 1>; }
    if (tr == 1) { *name = "nj_path_bwd_pipe_kernel<4,1>"; return nj_path_bwd_pipe_kernel<4, 1>; }
    *name = "nj_path_bwd_pipe_kernel<4,2>"; return nj_path_bwd_pipe_kernel<4, 2>;
}
static nj_path_kern_t nj_stat_pick(int rg, int tr, bool bwd, const char** name) {
    if (!bwd) {
        if (rg == 1) { *name = "nj_stat_fwd_kernel<1,1>"; return nj_stat_fwd_kernel<1, 1>; }
        if (rg == 2) { *name = "nj_stat_fwd_kernel<2,1>"; return nj_stat_fwd_kernel<2, 1>; }
        if (tr == 1) { *name = "nj_stat_fwd_kernel<4,1>"; return nj_stat_fwd_kernel<4, 1>; }
        *name = "nj_stat_fwd_kernel<4,2>"; return nj_stat_fwd_kernel<4, 2>;
    }
    if (rg == 1) { *name = "nj_stat_bwd_kernel<1,1>"; return nj_stat_bwd_kernel<1, 1>; }
    if (rg == 2) { *name = "nj_stat_bwd_kernel<2,1>"; return nj_stat_bwd_kernel<2, 1>; }
    if (tr == 1) { *name = "nj_stat_bwd_kernel<4,1>"; return nj_stat_bwd_kernel<4, 1>; }
    *name = "nj_stat_bwd_kernel<4,2>"; return nj_stat_bwd_kernel<4, 2>;
}
static nj_path_kern_t nj_path_pick(int rg, int tr, bool bwd, const char** name) {
    if (!bwd) {
        if (rg == 1) { *name = "nj_path_fwd_kernel<1,1>"; return nj_path_fwd_kernel<1, 1>; }
        if (rg == 2) { *name = "nj_path_fwd_kernel<2,1>"; return nj_path_fwd_kernel<2, 1>; }
        if (tr == 1) { *name = "nj_path_fwd_kernel<4,1>"; return nj_path_fwd_kernel<4, 1>; }
        *name = "nj_path_fwd_kernel<4,2>"; return nj_path_fwd_kernel<4, 2>;
    }
    if (rg == 1) { *name = "nj_path_bwd_kernel<1,1>"; return nj_path_bwd_kernel<1, 1>; }
    if (rg == 2) { *name = "nj_path_bwd_kernel<2,1>"; return nj_path_bwd_kernel<2, 1>; }
    if (tr == 1) { *name = "nj_path_bwd_kernel<4,1>"; return nj_path_bwd_kernel<4, 1>; }
    *name = "nj_path_bwd_kernel<4,2>"; return nj_path_bwd_kernel<4, 2>;
}


cudaError_t nj_launch_tpn(const NjPlanOut& pl, const NjArgs& a, bool bwd, cudaStream_t st, const char** name);      // njode_api_tpn.cu

cudaError_t nj_launch_path(const NjPlanOut& pl, const NjArgs& a, bool bwd, cudaStream_t st, const char** name) {
    const NjPath& p = pl.path;
    if (p.tpn && (bwd || p.tpn_fwd)) return nj_launch_tpn(pl, a, bwd, st, name);
    nj_path_kern_t kern;
    if (!bwd) kern = p.stat ? nj_stat_pick(p.rg_f, p.tr_f, false, name) : nj_path_pick(p.rg_f, p.tr_f, false, name);
    else kern = p.stat ? nj_stat_pick(p.rg_b, p.tr_b, true, name)
                       : (p.pipe ? nj_pipe_pick(p.rg_b, p.tr_b, name) : nj_path_pick(p.rg_b, p.tr_b, true, name));
    const size_t smem = bwd ? pl.path_smem_b_bytes : pl.path_smem_f_bytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (!bwd) kern<<<pl.path_grid_f, (p.stat ? p.nw_s : p.nw_f) * 32, smem, st>>>(pl.fwd, p, a);
    else kern<<<pl.path_grid_b, p.nt_b, smem, st>>>(pl.bwd, p, a);
    return cudaGetLastError();
}

