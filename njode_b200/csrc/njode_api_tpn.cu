// njode_api_tpn.cu -- the thread-per-neuron kernels of small batches (njode_tpn.cuh: whole paths and segments) with their
// launch functions; a translation unit of its own (compiled concurrently with the other njode_api*.cu files)
#include <cuda_runtime.h>
#include "njode_plan.h"

extern __shared__ __align__(16) float nj_smem[];

// thread-per-neuron kernels of small whole-path batches (njode_tpn.cuh): F / T / D warps around a glue warp
typedef NjTpnDims<NJN_A_KC0, NJN_A_KCH, NJN_A_HC, 1> NjTpnA1;
typedef NjTpnDims<NJN_A_KC0, NJN_A_KCH, NJN_A_HC, 4> NjTpnA4;
typedef NjTpnDims<NJN_B_KC0, NJN_B_KCH, NJN_B_HC, 1> NjTpnB1;
typedef NjTpnDims<NJN_B_KC0, NJN_B_KCH, NJN_B_HC, 4> NjTpnB4;
template <class D>
__global__ void __launch_bounds__(NJN_NT_FWD) nj_tpn_fwd_kernel(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjPath path,
                                                                const __grid_constant__ NjArgs args) {
    nj_tpn_cta_forward<D>(cfg, path, args, nj_smem);
}
template <class D>
__global__ void __launch_bounds__(NJN_NT_BWD) nj_tpn_bwd_kernel(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjPath path,
                                                                const __grid_constant__ NjArgs args) {
    nj_tpn_cta_backward<D>(cfg, path, args, nj_smem, blockIdx.x);
}
// segment units of small batches on the same roles (nj_segtpn_*, tiles of 4 segments)
template <class D>
__global__ void __launch_bounds__(NJN_NT_FWD) nj_segtpn_fwd_kernel(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjSeg seg,
                                                                   const __grid_constant__ NjArgs args) {
    nj_segtpn_cta_forward<D>(cfg, seg, args, nj_smem);
}
template <class D>
__global__ void __launch_bounds__(NJN_NT_BWD) nj_segtpn_bwd_kernel(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjSeg seg,
                                                                   const __grid_constant__ NjArgs args) {
    nj_segtpn_cta_backward<D>(cfg, seg, args, nj_smem, blockIdx.x);
}
typedef void (*nj_path_kern_t)(const NjCfg, const NjPath, const NjArgs);
static nj_path_kern_t nj_tpn_pick(int cls, int R, bool bwd, const char** name) {
    if (!bwd) {
        if (cls == 1 && R == 1) { *name = "nj_tpn_fwd_kernel<A,1>"; return nj_tpn_fwd_kernel<NjTpnA1>; }
        if (cls == 1) { *name = "nj_tpn_fwd_kernel<A,4>"; return nj_tpn_fwd_kernel<NjTpnA4>; }
        if (R == 1) { *name = "nj_tpn_fwd_kernel<B,1>"; return nj_tpn_fwd_kernel<NjTpnB1>; }
        *name = "nj_tpn_fwd_kernel<B,4>"; return nj_tpn_fwd_kernel<NjTpnB4>;
    }
    if (cls == 1 && R == 1) { *name = "nj_tpn_bwd_kernel<A,1>"; return nj_tpn_bwd_kernel<NjTpnA1>; }
    if (cls == 1) { *name = "nj_tpn_bwd_kernel<A,4>"; return nj_tpn_bwd_kernel<NjTpnA4>; }
    if (R == 1) { *name = "nj_tpn_bwd_kernel<B,1>"; return nj_tpn_bwd_kernel<NjTpnB1>; }
    *name = "nj_tpn_bwd_kernel<B,4>"; return nj_tpn_bwd_kernel<NjTpnB4>;
}

cudaError_t nj_launch_tpn(const NjPlanOut& pl, const NjArgs& a, bool bwd, cudaStream_t st, const char** name) {
    const NjPath& p = pl.path;
    nj_path_kern_t kern = bwd ? nj_tpn_pick(p.tpn, p.rg_b * p.tr_b, true, name) : nj_tpn_pick(p.tpn, p.rg_f * p.tr_f, false, name);
    const size_t smem = bwd ? pl.path_smem_b_bytes : pl.path_smem_f_bytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (!bwd) kern<<<pl.path_grid_f, NJN_NT_FWD, smem, st>>>(pl.fwd, p, a);
    else kern<<<pl.path_grid_b, p.nt_b, smem, st>>>(pl.bwd, p, a);
    return cudaGetLastError();
}

cudaError_t nj_launch_segtpn(const NjPlanOut& pl, const NjArgs& a, bool bwd, cudaStream_t st, const char** name) {
    const NjSeg& s = pl.seg;
    typedef void (*kern_t)(const NjCfg, const NjSeg, const NjArgs);
    kern_t kern;
    if (!bwd) { kern = s.tpn == 1 ? nj_segtpn_fwd_kernel<NjTpnA4> : nj_segtpn_fwd_kernel<NjTpnB4>; *name = s.tpn == 1 ? "nj_segtpn_fwd_kernel<A>" : "nj_segtpn_fwd_kernel<B>"; }
    else { kern = s.tpn == 1 ? nj_segtpn_bwd_kernel<NjTpnA4> : nj_segtpn_bwd_kernel<NjTpnB4>; *name = s.tpn == 1 ? "nj_segtpn_bwd_kernel<A>" : "nj_segtpn_bwd_kernel<B>"; }
    const size_t smem = bwd ? pl.seg_smem_b_bytes : pl.seg_smem_f_bytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (!bwd) kern<<<pl.seg_grid_f, NJN_NT_FWD, smem, st>>>(pl.fwd, s, a);
    else kern<<<pl.seg_grid_b, s.nt_b, smem, st>>>(pl.bwd, s, a);
    return cudaGetLastError();
}
