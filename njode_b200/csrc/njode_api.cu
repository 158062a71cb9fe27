// njode_api.cu -- kernels wrapping njode_core.cuh and the C ABI declared in include/njode_b200.h
#include <cuda_runtime.h>
#include <stdio.h>
#include <string>
#include <mutex>
#include "njode_plan.h"

static thread_local std::string g_err;
static int nj_fail(int code, const std::string& msg) { g_err = msg; return code; }
int nj_set_error(int code, const char* msg) { g_err = msg; return code; }      // used by njode_sde.cu
#define NJ_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess)                                                                \
            return nj_fail(-2, std::string(#call) + ": " + cudaGetErrorString(_e));           \
    } while (0)

extern "C" const char* njode_last_error(void) { return g_err.c_str(); }
extern "C" int njode_abi_version(void) { return NJODE_ABI_VERSION; }

// number of kernels this library has launched (bench.py reports it as gpu_launches)
static long long g_launches = 0;
extern "C" long long njode_launch_count(void) { return g_launches; }
#define NJ_LAUNCHED(n) (g_launches += (n))
void nj_count_launches(int n) { g_launches += n; }          // used by njode_wide.cu

// name of the main kernel the most recent forward (0) / backward (1) call launched: bench.py labels its roofline with
// what actually ran instead of re-deriving the planner's dispatch
static const char* g_last_kernel[2] = {"", ""};
void nj_set_last_kernel(int which, const char* name) { g_last_kernel[which & 1] = name; }
extern "C" const char* njode_last_kernel(int which) { return g_last_kernel[which & 1]; }

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
extern __shared__ __align__(16) float nj_smem[];

__global__ void __launch_bounds__(256) nj_fwd_kernel(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjArgs args) {
    nj_cta_forward(cfg, args, nj_smem, blockIdx.x, gridDim.x);
}

__global__ void __launch_bounds__(256) nj_bwd_kernel(const __grid_constant__ NjCfg cfg, const __grid_constant__ NjArgs args) {
    nj_cta_backward(cfg, args, nj_smem, blockIdx.x, gridDim.x);
}

// the segment kernels (njode_seg.cuh) live in njode_api_seg.cu, compiled concurrently with this file
cudaError_t nj_launch_seg(const NjPlanOut& pl, const NjArgs& a, bool bwd, cudaStream_t st, const char** name);

// the kernels of whole-path units and of small batches (njode_path.cuh, njode_tpn.cuh) live in their own translation unit,
// njode_api_path.cu, compiled concurrently with this one
cudaError_t nj_launch_path(const NjPlanOut& pl, const NjArgs& a, bool bwd, cudaStream_t st, const char** name);
cudaError_t nj_launch_segtpn(const NjPlanOut& pl, const NjArgs& a, bool bwd, cudaStream_t st, const char** name);

// flat parameters -> zero-padded image; one block per (net, layer)
__global__ void nj_pack_kernel(const __grid_constant__ NjCfg cfg, const float* __restrict__ params, float* __restrict__ image) {
    const int n = blockIdx.x / NJODE_MAX_LINEAR, l = blockIdx.x % NJODE_MAX_LINEAR;
    const NjNet& N = cfg.net[n];
    if (l >= N.n) return;
    const int K = N.dim[l], O = N.dim[l + 1], ks = N.ks[l], rows = N.rp[l];
    for (int i = threadIdx.x; i < rows * ks; i += blockDim.x) {
        const int o = i / ks, k = i % ks;
        image[N.w_img[l] + i] = (o < O && k < K) ? params[N.w_src[l] + (long long)o * K + k] : 0.f;
    }
    for (int o = threadIdx.x; o < rows; o += blockDim.x)
        image[N.b_img[l] + o] = (o < O && N.b_src[l] >= 0) ? params[N.b_src[l] + o] : 0.f;
}

// gradient partial images -> flat gradient buffer (fixed summation order: deterministic)
__global__ void nj_grad_reduce_kernel(const __grid_constant__ NjCfg cfg, const float* __restrict__ partials, int nparts,
                                      float* __restrict__ grads) {
    const int n = blockIdx.y / NJODE_MAX_LINEAR, l = blockIdx.y % NJODE_MAX_LINEAR;
    const NjNet& N = cfg.net[n];
    if (l >= N.n) return;
    const int K = N.dim[l], O = N.dim[l + 1], ks = N.ks[l];
    const int total = O * K + (N.b_src[l] >= 0 ? O : 0);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        long long dst; int src;
        if (i < O * K) { const int o = i / K, k = i % K; src = N.w_img[l] + o * ks + k; dst = N.w_src[l] + i; }
        else { const int o = i - O * K; src = N.b_img[l] + o; dst = N.b_src[l] + o; }
        float s = 0.f;
        for (int p = 0; p < nparts; ++p) s += partials[(size_t)p * cfg.img_floats + src];
        grads[dst] = s;
    }
}

// loss = (sum_r row_loss[r]) / batch_size   (single block, fixed summation order, fp64 accumulation;
// 1024 threads x 4 independent accumulators so that the latency of the loads overlaps)
__global__ void __launch_bounds__(1024) nj_loss_reduce_kernel(const float* __restrict__ row_loss, int N, float inv_b, float* __restrict__ loss) {
    __shared__ double sh[1024];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int i = threadIdx.x;
    for (; i + 3 * 1024 < N; i += 4 * 1024) {
        s0 += (double)row_loss[i]; s1 += (double)row_loss[i + 1024];
        s2 += (double)row_loss[i + 2048]; s3 += (double)row_loss[i + 3072];
    }
    for (; i < N; i += 1024) s0 += (double)row_loss[i];
    sh[threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = (float)(sh[0] * (double)inv_b);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// optional device-side timing of the two main kernels (NJODE_TIMING=1 or njode_set_timing(1)):
// cudaEvents recorded on the launching stream right around nj_fwd_kernel / nj_bwd_kernel.
static int g_timing = -1;
static cudaEvent_t g_ev[4];
static bool g_ev_ok = false, g_ev_rec[2] = {false, false};
static bool nj_timing_on() {
    if (g_timing < 0) { const char* e = getenv("NJODE_TIMING"); g_timing = (e && atoi(e)) ? 1 : 0; }
    if (g_timing && !g_ev_ok) { for (int i = 0; i < 4; ++i) cudaEventCreate(&g_ev[i]); g_ev_ok = true; }
    return g_timing != 0;
}
extern "C" void njode_set_timing(int on) { g_timing = on ? 1 : 0; }
int nj_timing_flag() { if (g_timing < 0) { const char* e = getenv("NJODE_TIMING"); g_timing = (e && atoi(e)) ? 1 : 0; } return g_timing; }   // njode_wide.cu
// elapsed milliseconds of the most recent forward / backward main kernel (-1: none recorded)
extern "C" int njode_get_timing(float* fwd_ms, float* bwd_ms) {
    float f = -1.f, b = -1.f;
    if (g_ev_ok && g_ev_rec[0]) { cudaEventSynchronize(g_ev[1]); cudaEventElapsedTime(&f, g_ev[0], g_ev[1]); }
    if (g_ev_ok && g_ev_rec[1]) { cudaEventSynchronize(g_ev[3]); cudaEventElapsedTime(&b, g_ev[2], g_ev[3]); }
    if (fwd_ms) *fwd_ms = f;
    if (bwd_ms) *bwd_ms = b;
    return 0;
}

struct NjDev { int sms; size_t smem_optin; bool ok; };
static NjDev g_dev[64];
static std::mutex g_mu;

static int nj_device_info(int dev, NjDev& out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (dev < 0 || dev >= 64) return nj_fail(-1, "bad device ordinal");
    if (!g_dev[dev].ok) {
        int sms = 0, optin = 0;
        NJ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        NJ_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        g_dev[dev].sms = sms; g_dev[dev].smem_optin = (size_t)optin; g_dev[dev].ok = true;
    }
    out = g_dev[dev];
    return 0;
}

static int nj_plan_for(const njode_model_t* model, const njode_batch_t* b, int dev, NjPlanOut& out) {
    if (!model || !b) return nj_fail(-1, "null model/batch");
    NjDev di;
    if (int rc = nj_device_info(dev, di)) return rc;
    std::string err;
    const char* fp = getenv("NJODE_FORCE_TILE");
    if (!nj_plan_all(*model, *b, di.sms, di.smem_optin, fp ? atoi(fp) : 0, out, err))
        return nj_fail(-3, err);
    // gradient partials: sized for the largest grid any backward launch of this model may use
    const size_t cap = (size_t)di.sms * 2;
    out.grid_bwd = (int)std::min<size_t>(out.grid_bwd, cap);
    out.ws_bytes = out.ws_partials_off + cap * std::max(out.fwd.img_floats, out.bwd.img_floats) * sizeof(float);
    // recompute mode of the segment backward: [grid][S][P_b][sH] floats, offered up to 1 GiB
    out.ws_scratch_off = (out.ws_bytes + 255) & ~(size_t)255;
    out.scratch_bytes = 0;
    if (out.seg.ok) {
        const size_t sb = (size_t)out.seg_grid_b * (size_t)std::max(1, (int)b->S) * out.seg.P_b * out.seg.sH * sizeof(float);
        if (sb <= ((size_t)1 << 30)) { out.scratch_bytes = sb; out.ws_bytes = out.ws_scratch_off + sb; }
    }
    return 0;
}

extern "C" int njode_plan(const njode_model_t* model, const njode_batch_t* batch_shape, int device,
                          njode_plan_t* p) {
    NjPlanOut o;
    if (int rc = nj_plan_for(model, batch_shape, device, o)) return rc;
    p->tile_paths = o.fwd.P; p->threads = o.fwd.nt; p->grid_fwd = o.grid_fwd; p->grid_bwd = o.grid_bwd;
    p->weights_in_smem = o.fwd.w_smem; p->grads_in_smem = o.bwd.dw_smem;
    p->smem_fwd_bytes = (int64_t)o.smem_fwd_bytes; p->smem_bwd_bytes = (int64_t)o.smem_bwd_bytes;
    p->image_floats = o.fwd.img_floats; p->workspace_bytes = (int64_t)o.ws_bytes;
    p->recompute_bytes = (int64_t)o.scratch_bytes;
    p->act_bytes = (int64_t)o.act_bytes;
    return 0;
}

static void nj_fill_args(NjArgs& a, const njode_batch_t* b, const NjPlanOut& pl, char* ws) {
    memset(&a, 0, sizeof(a));
    a.b = *b;
    a.image = reinterpret_cast<const float*>(ws + pl.ws_image_off);
    a.row_loss = reinterpret_cast<float*>(ws + pl.ws_rowloss_off);
    a.partials = reinterpret_cast<float*>(ws + pl.ws_partials_off);
    a.counter = reinterpret_cast<int*>(ws + pl.ws_counter_off);
}

extern "C" int njode_forward(const njode_model_t* model, const njode_batch_t* batch, const float* params,
                             float* hT, float* loss, float* path_h, float* path_y,
                             const njode_saved_t* saved, void* workspace, void* stream) {
    int dev = 0;
    NJ_CUDA(cudaGetDevice(&dev));
    NjPlanOut pl;
    if (int rc = nj_plan_for(model, batch, dev, pl)) return rc;
    if (!workspace || !params || !hT) return nj_fail(-1, "null buffer");
    if (batch->E > 0 && (!path_h || !path_y)) return nj_fail(-1, "return_path needs path_h/path_y");
    if (loss && !batch->n_obs_ot) return nj_fail(-1, "loss requested without n_obs_ot");
    if (model->masked && !batch->M) return nj_fail(-1, "masked model needs M");        // NJODE/models.py:263
    cudaStream_t st = (cudaStream_t)stream;
    NjArgs a;
    nj_fill_args(a, batch, pl, (char*)workspace);
    a.hT = hT; a.path_h = path_h; a.path_y = path_y;
    a.h_hist = saved ? saved->h_hist : nullptr;
    a.h_before = saved ? saved->h_before : nullptr;
    a.y_after = saved ? saved->y_after : nullptr;
    a.act_hist = (saved && pl.act_bytes > 0) ? saved->act_hist : nullptr; a.act_nh = pl.act_nh; a.act_wp = pl.act_wp;
    a.get_loss = loss ? 1 : 0;
    a.n_tiles = pl.n_tiles;
    nj_pack_kernel<<<NJODE_NUM_NETS * NJODE_MAX_LINEAR, 256, 0, st>>>(pl.fwd, params, const_cast<float*>(a.image));
    NJ_LAUNCHED(1 + (batch->n_units > 0 ? 1 : 0) + (loss ? 1 : 0));
    if (loss && batch->N > 0) NJ_CUDA(cudaMemsetAsync(a.row_loss, 0, (size_t)batch->N * 4, st));
    const bool tm = nj_timing_on();
    if (pl.seg.ok) {
        NJ_CUDA(cudaMemsetAsync(a.counter, 0, 4, st));
        if (tm) cudaEventRecord(g_ev[0], st);
        if (pl.seg.tpn) {
            const char* name = "";
            NJ_CUDA(nj_launch_segtpn(pl, a, false, st, &name));
            nj_set_last_kernel(0, name);
        } else {
            const char* name = "";
            NJ_CUDA(nj_launch_seg(pl, a, false, st, &name));
            nj_set_last_kernel(0, name);
        }
    } else if (pl.path.ok) {
        NJ_CUDA(cudaMemsetAsync(a.counter, 0, 4, st));
        const char* name = "";
        if (tm) cudaEventRecord(g_ev[0], st);
        NJ_CUDA(nj_launch_path(pl, a, false, st, &name));
        nj_set_last_kernel(0, name);
    } else {
        nj_set_last_kernel(0, "nj_fwd_kernel");
        auto kern = nj_fwd_kernel;
        NJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_fwd_bytes));
        if (tm) cudaEventRecord(g_ev[0], st);
        if (batch->n_units > 0)
            kern<<<pl.grid_fwd, pl.fwd.nt, pl.smem_fwd_bytes, st>>>(pl.fwd, a);
    }
    if (tm) { cudaEventRecord(g_ev[1], st); g_ev_rec[0] = true; }
    if (loss) nj_loss_reduce_kernel<<<1, 1024, 0, st>>>(a.row_loss, batch->N, 1.f / (float)batch->batch_size_norm, loss);
    NJ_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int njode_backward(const njode_model_t* model, const njode_batch_t* batch, const float* params,
                              const njode_saved_t* saved, const float* grad_loss, const float* grad_hT,
                              float* grads, void* workspace, void* stream) {
    int dev = 0;
    NJ_CUDA(cudaGetDevice(&dev));
    NjPlanOut pl;
    if (int rc = nj_plan_for(model, batch, dev, pl)) return rc;
    if (!workspace || !params || !grads || !grad_loss || !saved) return nj_fail(-1, "null buffer");
    // nothing saved by the forward pass: the segment backward recomputes every segment from its checkpoint
    const bool recompute = !saved->h_hist && !saved->h_before && pl.seg.ok && pl.scratch_bytes > 0;
    if (!recompute) {
        if (batch->S > 0 && !saved->h_hist) return nj_fail(-1, "backward needs the h history of the forward pass (njode_plan: recompute_bytes == 0)");
        if (batch->N > 0 && (!saved->h_before || (model->masked && !saved->y_after))) return nj_fail(-1, "backward needs h_before / y_after");
    }
    cudaStream_t st = (cudaStream_t)stream;
    NjArgs a;
    nj_fill_args(a, batch, pl, (char*)workspace);
    a.h_hist = saved->h_hist; a.h_before = saved->h_before; a.y_after = saved->y_after;
    a.act_hist = pl.act_bytes > 0 ? saved->act_hist : nullptr; a.act_nh = pl.act_nh; a.act_wp = pl.act_wp;
    if (recompute) a.scratch = reinterpret_cast<float*>((char*)workspace + pl.ws_scratch_off);
    a.grad_loss = grad_loss; a.grad_hT = grad_hT;
    a.get_loss = 1;
    a.n_tiles = pl.n_tiles;
    // the image is rebuilt: backward may run after an optimizer that shares the workspace
    nj_pack_kernel<<<NJODE_NUM_NETS * NJODE_MAX_LINEAR, 256, 0, st>>>(pl.bwd, params, const_cast<float*>(a.image));
    NJ_LAUNCHED(2 + (batch->n_units > 0 ? 1 : 0));
    int nparts = 0;
    const bool tm = nj_timing_on();
    if (pl.seg.ok) {
        NJ_CUDA(cudaMemsetAsync(a.counter, 0, 4, st));
        if (tm) cudaEventRecord(g_ev[2], st);
        nparts = pl.seg_grid_b;
        if (pl.seg.tpn) {
            const char* name = "";
            NJ_CUDA(nj_launch_segtpn(pl, a, true, st, &name));
            nj_set_last_kernel(1, name);
        } else {
            const char* name = "";
            NJ_CUDA(nj_launch_seg(pl, a, true, st, &name));
            nj_set_last_kernel(1, name);
        }
    } else if (pl.path.ok) {
        NJ_CUDA(cudaMemsetAsync(a.counter, 0, 4, st));
        const char* name = "";
        if (tm) cudaEventRecord(g_ev[2], st);
        nparts = pl.path_grid_b;
        NJ_CUDA(nj_launch_path(pl, a, true, st, &name));
        nj_set_last_kernel(1, name);
    } else {
        nj_set_last_kernel(1, "nj_bwd_kernel");
        auto kern = nj_bwd_kernel;
        NJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bwd_bytes));
        if (tm) cudaEventRecord(g_ev[2], st);
        if (batch->n_units > 0) {
            nparts = pl.grid_bwd;
            kern<<<pl.grid_bwd, pl.bwd.nt, pl.smem_bwd_bytes, st>>>(pl.bwd, a);
        }
    }
    if (tm) { cudaEventRecord(g_ev[3], st); g_ev_rec[1] = true; }
    dim3 g((unsigned)std::max(1, std::min(64, (int)((model->n_params + 255) / 256))), NJODE_NUM_NETS * NJODE_MAX_LINEAR);
    nj_grad_reduce_kernel<<<g, 256, 0, st>>>(pl.bwd, a.partials, nparts, grads);
    NJ_CUDA(cudaGetLastError());
    return 0;
}
