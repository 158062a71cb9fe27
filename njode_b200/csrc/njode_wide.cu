// njode_wide.cu -- host side + kernels of the tensor-core path (njode_wide.cuh): model validation,
// bf16 weight image packing, the three passes (encoder / Euler chain / readout + loss) and the C ABI
// entry points njode_wide_* declared in include/njode_b200.h.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <algorithm>
#include "njode_wide.cuh"

int nj_set_error(int code, const char* msg);        // njode_api.cu
void nj_count_launches(int n);                      // njode_api.cu
int nj_timing_flag();                               // njode_api.cu

#define NJW_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess)                                                                \
            return nj_set_error(-2, (std::string(#call) + ": " + cudaGetErrorString(_e)).c_str()); \
    } while (0)

namespace njw {

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// fills the launch-constant description; false (+ reason) when the tensor-core path does not serve the model
static bool make_cfg(const njode_model_t& m, WCfg& c, std::string& why) {
    memset(&c, 0, sizeof(c));
    const int d = m.input_size, H = m.hidden_size;
    if (m.masked) { why = "masked model"; return false; }
    if (m.output_size != d) { why = "output_size != input_size"; return false; }
    if (!(d == 1 || d == 2 || d == 4 || d == 8 || d == 16)) { why = "input_size must be a power of two <= 16"; return false; }
    if (H < 16 || H > MAX_W || (H % 16)) { why = "hidden_size must be a multiple of 16 in [16, 256]"; return false; }
    c.d = d; c.H = H; c.dout = d; c.curt = m.input_current_t ? 1 : 0;
    c.loss_kind = m.loss_kind; c.residual = m.residual ? 1 : 0; c.training = m.training ? 1 : 0;
    c.w = m.weight;
    const float p = m.dropout_p;
    c.has_drop = (m.training && p > 0.f) ? 1 : 0;
    c.keep_scale = (p < 1.f) ? 1.f / (1.f - p) : 0.f;
    const double thr = (double)p * 65536.0;
    c.thr = thr >= 65536.0 ? 65536u : (unsigned)thr;
    c.seed_lo = (unsigned)(m.dropout_seed & 0xFFFFFFFFull); c.seed_hi = (unsigned)(m.dropout_seed >> 32);
    unsigned off = 0; int boff = 0;
    for (int n = 0; n < 3; ++n) {
        const njode_mlp_t& s = m.net[n];
        WNet& N = c.net[n];
        if (s.n_linear < 1 || s.n_linear > NJODE_MAX_LINEAR) { why = "n_linear out of range"; return false; }
        const int in0 = n == NJODE_NET_ODE ? d + H + 2 + c.curt : (n == NJODE_NET_ENC ? d : H);
        const int out = n == NJODE_NET_RO ? d : H;
        if (s.dims[0] != in0 || s.dims[s.n_linear] != out) { why = "unexpected network input/output width"; return false; }
        N.n = s.n_linear;
        for (int l = 0; l < s.n_linear; ++l) {
            WLayer& L = N.l[l];
            L.in_dim = s.dims[l]; L.n = s.dims[l + 1]; L.n16 = ceil_div(L.n, 16) * 16;
            if (L.n < 1 || L.n > MAX_W) { why = "layer wider than 256"; return false; }
            const bool last = l == s.n_linear - 1;
            L.act = last ? 0 : s.act[l];
            if (!last && L.act != NJODE_ACT_TANH && L.act != NJODE_ACT_RELU) { why = "unknown activation"; return false; }
            L.drop = (!last && c.has_drop) ? 1 : 0;
            L.kind = KIND_PLAIN; L.has_aux = 0; L.aux_ksteps = 0;
            if (l == 0 && n == NJODE_NET_ODE) { L.kind = KIND_ODE0; L.kb_main = ceil_div(H, 64); L.has_aux = 1; }
            else if (l == 0 && n == NJODE_NET_ENC) { L.kind = KIND_ENC0; L.kb_main = 0; L.has_aux = 1; }
            else L.kb_main = ceil_div(L.in_dim, 64);
            if (L.has_aux) L.aux_ksteps = ceil_div(AUX_X0 + d, 16);
            L.img_off = off; off += (unsigned)(L.kb_main + L.has_aux) * (unsigned)L.n16 * 128u;
            L.bias_off = boff; boff += MAX_W;
            L.w_src = s.w_off[l]; L.b_src = s.b_off[l];
        }
    }
    c.img_bytes = off; c.bias_floats = boff;
    return true;
}

// flat fp32 parameters -> bf16 image in the shared-memory operand layout (one [n16 x 128 B] SWIZZLE_128B block per
// K-block) + fp32 bias image.  grid = 3 nets x 8 layers x 5 K-blocks.
__global__ void nj_wide_pack_kernel(const __grid_constant__ WCfg c, const float* __restrict__ params,
                                    unsigned char* __restrict__ img, float* __restrict__ bias) {
    const int n = blockIdx.x / (NJODE_MAX_LINEAR * A_BLOCKS), l = (blockIdx.x / A_BLOCKS) % NJODE_MAX_LINEAR, kb = blockIdx.x % A_BLOCKS;
    const WNet& N = c.net[n];
    if (l >= N.n) return;
    const WLayer& L = N.l[l];
    if (kb >= L.kb_main + L.has_aux) return;
    const bool aux = kb >= L.kb_main;
    const int d = c.d, H = c.H;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(img + L.img_off + (size_t)kb * L.n16 * 128);
    for (int i = threadIdx.x; i < L.n16 * 64; i += blockDim.x) {
        const int o = i >> 6, k = i & 63;
        int col = -1;
        if (!aux) {
            const int kk = kb * 64 + k;
            if (L.kind == KIND_ODE0) col = kk < H ? d + kk : -1;           // tanh(h) columns of cat[x, h, tau, tdiff(, t)]
            else col = kk < L.in_dim ? kk : -1;
        } else {
            if (k >= AUX_X0 && k < AUX_X0 + d) col = k - AUX_X0;           // tanh(x)
            else if (L.kind == KIND_ODE0 && k < 6) {
                const int which = k >> 1;                                  // hi and lo halves share the weight column
                if (which < 2 || c.curt) col = d + H + which;
            }
        }
        float v = 0.f;
        if (o < L.n && col >= 0) v = params[L.w_src + (long long)o * L.in_dim + col];
        dst[(size_t)o * 64 + ((((k >> 3) ^ (o & 7)) << 3) | (k & 7))] = __float2bfloat16_rn(v);
    }
    if (kb == 0)
        for (int o = threadIdx.x; o < MAX_W; o += blockDim.x)
            bias[L.bias_off + o] = (o < L.n && L.b_src >= 0) ? params[L.b_src + o] : 0.f;
}

extern __shared__ __align__(16) unsigned char njw_smem[];
__global__ void __launch_bounds__(NUM_THREADS, 1) nj_wide_kernel(const __grid_constant__ WCfg c, const __grid_constant__ WArgs a) {
    wide_cta(c, a, njw_smem);
}

// loss = (sum_r row_loss[r]) / batch_size, fixed order, fp64 accumulation (same as nj_loss_reduce_kernel)
__global__ void __launch_bounds__(1024) nj_wide_loss_reduce_kernel(const float* __restrict__ row_loss, int N, float inv_b, float* __restrict__ loss) {
    __shared__ double sh[1024];
    double s = 0.0;
    for (int i = threadIdx.x; i < N; i += 1024) s += (double)row_loss[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = (float)(sh[0] * (double)inv_b);
}

struct Ws { size_t img, bias, h_start, row_unit, row_loss, total; };
static Ws ws_layout(const WCfg& c, const njode_batch_t& b) {
    Ws w; size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 1023) & ~(size_t)1023; return r; };
    w.img = take(c.img_bytes);
    w.bias = take((size_t)c.bias_floats * 4);
    w.h_start = take((size_t)std::max(b.n_units, 1) * c.H * 4);
    w.row_unit = take((size_t)std::max(b.N, 1) * 4);
    w.row_loss = take((size_t)std::max(b.N, 1) * 4);
    w.total = o + 1024;
    return w;
}

static cudaEvent_t g_ev[6];
static bool g_ev_ok = false, g_ev_rec = false;

}  // namespace njw

using namespace njw;

extern "C" int njode_wide_supported(const njode_model_t* model) {
    if (!model) return 0;
    WCfg c; std::string why;
    if (!make_cfg(*model, c, why)) { nj_set_error(0, why.c_str()); return 0; }
    return 1;
}

extern "C" int64_t njode_wide_workspace_bytes(const njode_model_t* model, const njode_batch_t* batch) {
    if (!model || !batch) return -1;
    WCfg c; std::string why;
    if (!make_cfg(*model, c, why)) return nj_set_error(-3, why.c_str());
    return (int64_t)ws_layout(c, *batch).total;
}

// byte offsets {image, bias, h_start, row_unit, row_loss} inside the (1024-aligned) workspace -- tests / debugging
extern "C" int njode_wide_ws_offsets(const njode_model_t* model, const njode_batch_t* batch, int64_t* out5) {
    if (!model || !batch || !out5) return nj_set_error(-1, "null argument");
    WCfg c; std::string why;
    if (!make_cfg(*model, c, why)) return nj_set_error(-3, why.c_str());
    const Ws w = ws_layout(c, *batch);
    out5[0] = (int64_t)w.img; out5[1] = (int64_t)w.bias; out5[2] = (int64_t)w.h_start; out5[3] = (int64_t)w.row_unit; out5[4] = (int64_t)w.row_loss;
    return 0;
}

extern "C" int njode_wide_forward(const njode_model_t* model, const njode_batch_t* batch, const float* params,
                                  float* hT, float* loss, const njode_saved_t* saved, void* workspace, void* stream) {
    if (!model || !batch || !params || !hT || !workspace) return nj_set_error(-1, "null buffer");
    WCfg c; std::string why;
    if (!make_cfg(*model, c, why)) return nj_set_error(-3, ("tensor-core path unavailable: " + why).c_str());
    if (batch->unit_kind != 1) return nj_set_error(-3, "tensor-core path needs segment units (unit_kind 1)");
    if (batch->E > 0) return nj_set_error(-3, "tensor-core path does not record paths");
    if (loss && !batch->n_obs_ot) return nj_set_error(-1, "loss requested without n_obs_ot");
    int dev = 0, sms = 0;
    NJW_CUDA(cudaGetDevice(&dev));
    NJW_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t st = (cudaStream_t)stream;
    const Ws w = ws_layout(c, *batch);
    char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    WArgs a;
    memset(&a, 0, sizeof(a));
    a.b = *batch;
    a.wimg = reinterpret_cast<const unsigned char*>(base + w.img);
    a.bias = reinterpret_cast<const float*>(base + w.bias);
    a.h_start = reinterpret_cast<float*>(base + w.h_start);
    a.row_unit = reinterpret_cast<int*>(base + w.row_unit);
    a.row_loss = reinterpret_cast<float*>(base + w.row_loss);
    a.hT = hT;
    a.h_hist = saved ? saved->h_hist : nullptr;
    a.h_before = saved ? saved->h_before : nullptr;
    a.y_after = saved ? saved->y_after : nullptr;
    a.get_loss = loss ? 1 : 0;
    // h_before feeds the readout pass: without a caller buffer there is nothing to read it from
    if (loss && batch->N > 0 && !a.h_before) return nj_set_error(-1, "the tensor-core forward needs saved->h_before when a loss is requested");
    const int n_loss = batch->n_loss_units, n_tail = batch->n_units - n_loss;
    a.n_tiles_loss = ceil_div(n_loss, TILE_M);
    const int n_tiles_all = a.n_tiles_loss + ceil_div(n_tail, TILE_M);
    const bool timing = nj_timing_flag() != 0;
    if (timing && !g_ev_ok) { for (int i = 0; i < 6; ++i) cudaEventCreate(&g_ev[i]); g_ev_ok = true; }

    nj_wide_pack_kernel<<<3 * NJODE_MAX_LINEAR * A_BLOCKS, 256, 0, st>>>(c, params, const_cast<unsigned char*>(a.wimg), const_cast<float*>(a.bias));
    int launches = 1;
    NJW_CUDA(cudaFuncSetAttribute(nj_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    if (loss && batch->N > 0) NJW_CUDA(cudaMemsetAsync(a.row_loss, 0, (size_t)batch->N * 4, st));
    if (n_tiles_all > 0) {
        a.mode = MODE_ENC; a.n_tiles = n_tiles_all;
        if (timing) cudaEventRecord(g_ev[0], st);
        nj_wide_kernel<<<std::min(n_tiles_all, sms), NUM_THREADS, SMEM_BYTES, st>>>(c, a);
        if (timing) cudaEventRecord(g_ev[1], st);
        const char* mp = getenv("NJODE_WIDE_MAXPASS");          // debugging: stop after the first n passes
        const int maxpass = mp ? atoi(mp) : 3;
        if (maxpass < 2) { nj_count_launches(launches + 1); NJW_CUDA(cudaGetLastError()); return 0; }
        a.mode = MODE_ODE;
        if (timing) cudaEventRecord(g_ev[2], st);
        nj_wide_kernel<<<std::min(n_tiles_all, sms), NUM_THREADS, SMEM_BYTES, st>>>(c, a);
        if (timing) cudaEventRecord(g_ev[3], st);
        launches += 2;
    }
    { const char* mp = getenv("NJODE_WIDE_MAXPASS"); if (mp && atoi(mp) < 3) { nj_count_launches(launches); NJW_CUDA(cudaGetLastError()); return 0; } }
    if (loss) {
        if (a.n_tiles_loss > 0) {
            a.mode = MODE_RO; a.n_tiles = a.n_tiles_loss;
            if (timing) cudaEventRecord(g_ev[4], st);
            nj_wide_kernel<<<std::min(a.n_tiles_loss, sms), NUM_THREADS, SMEM_BYTES, st>>>(c, a);
            if (timing) { cudaEventRecord(g_ev[5], st); g_ev_rec = true; }
            ++launches;
        }
        nj_wide_loss_reduce_kernel<<<1, 1024, 0, st>>>(a.row_loss, batch->N, 1.f / (float)batch->batch_size_norm, loss);
        ++launches;
    }
    nj_count_launches(launches);
    NJW_CUDA(cudaGetLastError());
    return 0;
}

// elapsed ms of the encoder / Euler-chain / readout passes of the most recent njode_wide_forward (NJODE_TIMING=1)
extern "C" int njode_wide_get_timing(float* enc_ms, float* ode_ms, float* ro_ms) {
    float e = -1.f, o = -1.f, r = -1.f;
    if (g_ev_ok && g_ev_rec) {
        cudaEventSynchronize(g_ev[5]);
        cudaEventElapsedTime(&e, g_ev[0], g_ev[1]);
        cudaEventElapsedTime(&o, g_ev[2], g_ev[3]);
        cudaEventElapsedTime(&r, g_ev[4], g_ev[5]);
    }
    if (enc_ms) *enc_ms = e;
    if (ode_ms) *ode_ms = o;
    if (ro_ms) *ro_ms = r;
    return 0;
}
