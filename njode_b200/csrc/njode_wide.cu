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
void nj_set_last_kernel(int which, const char* name);   // njode_api.cu

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
    if (m.use_rnn) { why = "use_rnn (GRU jump) runs on the fp32 whole-path kernels"; return false; }
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
    unsigned off = 0, toff = 0; int boff = 0;
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
            // backward / spill geometry
            L.nt = L.kind == KIND_ODE0 ? H : (L.kind == KIND_ENC0 ? 0 : L.in_dim);
            L.nt16 = std::max(16, ceil_div(L.nt, 16) * 16);
            L.kt_blocks = ceil_div(L.n16, 64);
            L.kt_last_ksteps = ceil_div(L.n16 - 64 * (L.kt_blocks - 1), 16);
            L.wt_off = toff;
            if (L.kind != KIND_ENC0) toff += (unsigned)L.kt_blocks * (unsigned)L.nt16 * 128u;
            L.act_off = c.act_rec[n]; c.act_rec[n] += (L.kb_main + L.has_aux) * A_BLOCK_BYTES;
            L.g_off = c.g_rec[n]; c.g_rec[n] += L.kt_blocks * A_BLOCK_BYTES;
            L.dw_slot0[0] = L.dw_slot0[1] = -1; L.dw_J[0] = L.dw_J[1] = 0;
        }
    }
    c.img_bytes = off; c.bias_floats = boff; c.wt_bytes = toff;
    return true;
}

// distributes the CTAs of the dW pass: every (net, layer, part) item gets a share of the SMs proportional to the
// bytes it streams, at most one CTA per record
static void plan_dw(WCfg& c, const njode_batch_t& b, int sms) {
    const int n_loss = b.n_loss_units, n_tail = b.n_units - n_loss;
    const int tiles_loss = ceil_div(n_loss, TILE_M), tiles_all = tiles_loss + ceil_div(n_tail, TILE_M);
    const long long recs[3] = {std::max<long long>(1, ((long long)b.B * b.S) / TILE_M), tiles_all, 2LL * tiles_loss};   // ODE, ENC, RO
    int slot = 0, items = 0;
    for (int n = 0; n < 3; ++n) {
        double tot = 0;
        for (int l = 0; l < c.net[n].n; ++l) {
            const WLayer& L = c.net[n].l[l];
            if (L.kb_main > 0) tot += L.kt_blocks + L.kb_main;
            if (L.has_aux) tot += L.kt_blocks + 1;
        }
        for (int l = 0; l < c.net[n].n; ++l) {
            WLayer& L = c.net[n].l[l];
            for (int part = 0; part < 2; ++part) {
                const bool have = part == 0 ? L.kb_main > 0 : L.has_aux != 0;
                if (!have) { L.dw_slot0[part] = -1; L.dw_J[part] = 0; continue; }
                const double w = L.kt_blocks + (part == 0 ? L.kb_main : 1);
                long long J = (long long)(sms * w / tot);
                J = std::max<long long>(1, std::min<long long>(J, recs[n]));
                L.dw_slot0[part] = slot; L.dw_J[part] = (int)J;
                slot += (int)J; ++items;
            }
        }
    }
    c.dw_items = items; c.dw_slots = slot;
}

// flat fp32 parameters -> bf16 image in the shared-memory operand layout: per layer, for each 128-row half of the
// outputs, one [rows x 128 B] SWIZZLE_128B block per K-block (the order the MMA issuer consumes them) + fp32 bias
// image.  grid = 3 nets x 8 layers x 2 halves x 5 K-blocks.
__global__ void nj_wide_pack_kernel(const __grid_constant__ WCfg c, const float* __restrict__ params,
                                    unsigned char* __restrict__ img, float* __restrict__ bias) {
    int b = blockIdx.x;
    const int kb = b % 5; b /= 5;
    const int p = b % 2; b /= 2;
    const int l = b % NJODE_MAX_LINEAR, n = b / NJODE_MAX_LINEAR;
    const WNet& N = c.net[n];
    if (l >= N.n) return;
    const WLayer& L = N.l[l];
    const int nkb = L.kb_main + L.has_aux;
    if (kb >= nkb || p * 128 >= L.n16) return;
    const bool aux = kb >= L.kb_main;
    const int d = c.d, H = c.H;
    const int rows = min(128, L.n16 - 128 * p);
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(img + L.img_off + (size_t)p * 128 * 128 * nkb + (size_t)kb * rows * 128);
    for (int i = threadIdx.x; i < rows * 64; i += blockDim.x) {
        const int ol = i >> 6, k = i & 63, o = 128 * p + ol;
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
        dst[(size_t)ol * 64 + ((((k >> 3) ^ (ol & 7)) << 3) | (k & 7))] = __float2bfloat16_rn(v);
    }
    if (kb == 0 && p == 0)
        for (int o = threadIdx.x; o < MAX_W; o += blockDim.x)
            bias[L.bias_off + o] = (o < L.n && L.b_src >= 0) ? params[L.b_src + o] : 0.f;
}

extern __shared__ __align__(16) unsigned char njw_smem[];
__global__ void __launch_bounds__(NUM_THREADS, 1) nj_wide_kernel(const __grid_constant__ WCfg c, const __grid_constant__ WArgs a) {
    wide_cta(c, a, njw_smem);
}

// W_l^T images for the backward GEMMs: rows = the layer's main inputs (ODE layer 0: the tanh(h) columns) in halves of
// 128, K = outputs.  grid = 3 nets x 8 layers x 2 halves x 4 K-blocks.
__global__ void nj_wide_pack_t_kernel(const __grid_constant__ WCfg c, const float* __restrict__ params, unsigned char* __restrict__ img) {
    int b = blockIdx.x;
    const int kb = b % 4; b /= 4;
    const int p = b % 2; b /= 2;
    const int l = b % NJODE_MAX_LINEAR, n = b / NJODE_MAX_LINEAR;
    const WNet& N = c.net[n];
    if (l >= N.n) return;
    const WLayer& L = N.l[l];
    if (L.kind == KIND_ENC0 || kb >= L.kt_blocks || p * 128 >= L.nt16) return;
    const int rows = min(128, L.nt16 - 128 * p);
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(img + L.wt_off + (size_t)p * 128 * 128 * L.kt_blocks + (size_t)kb * rows * 128);
    for (int i = threadIdx.x; i < rows * 64; i += blockDim.x) {
        const int kl = i >> 6, oo = i & 63, o = kb * 64 + oo, k = 128 * p + kl;
        float v = 0.f;
        if (o < L.n && k < L.nt) v = params[L.w_src + (long long)o * L.in_dim + (L.kind == KIND_ODE0 ? c.d + k : k)];
        dst[(size_t)kl * 64 + ((((oo >> 3) ^ (kl & 7)) << 3) | (oo & 7))] = __float2bfloat16_rn(v);
    }
}

// tile_base[t] = number of Euler-chain records before tile t (exclusive prefix sum of the tiles' step counts)
__global__ void __launch_bounds__(1024) nj_wide_tile_base_kernel(const __grid_constant__ WCfg c, const __grid_constant__ WArgs a, int n_tiles) {
    __shared__ int sh[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int t = base + threadIdx.x;
        int v = 0;
        if (t < n_tiles) {
            const int nl = a.b.n_loss_units;
            const int u0 = t < a.n_tiles_loss ? t * TILE_M : nl + (t - a.n_tiles_loss) * TILE_M;
            const int32_t* dsc = a.b.unit_desc + (size_t)u0 * 6;
            v = dsc[2] - dsc[1];
        }
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int w = 1; w < 1024; w <<= 1) {
            const int add = threadIdx.x >= w ? sh[threadIdx.x - w] : 0;
            __syncthreads();
            sh[threadIdx.x] += add;
            __syncthreads();
        }
        if (t < n_tiles) a.tile_base[t] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.tile_base[n_tiles] = carry;
}

__global__ void __launch_bounds__(NUM_THREADS, 1) nj_wide_bwd_kernel(const __grid_constant__ WCfg c, const __grid_constant__ WArgs a) {
    wide_bwd_cta(c, a, njw_smem);
}

// one launch per net (a.mode = net id): CTA -> (layer, part, share) from the per-layer CTA counts
__global__ void __launch_bounds__(DW_THREADS, 1) nj_wide_dw_kernel(const __grid_constant__ WCfg c, const __grid_constant__ WArgs a) {
    __shared__ DwItem it;
    if (threadIdx.x == 0) {
        int b = blockIdx.x;
        it.net = a.mode; it.layer = -1;
        for (int l = 0; l < c.net[a.mode].n && it.layer < 0; ++l)
            for (int part = 0; part < 2; ++part) {
                const int J = c.net[a.mode].l[l].dw_J[part];
                if (b < J) { it.layer = l; it.part = part; it.j = b; it.J = J; it.slot = c.net[a.mode].l[l].dw_slot0[part] + b; break; }
                b -= J;
            }
    }
    __syncthreads();
    wide_dw_cta(c, a, it, njw_smem);
}

// partial accumulators -> flat gradient buffer (inverse of the column maps of nj_wide_pack_kernel); fixed order
__global__ void nj_wide_dw_reduce_kernel(const __grid_constant__ WCfg c, const float* __restrict__ dw_part, float* __restrict__ grads) {
    const int n = blockIdx.x / (NJODE_MAX_LINEAR * 2), l = (blockIdx.x / 2) % NJODE_MAX_LINEAR, part = blockIdx.x % 2;
    const WNet& N = c.net[n];
    if (l >= N.n) return;
    const WLayer& L = N.l[l];
    const int J = L.dw_J[part], s0 = L.dw_slot0[part];
    if (J <= 0) return;
    const int d = c.d, H = c.H;
    const float* bpart = dw_part + (size_t)c.dw_slots * (256 * 256);
    const int ncol = part == 0 ? L.nt : L.in_dim;          // aux: loop over the original columns it serves
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < L.n * ncol; i += gridDim.y * blockDim.x) {
        const int o = i / ncol, q = i % ncol;
        int col, k0, k1 = -1;
        if (part == 0) { col = L.kind == KIND_ODE0 ? d + q : q; k0 = q; }
        else {
            col = q;
            if (q < d) k0 = AUX_X0 + q;                                        // tanh(x) columns
            else if (L.kind == KIND_ODE0 && q >= d + H) { k0 = 2 * (q - d - H); k1 = k0 + 1; }   // hi + lo halves
            else continue;                                                     // served by the main part
        }
        float sum = 0.f;
        for (int j = 0; j < J; ++j) {
            const float* p = dw_part + (size_t)(s0 + j) * (256 * 256) + (size_t)o * 256;
            sum += p[k0];
            if (k1 >= 0) sum += p[k1];
        }
        grads[L.w_src + (long long)o * L.in_dim + col] = sum;
    }
    const bool do_bias = (part == 1 ? L.kb_main == 0 : true) && L.b_src >= 0;
    if (do_bias && blockIdx.y == 0)
        for (int o = threadIdx.x; o < L.n; o += blockDim.x) {
            float sum = 0.f;
            for (int j = 0; j < J; ++j) sum += bpart[(size_t)(s0 + j) * 256 + o];
            grads[L.b_src + o] = sum;
        }
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

// ---- buffer layouts ---------------------------------------------------------------------------
static inline int tiles_loss_of(const njode_batch_t& b) { return ceil_div(b.n_loss_units, TILE_M); }
static inline int tiles_all_of(const njode_batch_t& b) { return tiles_loss_of(b) + ceil_div(b.n_units - b.n_loss_units, TILE_M); }
// upper bound of the Euler-chain records: units are sorted longest first inside each of the two runs, so the
// padding of a tile is bounded by the length of the run's longest unit
static inline long long ode_recs_bound(const njode_batch_t& b) { return ((long long)b.B * b.S) / TILE_M + 2LL * b.S + 4; }

struct Lay {
    size_t o = 0;
    size_t take(size_t bytes) { size_t r = o; o += (bytes + 1023) & ~(size_t)1023; return r; }
};
// forward scratch (always) -- when no saved blob is given it also holds the buffers the readout pass needs
struct Ws { size_t img, bias, row_loss, h_start, row_unit, h_before, total; };
static Ws ws_layout(const WCfg& c, const njode_batch_t& b) {
    Ws w; Lay L;
    w.img = L.take(c.img_bytes);
    w.bias = L.take((size_t)c.bias_floats * 4);
    w.row_loss = L.take((size_t)std::max(b.N, 1) * 4);
    w.h_start = L.take((size_t)std::max(b.n_units, 1) * c.H * 4);
    w.row_unit = L.take((size_t)std::max(b.N, 1) * 4);
    w.h_before = L.take((size_t)std::max(b.N, 1) * c.H * 4);
    w.total = L.o + 1024;
    return w;
}
// everything the backward pass re-reads (owned by the caller, one blob per forward call)
struct Sv { size_t h_before, y_after, y_before, h_start, row_unit, tile_base, act_enc, act_ro, act_ode, total; };
static Sv sv_layout(const WCfg& c, const njode_batch_t& b) {
    Sv v; Lay L;
    const size_t N = std::max(b.N, 1), U = std::max(b.n_units, 1);
    v.h_before = L.take(N * c.H * 4);
    v.y_after = L.take(N * c.d * 4);
    v.y_before = L.take(N * c.d * 4);
    v.h_start = L.take(U * c.H * 4);
    v.row_unit = L.take(N * 4);
    v.tile_base = L.take((size_t)(tiles_all_of(b) + 1) * 4);
    v.act_enc = L.take((size_t)tiles_all_of(b) * c.act_rec[NJODE_NET_ENC]);
    v.act_ro = L.take((size_t)2 * tiles_loss_of(b) * c.act_rec[NJODE_NET_RO]);
    v.act_ode = L.take((size_t)ode_recs_bound(b) * c.act_rec[NJODE_NET_ODE]);
    v.total = L.o + 1024;
    return v;
}
// backward scratch
struct Wb { size_t wt, g_before, g_start, g_enc, g_ro, g_ode, dw_part, total; };
static Wb wb_layout(const WCfg& c, const njode_batch_t& b) {
    Wb w; Lay L;
    w.wt = L.take(c.wt_bytes);
    w.g_before = L.take((size_t)std::max(b.N, 1) * c.H * 4);
    w.g_start = L.take((size_t)std::max(b.n_units, 1) * c.H * 4);
    w.g_enc = L.take((size_t)tiles_all_of(b) * c.g_rec[NJODE_NET_ENC]);
    w.g_ro = L.take((size_t)2 * tiles_loss_of(b) * c.g_rec[NJODE_NET_RO]);
    w.g_ode = L.take((size_t)ode_recs_bound(b) * c.g_rec[NJODE_NET_ODE]);
    w.dw_part = L.take((size_t)c.dw_slots * (256 * 256 + 256) * 4);
    w.total = L.o + 1024;
    return w;
}
static inline char* align1k(void* p) { return reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023); }

static unsigned long long* g_prof = nullptr;      // debugging: device buffer of clock stamps (njode_wide_set_profile)
static cudaEvent_t g_ev[10];
static bool g_ev_ok = false, g_ev_rec = false, g_evb_rec = false;

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
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    plan_dw(c, *batch, sms);
    return (int64_t)std::max(ws_layout(c, *batch).total, wb_layout(c, *batch).total);
}

extern "C" int64_t njode_wide_saved_bytes(const njode_model_t* model, const njode_batch_t* batch) {
    if (!model || !batch) return -1;
    WCfg c; std::string why;
    if (!make_cfg(*model, c, why)) return nj_set_error(-3, why.c_str());
    return (int64_t)sv_layout(c, *batch).total;
}

extern "C" int njode_wide_forward(const njode_model_t* model, const njode_batch_t* batch, const float* params,
                                  float* hT, float* loss, const njode_saved_t* saved, void* wide_saved,
                                  void* workspace, void* stream) {
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
    char* base = align1k(workspace);
    WArgs a;
    memset(&a, 0, sizeof(a));
    a.b = *batch;
    a.wimg = reinterpret_cast<const unsigned char*>(base + w.img);
    a.bias = reinterpret_cast<const float*>(base + w.bias);
    a.row_loss = reinterpret_cast<float*>(base + w.row_loss);
    a.hT = hT;
    a.h_hist = saved ? saved->h_hist : nullptr;
    a.get_loss = loss ? 1 : 0;
    unsigned char* act_base[3] = {nullptr, nullptr, nullptr};
    if (wide_saved) {
        const Sv v = sv_layout(c, *batch);
        char* sb = align1k(wide_saved);
        a.h_before = reinterpret_cast<float*>(sb + v.h_before);
        a.y_after = reinterpret_cast<float*>(sb + v.y_after);
        a.y_before = reinterpret_cast<float*>(sb + v.y_before);
        a.h_start = reinterpret_cast<float*>(sb + v.h_start);
        a.row_unit = reinterpret_cast<int*>(sb + v.row_unit);
        a.tile_base = reinterpret_cast<int*>(sb + v.tile_base);
        act_base[NJODE_NET_ENC] = reinterpret_cast<unsigned char*>(sb + v.act_enc);
        act_base[NJODE_NET_RO] = reinterpret_cast<unsigned char*>(sb + v.act_ro);
        act_base[NJODE_NET_ODE] = reinterpret_cast<unsigned char*>(sb + v.act_ode);
        a.spill = 1;
    } else {
        a.h_start = reinterpret_cast<float*>(base + w.h_start);
        a.row_unit = reinterpret_cast<int*>(base + w.row_unit);
        a.h_before = (saved && saved->h_before) ? saved->h_before : reinterpret_cast<float*>(base + w.h_before);
        a.y_after = saved ? saved->y_after : nullptr;
    }
    a.n_tiles_loss = tiles_loss_of(*batch);
    const int n_tiles_all = tiles_all_of(*batch);
    a.n_tiles_all = n_tiles_all;
    const bool timing = nj_timing_flag() != 0;
    if (timing && !g_ev_ok) { for (int i = 0; i < 10; ++i) cudaEventCreate(&g_ev[i]); g_ev_ok = true; }

    nj_wide_pack_kernel<<<3 * NJODE_MAX_LINEAR * 2 * 5, 256, 0, st>>>(c, params, const_cast<unsigned char*>(a.wimg), const_cast<float*>(a.bias));
    int launches = 1;
    NJW_CUDA(cudaFuncSetAttribute(nj_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    if (loss && batch->N > 0) NJW_CUDA(cudaMemsetAsync(a.row_loss, 0, (size_t)batch->N * 4, st));
    if (n_tiles_all > 0) {
        if (a.spill) { nj_wide_tile_base_kernel<<<1, 1024, 0, st>>>(c, a, n_tiles_all); ++launches; }
        a.mode = MODE_ENC; a.n_tiles = n_tiles_all; a.act = act_base[NJODE_NET_ENC];
        if (timing) cudaEventRecord(g_ev[0], st);
        nj_wide_kernel<<<std::min(n_tiles_all, sms), NUM_THREADS, SMEM_BYTES, st>>>(c, a);
        if (timing) cudaEventRecord(g_ev[1], st);
        a.mode = MODE_ODE; a.act = act_base[NJODE_NET_ODE];
        a.prof = g_prof;
        if (timing) cudaEventRecord(g_ev[2], st);
        nj_wide_kernel<<<std::min(n_tiles_all, sms), NUM_THREADS, SMEM_BYTES, st>>>(c, a);
        if (timing) cudaEventRecord(g_ev[3], st);
        a.prof = nullptr;
        launches += 2;
    }
    if (loss) {
        if (a.n_tiles_loss > 0) {
            a.mode = MODE_RO; a.n_tiles = a.n_tiles_loss; a.act = act_base[NJODE_NET_RO];
            if (timing) cudaEventRecord(g_ev[4], st);
            nj_wide_kernel<<<std::min(a.n_tiles_loss, sms), NUM_THREADS, SMEM_BYTES, st>>>(c, a);
            if (timing) { cudaEventRecord(g_ev[5], st); g_ev_rec = true; }
            ++launches;
        }
        nj_wide_loss_reduce_kernel<<<1, 1024, 0, st>>>(a.row_loss, batch->N, 1.f / (float)batch->batch_size_norm, loss);
        ++launches;
    }
    nj_count_launches(launches);
    nj_set_last_kernel(0, "nj_wide_kernel");
    NJW_CUDA(cudaGetLastError());
    return 0;
}

// reverse pass of njode_wide_forward for all parameter tensors (the autograd tape replay of loss.backward(),
// NJODE/train.py:522): readout chains -> Euler-chain adjoint -> encoder chain -> dW pass -> flat gradients
extern "C" int njode_wide_backward(const njode_model_t* model, const njode_batch_t* batch, const float* params,
                                   void* wide_saved, const float* grad_loss, const float* grad_hT, float* grads,
                                   void* workspace, void* stream) {
    if (!model || !batch || !params || !wide_saved || !grads || !workspace) return nj_set_error(-1, "null buffer");
    WCfg c; std::string why;
    if (!make_cfg(*model, c, why)) return nj_set_error(-3, ("tensor-core path unavailable: " + why).c_str());
    int dev = 0, sms = 0;
    NJW_CUDA(cudaGetDevice(&dev));
    NJW_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    plan_dw(c, *batch, sms);
    cudaStream_t st = (cudaStream_t)stream;
    const Sv v = sv_layout(c, *batch);
    const Wb w = wb_layout(c, *batch);
    char* sb = align1k(wide_saved);
    char* base = align1k(workspace);
    WArgs a;
    memset(&a, 0, sizeof(a));
    a.b = *batch;
    a.h_before = reinterpret_cast<float*>(sb + v.h_before);
    a.y_after = reinterpret_cast<float*>(sb + v.y_after);
    a.y_before = reinterpret_cast<float*>(sb + v.y_before);
    a.h_start = reinterpret_cast<float*>(sb + v.h_start);
    a.row_unit = reinterpret_cast<int*>(sb + v.row_unit);
    a.tile_base = reinterpret_cast<int*>(sb + v.tile_base);
    unsigned char* act_base[3]; unsigned char* g_base[3];
    act_base[NJODE_NET_ENC] = reinterpret_cast<unsigned char*>(sb + v.act_enc);
    act_base[NJODE_NET_RO] = reinterpret_cast<unsigned char*>(sb + v.act_ro);
    act_base[NJODE_NET_ODE] = reinterpret_cast<unsigned char*>(sb + v.act_ode);
    g_base[NJODE_NET_ENC] = reinterpret_cast<unsigned char*>(base + w.g_enc);
    g_base[NJODE_NET_RO] = reinterpret_cast<unsigned char*>(base + w.g_ro);
    g_base[NJODE_NET_ODE] = reinterpret_cast<unsigned char*>(base + w.g_ode);
    a.wt = reinterpret_cast<const unsigned char*>(base + w.wt);
    a.g_before = reinterpret_cast<float*>(base + w.g_before);
    a.g_start = reinterpret_cast<float*>(base + w.g_start);
    a.dw_part = reinterpret_cast<float*>(base + w.dw_part);
    a.grad_loss = grad_loss; a.grad_hT = grad_hT;
    a.n_tiles_loss = tiles_loss_of(*batch);
    a.n_tiles_all = tiles_all_of(*batch);
    a.spill = 1;
    const bool timing = nj_timing_flag() != 0;
    if (timing && !g_ev_ok) { for (int i = 0; i < 10; ++i) cudaEventCreate(&g_ev[i]); g_ev_ok = true; }
    int launches = 0;
    nj_wide_pack_t_kernel<<<3 * NJODE_MAX_LINEAR * 2 * 4, 256, 0, st>>>(c, params, const_cast<unsigned char*>(a.wt)); ++launches;
    NJW_CUDA(cudaFuncSetAttribute(nj_wide_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    NJW_CUDA(cudaFuncSetAttribute(nj_wide_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    // rows without a loss contribution (none today) and units never reached by the readout chains start from zero
    NJW_CUDA(cudaMemsetAsync(a.g_before, 0, (size_t)std::max(batch->N, 1) * c.H * 4, st));
    if (timing) cudaEventRecord(g_ev[6], st);
    if (a.n_tiles_loss > 0) {
        a.mode = MODE_RO; a.n_tiles = a.n_tiles_loss; a.act = act_base[NJODE_NET_RO]; a.gsp = g_base[NJODE_NET_RO];
        nj_wide_bwd_kernel<<<std::min(a.n_tiles_loss, sms), NUM_THREADS, SMEM_BYTES, st>>>(c, a); ++launches;
    }
    if (a.n_tiles_all > 0) {
        a.mode = MODE_ODE; a.n_tiles = a.n_tiles_all; a.act = act_base[NJODE_NET_ODE]; a.gsp = g_base[NJODE_NET_ODE];
        nj_wide_bwd_kernel<<<std::min(a.n_tiles_all, sms), NUM_THREADS, SMEM_BYTES, st>>>(c, a); ++launches;
        a.mode = MODE_ENC; a.act = act_base[NJODE_NET_ENC]; a.gsp = g_base[NJODE_NET_ENC];
        nj_wide_bwd_kernel<<<std::min(a.n_tiles_all, sms), NUM_THREADS, SMEM_BYTES, st>>>(c, a); ++launches;
    }
    if (timing) cudaEventRecord(g_ev[7], st);
    for (int n = 0; n < 3; ++n) {
        int ctas = 0;
        for (int l = 0; l < c.net[n].n; ++l) ctas += c.net[n].l[l].dw_J[0] + c.net[n].l[l].dw_J[1];
        if (ctas == 0) continue;
        a.mode = n; a.act = act_base[n]; a.gsp = g_base[n];
        nj_wide_dw_kernel<<<ctas, DW_THREADS, SMEM_BYTES, st>>>(c, a); ++launches;
    }
    nj_wide_dw_reduce_kernel<<<dim3(3 * NJODE_MAX_LINEAR * 2, 16), 256, 0, st>>>(c, a.dw_part, grads); ++launches;
    if (timing) { cudaEventRecord(g_ev[8], st); g_evb_rec = true; }
    nj_count_launches(launches);
    nj_set_last_kernel(1, "nj_wide_bwd_kernel");
    NJW_CUDA(cudaGetLastError());
    return 0;
}

// elapsed ms of the encoder / Euler-chain / readout passes of the most recent njode_wide_forward (timing on)
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
// elapsed ms of the chain passes and of the dW pass of the most recent njode_wide_backward (timing on)
extern "C" int njode_wide_get_timing_bwd(float* chain_ms, float* dw_ms) {
    float ch = -1.f, dw = -1.f;
    if (g_ev_ok && g_evb_rec) {
        cudaEventSynchronize(g_ev[8]);
        cudaEventElapsedTime(&ch, g_ev[6], g_ev[7]);
        cudaEventElapsedTime(&dw, g_ev[7], g_ev[8]);
    }
    if (chain_ms) *chain_ms = ch;
    if (dw_ms) *dw_ms = dw;
    return 0;
}

// debugging aid: the forward Euler-chain kernel writes clock64 stamps of CTA 0 into `buf` (4 x 256 uint64) when set
extern "C" void njode_wide_set_profile(void* buf) { njw::g_prof = reinterpret_cast<unsigned long long*>(buf); }
