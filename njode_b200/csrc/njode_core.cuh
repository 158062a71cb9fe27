// njode_core.cuh -- CTA-level NJ-ODE forward / backward for sm_100a (generic fp32 path).
//
// One CTA marches a tile of P work units (paths, or (path, inter-observation segment) pairs) in
// lockstep through the batch-global Euler schedule.  All state lives in shared memory as row-major
// [unit][feature] matrices; every Linear layer is a register-tiled shared-memory GEMM over the tile
// (4x4 micro-tiles, float4 loads along the reduction dimension), the parameter image stays
// resident in shared memory for the whole launch when it fits, and parameter gradients accumulate
// in a thread-owned shared-memory image that is flushed once per CTA.
//
// Semantics restated from the reference (paths relative to /root/reference):
//   NJODE.forward            NJODE/models.py:379-518      ode_step     NJODE/models.py:369-377
//   ODEFunc.forward          NJODE/models.py:188-199      FFNN.forward NJODE/models.py:261-276
//   compute_loss(_2)         NJODE/models.py:71-126       get_ffnn     NJODE/models.py:140-166
//
// The file is written as barrier-separated phases (NJ_THREADS ... NJ_SYNC) with all cross-thread
// state in "shared" arrays, so that the identical source also compiles as a sequential host
// simulation (-DNJODE_HOST_SIM, tests only; the product never loads that build).
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>
#include "../../include/njode_b200.h"

#if defined(NJODE_HOST_SIM)
#define NJ_HD inline
#define NJ_HDN inline
#define NJ_UNROLL4
#define NJ_THREADS(tid, nt) for (int tid = 0; tid < (nt); ++tid)
#define NJ_SYNC() ((void)0)
#define NJ_LDG(p) (*(p))
struct nj_f4 { float x, y, z, w; };
static inline nj_f4 nj_ld4(const float* p) { nj_f4 v; memcpy(&v, p, 16); return v; }
static inline void nj_st4(float* p, const nj_f4& v) { memcpy(p, &v, 16); }
#else
#define NJ_HD __device__ __forceinline__
#define NJ_HDN static __device__ __noinline__
#define NJ_UNROLL4 _Pragma("unroll 4")
#define NJ_THREADS(tid, nt) for (int tid = threadIdx.x, _nj_e = threadIdx.x + 1; tid < _nj_e; ++tid)
#define NJ_SYNC() __syncthreads()
#define NJ_LDG(p) __ldg(p)
typedef float4 nj_f4;
__device__ __forceinline__ nj_f4 nj_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void nj_st4(float* p, const nj_f4& v) { *reinterpret_cast<float4*>(p) = v; }
#endif

// ------------------------------------------------------------------------------------------------
// static configuration of one launch (passed to the kernels by value)
// ------------------------------------------------------------------------------------------------
struct NjNet {
    int n;                              // Linear layers
    int dim[NJODE_MAX_LINEAR + 1];
    int act[NJODE_MAX_LINEAR];
    int ks[NJODE_MAX_LINEAR];           // padded row stride of W_l inside the image
    int og[NJODE_MAX_LINEAR];           // ceil(out/4)
    int rp[NJODE_MAX_LINEAR];           // rows of W_l held by the image (zero padded) = nch * 8 * to >= out
    int to[NJODE_MAX_LINEAR];           // outputs per lane and chunk of the warp GEMM (njode_seg.cuh), 1..8
    int tol[NJODE_MAX_LINEAR];          // ... of the last chunk (<= to)
    int nch[NJODE_MAX_LINEAR];          // output chunks of 8 * to rows
    int w_img[NJODE_MAX_LINEAR];        // float offsets inside the image
    int b_img[NJODE_MAX_LINEAR];
    unsigned m_in[NJODE_MAX_LINEAR], m_out[NJODE_MAX_LINEAR];   // nj_magic of dim[l] / dim[l + 1]
    // row-loop slicing of the layer GEMMs for tiles of <= NJ_RL_MAXROWS rows (0 slices: micro-tile path); filled by the
    // planner for the tile height of the launch: forward (slices, float4 chunks per slice), dx (first thread, slices,
    // output groups per slice)
    int rl_f_ns[NJODE_MAX_LINEAR], rl_f_kc[NJODE_MAX_LINEAR];
    int rl_d_t0[NJODE_MAX_LINEAR], rl_d_ns[NJODE_MAX_LINEAR], rl_d_oc[NJODE_MAX_LINEAR];
    long long w_src[NJODE_MAX_LINEAR];  // float offsets inside the flat parameter buffer
    long long b_src[NJODE_MAX_LINEAR];  // -1: no bias
};

struct NjCfg {
    NjNet net[NJODE_NUM_NETS];          // ODE, ENC, RO (, GRU_IH, GRU_HH when use_rnn; n = 0 otherwise)
    int d, H, dout, inf, enc_in;
    int masked, curt, loss_kind, residual, training, has_drop, use_rnn, compact;
    float w, keep_scale, one_minus_p;
    unsigned thr, seed_lo, seed_hi;
    int P, nt;
    unsigned m_H, m_d, m_dout, m_inf, m_dH, m_3H;      // nj_magic of H, d, dout, inf, d + H, 3 H
    int img_floats;
    int w_smem, dw_smem;                // dw_smem: 0 none, 1 whole gradient image, 2 ODE network part only
    int dimg_floats;                    // floats of the gradient image held in shared memory
    // strides (floats) of the shared-memory matrices
    int sIN, sACT, nACT, sOUT, sH, sD, sDO, sG, s3H;
    // offsets (floats) into dynamic shared memory
    int o_img, o_dimg, o_IN, o_ACT, o_OUT, o_H, o_LX, o_XI, o_YBJ, o_YY, o_XH, o_EE;
    int o_GOUT, o_GTMP, o_GA, o_GB, o_GH, o_GX, o_GYBJ, o_F, o_I;
    int o_KS;                           // [NJ_RL_SCRATCH] partial sums of the row-loop GEMMs (nj_rl_*)
    int o_GI, o_GHH;                    // use_rnn: [P][s3H] gate buffers of the GRU jump
    int smem_floats_fwd, smem_floats_bwd;
};

// per-unit float scalars (row-major [slot][P]) and int scalars
enum { NJ_F_TAU = 0, NJ_F_DT, NJ_F_T, NJ_F_CA, NJ_F_CB, NJ_F_COUNT };
enum { NJ_I_PATH = 0, NJ_I_S0, NJ_I_LEN, NJ_I_CUR, NJ_I_C0, NJ_I_C1, NJ_I_START, NJ_I_FLAG, NJ_I_RK,
       NJ_I_JMAP, NJ_I_JROW, NJ_I_JJMP, NJ_I_PEND, NJ_I_NEXT, NJ_I_COUNT };
enum { NJ_CTL_NJ = 0, NJ_CTL_MAXLEN, NJ_CTL_NEXT, NJ_CTL_COUNT = 4 };

static inline int nj_stride_host(int n) {
    int s = (n + 3) & ~3;
    if (s == 0) s = 4;
    if ((s & 7) == 0) s += 4;      // stride == 4 (mod 8): float4 row accesses of 8 consecutive rows
    return s;                       // hit 8 distinct 16-byte bank groups
}

#include "njode_hash.cuh"

// division by a launch constant: q = umulhi(x, ceil(2^32 / d)), exact for x * d < 2^32 (indices here are < 2^16 and
// divisors < 2^10).  The element loops of the lockstep phases decode (row, column) from a flat index; a hardware-less
// 32-bit division costs ~20 instructions, and with few warps per SM every instruction of a phase is exposed latency.
static inline unsigned nj_magic_host(int d) { return d > 1 ? 0xFFFFFFFFu / (unsigned)d + 1u : 0u; }
#if defined(NJODE_HOST_SIM)
static inline int nj_div(int x, int d, unsigned m) { return d > 1 ? (int)(((unsigned long long)(unsigned)x * m) >> 32) : x; }
static inline unsigned nj_magic(int d) { return nj_magic_host(d); }
#else
__device__ __forceinline__ int nj_div(int x, int d, unsigned m) { return d > 1 ? (int)__umulhi((unsigned)x, m) : x; }
__device__ __forceinline__ unsigned nj_magic(int d) { return d > 1 ? 0xFFFFFFFFu / (unsigned)d + 1u : 0u; }
#endif

NJ_HD float nj_act(float v, int act) {
    if (act == NJODE_ACT_TANH) return nj_tanh(v);
    if (act == NJODE_ACT_RELU) return v > 0.f ? v : 0.f;
    return v;
}

// shared-memory operand pointers of the GEMM inner loops: on the device a 32-bit shared-window address read
// with ld.shared.v4.f32 (the generic-pointer path lost the 128-bit vector width behind the non-inlined
// layer functions); on the host simulation a plain pointer.
#if defined(NJODE_HOST_SIM)
typedef const float* nj_sp;
static inline nj_sp nj_sp_of(const float* p) { return p; }
static inline nj_f4 nj_sp_ld4(nj_sp p) { return nj_ld4(p); }
#define NJ_SP_ADD(p, nfloats) ((p) + (nfloats))
#else
typedef unsigned nj_sp;
__device__ __forceinline__ nj_sp nj_sp_of(const float* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ nj_f4 nj_sp_ld4(nj_sp p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(p) : "memory");
    return v;
}
#define NJ_SP_ADD(p, nfloats) ((p) + 4u * (unsigned)(nfloats))
#endif

// weight operand of the micro-tile GEMMs: the shared-memory image (WS) or the image in global memory
template <bool WS> struct NjWPtr;
template <> struct NjWPtr<true> {
    nj_sp p;
    NJ_HD void set(const float* q) { p = nj_sp_of(q); }
    NJ_HD nj_f4 ld(int off) const { return nj_sp_ld4(NJ_SP_ADD(p, off)); }
};
template <> struct NjWPtr<false> {
    const float* p;
    NJ_HD void set(const float* q) { p = q; }
    NJ_HD nj_f4 ld(int off) const { return nj_ld4(p + off); }
};

// ------------------------------------------------------------------------------------------------
// tile GEMMs.  All matrices row-major, rows 16-byte aligned, padding columns/rows zero (weights) or
// finite (activations; they only ever meet zero weights).
// ------------------------------------------------------------------------------------------------
struct NjLin {
    const float* in; int in_s;      // [nrows][in_s]
    int K4;                         // float4 chunks along the reduction dimension
    const float* W; int w_s;        // image rows [4*OG][w_s]
    const float* bias;              // [4*OG] or nullptr
    int O, OG; unsigned mO;
    float* out; int out_s;
    int nrows;
    int act;                        // activation applied in the epilogue (NONE for the last layer)
    int drop; unsigned thr; float keep_scale; const int* rk; unsigned tag;
};

// out[r][o] = act(bias[o] + sum_k in[r][k] W[o][k]) (* dropout)
template <int MR, bool WS>
NJ_HD void nj_tile_fwd(const NjLin& L, int tid, int nt) {
    const int RG = (L.nrows + MR - 1) / MR;
    const int ntiles = RG * L.OG;
    if (tid >= ntiles) return;
    const unsigned mRG = nj_magic(RG);
    for (int tile = tid; tile < ntiles; tile += nt) {
        const int og = nj_div(tile, RG, mRG), rg = tile - og * RG;
        float acc[MR][4];
        nj_sp ap[MR];
        NjWPtr<WS> wp[4];
#pragma unroll
        for (int i = 0; i < MR; ++i) {
            int r = rg + RG * i; if (r >= L.nrows) r = L.nrows - 1;
            ap[i] = nj_sp_of(L.in + (size_t)r * L.in_s);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) wp[j].set(L.W + (size_t)(og + L.OG * j) * L.w_s);
#pragma unroll 2
        for (int k4 = 0; k4 < L.K4; ++k4) {
            nj_f4 a[MR], w[4];
#pragma unroll
            for (int i = 0; i < MR; ++i) a[i] = nj_sp_ld4(NJ_SP_ADD(ap[i], 4 * k4));
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = wp[j].ld(4 * k4);
#pragma unroll
            for (int i = 0; i < MR; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(a[i].x, w[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, w[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, w[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, w[j].w, acc[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < MR; ++i) {
            const int r = rg + RG * i;
            if (r >= L.nrows) continue;
            unsigned lk = 0;
            if (L.drop) lk = nj_layer_key((unsigned)L.rk[r], L.tag);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int o = og + L.OG * j;
                if (o >= L.O) continue;
                float v = acc[i][j] + (L.bias ? L.bias[o] : 0.f);
                v = nj_act(v, L.act);
                if (L.drop) v = nj_keep(lk, (unsigned)o, L.thr) ? v * L.keep_scale : 0.f;
                L.out[(size_t)r * L.out_s + o] = v;
            }
        }
    }
}

struct NjDx {
    const float* g; int g_s;        // gradient wrt the layer's pre-activation output [nrows][g_s]
    int O4;                         // float4 chunks along out (= OG of the layer)
    const float* W; int w_s;
    int Kin; unsigned mK;           // true input width (+ nj_magic)
    float* gin; int gin_s;
    int nrows;
    const float* aprev; int a_s; int act_prev;    // activations feeding this layer (nullptr: network input)
    int drop; unsigned thr; float keep_scale, one_minus_p; const int* rk; unsigned tag_prev;
};

// gin[r][k] = (sum_o g[r][o] W[o][k]) * act'(aprev[r][k]) * dropout factor
template <int MR, bool WS>
NJ_HD void nj_tile_dx(const NjDx& L, int tid, int nt) {
    const int RG = (L.nrows + MR - 1) / MR;
    const int KG = (L.Kin + 3) >> 2;
    const int ntiles = RG * KG;
    if (tid >= ntiles) return;
    const unsigned mRG = nj_magic(RG);
    for (int tile = tid; tile < ntiles; tile += nt) {
        const int kg = nj_div(tile, RG, mRG), rg = tile - kg * RG;
        float acc[MR][4];
        nj_sp gp[MR];
#pragma unroll
        for (int i = 0; i < MR; ++i) {
            int r = rg + RG * i; if (r >= L.nrows) r = L.nrows - 1;
            gp[i] = nj_sp_of(L.g + (size_t)r * L.g_s);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        }
        NjWPtr<WS> wp;
        wp.set(L.W + 4 * kg);
#pragma unroll 2
        for (int o4 = 0; o4 < L.O4; ++o4) {
            nj_f4 g[MR], w[4];
#pragma unroll
            for (int i = 0; i < MR; ++i) g[i] = nj_sp_ld4(NJ_SP_ADD(gp[i], 4 * o4));
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) w[jj] = wp.ld((4 * o4 + jj) * L.w_s);
#pragma unroll
            for (int i = 0; i < MR; ++i) {
                acc[i][0] = fmaf(g[i].x, w[0].x, fmaf(g[i].y, w[1].x, fmaf(g[i].z, w[2].x, fmaf(g[i].w, w[3].x, acc[i][0]))));
                acc[i][1] = fmaf(g[i].x, w[0].y, fmaf(g[i].y, w[1].y, fmaf(g[i].z, w[2].y, fmaf(g[i].w, w[3].y, acc[i][1]))));
                acc[i][2] = fmaf(g[i].x, w[0].z, fmaf(g[i].y, w[1].z, fmaf(g[i].z, w[2].z, fmaf(g[i].w, w[3].z, acc[i][2]))));
                acc[i][3] = fmaf(g[i].x, w[0].w, fmaf(g[i].y, w[1].w, fmaf(g[i].z, w[2].w, fmaf(g[i].w, w[3].w, acc[i][3]))));
            }
        }
#pragma unroll
        for (int i = 0; i < MR; ++i) {
            const int r = rg + RG * i;
            if (r >= L.nrows) continue;
            unsigned lk = 0;
            if (L.drop && L.aprev) lk = nj_layer_key((unsigned)L.rk[r], L.tag_prev);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = 4 * kg + j;
                if (k >= L.Kin) continue;
                float v = acc[i][j];
                if (L.aprev) {
                    float a = L.aprev[(size_t)r * L.a_s + k];
                    if (L.drop) {
                        if (nj_keep(lk, (unsigned)k, L.thr)) { a *= L.one_minus_p; v *= L.keep_scale; }
                        else { v = 0.f; a = 0.f; }
                    }
                    if (L.act_prev == NJODE_ACT_TANH) v *= (1.f - a * a);
                    else if (L.act_prev == NJODE_ACT_RELU) v = a > 0.f ? v : 0.f;
                }
                L.gin[(size_t)r * L.gin_s + k] = v;
            }
        }
    }
}

// ---- row-loop variants for tiles of few rows (whole-path batches: 14 records per CTA for a PhysioNet batch of 2000,
// one at the reference's batch size of 50).  The 1x4 micro-tiles above leave most threads idle there and every thread
// re-reads its weights for every row.  Here a thread owns ONE output (forward) / ONE input column (dx) and a slice of
// the reduction dimension, loads that weight slice into registers ONCE per call and then walks the rows: per row only
// broadcast loads of the activation slice + independent FMAs, so consecutive rows pipeline.  The KS partial sums of an
// output meet in shared memory (scratch[kq][r][o]); a second phase, one thread per output element, sums them and
// applies the epilogue.  ----
#define NJ_RL_MAXROWS 4             // measured on B200: from ~8 rows up the 1x4 micro-tiles are faster again
#define NJ_RL_SCRATCH 4096          // floats

// forward: thread (kq, o), lane = o.  KC = float4 chunks of the reduction dimension per slice (template bound).
template <int KC>
NJ_HD void nj_rl_fwd1(const NjLin& L, int KS, int kc, float* scratch, int tid) {
    const int O = L.O;
    if (tid >= KS * O) return;
    const int kq = nj_div(tid, O, L.mO), o = tid - kq * O;
    const int c0 = kq * kc;
    int nc = L.K4 - c0; nc = nc > kc ? kc : nc;
    nj_f4 w[KC];
    const float* wp = L.W + (size_t)o * L.w_s + 4 * c0;
#pragma unroll
    for (int c = 0; c < KC; ++c) {
        if (c < nc) w[c] = nj_ld4(wp + 4 * c);
        else { w[c].x = 0.f; w[c].y = 0.f; w[c].z = 0.f; w[c].w = 0.f; }
    }
    nj_sp ap = nj_sp_of(L.in + 4 * c0);
    float* sp = scratch + (size_t)kq * L.nrows * O + o;
    for (int r = 0; r < L.nrows; ++r, ap = NJ_SP_ADD(ap, L.in_s)) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            if (c < nc) {
                const nj_f4 a = nj_sp_ld4(NJ_SP_ADD(ap, 4 * c));
                s0 = fmaf(a.x, w[c].x, s0); s1 = fmaf(a.y, w[c].y, s1);
                s2 = fmaf(a.z, w[c].z, s2); s3 = fmaf(a.w, w[c].w, s3);
            }
        }
        sp[(size_t)r * O] = (s0 + s1) + (s2 + s3);
    }
}

NJ_HD void nj_rl_fwd2(const NjLin& L, int KS, const float* scratch, int tid, int nt) {
    const int O = L.O, n = L.nrows * O;
    for (int idx = tid; idx < n; idx += nt) {
        const int r = nj_div(idx, O, L.mO), o = idx - r * O;
        float v = L.bias ? L.bias[o] : 0.f;
        for (int kq = 0; kq < KS; ++kq) v += scratch[(size_t)kq * n + idx];
        v = nj_act(v, L.act);
        if (L.drop) v = nj_keep(nj_layer_key((unsigned)L.rk[r], L.tag), (unsigned)o, L.thr) ? v * L.keep_scale : 0.f;
        L.out[(size_t)r * L.out_s + o] = v;
    }
}

// dx: thread (oq, k), lane = k.  OC = groups of 4 outputs per slice (template bound).
template <int OC>
NJ_HD void nj_rl_dx1(const NjDx& L, int OS, int oc, float* scratch, int tid) {
    const int K = L.Kin;
    if (tid >= OS * K) return;
    const int oq = nj_div(tid, K, L.mK), k = tid - oq * K;
    const int c0 = oq * oc;
    int nc = L.O4 - c0; nc = nc > oc ? oc : nc;
    float w[OC][4];
    const float* wp = L.W + (size_t)(4 * c0) * L.w_s + k;
#pragma unroll
    for (int c = 0; c < OC; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) w[c][i] = (c < nc) ? wp[(size_t)(4 * c + i) * L.w_s] : 0.f;
    nj_sp gp = nj_sp_of(L.g + 4 * c0);
    float* sp = scratch + (size_t)oq * L.nrows * K + k;
    for (int r = 0; r < L.nrows; ++r, gp = NJ_SP_ADD(gp, L.g_s)) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int c = 0; c < OC; ++c) {
            if (c < nc) {
                const nj_f4 g = nj_sp_ld4(NJ_SP_ADD(gp, 4 * c));
                s0 = fmaf(g.x, w[c][0], s0); s1 = fmaf(g.y, w[c][1], s1);
                s2 = fmaf(g.z, w[c][2], s2); s3 = fmaf(g.w, w[c][3], s3);
            }
        }
        sp[(size_t)r * K] = (s0 + s1) + (s2 + s3);
    }
}

NJ_HD void nj_rl_dx2(const NjDx& L, int OS, const float* scratch, int tid, int nt) {
    const int K = L.Kin, n = L.nrows * K;
    for (int idx = tid; idx < n; idx += nt) {
        const int r = nj_div(idx, K, L.mK), k = idx - r * K;
        float v = 0.f;
        for (int oq = 0; oq < OS; ++oq) v += scratch[(size_t)oq * n + idx];
        if (L.aprev) {
            float a = L.aprev[(size_t)r * L.a_s + k];
            if (L.drop) {
                if (nj_keep(nj_layer_key((unsigned)L.rk[r], L.tag_prev), (unsigned)k, L.thr)) { a *= L.one_minus_p; v *= L.keep_scale; }
                else { v = 0.f; a = 0.f; }
            }
            if (L.act_prev == NJODE_ACT_TANH) v *= (1.f - a * a);
            else if (L.act_prev == NJODE_ACT_RELU) v = a > 0.f ? v : 0.f;
        }
        L.gin[(size_t)r * L.gin_s + k] = v;
    }
}

// dW[o][k] += sum_r g[r][o] a[r][k] ; db[o] += sum_r g[r][o]   (thread-owned 4x4 blocks of the image)
NJ_HD void nj_tile_dw(const float* g, int g_s, int OG, const float* a, int a_s, int K4, int nrows,
                      float* dW, int w_s, float* db, int tid, int nt) {
    const int ntiles = OG * K4;
    if (tid >= ntiles) return;
    const unsigned mK4 = nj_magic(K4);
    for (int tile = tid; tile < ntiles; tile += nt) {
        const int og = nj_div(tile, K4, mK4), kg = tile - og * K4;
        float acc[4][4];
        float bs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        nj_sp gp = nj_sp_of(g + 4 * og), apx = nj_sp_of(a + 4 * kg);
#pragma unroll 2
        for (int r = 0; r < nrows; ++r) {
            const nj_f4 gv = nj_sp_ld4(gp);
            const nj_f4 av = nj_sp_ld4(apx);
            gp = NJ_SP_ADD(gp, g_s); apx = NJ_SP_ADD(apx, a_s);
            acc[0][0] = fmaf(gv.x, av.x, acc[0][0]); acc[0][1] = fmaf(gv.x, av.y, acc[0][1]);
            acc[0][2] = fmaf(gv.x, av.z, acc[0][2]); acc[0][3] = fmaf(gv.x, av.w, acc[0][3]);
            acc[1][0] = fmaf(gv.y, av.x, acc[1][0]); acc[1][1] = fmaf(gv.y, av.y, acc[1][1]);
            acc[1][2] = fmaf(gv.y, av.z, acc[1][2]); acc[1][3] = fmaf(gv.y, av.w, acc[1][3]);
            acc[2][0] = fmaf(gv.z, av.x, acc[2][0]); acc[2][1] = fmaf(gv.z, av.y, acc[2][1]);
            acc[2][2] = fmaf(gv.z, av.z, acc[2][2]); acc[2][3] = fmaf(gv.z, av.w, acc[2][3]);
            acc[3][0] = fmaf(gv.w, av.x, acc[3][0]); acc[3][1] = fmaf(gv.w, av.y, acc[3][1]);
            acc[3][2] = fmaf(gv.w, av.z, acc[3][2]); acc[3][3] = fmaf(gv.w, av.w, acc[3][3]);
            if (kg == 0) { bs[0] += gv.x; bs[1] += gv.y; bs[2] += gv.z; bs[3] += gv.w; }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float* p = dW + (size_t)(4 * og + i) * w_s + 4 * kg;
            nj_f4 v = nj_ld4(p);
            v.x += acc[i][0]; v.y += acc[i][1]; v.z += acc[i][2]; v.w += acc[i][3];
            nj_st4(p, v);
        }
        if (kg == 0 && db) {
            nj_f4 v = nj_ld4(db + 4 * og);
            v.x += bs[0]; v.y += bs[1]; v.z += bs[2]; v.w += bs[3];
            nj_st4(db + 4 * og, v);
        }
    }
}

NJ_HD int nj_pick_mr(int nrows, int groups, int nt) {
    if (((nrows + 3) >> 2) * groups >= (nt >> 1)) return 4;
    if (((nrows + 1) >> 1) * groups >= (nt >> 1)) return 2;
    return 1;
}

// ------------------------------------------------------------------------------------------------
// the CTA context: pointers into shared memory + launch arguments
// ------------------------------------------------------------------------------------------------
struct NjArgs {
    njode_batch_t b;
    const float* image;        // padded parameter image in global memory
    float* hT; float* row_loss; float* path_h; float* path_y;
    float* h_hist; float* h_before; float* y_after;      // forward: written; backward: read
    const float* grad_loss; const float* grad_hT;
    float* partials;           // [grid][img_floats] gradient partial images (backward)
    int* counter;              // tile counter of the segment kernels (zeroed before every launch)
    int n_tiles;
    int get_loss;
    // segment backward without anything saved by the forward pass ("recompute forward segments from the checkpoints at
    // the observation times"): per-CTA scratch [S][P_b][sH] for the h chain of the tile being reversed, or NULL
    float* scratch;
    // hidden activations of the ODE network at every Euler step, [S][B][act_nh][act_wp] (dropped units as -0.0f), or NULL:
    // written by the segment forward, read by the segment backward instead of recomputing the two hidden layers
    float* act_hist;
    int act_nh, act_wp;
};

struct NjCta {
    const NjCfg* c;
    const float* wimg;         // parameter image (shared or global)
    float* dimg;               // gradient image: shared-memory part [0, dimg_floats) ...
    float* gpart;              // ... and this CTA's partial image in global memory for the rest
    float *IN, *ACT, *OUT, *Hs, *LX, *XI, *YBJ, *YY, *XH, *EE, *GOUT, *GTMP, *GA, *GB, *GH, *GX, *GYBJ, *F;
    float *GI, *GHH, *KSB;
    int *I, *CTL;
    int tid0, nt;              // nt: threads; (tid comes from NJ_THREADS)
};

NJ_HD void nj_cta_bind(NjCta& t, const NjCfg& c, float* smem, bool bwd) {
    t.c = &c; t.nt = c.nt;
    t.IN = smem + c.o_IN; t.ACT = smem + c.o_ACT; t.OUT = smem + c.o_OUT; t.Hs = smem + c.o_H;
    t.LX = smem + c.o_LX; t.XI = smem + c.o_XI; t.YBJ = smem + c.o_YBJ; t.YY = smem + c.o_YY;
    t.XH = smem + c.o_XH; t.EE = smem + c.o_EE; t.F = smem + c.o_F;
    t.GI = smem + c.o_GI; t.GHH = smem + c.o_GHH; t.KSB = smem + c.o_KS;
    t.I = reinterpret_cast<int*>(smem + c.o_I); t.CTL = t.I + NJ_I_COUNT * c.P;
    t.GOUT = t.GTMP = t.GA = t.GB = t.GH = t.GX = t.GYBJ = nullptr;
    if (bwd) {
        t.GOUT = smem + c.o_GOUT; t.GTMP = smem + c.o_GTMP; t.GA = smem + c.o_GA; t.GB = smem + c.o_GB;
        t.GH = smem + c.o_GH; t.GX = smem + c.o_GX; t.GYBJ = smem + c.o_GYBJ;
    }
}

#define NJ_IU(t, slot, u) ((t).I[(slot) * (t).c->P + (u)])
#define NJ_FU(t, slot, u) ((t).F[(slot) * (t).c->P + (u)])

// residual connections of FFNN.forward (NJODE/models.py:268-276)
NJ_HD float nj_resid(const float* x, int in_sz, int out_sz, int c) {
    if (in_sz <= out_sz) return x[c % in_sz];
    const int mult = in_sz / out_sz;
    float s = 0.f;
    for (int k = 0; k < mult; ++k) s += x[k * out_sz + c];
    return s / (float)mult;
}
// d(residual)/d x[cp] contracted with g (width out_sz)
NJ_HD float nj_resid_bwd(const float* g, int in_sz, int out_sz, int cp) {
    if (in_sz <= out_sz) {
        float s = 0.f;
        for (int k = 0; k < out_sz / in_sz; ++k) s += g[k * in_sz + cp];
        return s;
    }
    return g[cp % out_sz] / (float)(in_sz / out_sz);
}

// MLP forward over rows [0, nrows) of IN -> OUT (raw output of the last Linear); hidden activations
// are kept in ACT[l].  skip_last: backward-side recompute that only needs the hidden activations.
NJ_HDN void nj_mlp_forward(NjCta& t, int netid, int nrows, bool skip_last) {
    const NjCfg& c = *t.c;
    const NjNet& N = c.net[netid];
    const float* in = t.IN; int in_s = c.sIN;
    for (int l = 0; l < N.n; ++l) {
        const bool last = (l == N.n - 1);
        if (last && skip_last) break;
        NjLin L;
        L.in = in; L.in_s = in_s; L.K4 = (N.dim[l] + 3) >> 2;
        L.W = t.wimg + N.w_img[l]; L.w_s = N.ks[l];
        L.bias = N.b_src[l] >= 0 ? t.wimg + N.b_img[l] : nullptr;
        L.O = N.dim[l + 1]; L.OG = N.og[l]; L.mO = N.m_out[l];
        L.out = last ? t.OUT : t.ACT + (size_t)l * c.P * c.sACT; L.out_s = last ? c.sOUT : c.sACT;
        L.nrows = nrows; L.act = last ? NJODE_ACT_NONE : N.act[l];
        L.drop = (!last) && c.has_drop; L.thr = c.thr; L.keep_scale = c.keep_scale;
        L.rk = t.I + NJ_I_RK * c.P; L.tag = (unsigned)(netid * 16 + l + 1);
        const int mr = nj_pick_mr(nrows, L.OG, t.nt);
        const int kc = N.rl_f_kc[l];
        const int ksl = nrows <= NJ_RL_MAXROWS ? N.rl_f_ns[l] : 0;
        if (ksl > 0) {
            NJ_THREADS(tid, t.nt) {
                if (kc <= 1) nj_rl_fwd1<1>(L, ksl, kc, t.KSB, tid);
                else if (kc <= 2) nj_rl_fwd1<2>(L, ksl, kc, t.KSB, tid);
                else if (kc <= 4) nj_rl_fwd1<4>(L, ksl, kc, t.KSB, tid);
                else nj_rl_fwd1<8>(L, ksl, kc, t.KSB, tid);
            }
            NJ_SYNC();
            NJ_THREADS(tid, t.nt) { nj_rl_fwd2(L, ksl, t.KSB, tid, t.nt); }
        } else {
            NJ_THREADS(tid, t.nt) {
                if (c.w_smem) {
                    if (mr == 4) nj_tile_fwd<4, true>(L, tid, t.nt);
                    else if (mr == 2) nj_tile_fwd<2, true>(L, tid, t.nt);
                    else nj_tile_fwd<1, true>(L, tid, t.nt);
                } else {
                    if (mr == 4) nj_tile_fwd<4, false>(L, tid, t.nt);
                    else if (mr == 2) nj_tile_fwd<2, false>(L, tid, t.nt);
                    else nj_tile_fwd<1, false>(L, tid, t.nt);
                }
            }
        }
        NJ_SYNC();
        in = L.out; in_s = L.out_s;
    }
}

// MLP backward: gradient wrt the raw output in GOUT (width = out dim); accumulates dW/db into the
// gradient image; returns the buffer holding the gradient wrt the network input (already multiplied
// by nothing: the caller applies the derivative of its own input transform), or nullptr.
NJ_HDN float* nj_mlp_backward(NjCta& t, int netid, int nrows, bool need_in_grad) {
    const NjCfg& c = *t.c;
    const NjNet& N = c.net[netid];
    const float* g = t.GOUT; int g_s = c.sOUT;
    float* nxt = t.GA;
    float* res = nullptr;
    for (int l = N.n - 1; l >= 0; --l) {
        const float* inp = l > 0 ? t.ACT + (size_t)(l - 1) * c.P * c.sACT : t.IN;
        const int inp_s = l > 0 ? c.sACT : c.sIN;
        const bool dx = (l > 0) || need_in_grad;
        NjDx D;
        D.g = g; D.g_s = g_s; D.O4 = N.og[l]; D.W = t.wimg + N.w_img[l]; D.w_s = N.ks[l];
        D.Kin = N.dim[l]; D.mK = N.m_in[l]; D.gin = nxt; D.gin_s = c.sG; D.nrows = nrows;
        D.aprev = l > 0 ? inp : nullptr; D.a_s = inp_s; D.act_prev = l > 0 ? N.act[l - 1] : NJODE_ACT_NONE;
        D.drop = c.has_drop; D.thr = c.thr; D.keep_scale = c.keep_scale; D.one_minus_p = c.one_minus_p;
        D.rk = t.I + NJ_I_RK * c.P; D.tag_prev = (unsigned)(netid * 16 + l);
        const int mr = nj_pick_mr(nrows, (N.dim[l] + 3) >> 2, t.nt);
        float* dbase = (N.b_img[l] + N.rp[l] <= c.dimg_floats) ? t.dimg : t.gpart;
        float* dW = dbase + N.w_img[l];
        float* db = N.b_src[l] >= 0 ? dbase + N.b_img[l] : nullptr;
        // small tiles: the dW micro-tiles take the first threads, the row-loop dx slices the rest of the CTA
        const int dx_t0 = N.rl_d_t0[l], oc = N.rl_d_oc[l];
        const int osl = (dx && nrows <= NJ_RL_MAXROWS) ? N.rl_d_ns[l] : 0;
        NJ_THREADS(tid, t.nt) {
            nj_tile_dw(g, g_s, N.og[l], inp, inp_s, (N.dim[l] + 3) >> 2, nrows, dW, N.ks[l], db, tid, t.nt);
            if (dx) {
                if (osl > 0) {
                    const int tid2 = tid - dx_t0;
                    if (tid2 < 0) { }
                    else if (oc <= 1) nj_rl_dx1<1>(D, osl, oc, t.KSB, tid2);
                    else if (oc <= 2) nj_rl_dx1<2>(D, osl, oc, t.KSB, tid2);
                    else if (oc <= 4) nj_rl_dx1<4>(D, osl, oc, t.KSB, tid2);
                    else nj_rl_dx1<8>(D, osl, oc, t.KSB, tid2);
                }
                else if (c.w_smem) {
                    if (mr == 4) nj_tile_dx<4, true>(D, tid, t.nt);
                    else if (mr == 2) nj_tile_dx<2, true>(D, tid, t.nt);
                    else nj_tile_dx<1, true>(D, tid, t.nt);
                } else {
                    if (mr == 4) nj_tile_dx<4, false>(D, tid, t.nt);
                    else if (mr == 2) nj_tile_dx<2, false>(D, tid, t.nt);
                    else nj_tile_dx<1, false>(D, tid, t.nt);
                }
            }
        }
        NJ_SYNC();
        if (osl > 0) {
            NJ_THREADS(tid, t.nt) { nj_rl_dx2(D, osl, t.KSB, tid, t.nt); }
            NJ_SYNC();
        }
        if (dx) { res = nxt; g = nxt; g_s = c.sG; nxt = (nxt == t.GA) ? t.GB : t.GA; }
    }
    return need_in_grad ? res : nullptr;
}

// ------------------------------------------------------------------------------------------------
// shared helpers of forward and backward
// ------------------------------------------------------------------------------------------------
NJ_HDN void nj_load_units(NjCta& t, const NjArgs& a, int tile, bool reverse) {
    const NjCfg& c = *t.c;
    const int u0 = tile * c.P;
    NJ_THREADS(tid, t.nt) {
        if (tid < c.P) {
            const int u = u0 + tid;
            if (u < a.b.n_units) {
                const int32_t* dsc = a.b.unit_desc + (size_t)u * 6;
                NJ_IU(t, NJ_I_PATH, tid) = dsc[0]; NJ_IU(t, NJ_I_S0, tid) = dsc[1];
                NJ_IU(t, NJ_I_LEN, tid) = dsc[2] - dsc[1];
                NJ_IU(t, NJ_I_C0, tid) = dsc[3]; NJ_IU(t, NJ_I_C1, tid) = dsc[4];
                NJ_IU(t, NJ_I_CUR, tid) = reverse ? dsc[4] : dsc[3];
                const int sc = dsc[5];      // (start_row + 1) | NJODE_UNIT_WRITES_HT
                NJ_IU(t, NJ_I_FLAG, tid) = (sc & NJODE_UNIT_WRITES_HT) ? 1 : 0;
                NJ_IU(t, NJ_I_START, tid) = (sc & ~NJODE_UNIT_WRITES_HT) - 1;
            } else {
                NJ_IU(t, NJ_I_PATH, tid) = -1; NJ_IU(t, NJ_I_S0, tid) = 0; NJ_IU(t, NJ_I_LEN, tid) = -1;
                NJ_IU(t, NJ_I_C0, tid) = 0; NJ_IU(t, NJ_I_C1, tid) = 0; NJ_IU(t, NJ_I_CUR, tid) = 0;
                NJ_IU(t, NJ_I_FLAG, tid) = 0; NJ_IU(t, NJ_I_START, tid) = -1;
            }
        }
    }
    NJ_SYNC();
    NJ_THREADS(tid, t.nt) {
        if (tid == 0) {
            int m = 0;
            for (int u = 0; u < c.P; ++u) m = NJ_IU(t, NJ_I_LEN, u) > m ? NJ_IU(t, NJ_I_LEN, u) : m;
            t.CTL[NJ_CTL_MAXLEN] = m;
        }
    }
    NJ_SYNC();
}

// (last_X, tau) valid after the jump of row `prev` (or at the path start when prev < 0)
NJ_HD void nj_state_from_row(NjCta& t, const NjArgs& a, int u, int c_, int prev) {
    const NjCfg& c = *t.c;
    const int p = NJ_IU(t, NJ_I_PATH, u);
    float v;
    if (prev < 0) v = NJ_LDG(a.b.start_X + (size_t)p * c.d + c_);
    else if (c.masked) v = a.y_after[(size_t)prev * c.dout + c_];
    else v = NJ_LDG(a.b.X + (size_t)prev * c.d + c_);
    t.LX[u * c.sD + c_] = v;
    if (c_ == 0) NJ_FU(t, NJ_F_TAU, u) = prev < 0 ? 0.f : NJ_LDG(a.b.jump_tau + NJ_LDG(a.b.row_jump + prev));
}

// ODE network input row (ODEFunc.forward, NJODE/models.py:192-197) for unit u at Euler step k
NJ_HD void nj_build_ode_input(NjCta& t, int u, int c_, float hval, float tcur) {
    const NjCfg& c = *t.c;
    float* row = t.IN + (size_t)u * c.sIN;
    const float tau = NJ_FU(t, NJ_F_TAU, u);
    if (c_ < c.d) row[c_] = nj_tanh(t.LX[u * c.sD + c_]);
    else if (c_ < c.d + c.H) row[c_] = nj_tanh(hval);
    else if (c_ == c.d + c.H) row[c_] = tau;
    else if (c_ == c.d + c.H + 1) row[c_] = tcur - tau;
    else row[c_] = tau + (tcur - tau);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
NJ_HDN void nj_record(NjCta& t, const NjArgs& a, int nu, int e, unsigned event_key) {
    const NjCfg& c = *t.c;
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nu * c.H; idx += t.nt) {
            const int u = nj_div(idx, c.H, c.m_H), c_ = idx - u * c.H;
            const int p = NJ_IU(t, NJ_I_PATH, u);
            const float h = t.Hs[u * c.sH + c_];
            t.IN[(size_t)u * c.sIN + c_] = nj_tanh(h);
            a.path_h[((size_t)e * a.b.B + p) * c.H + c_] = h;
            if (c_ == 0) NJ_IU(t, NJ_I_RK, u) = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(p + a.b.path_id_offset), event_key);
        }
    }
    NJ_SYNC();
    nj_mlp_forward(t, NJODE_NET_RO, nu, false);
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nu * c.dout; idx += t.nt) {
            const int u = nj_div(idx, c.dout, c.m_dout), c_ = idx - u * c.dout;
            const int p = NJ_IU(t, NJ_I_PATH, u);
            float y = t.OUT[u * c.sOUT + c_];
            if (c.residual) y += nj_resid(t.Hs + u * c.sH, c.H, c.dout, c_);
            a.path_y[((size_t)e * a.b.B + p) * c.dout + c_] = y;
        }
    }
    NJ_SYNC();
}

// lockstep iteration (relative Euler-step index) at which each unit meets its next pending jump, and the first such
// iteration over the tile (forward: min, reverse: max; none: INT_MAX / -1) in CTL[NJ_CTL_NEXT].  Refreshed when a
// tile is loaded and after every jump, so the march only looks for jumps (nj_find_jumps: compaction, dependent global
// loads, two barriers) at the iterations that have one.
NJ_HDN void nj_next_jump(NjCta& t, const NjArgs& a, int j, bool reverse) {
    const NjCfg& c = *t.c;
    NJ_THREADS(tid, t.nt) {
        if (tid < c.P) {
            int nxt = reverse ? -1 : 0x7FFFFFFF;
            const int len = NJ_IU(t, NJ_I_LEN, tid);
            if (len >= 0) {
                const int cur = NJ_IU(t, NJ_I_CUR, tid);
                const int idx = reverse ? cur - 1 : cur;
                const bool has = reverse ? (idx >= NJ_IU(t, NJ_I_C0, tid)) : (idx < NJ_IU(t, NJ_I_C1, tid));
                if (has) {
                    const int rel = NJ_LDG(a.b.jump_step + NJ_LDG(a.b.row_jump + NJ_LDG(a.b.path_rows + idx))) - NJ_IU(t, NJ_I_S0, tid);
                    if (rel >= 0 && rel <= len && (reverse ? rel <= j : rel >= j)) nxt = rel;
                }
            }
            NJ_IU(t, NJ_I_NEXT, tid) = nxt;
        }
    }
    NJ_SYNC();
    NJ_THREADS(tid, t.nt) {
        if (tid == 0) {
            int m = reverse ? -1 : 0x7FFFFFFF;
            for (int u = 0; u < c.P; ++u) {
                const int v = NJ_IU(t, NJ_I_NEXT, u);
                m = reverse ? (v > m ? v : m) : (v < m ? v : m);
            }
            t.CTL[NJ_CTL_NEXT] = m;
        }
    }
    NJ_SYNC();
}

// find the units that jump now; compacts them into JMAP/JROW/JJMP, CTL[NJ]=count.
//   j: lockstep iteration; only_jump >= 0: restrict to that observation-time index (return_path)
NJ_HDN void nj_find_jumps(NjCta& t, const NjArgs& a, int j, int only_jump, bool reverse) {
    const NjCfg& c = *t.c;
    NJ_THREADS(tid, t.nt) {
        if (tid < c.P) {
            int pend = -1;
            const int len = NJ_IU(t, NJ_I_LEN, tid);
            if (len >= 0 && j <= len) {
                const int cur = NJ_IU(t, NJ_I_CUR, tid);
                const int idx = reverse ? cur - 1 : cur;
                const bool has = reverse ? (idx >= NJ_IU(t, NJ_I_C0, tid)) : (idx < NJ_IU(t, NJ_I_C1, tid));
                if (has) {
                    const int r = NJ_LDG(a.b.path_rows + idx);
                    const int i = NJ_LDG(a.b.row_jump + r);
                    if ((j < 0 || NJ_LDG(a.b.jump_step + i) == NJ_IU(t, NJ_I_S0, tid) + j) && (only_jump < 0 || i == only_jump))
                        pend = r;
                }
            }
            NJ_IU(t, NJ_I_PEND, tid) = pend;
        }
    }
    NJ_SYNC();
    NJ_THREADS(tid, t.nt) {
        if (tid == 0) {
            int n = 0;
            for (int u = 0; u < c.P; ++u) {
                const int r = NJ_IU(t, NJ_I_PEND, u);
                if (r >= 0) { NJ_IU(t, NJ_I_JMAP, n) = u; NJ_IU(t, NJ_I_JROW, n) = r; NJ_IU(t, NJ_I_JJMP, n) = NJ_LDG(a.b.row_jump + r); ++n; }
            }
            t.CTL[NJ_CTL_NJ] = n;
        }
    }
    NJ_SYNC();
}

NJ_HD void nj_set_jump_keys(NjCta& t, const NjArgs& a, int nj, int which, int tid) {
    const NjCfg& c = *t.c;
    if (tid < nj) {
        const int u = NJ_IU(t, NJ_I_JMAP, tid);
        const unsigned ev = NJ_EVENT_JUMP_BASE + 3u * (unsigned)NJ_IU(t, NJ_I_JJMP, tid) + (unsigned)which;
        NJ_IU(t, NJ_I_RK, tid) = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(NJ_IU(t, NJ_I_PATH, u) + a.b.path_id_offset), ev);
    }
}

NJ_HD float nj_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// GRU jump (use_rnn=True; NJODE/models.py:202-217 -> torch.nn.GRUCell, gate order r, z, n) for the compacted
// rows [0, nj); h_old = XH[jr]:
//   gi = W_ih tanh(X_obs) + b_ih ; gh = W_hh tanh(h_old) + b_hh
//   r = sig(gi_r + gh_r), z = sig(gi_z + gh_z), n = tanh(gi_n + r gh_n), h' = (1 - z) n + z tanh(h_old)
// leaves (r, z, n) in GI, tanh(h_old) in GHH[0, H), gh_n in GHH[2H, 3H), h' in EE, X_obs in XI
NJ_HDN void nj_gru_forward(NjCta& t, const NjArgs& a, int nj) {
    const NjCfg& c = *t.c;
    const int H = c.H;
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.d; idx += t.nt) {
            const int jr = nj_div(idx, c.d, c.m_d), c_ = idx - jr * c.d;
            const float x = NJ_LDG(a.b.X + (size_t)NJ_IU(t, NJ_I_JROW, jr) * c.d + c_);
            t.XI[jr * c.sD + c_] = x;
            t.IN[(size_t)jr * c.sIN + c_] = nj_tanh(x);
        }
    }
    NJ_SYNC();
    nj_mlp_forward(t, NJODE_NET_GRU_IH, nj, false);
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * 3 * H; idx += t.nt) {
            const int jr = nj_div(idx, (3 * H), c.m_3H), k = idx - jr * (3 * H);
            t.GI[jr * c.s3H + k] = t.OUT[jr * c.sOUT + k];
        }
        for (int idx = tid; idx < nj * H; idx += t.nt) {
            const int jr = nj_div(idx, H, c.m_H), c_ = idx - jr * H;
            t.IN[(size_t)jr * c.sIN + c_] = nj_tanh(t.XH[jr * c.sH + c_]);
        }
    }
    NJ_SYNC();
    nj_mlp_forward(t, NJODE_NET_GRU_HH, nj, false);
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * H; idx += t.nt) {
            const int jr = nj_div(idx, H, c.m_H), c_ = idx - jr * H;
            float* gi = t.GI + jr * c.s3H;
            float* gh = t.GHH + jr * c.s3H;
            const float* o = t.OUT + jr * c.sOUT;
            const float hh = t.IN[(size_t)jr * c.sIN + c_];
            const float r = nj_sigmoid(gi[c_] + o[c_]);
            const float z = nj_sigmoid(gi[H + c_] + o[H + c_]);
            const float ghn = o[2 * H + c_];
            const float n = nj_tanh(fmaf(r, ghn, gi[2 * H + c_]));
            gi[c_] = r; gi[H + c_] = z; gi[2 * H + c_] = n;
            gh[c_] = hh; gh[2 * H + c_] = ghn;
            t.EE[jr * c.sH + c_] = fmaf(z, hh - n, n);
        }
    }
    NJ_SYNC();
}

// the jump of NJODE/models.py:449-489 for the compacted rows [0, nj)
NJ_HDN void nj_jump_forward(NjCta& t, const NjArgs& a, int nj) {
    const NjCfg& c = *t.c;
    // (a) readout input = tanh(h before the jump)
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.H; idx += t.nt) {
            const int jr = nj_div(idx, c.H, c.m_H), c_ = idx - jr * c.H;
            const int u = NJ_IU(t, NJ_I_JMAP, jr);
            const float h = t.Hs[u * c.sH + c_];
            t.IN[(size_t)jr * c.sIN + c_] = nj_tanh(h);
            t.XH[jr * c.sH + c_] = h;
            if (a.h_before) a.h_before[(size_t)NJ_IU(t, NJ_I_JROW, jr) * c.H + c_] = h;
        }
        nj_set_jump_keys(t, a, nj, 0, tid);
    }
    NJ_SYNC();
    nj_mlp_forward(t, NJODE_NET_RO, nj, false);
    // (c) Y_bj, imputation, encoder input
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.dout; idx += t.nt) {
            const int jr = nj_div(idx, c.dout, c.m_dout), c_ = idx - jr * c.dout;
            float y = t.OUT[jr * c.sOUT + c_];
            if (c.residual) y += nj_resid(t.XH + jr * c.sH, c.H, c.dout, c_);
            t.YBJ[jr * c.sDO + c_] = y;
        }
    }
    NJ_SYNC();
    if (c.use_rnn) {
        // (d') h[i_obs] = GRUCell(tanh(X_obs), tanh(h[i_obs]))  (NJODE/models.py:460-461)
        nj_gru_forward(t, a, nj);
        NJ_THREADS(tid, t.nt) {
            for (int idx = tid; idx < nj * c.H; idx += t.nt) {
                const int jr = nj_div(idx, c.H, c.m_H), c_ = idx - jr * c.H;
                const float e = t.EE[jr * c.sH + c_];
                t.Hs[NJ_IU(t, NJ_I_JMAP, jr) * c.sH + c_] = e;
                t.IN[(size_t)jr * c.sIN + c_] = nj_tanh(e);
            }
            nj_set_jump_keys(t, a, nj, 2, tid);
        }
        NJ_SYNC();
    } else {
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.d; idx += t.nt) {
            const int jr = nj_div(idx, c.d, c.m_d), c_ = idx - jr * c.d;
            const int r = NJ_IU(t, NJ_I_JROW, jr);
            float x = NJ_LDG(a.b.X + (size_t)r * c.d + c_);
            if (c.masked) {
                const float m = NJ_LDG(a.b.M + (size_t)r * c.d + c_);
                x = x * m + (1.f - m) * t.YBJ[jr * c.sDO + c_];
                t.IN[(size_t)jr * c.sIN + c.d + c_] = m;
            }
            t.XI[jr * c.sD + c_] = x;
            t.IN[(size_t)jr * c.sIN + c_] = nj_tanh(x);
        }
        nj_set_jump_keys(t, a, nj, 1, tid);
    }
    NJ_SYNC();
    nj_mlp_forward(t, NJODE_NET_ENC, nj, false);
    // (e) new hidden state, readout input
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.H; idx += t.nt) {
            const int jr = nj_div(idx, c.H, c.m_H), c_ = idx - jr * c.H;
            const int u = NJ_IU(t, NJ_I_JMAP, jr);
            float e = t.OUT[jr * c.sOUT + c_];
            if (c.residual) e += nj_resid(t.XI + jr * c.sD, c.d, c.H, c_);
            t.Hs[u * c.sH + c_] = e;
            t.IN[(size_t)jr * c.sIN + c_] = nj_tanh(e);
        }
        nj_set_jump_keys(t, a, nj, 2, tid);
    }
    NJ_SYNC();
    }
    nj_mlp_forward(t, NJODE_NET_RO, nj, false);
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.dout; idx += t.nt) {
            const int jr = nj_div(idx, c.dout, c.m_dout), c_ = idx - jr * c.dout;
            const int u = NJ_IU(t, NJ_I_JMAP, jr);
            float y = t.OUT[jr * c.sOUT + c_];
            if (c.residual) y += nj_resid(t.Hs + u * c.sH, c.H, c.dout, c_);
            t.YY[jr * c.sDO + c_] = y;
            if (a.y_after) a.y_after[(size_t)NJ_IU(t, NJ_I_JROW, jr) * c.dout + c_] = y;
        }
    }
    NJ_SYNC();
    // (g) loss term of the row, last_X / tau update, cursor advance
    NJ_THREADS(tid, t.nt) {
        if (tid < nj) {
            const int jr = tid, u = NJ_IU(t, NJ_I_JMAP, jr), r = NJ_IU(t, NJ_I_JROW, jr);
            if (a.get_loss) {
                float sa = 0.f, sb = 0.f;
                for (int c_ = 0; c_ < c.dout; ++c_) {
                    const float x = NJ_LDG(a.b.X + (size_t)r * c.d + c_);
                    const float m = c.masked ? NJ_LDG(a.b.M + (size_t)r * c.d + c_) : 1.f;
                    const float y = t.YY[jr * c.sDO + c_], yb = t.YBJ[jr * c.sDO + c_];
                    const float da = x - y, db = (c.loss_kind == NJODE_LOSS_STANDARD) ? (yb - y) : (yb - x);
                    sa = fmaf(m * da, da, sa); sb = fmaf(m * db, db, sb);
                }
                const float ra = sqrtf(sa + 1e-10f), rb = sqrtf(sb + 1e-10f);
                const float s = (c.loss_kind == NJODE_LOSS_STANDARD) ? (2.f * c.w * ra + 2.f * (1.f - c.w) * rb)
                                                                     : (c.w * ra + (1.f - c.w) * rb);
                a.row_loss[r] = s * s / NJ_LDG(a.b.n_obs_ot + NJ_IU(t, NJ_I_PATH, u));
            }
            NJ_FU(t, NJ_F_TAU, u) = NJ_LDG(a.b.jump_tau + NJ_IU(t, NJ_I_JJMP, jr));
            NJ_IU(t, NJ_I_CUR, u) += 1;
        }
        for (int idx = tid; idx < nj * c.d; idx += t.nt) {
            const int jr = nj_div(idx, c.d, c.m_d), c_ = idx - jr * c.d;
            const int u = NJ_IU(t, NJ_I_JMAP, jr);
            t.LX[u * c.sD + c_] = c.masked ? t.YY[jr * c.sDO + c_] : NJ_LDG(a.b.X + (size_t)NJ_IU(t, NJ_I_JROW, jr) * c.d + c_);
        }
    }
    NJ_SYNC();
}

NJ_HD void nj_zero(float* p, int n, int nt) {
    NJ_THREADS(tid, nt) { for (int i = tid; i < n; i += nt) p[i] = 0.f; }
}

NJ_HD void nj_cta_forward(const NjCfg& c, const NjArgs& a, float* smem, int cta, int ncta) {
    NjCta t;
    nj_cta_bind(t, c, smem, false);
    if (c.w_smem) {
        float* simg = smem + c.o_img;
        NJ_THREADS(tid, t.nt) { for (int i = tid; i < c.img_floats / 4; i += t.nt) nj_st4(simg + 4 * i, nj_ld4(a.image + 4 * i)); }
        t.wimg = simg;
    } else t.wimg = a.image;
    t.dimg = nullptr;
    nj_zero(smem + c.o_IN, c.o_I - c.o_IN, t.nt);
    NJ_SYNC();
    const bool rec = a.b.E > 0;
    // whole-path units jump at a few of their thousands of lockstep iterations: look for jumps only where the tile
    // has one (nj_next_jump).  Segment units jump at almost every iteration of a tile: look every time.
    const bool gate = !rec && a.b.unit_kind == 0;
    // segment units end with their one jump and nothing follows it: all jumps of a tile are applied together after the
    // march (one dense batch of rows through readout / encoder / readout) instead of one sparse event per unit
    const bool defer = !rec && a.b.unit_kind == 1;
    for (int tile = cta; tile < a.n_tiles; tile += ncta) {
        nj_load_units(t, a, tile, false);
        if (gate) nj_next_jump(t, a, 0, false);
        const int nu = (a.b.n_units - tile * c.P) < c.P ? (a.b.n_units - tile * c.P) : c.P;
        const int maxlen = t.CTL[NJ_CTL_MAXLEN];
        // ---- start: h = encoder(start value)  (NJODE/models.py:411-419) ----
        NJ_THREADS(tid, t.nt) {
            for (int idx = tid; idx < nu * c.d; idx += t.nt) {
                const int u = nj_div(idx, c.d, c.m_d), c_ = idx - u * c.d;
                nj_state_from_row(t, a, u, c_, NJ_IU(t, NJ_I_START, u));
                // the start value of a segment is the *observation* X[row] (non-masked only)
                const int sr = NJ_IU(t, NJ_I_START, u);
                const float x = sr < 0 ? NJ_LDG(a.b.start_X + (size_t)NJ_IU(t, NJ_I_PATH, u) * c.d + c_)
                                       : NJ_LDG(a.b.X + (size_t)sr * c.d + c_);
                t.XI[u * c.sD + c_] = x;
                t.IN[(size_t)u * c.sIN + c_] = nj_tanh(x);
                if (c.masked) t.IN[(size_t)u * c.sIN + c.d + c_] = 0.f;
            }
            if (tid < nu) {
                const int sr = NJ_IU(t, NJ_I_START, tid);
                const unsigned ev = sr < 0 ? NJ_EVENT_INIT : NJ_EVENT_JUMP_BASE + 3u * (unsigned)NJ_LDG(a.b.row_jump + sr) + 1u;
                NJ_IU(t, NJ_I_RK, tid) = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(NJ_IU(t, NJ_I_PATH, tid) + a.b.path_id_offset), ev);
            }
        }
        NJ_SYNC();
        nj_mlp_forward(t, NJODE_NET_ENC, nu, false);
        NJ_THREADS(tid, t.nt) {
            for (int idx = tid; idx < nu * c.H; idx += t.nt) {
                const int u = nj_div(idx, c.H, c.m_H), c_ = idx - u * c.H;
                float e = t.OUT[u * c.sOUT + c_];
                if (c.residual) e += nj_resid(t.XI + u * c.sD, c.d, c.H, c_);
                t.Hs[u * c.sH + c_] = e;
            }
        }
        NJ_SYNC();
        if (rec) nj_record(t, a, nu, 0, NJ_EVENT_INIT);
        int gi = 0;    // global jump cursor (return_path: all units are whole paths with s0 = 0)
        for (int j = 0; j <= maxlen; ++j) {
            // ---- jumps that fall before Euler step s0 + j ----
            for (; !defer;) {
                int only = -1;
                if (rec) {
                    if (!(gi < a.b.K && NJ_LDG(a.b.jump_step + gi) == j)) break;
                    only = gi;
                } else if (gate && t.CTL[NJ_CTL_NEXT] != j) break;  // no unit of the tile jumps before step s0 + j
                nj_find_jumps(t, a, j, only, false);
                const int nj = t.CTL[NJ_CTL_NJ];
                if (nj > 0) nj_jump_forward(t, a, nj);
                if (rec) { nj_record(t, a, nu, NJ_LDG(a.b.jump_event + gi), NJ_EVENT_JUMP_BASE + 3u * (unsigned)gi + 2u); ++gi; }
                else {
                    if (nj == 0) break;
                    if (gate) nj_next_jump(t, a, j, false);
                }
                NJ_SYNC();
            }
            if (j == maxlen) break;
            // ---- Euler step k = s0 + j of every unit that still has one  (NJODE/models.py:369-377) ----
            NJ_THREADS(tid, t.nt) {
                for (int idx = tid; idx < nu * c.inf; idx += t.nt) {
                    const int u = nj_div(idx, c.inf, c.m_inf), c_ = idx - u * c.inf;
                    const bool active = j < NJ_IU(t, NJ_I_LEN, u);
                    const int k = NJ_IU(t, NJ_I_S0, u) + j;
                    float hval = 0.f;
                    if (c_ >= c.d && c_ < c.d + c.H) {
                        hval = t.Hs[u * c.sH + c_ - c.d];
                        if (active && a.h_hist) a.h_hist[((size_t)k * a.b.B + NJ_IU(t, NJ_I_PATH, u)) * c.H + c_ - c.d] = hval;
                    }
                    nj_build_ode_input(t, u, c_, hval, active ? NJ_LDG(a.b.step_t + k) : 0.f);
                    if (c_ == 0) {
                        NJ_FU(t, NJ_F_DT, u) = active ? NJ_LDG(a.b.step_dt + k) : 0.f;
                        NJ_IU(t, NJ_I_RK, u) = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(NJ_IU(t, NJ_I_PATH, u) + a.b.path_id_offset), (unsigned)k);
                    }
                }
            }
            NJ_SYNC();
            nj_mlp_forward(t, NJODE_NET_ODE, nu, false);
            NJ_THREADS(tid, t.nt) {
                for (int idx = tid; idx < nu * c.H; idx += t.nt) {
                    const int u = nj_div(idx, c.H, c.m_H), c_ = idx - u * c.H;
                    if (j < NJ_IU(t, NJ_I_LEN, u))
                        t.Hs[u * c.sH + c_] = fmaf(NJ_FU(t, NJ_F_DT, u), t.OUT[u * c.sOUT + c_], t.Hs[u * c.sH + c_]);
                }
            }
            NJ_SYNC();
            if (rec) nj_record(t, a, nu, NJ_LDG(a.b.step_event + j), NJ_EVENT_PATH_RO_BASE + (unsigned)j);
        }
        if (defer) {
            nj_find_jumps(t, a, -1, -1, false);
            const int nj = t.CTL[NJ_CTL_NJ];
            if (nj > 0) nj_jump_forward(t, a, nj);
            NJ_SYNC();
        }
        // ---- hT ----
        NJ_THREADS(tid, t.nt) {
            for (int idx = tid; idx < nu * c.H; idx += t.nt) {
                const int u = nj_div(idx, c.H, c.m_H), c_ = idx - u * c.H;
                if (NJ_IU(t, NJ_I_FLAG, u)) a.hT[(size_t)NJ_IU(t, NJ_I_PATH, u) * c.H + c_] = t.Hs[u * c.sH + c_];
            }
        }
        NJ_SYNC();
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
NJ_HDN void nj_reload_state(NjCta& t, const NjArgs& a, int nu) {
    const NjCfg& c = *t.c;
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nu * c.d; idx += t.nt) {
            const int u = nj_div(idx, c.d, c.m_d), c_ = idx - u * c.d;
            const int cur = NJ_IU(t, NJ_I_CUR, u);
            const int p = NJ_IU(t, NJ_I_PATH, u);
            const int prev = (cur - 1 >= NJ_LDG(a.b.path_ptr + p)) ? NJ_LDG(a.b.path_rows + cur - 1) : -1;
            nj_state_from_row(t, a, u, c_, prev);
        }
    }
    NJ_SYNC();
}

NJ_HDN void nj_jump_backward(NjCta& t, const NjArgs& a, int nj) {
    const NjCfg& c = *t.c;
    const float gl = NJ_LDG(a.grad_loss);
    // 1. readout at h_before -> Y_bj
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.H; idx += t.nt) {
            const int jr = nj_div(idx, c.H, c.m_H), c_ = idx - jr * c.H;
            const float h = a.h_before[(size_t)NJ_IU(t, NJ_I_JROW, jr) * c.H + c_];
            t.XH[jr * c.sH + c_] = h;
            t.IN[(size_t)jr * c.sIN + c_] = nj_tanh(h);
        }
        nj_set_jump_keys(t, a, nj, 0, tid);
    }
    NJ_SYNC();
    nj_mlp_forward(t, NJODE_NET_RO, nj, false);
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.dout; idx += t.nt) {
            const int jr = nj_div(idx, c.dout, c.m_dout), c_ = idx - jr * c.dout;
            float y = t.OUT[jr * c.sOUT + c_];
            if (c.residual) y += nj_resid(t.XH + jr * c.sH, c.H, c.dout, c_);
            t.YBJ[jr * c.sDO + c_] = y;
        }
    }
    NJ_SYNC();
    // 2. encoder at the (imputed) observation -> E   (use_rnn: the GRU cell at (X_obs, h_before) -> E)
    if (c.use_rnn) {
        nj_gru_forward(t, a, nj);
        NJ_THREADS(tid, t.nt) {
            for (int idx = tid; idx < nj * c.H; idx += t.nt) {
                const int jr = nj_div(idx, c.H, c.m_H), c_ = idx - jr * c.H;
                t.IN[(size_t)jr * c.sIN + c_] = nj_tanh(t.EE[jr * c.sH + c_]);
            }
            nj_set_jump_keys(t, a, nj, 2, tid);
        }
        NJ_SYNC();
    } else {
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.d; idx += t.nt) {
            const int jr = nj_div(idx, c.d, c.m_d), c_ = idx - jr * c.d;
            const int r = NJ_IU(t, NJ_I_JROW, jr);
            float x = NJ_LDG(a.b.X + (size_t)r * c.d + c_);
            if (c.masked) {
                const float m = NJ_LDG(a.b.M + (size_t)r * c.d + c_);
                x = x * m + (1.f - m) * t.YBJ[jr * c.sDO + c_];
                t.IN[(size_t)jr * c.sIN + c.d + c_] = m;
            }
            t.XI[jr * c.sD + c_] = x;
            t.IN[(size_t)jr * c.sIN + c_] = nj_tanh(x);
        }
        nj_set_jump_keys(t, a, nj, 1, tid);
    }
    NJ_SYNC();
    nj_mlp_forward(t, NJODE_NET_ENC, nj, false);
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.H; idx += t.nt) {
            const int jr = nj_div(idx, c.H, c.m_H), c_ = idx - jr * c.H;
            float e = t.OUT[jr * c.sOUT + c_];
            if (c.residual) e += nj_resid(t.XI + jr * c.sD, c.d, c.H, c_);
            t.EE[jr * c.sH + c_] = e;
            t.IN[(size_t)jr * c.sIN + c_] = nj_tanh(e);
        }
        nj_set_jump_keys(t, a, nj, 2, tid);
    }
    NJ_SYNC();
    }
    // 3. readout at E -> Y (activations kept for its backward)
    nj_mlp_forward(t, NJODE_NET_RO, nj, false);
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.dout; idx += t.nt) {
            const int jr = nj_div(idx, c.dout, c.m_dout), c_ = idx - jr * c.dout;
            float y = t.OUT[jr * c.sOUT + c_];
            if (c.residual) y += nj_resid(t.EE + jr * c.sH, c.H, c.dout, c_);
            t.YY[jr * c.sDO + c_] = y;
        }
    }
    NJ_SYNC();
    // loss derivative coefficients per row (compute_loss / compute_loss_2, NJODE/models.py:71-126)
    NJ_THREADS(tid, t.nt) {
        if (tid < nj) {
            const int jr = tid, u = NJ_IU(t, NJ_I_JMAP, jr), r = NJ_IU(t, NJ_I_JROW, jr);
            float sa = 0.f, sb = 0.f;
            for (int c_ = 0; c_ < c.dout; ++c_) {
                const float x = NJ_LDG(a.b.X + (size_t)r * c.d + c_);
                const float m = c.masked ? NJ_LDG(a.b.M + (size_t)r * c.d + c_) : 1.f;
                const float y = t.YY[jr * c.sDO + c_], yb = t.YBJ[jr * c.sDO + c_];
                const float da = x - y, db = (c.loss_kind == NJODE_LOSS_STANDARD) ? (yb - y) : (yb - x);
                sa = fmaf(m * da, da, sa); sb = fmaf(m * db, db, sb);
            }
            const float ra = sqrtf(sa + 1e-10f), rb = sqrtf(sb + 1e-10f);
            const float wa = (c.loss_kind == NJODE_LOSS_STANDARD) ? 2.f * c.w : c.w;
            const float wb = (c.loss_kind == NJODE_LOSS_STANDARD) ? 2.f * (1.f - c.w) : (1.f - c.w);
            const float s = wa * ra + wb * rb;
            const float cf = gl * 2.f * s / (NJ_LDG(a.b.n_obs_ot + NJ_IU(t, NJ_I_PATH, u)) * (float)a.b.batch_size_norm);
            NJ_FU(t, NJ_F_CA, jr) = cf * wa / ra;
            NJ_FU(t, NJ_F_CB, jr) = cf * wb / rb;
        }
    }
    NJ_SYNC();
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.dout; idx += t.nt) {
            const int jr = nj_div(idx, c.dout, c.m_dout), c_ = idx - jr * c.dout;
            const int u = NJ_IU(t, NJ_I_JMAP, jr), r = NJ_IU(t, NJ_I_JROW, jr);
            const float x = NJ_LDG(a.b.X + (size_t)r * c.d + c_);
            const float m = c.masked ? NJ_LDG(a.b.M + (size_t)r * c.d + c_) : 1.f;
            const float y = t.YY[jr * c.sDO + c_], yb = t.YBJ[jr * c.sDO + c_];
            const float ca = NJ_FU(t, NJ_F_CA, jr), cb = NJ_FU(t, NJ_F_CB, jr);
            float gy, gyb;
            if (c.loss_kind == NJODE_LOSS_STANDARD) { gy = -ca * m * (x - y) - cb * m * (yb - y); gyb = cb * m * (yb - y); }
            else { gy = -ca * m * (x - y); gyb = cb * m * (yb - x); }
            if (c.masked) gy += t.GX[u * c.sD + c_];          // last_X = Y[i_obs]  (NJODE/models.py:483-484)
            t.GOUT[jr * c.sOUT + c_] = gy;
            t.GYBJ[jr * c.sDO + c_] = gyb;
        }
    }
    NJ_SYNC();
    // 4. readout backward at E
    float* gin = nj_mlp_backward(t, NJODE_NET_RO, nj, true);
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.H; idx += t.nt) {
            const int jr = nj_div(idx, c.H, c.m_H), c_ = idx - jr * c.H;
            const int u = NJ_IU(t, NJ_I_JMAP, jr);
            const float th = t.IN[(size_t)jr * c.sIN + c_];
            float ge = t.GH[u * c.sH + c_] + gin[jr * c.sG + c_] * (1.f - th * th);
            if (c.residual) ge += nj_resid_bwd(t.GOUT + jr * c.sOUT, c.H, c.dout, c_);
            t.GTMP[jr * c.sOUT + c_] = ge;
        }
    }
    NJ_SYNC();
    // 5. encoder recompute + backward   (use_rnn: backward of the GRU cell; EE <- gradient wrt h_before through it)
    if (c.use_rnn) {
        const int H = c.H;
        NJ_THREADS(tid, t.nt) {
            for (int idx = tid; idx < nj * H; idx += t.nt) {
                const int jr = nj_div(idx, H, c.m_H), c_ = idx - jr * H;
                float* gi = t.GI + jr * c.s3H;
                const float* gh = t.GHH + jr * c.s3H;
                float* go = t.GOUT + jr * c.sOUT;
                const float ge = t.GTMP[jr * c.sOUT + c_];
                const float r = gi[c_], z = gi[H + c_], n = gi[2 * H + c_], hh = gh[c_], ghn = gh[2 * H + c_];
                const float dpn = ge * (1.f - z) * (1.f - n * n);          // wrt the pre-activation of n
                const float dpr = dpn * ghn * r * (1.f - r);
                const float dpz = ge * (hh - n) * z * (1.f - z);
                go[c_] = dpr; go[H + c_] = dpz; go[2 * H + c_] = dpn * r;  // wrt gh = W_hh tanh(h_old) + b_hh
                gi[c_] = dpr; gi[H + c_] = dpz; gi[2 * H + c_] = dpn;      // wrt gi = W_ih tanh(x) + b_ih
                t.EE[jr * c.sH + c_] = ge * z;                             // direct path h' = ... + z tanh(h_old)
                t.IN[(size_t)jr * c.sIN + c_] = hh;
            }
        }
        NJ_SYNC();
        gin = nj_mlp_backward(t, NJODE_NET_GRU_HH, nj, true);
        NJ_THREADS(tid, t.nt) {
            for (int idx = tid; idx < nj * H; idx += t.nt) {
                const int jr = nj_div(idx, H, c.m_H), c_ = idx - jr * H;
                const float hh = t.GHH[jr * c.s3H + c_];
                t.EE[jr * c.sH + c_] = (t.EE[jr * c.sH + c_] + gin[jr * c.sG + c_]) * (1.f - hh * hh);
            }
            for (int idx = tid; idx < nj * 3 * H; idx += t.nt) {
                const int jr = nj_div(idx, (3 * H), c.m_3H), k = idx - jr * (3 * H);
                t.GOUT[jr * c.sOUT + k] = t.GI[jr * c.s3H + k];
            }
        }
        NJ_SYNC();
        NJ_THREADS(tid, t.nt) {
            for (int idx = tid; idx < nj * c.d; idx += t.nt) {
                const int jr = nj_div(idx, c.d, c.m_d), c_ = idx - jr * c.d;
                t.IN[(size_t)jr * c.sIN + c_] = nj_tanh(t.XI[jr * c.sD + c_]);
            }
        }
        NJ_SYNC();
        nj_mlp_backward(t, NJODE_NET_GRU_IH, nj, false);
    } else {
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.d; idx += t.nt) {
            const int jr = nj_div(idx, c.d, c.m_d), c_ = idx - jr * c.d;
            t.IN[(size_t)jr * c.sIN + c_] = nj_tanh(t.XI[jr * c.sD + c_]);
            if (c.masked) t.IN[(size_t)jr * c.sIN + c.d + c_] = NJ_LDG(a.b.M + (size_t)NJ_IU(t, NJ_I_JROW, jr) * c.d + c_);
        }
        for (int idx = tid; idx < nj * c.H; idx += t.nt) {
            const int jr = nj_div(idx, c.H, c.m_H), c_ = idx - jr * c.H;
            t.GOUT[jr * c.sOUT + c_] = t.GTMP[jr * c.sOUT + c_];
        }
        nj_set_jump_keys(t, a, nj, 1, tid);
    }
    NJ_SYNC();
    nj_mlp_forward(t, NJODE_NET_ENC, nj, true);
    gin = nj_mlp_backward(t, NJODE_NET_ENC, nj, c.masked != 0);
    if (c.masked) {
        NJ_THREADS(tid, t.nt) {
            for (int idx = tid; idx < nj * c.d; idx += t.nt) {
                const int jr = nj_div(idx, c.d, c.m_d), c_ = idx - jr * c.d;
                const float tx = t.IN[(size_t)jr * c.sIN + c_];
                float gx = gin[jr * c.sG + c_] * (1.f - tx * tx);
                if (c.residual) gx += nj_resid_bwd(t.GOUT + jr * c.sOUT, c.d, c.H, c_);
                const float m = NJ_LDG(a.b.M + (size_t)NJ_IU(t, NJ_I_JROW, jr) * c.d + c_);
                t.GYBJ[jr * c.sDO + c_] += (1.f - m) * gx;
            }
        }
        NJ_SYNC();
    }
    }
    // 6. readout recompute at h_before + backward -> gradient wrt h before the jump
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.H; idx += t.nt) {
            const int jr = nj_div(idx, c.H, c.m_H), c_ = idx - jr * c.H;
            t.IN[(size_t)jr * c.sIN + c_] = nj_tanh(t.XH[jr * c.sH + c_]);
        }
        for (int idx = tid; idx < nj * c.dout; idx += t.nt) {
            const int jr = nj_div(idx, c.dout, c.m_dout), c_ = idx - jr * c.dout;
            t.GOUT[jr * c.sOUT + c_] = t.GYBJ[jr * c.sDO + c_];
        }
        nj_set_jump_keys(t, a, nj, 0, tid);
    }
    NJ_SYNC();
    nj_mlp_forward(t, NJODE_NET_RO, nj, true);
    gin = nj_mlp_backward(t, NJODE_NET_RO, nj, true);
    NJ_THREADS(tid, t.nt) {
        for (int idx = tid; idx < nj * c.H; idx += t.nt) {
            const int jr = nj_div(idx, c.H, c.m_H), c_ = idx - jr * c.H;
            const int u = NJ_IU(t, NJ_I_JMAP, jr);
            const float th = t.IN[(size_t)jr * c.sIN + c_];
            float gh = gin[jr * c.sG + c_] * (1.f - th * th);
            if (c.residual) gh += nj_resid_bwd(t.GOUT + jr * c.sOUT, c.H, c.dout, c_);
            if (c.use_rnn) gh += t.EE[jr * c.sH + c_];
            t.GH[u * c.sH + c_] = gh;
        }
        for (int idx = tid; idx < nj * c.d; idx += t.nt) {
            const int jr = nj_div(idx, c.d, c.m_d), c_ = idx - jr * c.d;
            t.GX[NJ_IU(t, NJ_I_JMAP, jr) * c.sD + c_] = 0.f;
        }
        if (tid < nj) NJ_IU(t, NJ_I_CUR, NJ_IU(t, NJ_I_JMAP, tid)) -= 1;
    }
    NJ_SYNC();
}

NJ_HD void nj_cta_backward(const NjCfg& c, const NjArgs& a, float* smem, int cta, int ncta) {
    NjCta t;
    nj_cta_bind(t, c, smem, true);
    if (c.w_smem) {
        float* simg = smem + c.o_img;
        NJ_THREADS(tid, t.nt) { for (int i = tid; i < c.img_floats / 4; i += t.nt) nj_st4(simg + 4 * i, nj_ld4(a.image + 4 * i)); }
        t.wimg = simg;
    } else t.wimg = a.image;
    float* gpart = a.partials + (size_t)cta * c.img_floats;
    t.dimg = smem + c.o_dimg; t.gpart = gpart;
    nj_zero(t.dimg, c.dimg_floats, t.nt);
    if (c.dimg_floats < c.img_floats) nj_zero(gpart + c.dimg_floats, c.img_floats - c.dimg_floats, t.nt);
    nj_zero(smem + c.o_IN, c.o_I - c.o_IN, t.nt);
    NJ_SYNC();
    const bool gate = a.b.unit_kind == 0;          // see nj_cta_forward
    const bool defer = a.b.unit_kind == 1;
    for (int tile = cta; tile < a.n_tiles; tile += ncta) {
        nj_load_units(t, a, tile, true);
        const int nu = (a.b.n_units - tile * c.P) < c.P ? (a.b.n_units - tile * c.P) : c.P;
        const int maxlen = t.CTL[NJ_CTL_MAXLEN];
        if (gate) nj_next_jump(t, a, maxlen, true);
        NJ_THREADS(tid, t.nt) {
            for (int idx = tid; idx < nu * c.H; idx += t.nt) {
                const int u = nj_div(idx, c.H, c.m_H), c_ = idx - u * c.H;
                t.GH[u * c.sH + c_] = (NJ_IU(t, NJ_I_FLAG, u) && a.grad_hT) ? NJ_LDG(a.grad_hT + (size_t)NJ_IU(t, NJ_I_PATH, u) * c.H + c_) : 0.f;
            }
            for (int idx = tid; idx < nu * c.d; idx += t.nt) { const int u = nj_div(idx, c.d, c.m_d); t.GX[u * c.sD + idx - u * c.d] = 0.f; }
        }
        NJ_SYNC();
        nj_reload_state(t, a, nu);
        if (defer) {
            // reverse of the deferred jumps: every unit's jump precedes (in reverse order) all of its Euler steps
            nj_find_jumps(t, a, -1, -1, true);
            const int nj = t.CTL[NJ_CTL_NJ];
            if (nj > 0) { nj_jump_backward(t, a, nj); nj_reload_state(t, a, nu); }
        }
        for (int j = maxlen; j >= 0; --j) {
            if (j < maxlen) {
                // ---- reverse of Euler step k = s0 + j ----
                NJ_THREADS(tid, t.nt) {
                    for (int idx = tid; idx < nu * c.inf; idx += t.nt) {
                        const int u = nj_div(idx, c.inf, c.m_inf), c_ = idx - u * c.inf;
                        const bool active = j < NJ_IU(t, NJ_I_LEN, u);
                        const int k = NJ_IU(t, NJ_I_S0, u) + j;
                        float hval = 0.f;
                        if (active && c_ >= c.d && c_ < c.d + c.H)
                            hval = a.h_hist[((size_t)k * a.b.B + NJ_IU(t, NJ_I_PATH, u)) * c.H + c_ - c.d];
                        nj_build_ode_input(t, u, c_, hval, active ? NJ_LDG(a.b.step_t + k) : 0.f);
                        if (c_ == 0) {
                            NJ_FU(t, NJ_F_DT, u) = active ? NJ_LDG(a.b.step_dt + k) : 0.f;
                            NJ_IU(t, NJ_I_RK, u) = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(NJ_IU(t, NJ_I_PATH, u) + a.b.path_id_offset), (unsigned)k);
                        }
                    }
                }
                NJ_SYNC();
                NJ_THREADS(tid, t.nt) {
                    for (int idx = tid; idx < nu * c.H; idx += t.nt) {
                        const int u = nj_div(idx, c.H, c.m_H), c_ = idx - u * c.H;
                        t.GOUT[u * c.sOUT + c_] = NJ_FU(t, NJ_F_DT, u) * t.GH[u * c.sH + c_];
                    }
                }
                NJ_SYNC();
                nj_mlp_forward(t, NJODE_NET_ODE, nu, true);
                float* gin = nj_mlp_backward(t, NJODE_NET_ODE, nu, true);
                NJ_THREADS(tid, t.nt) {
                    for (int idx = tid; idx < nu * (c.d + c.H); idx += t.nt) {
                        const int u = nj_div(idx, (c.d + c.H), c.m_dH), c_ = idx - u * (c.d + c.H);
                        if (j >= NJ_IU(t, NJ_I_LEN, u)) continue;
                        const float th = t.IN[(size_t)u * c.sIN + c_];
                        const float g = gin[u * c.sG + c_] * (1.f - th * th);
                        if (c_ >= c.d) t.GH[u * c.sH + c_ - c.d] += g;
                        else if (c.masked) t.GX[u * c.sD + c_] += g;
                    }
                }
                NJ_SYNC();
            }
            // ---- reverse of the jumps that fall before step s0 + j ----
            for (; !defer;) {
                if (gate && t.CTL[NJ_CTL_NEXT] != j) break;            // no unit of the tile jumped before step s0 + j
                nj_find_jumps(t, a, j, -1, true);
                const int nj = t.CTL[NJ_CTL_NJ];
                if (nj == 0) break;
                nj_jump_backward(t, a, nj);
                nj_reload_state(t, a, nu);
                if (gate) nj_next_jump(t, a, j, true);
            }
        }
        // ---- reverse of the start encoder ----
        NJ_THREADS(tid, t.nt) {
            for (int idx = tid; idx < nu * c.d; idx += t.nt) {
                const int u = nj_div(idx, c.d, c.m_d), c_ = idx - u * c.d;
                const int sr = NJ_IU(t, NJ_I_START, u);
                const float x = sr < 0 ? NJ_LDG(a.b.start_X + (size_t)NJ_IU(t, NJ_I_PATH, u) * c.d + c_)
                                       : NJ_LDG(a.b.X + (size_t)sr * c.d + c_);
                t.IN[(size_t)u * c.sIN + c_] = nj_tanh(x);
                if (c.masked) t.IN[(size_t)u * c.sIN + c.d + c_] = 0.f;
            }
            for (int idx = tid; idx < nu * c.H; idx += t.nt) {
                const int u = nj_div(idx, c.H, c.m_H), c_ = idx - u * c.H;
                t.GOUT[u * c.sOUT + c_] = t.GH[u * c.sH + c_];
            }
            if (tid < nu) {
                const int sr = NJ_IU(t, NJ_I_START, tid);
                const unsigned ev = sr < 0 ? NJ_EVENT_INIT : NJ_EVENT_JUMP_BASE + 3u * (unsigned)NJ_LDG(a.b.row_jump + sr) + 1u;
                NJ_IU(t, NJ_I_RK, tid) = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(NJ_IU(t, NJ_I_PATH, tid) + a.b.path_id_offset), ev);
            }
        }
        NJ_SYNC();
        nj_mlp_forward(t, NJODE_NET_ENC, nu, true);
        nj_mlp_backward(t, NJODE_NET_ENC, nu, false);
    }
    if (c.dimg_floats) {
        NJ_THREADS(tid, t.nt) { for (int i = tid; i < c.dimg_floats / 4; i += t.nt) nj_st4(gpart + 4 * i, nj_ld4(t.dimg + 4 * i)); }
    }
}
