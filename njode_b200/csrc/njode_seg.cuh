// njode_seg.cuh -- fast path for the non-masked training call: every (path, inter-observation
// segment) is an independent work unit (NJODE/models.py:463-470: h after a jump = enc(X_obs) does not
// depend on h before it), so units are marched as dense tiles without any jump compaction.
//
// Design (sm_100a, fp32 FMA pipe; the demo nets are 13->50->50->10, far below a tensor-core tile):
//   * every Linear layer over a warp's R = 4*TR rows is a register-tiled GEMM: lane (rg, og) owns the
//     TR x TO accumulators of rows {rg + 4i} x outputs {og + 8j}; operands come from shared memory as
//     float4 along the reduction dimension (TR + TO LDS.128 per 4*TR*TO FFMA).  Activation rows use
//     a stride = 8 (mod 16) floats and weight rows a stride = 4 (mod 8) floats, which makes every
//     LDS.128 of the inner loop a single conflict-free wavefront.
//   * forward: warp-autonomous.  A warp owns R units from the start encoder to the loss row; there
//     is no CTA barrier after the weight image is in shared memory.  Tiles (sorted longest first)
//     are handed out by an atomic counter, i.e. greedy longest-processing-time scheduling.
//   * backward: a CTA owns P = NW*R units.  Per Euler step every warp recomputes the hidden
//     activations of its R rows and back-propagates through the layers (warp-local phase A), then
//     all threads accumulate dW += G^T A over the P rows of the CTA into 4x4 tiles that stay in
//     REGISTERS for the whole launch (phase B) -- two CTA barriers per step.
//   * dropout: a dropped activation is stored as -0.0f, so the backward pass reads the keep bit
//     from the recomputed activation instead of hashing again.
//
// Same dual-compilation scheme as njode_core.cuh: phases separated by NJ_SYNC / NJ_SYNCWARP with all
// cross-lane state in shared arrays, so -DNJODE_HOST_SIM runs it sequentially on the host (tests).
#pragma once
#include "njode_core.cuh"

#if defined(NJODE_HOST_SIM)
#define NJ_WARPS(w, nw) for (int w = 0; w < (nw); ++w)
#define NJ_LANES(lane) for (int lane = 0; lane < 32; ++lane)
#define NJ_SYNCWARP() ((void)0)
static inline int nj_atomic_inc(int* p) { return (*p)++; }
static inline unsigned nj_f2u(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float nj_u2f(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
#else
#define NJ_WARPS(w, nw) for (int w = (int)(threadIdx.x >> 5), _nj_we = w + 1; w < _nj_we; ++w)
#define NJ_LANES(lane) for (int lane = (int)(threadIdx.x & 31), _nj_le = lane + 1; lane < _nj_le; ++lane)
#define NJ_SYNCWARP() __syncwarp()
__device__ __forceinline__ int nj_atomic_inc(int* p) { return atomicAdd(p, 1); }
__device__ __forceinline__ unsigned nj_f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float nj_u2f(unsigned u) { return __uint_as_float(u); }
#endif

// parameter image global -> shared: one TMA bulk copy (cp.async.bulk, completion on an mbarrier) issued by
// one thread instead of a cooperative load/store loop; the caller must NJ_SYNC() afterwards.
#if defined(NJODE_HOST_SIM)
static inline void nj_stage_image(float* dst, const float* src, int nfloats, int nt) {
    (void)nt;
    memcpy(dst, src, (size_t)nfloats * 4);
}
#else
__device__ __forceinline__ void nj_stage_image(float* dst, const float* src, int nfloats, int nt) {
    (void)nt;
    __shared__ __align__(8) unsigned long long nj_img_bar;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&nj_img_bar);
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned bytes = (unsigned)nfloats * 4u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        for (unsigned off = 0; off < bytes; off += 32768u) {
            const unsigned n = bytes - off < 32768u ? bytes - off : 32768u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(d + off), "l"(reinterpret_cast<const char*>(src) + off), "r"(n), "r"(bar) : "memory");
        }
    }
    __syncthreads();                       // the barrier is initialised and armed for every waiter
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar) : "memory");
}
#endif

#define NJ_SEG_NT_MAX 2            // dW tiles (4x4 + bias) a thread may own in registers
#define NJ_SEG_ACC (NJ_SEG_NT_MAX * 20)
#define NJ_DROPPED 0x80000000u     // bit pattern (-0.0f) of a dropped activation

// launch-constant layout of the segment kernels (floats unless noted)
struct NjSeg {
    int ok;                         // 0: not eligible, use the generic kernels
    int tr_f, tr_b;                 // rows per lane group: a warp owns 4*tr rows (forward / backward)
    int nw_f, nw_b;                 // warps per CTA that own rows (forward / backward)
    int nt_b;                       // threads of a backward CTA: 32 * nw_b + helper warps that only join the dW phases
    int sI, sA, sO, sH, sD;         // strides: first-layer input, hidden activations, last-layer output, H, d
    int nA;                         // hidden-activation buffers kept in backward (max n_linear - 1)
    // forward: per-warp region
    int f_region, f_IN, f_A0, f_A1, f_OUT, f_HS, f_LX, f_TX, f_XI, f_YBJ, f_F, f_I;
    int f_img, f_warp0, f_smem_floats;
    // backward: CTA-level [P][stride] arrays
    int b_img, b_IN, b_A, b_G, b_GOUT, b_GZ, b_OUT, b_GH, b_HB, b_EE, b_GE, b_XI, b_LX, b_TX, b_YBJ, b_YY, b_GYBJ, b_F, b_I;
    int b_smem_floats;
    int P_b;                        // rows per backward CTA
    int n_tiles_f, n_tiles_b;       // warp tiles (forward), CTA tiles (backward)
    int tile_base[3][NJODE_MAX_LINEAR];   // first dW tile id of (net, layer); order ODE, RO, ENC
    int tiles_total, nt_slots;
    // work tiles come in classes of equal height: long units get low tiles (TR = 1), short ones tall tiles,
    // so that no tile's sequential chain of Euler steps dominates the makespan.  Class i serves the units
    // [u0, u1) in tiles of 4*tr (forward, per warp) or 4*tr*nw_b (backward, per CTA) rows; t0 = first tile id.
    int f_ncls, f_t0[7], f_u0[6], f_u1[6], f_tr[6];
    int b_ncls, b_t0[7], b_u0[6], b_u1[6], b_tr[6];
    // thread-per-neuron kernels of small batches (njode_tpn.cuh): dimension class, operand buffers three times (b_copy apart),
    // dW tile table, prefetch slots, gradient image of the jump networks in shared memory
    int tpn, b_copy, b_TD, b_PRE, b_GIMG, f_MB, b_MB;
};

// tile id -> (class, first unit, one-past-last unit)
NJ_HD int nj_seg_tile_lookup(int ncls, const int* t0, const int* u0, const int* u1, const int* tr, int rows_per_tr,
                             int tile, int& ub, int& ue) {
    int ci = 0;
    for (int i = 1; i < ncls; ++i) if (tile >= t0[i]) ci = i;
    const int rows = rows_per_tr * tr[ci];
    ub = u0[ci] + (tile - t0[ci]) * rows;
    ue = ub + rows < u1[ci] ? ub + rows : u1[ci];
    return tr[ci];
}

enum { NJS_I_PATH = 0, NJS_I_S0, NJS_I_LEN, NJS_I_ROW, NJS_I_START, NJS_I_FLAG, NJS_I_RK, NJS_I_COUNT };
enum { NJS_F_TAU = 0, NJS_F_DT, NJS_F_CA, NJS_F_CB, NJS_F_COUNT };

static inline int nj_stride_act(int n) { return ((n + 7) / 16) * 16 + 8; }    // = 8 (mod 16), >= n

// ------------------------------------------------------------------------------------------------
// warp GEMM primitives
// ------------------------------------------------------------------------------------------------
struct NjWL {                       // one Linear layer evaluated by a warp
    const float* in; int in_s; int K4;
    const float* W; int w_s; const float* bias;
    int o_base;
    float* out; int out_s;
    int act; int drop; unsigned thr; float keep_scale; const int* rk; unsigned tag;
};

// out[r][o] = act(b[o] + sum_k in[r][k] W[o][k]) (* dropout), rows r = rg + 4i, outputs o = o_base + og + 8j.
// Every o < o_base + 8*TO is stored: the image rows >= O are zero and the out stride covers them.
template <int TR, int TO>
NJ_HD void nj_wg_fwd(const NjWL& L, int lane) {
    const int rg = lane >> 3, og = lane & 7;
    float acc[TR][TO];
#pragma unroll
    for (int j = 0; j < TO; ++j) {
        const float b = L.bias ? L.bias[L.o_base + og + 8 * j] : 0.f;
#pragma unroll
        for (int i = 0; i < TR; ++i) acc[i][j] = b;
    }
    nj_sp ap[TR], wp[TO];
#pragma unroll
    for (int i = 0; i < TR; ++i) ap[i] = nj_sp_of(L.in + (size_t)(rg + 4 * i) * L.in_s);
#pragma unroll
    for (int j = 0; j < TO; ++j) wp[j] = nj_sp_of(L.W + (size_t)(L.o_base + og + 8 * j) * L.w_s);
#pragma unroll 2
    for (int k4 = 0; k4 < L.K4; ++k4) {
        nj_f4 a[TR], w[TO];
#pragma unroll
        for (int i = 0; i < TR; ++i) a[i] = nj_sp_ld4(NJ_SP_ADD(ap[i], 4 * k4));
#pragma unroll
        for (int j = 0; j < TO; ++j) w[j] = nj_sp_ld4(NJ_SP_ADD(wp[j], 4 * k4));
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int j = 0; j < TO; ++j) {
                acc[i][j] = fmaf(a[i].x, w[j].x, acc[i][j]);
                acc[i][j] = fmaf(a[i].y, w[j].y, acc[i][j]);
                acc[i][j] = fmaf(a[i].z, w[j].z, acc[i][j]);
                acc[i][j] = fmaf(a[i].w, w[j].w, acc[i][j]);
            }
    }
    const unsigned obase16 = (unsigned)(L.o_base + og);
#pragma unroll
    for (int i = 0; i < TR; ++i) {
        const int r = rg + 4 * i;
        float* orow = L.out + (size_t)r * L.out_s + L.o_base + og;
        if (L.drop) {
            const unsigned lk = nj_layer_key((unsigned)L.rk[r], L.tag);
#pragma unroll
            for (int j = 0; j < TO; j += 2) {
                // neurons o and o ^ 8 share a hash word (nj_keep); o_base is a multiple of 16 (planner),
                // so outputs j (even) and j + 1 of this lane are such a pair
                const unsigned o = obase16 + 8u * j;
                const unsigned word = nj_keep_word(lk, (o & 7u) | ((o >> 4) << 3));
                const float v0 = nj_act(acc[i][j], L.act) * L.keep_scale;
                orow[8 * j] = (word & 0xFFFFu) >= L.thr ? v0 : nj_u2f(NJ_DROPPED);
                if (j + 1 < TO) {
                    const float v1 = nj_act(acc[i][j + 1], L.act) * L.keep_scale;
                    orow[8 * j + 8] = (word >> 16) >= L.thr ? v1 : nj_u2f(NJ_DROPPED);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < TO; ++j) orow[8 * j] = nj_act(acc[i][j], L.act);
        }
    }
}

#if defined(NJODE_HOST_SIM)
#define NJ_ASSUME_SHARED(p) ((void)0)
#else
#define NJ_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#endif

// one layer over the warp's rows: all output chunks.  Not inlined: one copy of the TO variants per TR.
template <int TR>
NJ_HDN void nj_seg_layer_fwd(NjWL L, int to0, int to_last, int nch) {
    NJ_ASSUME_SHARED(L.in); NJ_ASSUME_SHARED(L.W); NJ_ASSUME_SHARED(L.out); NJ_ASSUME_SHARED(L.rk);
    if (L.bias) NJ_ASSUME_SHARED(L.bias);
    for (int ch = 0; ch < nch; ++ch) {
        // chunks of 8 * to0 outputs, the last one narrower: the image holds exactly 8 * ceil(out / 8) rows
        L.o_base = ch * 8 * to0;
        const int to = ch < nch - 1 ? to0 : to_last;
        NJ_LANES(lane) {
            switch (to) {
                case 1: nj_wg_fwd<TR, 1>(L, lane); break;
                case 2: nj_wg_fwd<TR, 2>(L, lane); break;
                case 3: nj_wg_fwd<TR, 3>(L, lane); break;
                case 4: nj_wg_fwd<TR, 4>(L, lane); break;
                case 5: nj_wg_fwd<TR, 5>(L, lane); break;
                case 6: nj_wg_fwd<TR, 6>(L, lane); break;
                case 7: nj_wg_fwd<TR, 7>(L, lane); break;
                default: nj_wg_fwd<TR, 8>(L, lane); break;
            }
        }
    }
}

struct NjWD {                       // input-gradient of one Linear layer evaluated by a warp
    const float* g; int g_s; int O4;
    const float* W; int w_s;
    int kg_base, K4in;
    float* gin; int gin_s;
    const float* aprev; int a_s; int act_prev;
    int drop; float keep_scale, one_minus_p;
};

// gin[r][k] = (sum_o g[r][o] W[o][k]) * act'(aprev[r][k]) * dropout factor; rows rg + 4i,
// k-groups (float4) kg = kg_base + kq + 8*jk
template <int TR, int TK>
NJ_HD void nj_wg_dx(const NjWD& L, int lane) {
    const int rg = lane >> 3, kq = lane & 7;
    float acc[TR][TK][4];
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int jk = 0; jk < TK; ++jk)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][jk][c] = 0.f;
    nj_sp wp[TK], gp[TR];
#pragma unroll
    for (int jk = 0; jk < TK; ++jk) {
        int kg = L.kg_base + kq + 8 * jk;
        kg = kg < L.K4in ? kg : L.K4in - 1;          // clamped lanes compute a duplicate, never store
        wp[jk] = nj_sp_of(L.W + 4 * kg);
    }
#pragma unroll
    for (int i = 0; i < TR; ++i) gp[i] = nj_sp_of(L.g + (size_t)(rg + 4 * i) * L.g_s);
    const int ws = L.w_s;
    for (int o4 = 0; o4 < L.O4; ++o4) {
        nj_f4 gv[TR];
#pragma unroll
        for (int i = 0; i < TR; ++i) gv[i] = nj_sp_ld4(NJ_SP_ADD(gp[i], 4 * o4));
#pragma unroll
        for (int jk = 0; jk < TK; ++jk) {
            const nj_sp q = NJ_SP_ADD(wp[jk], 4 * o4 * ws);
            const nj_f4 w0 = nj_sp_ld4(q), w1 = nj_sp_ld4(NJ_SP_ADD(q, ws)), w2 = nj_sp_ld4(NJ_SP_ADD(q, 2 * ws)), w3 = nj_sp_ld4(NJ_SP_ADD(q, 3 * ws));
#pragma unroll
            for (int i = 0; i < TR; ++i) {
                acc[i][jk][0] = fmaf(gv[i].x, w0.x, fmaf(gv[i].y, w1.x, fmaf(gv[i].z, w2.x, fmaf(gv[i].w, w3.x, acc[i][jk][0]))));
                acc[i][jk][1] = fmaf(gv[i].x, w0.y, fmaf(gv[i].y, w1.y, fmaf(gv[i].z, w2.y, fmaf(gv[i].w, w3.y, acc[i][jk][1]))));
                acc[i][jk][2] = fmaf(gv[i].x, w0.z, fmaf(gv[i].y, w1.z, fmaf(gv[i].z, w2.z, fmaf(gv[i].w, w3.z, acc[i][jk][2]))));
                acc[i][jk][3] = fmaf(gv[i].x, w0.w, fmaf(gv[i].y, w1.w, fmaf(gv[i].z, w2.w, fmaf(gv[i].w, w3.w, acc[i][jk][3]))));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < TR; ++i) {
        const int r = rg + 4 * i;
#pragma unroll
        for (int jk = 0; jk < TK; ++jk) {
            const int kg = L.kg_base + kq + 8 * jk;
            if (kg >= L.K4in) continue;
            float v[4] = {acc[i][jk][0], acc[i][jk][1], acc[i][jk][2], acc[i][jk][3]};
            if (L.aprev) {
                const nj_f4 av = nj_sp_ld4(nj_sp_of(L.aprev + (size_t)r * L.a_s + 4 * kg));
                const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float a = a4[c];
                    if (L.drop) {
                        if (nj_f2u(a) == NJ_DROPPED) v[c] = 0.f;
                        else { a *= L.one_minus_p; v[c] *= L.keep_scale; }
                    }
                    if (L.act_prev == NJODE_ACT_TANH) v[c] *= (1.f - a * a);
                    else if (L.act_prev == NJODE_ACT_RELU) v[c] = a > 0.f ? v[c] : 0.f;
                }
            }
            nj_f4 o; o.x = v[0]; o.y = v[1]; o.z = v[2]; o.w = v[3];
            nj_st4(L.gin + (size_t)r * L.gin_s + 4 * kg, o);
        }
    }
}

template <int TR>
NJ_HDN void nj_seg_layer_dx(NjWD D) {
    NJ_ASSUME_SHARED(D.g); NJ_ASSUME_SHARED(D.W); NJ_ASSUME_SHARED(D.gin);
    if (D.aprev) NJ_ASSUME_SHARED(D.aprev);
    for (int kb = 0; kb < D.K4in; kb += 16) {
        D.kg_base = kb;
        const bool two = (D.K4in - kb) > 8;
        NJ_LANES(lane) {
            if (two) nj_wg_dx<TR, 2>(D, lane); else nj_wg_dx<TR, 1>(D, lane);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// per-warp views
// ------------------------------------------------------------------------------------------------
struct NjSegW {
    const NjCfg* c; const NjSeg* s;
    const float* wimg;
    float *IN, *A0, *A1, *OUT;          // forward: A0/A1 ping-pong; backward: A0 = first hidden buffer
    float *G0, *GOUT, *GZ;              // backward only
    int a_buf_stride, g_buf_stride;     // distance (floats) between consecutive hidden buffers
    int* RK;                            // dropout row keys of this warp's rows
    void* coop;                         // mailbox of the cooperative layer service (njode_tpn.cuh) or null
};

// cooperative evaluation of one layer by the F threads of a thread-per-neuron CTA, posted by the glue warp (njode_tpn.cuh)
NJ_HD void nj_coop_post_fwd(void* mailbox, const NjWL& L, int o_store);
NJ_HD void nj_coop_post_dx(void* mailbox, const NjWD& D);

// MLP forward over the warp's rows.  keep_all: hidden activations of layer l go to A0 + l*a_buf_stride
// (backward); otherwise they ping-pong between A0 and A1.  skip_last: stop after the hidden layers.
template <int TR, bool COOP = false>
NJ_HD void nj_seg_mlp_fwd(const NjSegW& w, int netid, bool keep_all, bool skip_last) {
    const NjCfg& c = *w.c;
    const NjNet& N = c.net[netid];
    const float* in = w.IN; int in_s = w.s->sI;
    for (int l = 0; l < N.n; ++l) {
        const bool last = (l == N.n - 1);
        if (last && skip_last) break;
        NjWL L;
        L.in = in; L.in_s = in_s; L.K4 = (N.dim[l] + 3) >> 2;
        L.W = w.wimg + N.w_img[l]; L.w_s = N.ks[l];
        L.bias = N.b_src[l] >= 0 ? w.wimg + N.b_img[l] : nullptr;
        if (last) { L.out = w.OUT; L.out_s = w.s->sO; }
        else { L.out = keep_all ? w.A0 + (size_t)l * w.a_buf_stride : ((l & 1) ? w.A1 : w.A0); L.out_s = w.s->sA; }
        L.act = last ? NJODE_ACT_NONE : N.act[l];
        L.drop = (!last) && c.has_drop; L.thr = c.thr; L.keep_scale = c.keep_scale;
        L.rk = w.RK; L.tag = (unsigned)(netid * 16 + l + 1);
        L.o_base = 0;
        if (COOP) nj_coop_post_fwd(w.coop, L, 8 * (N.to[l] * (N.nch[l] - 1) + N.tol[l]));
        else nj_seg_layer_fwd<TR>(L, N.to[l], N.tol[l], N.nch[l]);
        NJ_SYNCWARP();
        in = L.out; in_s = L.out_s;
    }
}

// MLP backward (input gradients only; dW is phase B).  g wrt the raw output is in GOUT; hidden
// activations in A0 + l*a_buf_stride; g wrt hidden pre-activation l goes to G0 + l*g_buf_stride;
// the gradient wrt the network input goes to GZ when need_in_grad.
template <int TR, bool COOP = false>
NJ_HD void nj_seg_mlp_dx(const NjSegW& w, int netid, bool need_in_grad) {
    const NjCfg& c = *w.c;
    const NjNet& N = c.net[netid];
    for (int l = N.n - 1; l >= 0; --l) {
        if (l == 0 && !need_in_grad) break;
        NjWD D;
        if (l == N.n - 1) { D.g = w.GOUT; D.g_s = w.s->sO; } else { D.g = w.G0 + (size_t)l * w.g_buf_stride; D.g_s = w.s->sA; }
        D.O4 = (N.dim[l + 1] + 3) >> 2;
        D.W = w.wimg + N.w_img[l]; D.w_s = N.ks[l];
        D.K4in = (N.dim[l] + 3) >> 2;
        if (l > 0) {
            D.gin = w.G0 + (size_t)(l - 1) * w.g_buf_stride; D.gin_s = w.s->sA;
            D.aprev = w.A0 + (size_t)(l - 1) * w.a_buf_stride; D.a_s = w.s->sA; D.act_prev = N.act[l - 1];
        } else {
            D.gin = w.GZ; D.gin_s = w.s->sI; D.aprev = nullptr; D.a_s = 0; D.act_prev = NJODE_ACT_NONE;
        }
        D.drop = c.has_drop; D.keep_scale = c.keep_scale; D.one_minus_p = c.one_minus_p;
        D.kg_base = 0;
        if (COOP) nj_coop_post_dx(w.coop, D);
        else nj_seg_layer_dx<TR>(D);
        NJ_SYNCWARP();
    }
}

NJ_HD unsigned nj_seg_event_of_start(const NjArgs& a, int sr) {
    return sr < 0 ? NJ_EVENT_INIT : NJ_EVENT_JUMP_BASE + 3u * (unsigned)NJ_LDG(a.b.row_jump + sr) + 1u;
}
NJ_HD unsigned nj_seg_event_of_jump(const NjArgs& a, int row, unsigned which) {
    return NJ_EVENT_JUMP_BASE + 3u * (unsigned)NJ_LDG(a.b.row_jump + row) + which;
}

// lane -> (row, first column, column step) of the elementwise phases: 32/R lanes per row
#define NJ_ROWMAP(R)                                              \
    const int LPR = 32 / (R);                                     \
    const int er = lane / LPR, ec0 = lane % LPR;                  \
    (void)er; (void)ec0

// ------------------------------------------------------------------------------------------------
// forward: one warp = R = 4*TR units
// ------------------------------------------------------------------------------------------------
// saved hidden activations (NjArgs::act_hist): rows of one warp <-> [k][p][layer][act_wp] records, float4 per lane,
// consecutive lanes on consecutive 16 bytes of a row.  kp[r] = step index of row r or -1 (row rests), pp[r] = its path.
template <int R>
NJ_HD void nj_seg_act_store(const NjArgs& a, int l, const float* buf, int buf_s, const int* kp, int kstride, const int* pp, int j) {
    const int w4 = a.act_wp >> 2;
    NJ_LANES(lane) {
        for (int idx = lane; idx < R * w4; idx += 32) {
            const int r = idx / w4, q = idx - r * w4;
            if (j >= kp[NJS_I_LEN * kstride + r]) continue;
            const size_t rec = ((size_t)(kp[NJS_I_S0 * kstride + r] + j) * a.b.B + pp[r]) * a.act_nh + l;
            nj_st4(a.act_hist + rec * a.act_wp + 4 * q, nj_ld4(buf + (size_t)r * buf_s + 4 * q));
        }
    }
}
template <int R>
NJ_HD void nj_seg_act_load(const NjArgs& a, int l, float* buf, int buf_s, const int* kp, int kstride, const int* pp, int j) {
    const int w4 = a.act_wp >> 2;
    NJ_LANES(lane) {
        for (int idx = lane; idx < R * w4; idx += 32) {
            const int r = idx / w4, q = idx - r * w4;
            if (j >= kp[NJS_I_LEN * kstride + r]) continue;
            const size_t rec = ((size_t)(kp[NJS_I_S0 * kstride + r] + j) * a.b.B + pp[r]) * a.act_nh + l;
            nj_st4(buf + (size_t)r * buf_s + 4 * q, nj_ld4(a.act_hist + rec * a.act_wp + 4 * q));
        }
    }
}

// forward of one tile of R = 4*TR segment units by one warp, in three parts (the weight-stationary kernels of small
// batches run begin / finish on warp 0 and replace the Euler steps by CTA-cooperative ones)
template <int TR, bool COOP = false>
struct NjSegFwd {
    static constexpr int R = 4 * TR;
    static constexpr int RS = 16;             // row-slot stride of the per-warp scalar arrays (tallest tile)
    const NjCfg& c; const NjSeg& s; const NjArgs& a;
    NjSegW w;
    float *HS, *LX, *TX, *XI, *YBJ, *F;
    int* I;
    int d4, H4, inf4, sI, sO, sH, sD;
    int maxlen, any_jump;

    NJ_HD NjSegFwd(const NjCfg& c_, const NjSeg& s_, const NjArgs& a_, float* reg, const float* wimg) : c(c_), s(s_), a(a_) {
        w.c = &c; w.s = &s; w.wimg = wimg;
        w.IN = reg + s.f_IN; w.A0 = reg + s.f_A0; w.A1 = reg + s.f_A1; w.OUT = reg + s.f_OUT;
        w.G0 = w.GOUT = w.GZ = nullptr; w.a_buf_stride = 0; w.g_buf_stride = 0; w.coop = nullptr;
        HS = reg + s.f_HS; LX = reg + s.f_LX; TX = reg + s.f_TX; XI = reg + s.f_XI;
        YBJ = reg + s.f_YBJ; F = reg + s.f_F;
        I = reinterpret_cast<int*>(reg + s.f_I);
        w.RK = I + NJS_I_RK * RS;
        d4 = ((c.d + 3) >> 2) << 2; H4 = ((c.H + 3) >> 2) << 2; inf4 = ((c.inf + 3) >> 2) << 2;
        sI = s.sI; sO = s.sO; sH = s.sH; sD = s.sD;
        maxlen = 0; any_jump = 0;
    }

    // unit descriptors, h = encoder(start value)
    NJ_HD void begin(int u0, int u1) {
    // ---- unit descriptors ----
    NJ_LANES(lane) {
        if (lane < R) {
            const int u = u0 + lane;
            if (u < u1) {
                const int32_t* dsc = a.b.unit_desc + (size_t)u * 6;
                const int sc = dsc[5];
                I[NJS_I_PATH * RS + lane] = dsc[0]; I[NJS_I_S0 * RS + lane] = dsc[1]; I[NJS_I_LEN * RS + lane] = dsc[2] - dsc[1];
                I[NJS_I_ROW * RS + lane] = dsc[4] > dsc[3] ? NJ_LDG(a.b.path_rows + dsc[3]) : -1;
                I[NJS_I_FLAG * RS + lane] = (sc & NJODE_UNIT_WRITES_HT) ? 1 : 0;
                I[NJS_I_START * RS + lane] = (sc & ~NJODE_UNIT_WRITES_HT) - 1;
            } else {
                I[NJS_I_PATH * RS + lane] = -1; I[NJS_I_S0 * RS + lane] = 0; I[NJS_I_LEN * RS + lane] = 0;
                I[NJS_I_ROW * RS + lane] = -1; I[NJS_I_FLAG * RS + lane] = 0; I[NJS_I_START * RS + lane] = -1;
            }
        }
    }
    NJ_SYNCWARP();
    maxlen = 0; any_jump = 0;
    for (int r = 0; r < R; ++r) {
        maxlen = I[NJS_I_LEN * RS + r] > maxlen ? I[NJS_I_LEN * RS + r] : maxlen;
        any_jump |= (I[NJS_I_ROW * RS + r] >= 0);
    }
    // ---- start: h = encoder(start value); last_X, tanh(last_X), tau stay fixed for the whole unit ----
    NJ_LANES(lane) {
        NJ_ROWMAP(R);
        const int p = I[NJS_I_PATH * RS + er], sr = I[NJS_I_START * RS + er];
        for (int c_ = ec0; c_ < d4; c_ += LPR) {
            float x = 0.f;
            if (p >= 0 && c_ < c.d) x = sr < 0 ? NJ_LDG(a.b.start_X + (size_t)p * c.d + c_) : NJ_LDG(a.b.X + (size_t)sr * c.d + c_);
            const float tx = c_ < c.d ? nj_tanh(x) : 0.f;
            XI[er * sD + c_] = x; LX[er * sD + c_] = x; TX[er * sD + c_] = tx;
            w.IN[(size_t)er * sI + c_] = tx;
        }
        if (ec0 == 0) {
            F[NJS_F_TAU * RS + er] = (p >= 0 && sr >= 0) ? NJ_LDG(a.b.jump_tau + NJ_LDG(a.b.row_jump + sr)) : 0.f;
            w.RK[er] = p >= 0 ? (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(p + a.b.path_id_offset), nj_seg_event_of_start(a, sr)) : 0;
        }
    }
    NJ_SYNCWARP();
    nj_seg_mlp_fwd<TR, COOP>(w, NJODE_NET_ENC, false, false);
    NJ_LANES(lane) {
        NJ_ROWMAP(R);
        for (int c_ = ec0; c_ < c.H; c_ += LPR) {
            float e = w.OUT[er * sO + c_];
            if (c.residual) e += nj_resid(XI + er * sD, c.d, c.H, c_);
            HS[er * sH + c_] = e;
        }
    }
    NJ_SYNCWARP();
    }

    // Euler step j of the tile (units shorter than j + 1 steps rest)
    NJ_HD void step(int j) {
        const int sA_ = s.sA;

        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const bool active = j < I[NJS_I_LEN * RS + er];
            const int k = I[NJS_I_S0 * RS + er] + j;
            const int p = I[NJS_I_PATH * RS + er];
            const float tau = F[NJS_F_TAU * RS + er];
            float* hh = (active && a.h_hist) ? a.h_hist + ((size_t)k * a.b.B + p) * c.H : nullptr;
            for (int c_ = ec0; c_ < inf4; c_ += LPR) {
                float v = 0.f;
                if (c_ < c.d) v = TX[er * sD + c_];
                else if (c_ < c.d + c.H) {
                    const float h = HS[er * sH + c_ - c.d];
                    if (hh) hh[c_ - c.d] = h;
                    v = nj_tanh(h);
                } else if (c_ < c.inf) {
                    const float tcur = active ? NJ_LDG(a.b.step_t + k) : 0.f;
                    if (c_ == c.d + c.H) v = tau;
                    else if (c_ == c.d + c.H + 1) v = tcur - tau;
                    else v = tau + (tcur - tau);
                }
                w.IN[(size_t)er * sI + c_] = v;
            }
            if (ec0 == 0) {
                F[NJS_F_DT * RS + er] = active ? NJ_LDG(a.b.step_dt + k) : 0.f;
                w.RK[er] = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(p + a.b.path_id_offset), (unsigned)k);
            }
        }
        NJ_SYNCWARP();
        nj_seg_mlp_fwd<TR, COOP>(w, NJODE_NET_ODE, false, false);
        if (a.act_hist) {                       // (the planner offers the buffer for ODE networks with <= 2 hidden layers: A0, A1)
            nj_seg_act_store<R>(a, 0, w.A0, sA_, I, RS, I + NJS_I_PATH * RS, j);
            if (a.act_nh > 1) nj_seg_act_store<R>(a, 1, w.A1, sA_, I, RS, I + NJS_I_PATH * RS, j);
        }
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            if (j < I[NJS_I_LEN * RS + er]) {
                const float dt = F[NJS_F_DT * RS + er];
                for (int c_ = ec0; c_ < c.H; c_ += LPR)
                    HS[er * sH + c_] = fmaf(dt, w.OUT[er * sO + c_], HS[er * sH + c_]);
            }
        }
        NJ_SYNCWARP();
        }

    // hT of the units that end their path, then the jump that ends the segment
    NJ_HD void finish() {
    // ---- hT of the units that end their path ----
    NJ_LANES(lane) {
        NJ_ROWMAP(R);
        if (I[NJS_I_FLAG * RS + er]) {
            float* dst = a.hT + (size_t)I[NJS_I_PATH * RS + er] * c.H;
            for (int c_ = ec0; c_ < c.H; c_ += LPR) dst[c_] = HS[er * sH + c_];
        }
    }
    if (!any_jump) { NJ_SYNCWARP(); return; }
    // ---- the jump that ends the segment (NJODE/models.py:449-489) ----
    NJ_LANES(lane) {
        NJ_ROWMAP(R);
        const int row = I[NJS_I_ROW * RS + er], p = I[NJS_I_PATH * RS + er];
        for (int c_ = ec0; c_ < H4; c_ += LPR) {
            float h = 0.f;
            if (row >= 0 && c_ < c.H) {
                h = HS[er * sH + c_];
                if (a.h_before) a.h_before[(size_t)row * c.H + c_] = h;
            }
            w.IN[(size_t)er * sI + c_] = c_ < c.H ? nj_tanh(h) : 0.f;
        }
        if (ec0 == 0)
            w.RK[er] = row >= 0 ? (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(p + a.b.path_id_offset), nj_seg_event_of_jump(a, row, 0u)) : 0;
    }
    NJ_SYNCWARP();
    nj_seg_mlp_fwd<TR, COOP>(w, NJODE_NET_RO, false, false);
    NJ_LANES(lane) {
        NJ_ROWMAP(R);
        const int row = I[NJS_I_ROW * RS + er], p = I[NJS_I_PATH * RS + er];
        for (int c_ = ec0; c_ < c.dout; c_ += LPR) {
            float y = w.OUT[er * sO + c_];
            if (c.residual) y += nj_resid(HS + er * sH, c.H, c.dout, c_);
            YBJ[er * sD + c_] = y;
        }
        for (int c_ = ec0; c_ < d4; c_ += LPR) {
            float x = 0.f;
            if (row >= 0 && c_ < c.d) x = NJ_LDG(a.b.X + (size_t)row * c.d + c_);
            XI[er * sD + c_] = x;
            w.IN[(size_t)er * sI + c_] = c_ < c.d ? nj_tanh(x) : 0.f;
        }
        if (ec0 == 0)
            w.RK[er] = row >= 0 ? (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(p + a.b.path_id_offset), nj_seg_event_of_jump(a, row, 1u)) : 0;
    }
    NJ_SYNCWARP();
    nj_seg_mlp_fwd<TR, COOP>(w, NJODE_NET_ENC, false, false);
    NJ_LANES(lane) {
        NJ_ROWMAP(R);
        const int row = I[NJS_I_ROW * RS + er], p = I[NJS_I_PATH * RS + er];
        for (int c_ = ec0; c_ < H4; c_ += LPR) {
            float e = 0.f;
            if (c_ < c.H) {
                e = w.OUT[er * sO + c_];
                if (c.residual) e += nj_resid(XI + er * sD, c.d, c.H, c_);
                HS[er * sH + c_] = e;             // the unit is finished: HS now holds enc(X_obs)
            }
            w.IN[(size_t)er * sI + c_] = c_ < c.H ? nj_tanh(e) : 0.f;
        }
        if (ec0 == 0)
            w.RK[er] = row >= 0 ? (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(p + a.b.path_id_offset), nj_seg_event_of_jump(a, row, 2u)) : 0;
    }
    NJ_SYNCWARP();
    nj_seg_mlp_fwd<TR, COOP>(w, NJODE_NET_RO, false, false);
    NJ_LANES(lane) {
        if (lane < R) {
            const int r = lane, row = I[NJS_I_ROW * RS + r];
            if (row >= 0) {
                float sa = 0.f, sb = 0.f;
                NJ_UNROLL4
                    for (int c_ = 0; c_ < c.dout; ++c_) {      // (unrolled: the independent loads of four features go out together)
                    float y = w.OUT[r * sO + c_];
                    if (c.residual) y += nj_resid(HS + r * sH, c.H, c.dout, c_);
                    if (a.y_after) a.y_after[(size_t)row * c.dout + c_] = y;
                    const float x = XI[r * sD + c_], yb = YBJ[r * sD + c_];
                    const float da = x - y, db = (c.loss_kind == NJODE_LOSS_STANDARD) ? (yb - y) : (yb - x);
                    sa = fmaf(da, da, sa); sb = fmaf(db, db, sb);
                }
                if (a.get_loss) {
                    const float ra = sqrtf(sa + 1e-10f), rb = sqrtf(sb + 1e-10f);
                    const float sm = (c.loss_kind == NJODE_LOSS_STANDARD) ? (2.f * c.w * ra + 2.f * (1.f - c.w) * rb)
                                                                          : (c.w * ra + (1.f - c.w) * rb);
                    a.row_loss[row] = sm * sm / NJ_LDG(a.b.n_obs_ot + I[NJS_I_PATH * RS + r]);
                }
            }
        }
    }
    NJ_SYNCWARP();
    }
};

template <int TR>
NJ_HD void nj_seg_forward_warp(const NjCfg& c, const NjSeg& s, const NjArgs& a, float* reg, const float* wimg, int u0, int u1) {
    NjSegFwd<TR> f(c, s, a, reg, wimg);
    f.begin(u0, u1);
    for (int j = 0; j < f.maxlen; ++j) f.step(j);
    f.finish();
}

NJ_HD void nj_seg_cta_forward(const NjCfg& c, const NjSeg& s, const NjArgs& a, float* smem) {
    float* simg = smem + s.f_img;
    nj_stage_image(simg, a.image, c.img_floats, s.nw_f * 32);
    nj_zero(smem + s.f_warp0, s.nw_f * s.f_region, s.nw_f * 32);
    NJ_SYNC();
    NJ_WARPS(wp, s.nw_f) {
        float* reg = smem + s.f_warp0 + (size_t)wp * s.f_region;
        int* slot = reinterpret_cast<int*>(reg + s.f_I) + NJS_I_COUNT * 16;
        for (;;) {
            NJ_LANES(lane) { if (lane == 0) *slot = nj_atomic_inc(a.counter); }
            NJ_SYNCWARP();
            const int wt = *slot;
            NJ_SYNCWARP();
            if (wt >= s.n_tiles_f) break;
            int ub, ue;
            const int tr = nj_seg_tile_lookup(s.f_ncls, s.f_t0, s.f_u0, s.f_u1, s.f_tr, 4, wt, ub, ue);
            if (tr == 4) nj_seg_forward_warp<4>(c, s, a, reg, simg, ub, ue);
            else if (tr == 2) nj_seg_forward_warp<2>(c, s, a, reg, simg, ub, ue);
            else nj_seg_forward_warp<1>(c, s, a, reg, simg, ub, ue);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct NjSegB {                      // CTA-level views
    float *IN, *A, *G, *GOUT, *GZ, *OUT, *GH, *HB, *EE, *GE, *XI, *LX, *TX, *YBJ, *YY, *GYBJ, *F;
    int* I;
};

NJ_HD void nj_segb_bind(NjSegB& t, const NjSeg& s, float* smem) {
    t.IN = smem + s.b_IN; t.A = smem + s.b_A; t.G = smem + s.b_G; t.GOUT = smem + s.b_GOUT; t.GZ = smem + s.b_GZ;
    t.OUT = smem + s.b_OUT; t.GH = smem + s.b_GH; t.HB = smem + s.b_HB; t.EE = smem + s.b_EE; t.GE = smem + s.b_GE;
    t.XI = smem + s.b_XI; t.LX = smem + s.b_LX; t.TX = smem + s.b_TX; t.YBJ = smem + s.b_YBJ; t.YY = smem + s.b_YY;
    t.GYBJ = smem + s.b_GYBJ; t.F = smem + s.b_F; t.I = reinterpret_cast<int*>(smem + s.b_I);
}

// phase B: dW[o][k] += sum_r g[r][o] a[r][k] over the P rows of the CTA, thread-owned 4x4 tiles
// (+ bias sums in the kg == 0 tiles) held in `acc` (registers) for the whole launch.
// (net, layer, og, kg) of dW tile T inside network `netid`; false when T belongs to another network
NJ_HD bool nj_seg_tile_decode(const NjCfg& c, const NjSeg& s, int netid, int T, int& l, int& og, int& kg) {
    const NjNet& N = c.net[netid];
    if (T < s.tile_base[netid][0]) return false;
    l = 0;
    for (int ll = N.n - 1; ll > 0; --ll) if (T >= s.tile_base[netid][ll]) { l = ll; break; }
    const int K4 = (N.dim[l] + 3) >> 2, O4 = (N.dim[l + 1] + 3) >> 2;
    const int tl = T - s.tile_base[netid][l];
    if (tl >= K4 * O4) return false;
    kg = tl % K4; og = tl / K4;
    return true;
}

// r[0..15] += sum_r g[r][4og..] (x) a[r][4kg..],  r[16..19] += sum_r g[r][4og..]   over Pt rows
NJ_HD void nj_seg_dw_rows(const NjCfg& c, const NjSeg& s, const NjSegB& t, int netid, int l, int og, int kg, int Pt, float* q) {
    const NjNet& N = c.net[netid];
    const int P = s.P_b;
    const float* g; int g_s;
    if (l == N.n - 1) { g = t.GOUT; g_s = s.sO; } else { g = t.G + (size_t)l * P * s.sA; g_s = s.sA; }
    const float* av; int a_s;
    if (l == 0) { av = t.IN; a_s = s.sI; } else { av = t.A + (size_t)(l - 1) * P * s.sA; a_s = s.sA; }
    nj_sp gq = nj_sp_of(g + 4 * og), aq = nj_sp_of(av + 4 * kg);
    float r00 = q[0], r01 = q[1], r02 = q[2], r03 = q[3], r10 = q[4], r11 = q[5], r12 = q[6], r13 = q[7];
    float r20 = q[8], r21 = q[9], r22 = q[10], r23 = q[11], r30 = q[12], r31 = q[13], r32 = q[14], r33 = q[15];
    float b0 = q[16], b1 = q[17], b2 = q[18], b3 = q[19];
#pragma unroll 4
    for (int r = 0; r < Pt; ++r) {
        const nj_f4 gv = nj_sp_ld4(gq);
        const nj_f4 x = nj_sp_ld4(aq);
        gq = NJ_SP_ADD(gq, g_s); aq = NJ_SP_ADD(aq, a_s);
        r00 = fmaf(gv.x, x.x, r00); r01 = fmaf(gv.x, x.y, r01); r02 = fmaf(gv.x, x.z, r02); r03 = fmaf(gv.x, x.w, r03);
        r10 = fmaf(gv.y, x.x, r10); r11 = fmaf(gv.y, x.y, r11); r12 = fmaf(gv.y, x.z, r12); r13 = fmaf(gv.y, x.w, r13);
        r20 = fmaf(gv.z, x.x, r20); r21 = fmaf(gv.z, x.y, r21); r22 = fmaf(gv.z, x.z, r22); r23 = fmaf(gv.z, x.w, r23);
        r30 = fmaf(gv.w, x.x, r30); r31 = fmaf(gv.w, x.y, r31); r32 = fmaf(gv.w, x.z, r32); r33 = fmaf(gv.w, x.w, r33);
        b0 += gv.x; b1 += gv.y; b2 += gv.z; b3 += gv.w;
    }
    q[0] = r00; q[1] = r01; q[2] = r02; q[3] = r03; q[4] = r10; q[5] = r11; q[6] = r12; q[7] = r13;
    q[8] = r20; q[9] = r21; q[10] = r22; q[11] = r23; q[12] = r30; q[13] = r31; q[14] = r32; q[15] = r33;
    q[16] = b0; q[17] = b1; q[18] = b2; q[19] = b3;
}

// adds one 4x4 tile (+ bias sums of the kg == 0 tile) into a gradient image; `add`: read-modify-write
NJ_HD void nj_seg_tile_store(const NjCfg& c, int netid, int l, int og, int kg, const float* q, float* gpart, bool add) {
    const NjNet& N = c.net[netid];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float* p = gpart + N.w_img[l] + (size_t)(4 * og + i) * N.ks[l] + 4 * kg;
        nj_f4 v; v.x = q[4 * i]; v.y = q[4 * i + 1]; v.z = q[4 * i + 2]; v.w = q[4 * i + 3];
        if (add) { const nj_f4 o = nj_ld4(p); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        nj_st4(p, v);
    }
    if (kg == 0 && N.b_src[l] >= 0) {
        float* p = gpart + N.b_img[l] + 4 * og;
        nj_f4 v; v.x = q[16]; v.y = q[17]; v.z = q[18]; v.w = q[19];
        if (add) { const nj_f4 o = nj_ld4(p); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        nj_st4(p, v);
    }
}

// phase B: dW[o][k] += sum_r g[r][o] a[r][k] over the Pt rows of the tile.  Tiles are thread-owned: the first
// NT_MAX * nt of them (the ODE network first) live in registers (`acc`) for the whole launch; the others are
// accumulated from zero and added into this CTA's partial image in global memory (L2 resident) right away.
// tiles beyond the register slots.  Out of line on purpose: inlined next to the 40 accumulator registers of the dW phase,
// ptxas 12.9 (-O3, sm_100a) has produced a wrong partial-image address for this read-modify-write (compute-sanitizer:
// out-of-bounds LDG; first seen in njode_path.cuh, then here once the recompute phase changed the register allocation of
// the kernel) while the same source is correct in the host simulation.
NJ_HDN void nj_seg_dw_overflow(const NjCfg* cp, const NjSeg* sp, const NjSegB* tp, int netid, float* gpart, int tid, int nt, int Pt) {
    const NjCfg& c = *cp; const NjSeg& s = *sp; const NjSegB& t = *tp;
    // the tiles of one network are a contiguous range of the tile numbering: walk the part of it beyond the register slots
    int lo = s.tile_base[netid][0], hi = s.tiles_total;
    for (int q = 0; q < 3; ++q) {
        const int b = s.tile_base[q][0];
        if (b > lo && b < hi) hi = b;
    }
    if (lo < s.nt_slots * nt) lo = s.nt_slots * nt;
    for (int T = lo + tid; T < hi; T += nt) {
        int l, og, kg;
        if (!nj_seg_tile_decode(c, s, netid, T, l, og, kg)) continue;
        float q[20];
#pragma unroll
        for (int i = 0; i < 20; ++i) q[i] = 0.f;
        nj_seg_dw_rows(c, s, t, netid, l, og, kg, Pt, q);
        nj_seg_tile_store(c, netid, l, og, kg, q, gpart, true);
    }
}

template <bool OVF = true>
NJ_HD void nj_seg_dw(const NjCfg& c, const NjSeg& s, const NjSegB& t, int netid, float* acc, float* gpart,
                     int tid, int nt, int Pt) {
#pragma unroll
    for (int slot = 0; slot < NJ_SEG_NT_MAX; ++slot) {
        if (slot >= s.nt_slots) break;
        int l, og, kg;
        if (!nj_seg_tile_decode(c, s, netid, slot * nt + tid, l, og, kg)) continue;
        nj_seg_dw_rows(c, s, t, netid, l, og, kg, Pt, acc + slot * 20);
    }
    // OVF = false: an instantiation for launches whose tiles all fit the register slots -- the mere presence of the
    // out-of-line call cost the backward of the 20 000-path demo batch 10 % (B200: 4.08 -> 4.55 ms)
    if (OVF && s.tiles_total > s.nt_slots * nt) nj_seg_dw_overflow(&c, &s, &t, netid, gpart, tid, nt, Pt);
}

// writes the register tiles into this CTA's partial gradient image (pre-zeroed by the caller)
NJ_HD void nj_seg_dw_flush(const NjCfg& c, const NjSeg& s, const float* acc, float* gpart, int tid, int nt) {
#pragma unroll
    for (int slot = 0; slot < NJ_SEG_NT_MAX; ++slot) {
        if (slot >= s.nt_slots) break;
        const int T = slot * nt + tid;
        if (T >= s.tiles_total) continue;
        for (int netid = 0; netid < 3; ++netid) {
            int l, og, kg;
            if (nj_seg_tile_decode(c, s, netid, T, l, og, kg)) { nj_seg_tile_store(c, netid, l, og, kg, acc + slot * 20, gpart, false); break; }
        }
    }
}

// thread-owned dW accumulators: registers on the device, one slice per simulated thread on the host
#if defined(NJODE_HOST_SIM)
#define NJ_ACC_DECL(nt) std::vector<float> nj_acc_store((size_t)(nt) * NJ_SEG_ACC, 0.f); float* nj_acc_base = nj_acc_store.data()
#define NJ_ACC_OFF(tid) ((size_t)(tid) * NJ_SEG_ACC)
#else
#define NJ_ACC_DECL(nt) float nj_acc_store[NJ_SEG_ACC]; _Pragma("unroll") for (int _i = 0; _i < NJ_SEG_ACC; ++_i) nj_acc_store[_i] = 0.f; float* nj_acc_base = nj_acc_store
#define NJ_ACC_OFF(tid) 0
#endif
#define NJ_ACC(tid) (nj_acc_base + NJ_ACC_OFF(tid))

#define NJ_SEGB_WARP_VIEW()                                                                                        \
    const int r0 = wp * R;                                                                                         \
    NjSegW w;                                                                                                      \
    w.c = &c; w.s = &s; w.wimg = simg;                                                                             \
    w.IN = t.IN + (size_t)r0 * sI; w.A0 = t.A + (size_t)r0 * s.sA; w.A1 = nullptr; w.OUT = t.OUT + (size_t)r0 * sO; \
    w.G0 = t.G + (size_t)r0 * s.sA; w.GOUT = t.GOUT + (size_t)r0 * sO; w.GZ = t.GZ + (size_t)r0 * sI;             \
    w.a_buf_stride = wa; w.g_buf_stride = wa; w.RK = t.I + NJS_I_RK * P + r0; w.coop = rev.mailbox()

#define NJ_SEGB_KEY(r, valid, ev)                                                                                  \
    t.I[NJS_I_RK * P + (r)] = (valid) ? (int)nj_row_key(c.seed_lo, c.seed_hi,                                      \
        (unsigned)(t.I[NJS_I_PATH * P + (r)] + a.b.path_id_offset), (ev)) : 0

// REV: how the Euler steps are reversed.  The default (NjSegWarpRev) is the warp-local recompute + dx below followed by the
// CTA-wide dW phase; the weight-stationary kernels of small batches (njode_path.cuh) pass a functor whose step(j) runs
// on all warps of the CTA, and then only warp 0 executes the warp-local sections of this function.
template <bool OVF>
struct NjSegWarpRevT {
    static constexpr bool stat = false;
    static constexpr bool glue = true;             // this instantiation contains the warp-local sections
    static constexpr bool ovf = OVF;               // dW tiles beyond the register slots exist (nj_seg_dw)
    static constexpr bool coop = false;            // layers of the warp-local sections: the warp's own GEMMs
    NJ_HD void run(int) const {}
    NJ_HD float* gpart(float* global_partial) const { return global_partial; }
    NJ_HD void* mailbox() const { return nullptr; }
    NJ_HD void serve(int) const {}
    NJ_HD void glue_done() const {}
};

typedef NjSegWarpRevT<true> NjSegWarpRev;

template <int TR, class REV = NjSegWarpRev>
NJ_HD void nj_seg_bwd_tile(const NjCfg& c, const NjSeg& s, const NjArgs& a, float* smem, const NjSegB& t, float* nj_acc_base,
                           int cta, int u0, int u1, const REV& rev = REV()) {
    constexpr int R = 4 * TR;
    const int P = s.P_b, nt = s.nt_b, Pt = R * s.nw_b;
    float* simg = smem + s.b_img;
    float* gpart = rev.gpart(a.partials + (size_t)cta * c.img_floats);
    const float gl = NJ_LDG(a.grad_loss);
    const int d4 = ((c.d + 3) >> 2) << 2, H4 = ((c.H + 3) >> 2) << 2, inf4 = ((c.inf + 3) >> 2) << 2, do4 = ((c.dout + 3) >> 2) << 2;
    const int wa = P * s.sA;
    const int sI = s.sI, sO = s.sO, sH = s.sH, sD = s.sD;
    {
        NJ_THREADS(tid, nt) {
            if (tid < Pt) {
                const int u = u0 + tid;
                if (u < u1) {
                    const int32_t* dsc = a.b.unit_desc + (size_t)u * 6;
                    const int sc = dsc[5];
                    t.I[NJS_I_PATH * P + tid] = dsc[0]; t.I[NJS_I_S0 * P + tid] = dsc[1]; t.I[NJS_I_LEN * P + tid] = dsc[2] - dsc[1];
                    t.I[NJS_I_ROW * P + tid] = dsc[4] > dsc[3] ? NJ_LDG(a.b.path_rows + dsc[3]) : -1;
                    t.I[NJS_I_FLAG * P + tid] = (sc & NJODE_UNIT_WRITES_HT) ? 1 : 0;
                    t.I[NJS_I_START * P + tid] = (sc & ~NJODE_UNIT_WRITES_HT) - 1;
                } else {
                    t.I[NJS_I_PATH * P + tid] = -1; t.I[NJS_I_S0 * P + tid] = 0; t.I[NJS_I_LEN * P + tid] = 0;
                    t.I[NJS_I_ROW * P + tid] = -1; t.I[NJS_I_FLAG * P + tid] = 0; t.I[NJS_I_START * P + tid] = -1;
                }
            }
        }
        NJ_SYNC();
        int maxlen = 0, any_jump = 0;
        for (int r = 0; r < Pt; ++r) {
            maxlen = t.I[NJS_I_LEN * P + r] > maxlen ? t.I[NJS_I_LEN * P + r] : maxlen;
            any_jump |= (t.I[NJS_I_ROW * P + r] >= 0);
        }
        // ================= the jump at the end of the segment, reversed =================
        // J1-J4 (warp-local): Y_bj = ro(h_before), E = enc(X_obs), Y = ro(E); loss gradients; ro backward at E
        NJ_WARPS(wp, s.nw_b) {
                if (REV::stat && (!REV::glue || wp != 0)) { rev.serve(wp); continue; }
            NJ_SEGB_WARP_VIEW();
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const int r = r0 + er;
                const int p = t.I[NJS_I_PATH * P + r], sr = t.I[NJS_I_START * P + r];
                // gh starts from grad_hT for the units that end their path, else 0
                const float* ght = (t.I[NJS_I_FLAG * P + r] && a.grad_hT) ? a.grad_hT + (size_t)p * c.H : nullptr;
                for (int c_ = ec0; c_ < c.H; c_ += LPR) t.GH[r * sH + c_] = ght ? NJ_LDG(ght + c_) : 0.f;
                for (int c_ = ec0; c_ < d4; c_ += LPR) {      // (last_X, tau) of the segment = its start observation
                    float x = 0.f;
                    if (p >= 0 && c_ < c.d) x = sr < 0 ? NJ_LDG(a.b.start_X + (size_t)p * c.d + c_) : NJ_LDG(a.b.X + (size_t)sr * c.d + c_);
                    t.LX[r * sD + c_] = x; t.TX[r * sD + c_] = c_ < c.d ? nj_tanh(x) : 0.f;
                }
                if (ec0 == 0) t.F[NJS_F_TAU * P + r] = (p >= 0 && sr >= 0) ? NJ_LDG(a.b.jump_tau + NJ_LDG(a.b.row_jump + sr)) : 0.f;
            }
            NJ_SYNCWARP();
            if (a.scratch) {
                // ---- recompute the forward of the tile's segments from their checkpoints: h at the start of a segment is
                // the encoder of its start observation (nothing saved by the forward pass).  The h chain goes to this CTA's
                // scratch [step][row][sH] (L2 resident), h at the segment end stays in HB for the jump reversal. ----
                float* sc = a.scratch + (size_t)cta * a.b.S * P * sH;
                NJ_LANES(lane) {
                    NJ_ROWMAP(R);
                    const int r = r0 + er;
                    const bool valid = t.I[NJS_I_PATH * P + r] >= 0;
                    for (int c_ = ec0; c_ < d4; c_ += LPR) {
                        t.XI[r * sD + c_] = t.LX[r * sD + c_];
                        t.IN[(size_t)r * sI + c_] = t.TX[r * sD + c_];
                    }
                    if (ec0 == 0) { NJ_SEGB_KEY(r, valid, nj_seg_event_of_start(a, t.I[NJS_I_START * P + r])); }
                }
                NJ_SYNCWARP();
                nj_seg_mlp_fwd<TR, REV::coop>(w, NJODE_NET_ENC, true, false);
                NJ_LANES(lane) {
                    NJ_ROWMAP(R);
                    const int r = r0 + er;
                    for (int c_ = ec0; c_ < c.H; c_ += LPR) {
                        float e = t.OUT[r * sO + c_];
                        if (c.residual) e += nj_resid(t.XI + r * sD, c.d, c.H, c_);
                        t.HB[r * sH + c_] = e;
                    }
                }
                NJ_SYNCWARP();
                for (int j = 0; j < maxlen; ++j) {
                    NJ_LANES(lane) {
                        NJ_ROWMAP(R);
                        const int r = r0 + er;
                        const bool active = j < t.I[NJS_I_LEN * P + r];
                        const int k = t.I[NJS_I_S0 * P + r] + j;
                        const float tau = t.F[NJS_F_TAU * P + r];
                        for (int c_ = ec0; c_ < inf4; c_ += LPR) {
                            float v = 0.f;
                            if (c_ < c.d) v = t.TX[r * sD + c_];
                            else if (c_ < c.d + c.H) {
                                const float h = t.HB[r * sH + c_ - c.d];
                                if (active) sc[((size_t)j * P + r) * sH + c_ - c.d] = h;
                                v = nj_tanh(h);
                            } else if (c_ < c.inf) {
                                const float tcur = active ? NJ_LDG(a.b.step_t + k) : 0.f;
                                if (c_ == c.d + c.H) v = tau;
                                else if (c_ == c.d + c.H + 1) v = tcur - tau;
                                else v = tau + (tcur - tau);
                            }
                            t.IN[(size_t)r * sI + c_] = v;
                        }
                        if (ec0 == 0) { NJ_SEGB_KEY(r, true, (unsigned)k); }
                    }
                    NJ_SYNCWARP();
                    nj_seg_mlp_fwd<TR, REV::coop>(w, NJODE_NET_ODE, true, false);
                    NJ_LANES(lane) {
                        NJ_ROWMAP(R);
                        const int r = r0 + er;
                        if (j < t.I[NJS_I_LEN * P + r]) {
                            const float dt = NJ_LDG(a.b.step_dt + t.I[NJS_I_S0 * P + r] + j);
                            for (int c_ = ec0; c_ < c.H; c_ += LPR)
                                t.HB[r * sH + c_] = fmaf(dt, t.OUT[r * sO + c_], t.HB[r * sH + c_]);
                        }
                    }
                    NJ_SYNCWARP();
                }
            }
            if (any_jump) {
                NJ_LANES(lane) {
                    NJ_ROWMAP(R);
                    const int r = r0 + er, row = t.I[NJS_I_ROW * P + r];
                    for (int c_ = ec0; c_ < H4; c_ += LPR) {
                        const float h = (row >= 0 && c_ < c.H) ? (a.scratch ? t.HB[r * sH + c_] : a.h_before[(size_t)row * c.H + c_]) : 0.f;
                        if (c_ < c.H) t.HB[r * sH + c_] = h;
                        t.IN[(size_t)r * sI + c_] = c_ < c.H ? nj_tanh(h) : 0.f;
                    }
                    if (ec0 == 0) { NJ_SEGB_KEY(r, row >= 0, nj_seg_event_of_jump(a, row >= 0 ? row : 0, 0u)); }
                }
                NJ_SYNCWARP();
                nj_seg_mlp_fwd<TR, REV::coop>(w, NJODE_NET_RO, true, false);
                NJ_LANES(lane) {
                    NJ_ROWMAP(R);
                    const int r = r0 + er, row = t.I[NJS_I_ROW * P + r];
                    for (int c_ = ec0; c_ < c.dout; c_ += LPR) {
                        float y = t.OUT[r * sO + c_];
                        if (c.residual) y += nj_resid(t.HB + r * sH, c.H, c.dout, c_);
                        t.YBJ[r * sD + c_] = y;
                    }
                    for (int c_ = ec0; c_ < d4; c_ += LPR) {
                        const float x = (row >= 0 && c_ < c.d) ? NJ_LDG(a.b.X + (size_t)row * c.d + c_) : 0.f;
                        t.XI[r * sD + c_] = x;
                        t.IN[(size_t)r * sI + c_] = c_ < c.d ? nj_tanh(x) : 0.f;
                    }
                    if (ec0 == 0) { NJ_SEGB_KEY(r, row >= 0, nj_seg_event_of_jump(a, row >= 0 ? row : 0, 1u)); }
                }
                NJ_SYNCWARP();
                nj_seg_mlp_fwd<TR, REV::coop>(w, NJODE_NET_ENC, true, false);
                NJ_LANES(lane) {
                    NJ_ROWMAP(R);
                    const int r = r0 + er, row = t.I[NJS_I_ROW * P + r];
                    for (int c_ = ec0; c_ < H4; c_ += LPR) {
                        float e = 0.f;
                        if (c_ < c.H) {
                            e = t.OUT[r * sO + c_];
                            if (c.residual) e += nj_resid(t.XI + r * sD, c.d, c.H, c_);
                            t.EE[r * sH + c_] = e;
                        }
                        t.IN[(size_t)r * sI + c_] = c_ < c.H ? nj_tanh(e) : 0.f;
                    }
                    if (ec0 == 0) { NJ_SEGB_KEY(r, row >= 0, nj_seg_event_of_jump(a, row >= 0 ? row : 0, 2u)); }
                }
                NJ_SYNCWARP();
                nj_seg_mlp_fwd<TR, REV::coop>(w, NJODE_NET_RO, true, false);
                // loss derivative (compute_loss / compute_loss_2, NJODE/models.py:71-126)
                NJ_LANES(lane) {
                    if (lane < R) {
                        const int r = r0 + lane, row = t.I[NJS_I_ROW * P + r];
                        float ca = 0.f, cb = 0.f;
                        if (row >= 0) {
                            float sa = 0.f, sb = 0.f;
                            NJ_UNROLL4
                    for (int c_ = 0; c_ < c.dout; ++c_) {      // (unrolled: the independent loads of four features go out together)
                                float y = t.OUT[r * sO + c_];
                                if (c.residual) y += nj_resid(t.EE + r * sH, c.H, c.dout, c_);
                                t.YY[r * sD + c_] = y;
                                const float x = t.XI[r * sD + c_], yb = t.YBJ[r * sD + c_];
                                const float da = x - y, db = (c.loss_kind == NJODE_LOSS_STANDARD) ? (yb - y) : (yb - x);
                                sa = fmaf(da, da, sa); sb = fmaf(db, db, sb);
                            }
                            const float ra = sqrtf(sa + 1e-10f), rb = sqrtf(sb + 1e-10f);
                            const float wa_ = (c.loss_kind == NJODE_LOSS_STANDARD) ? 2.f * c.w : c.w;
                            const float wb_ = (c.loss_kind == NJODE_LOSS_STANDARD) ? 2.f * (1.f - c.w) : (1.f - c.w);
                            const float sm = wa_ * ra + wb_ * rb;
                            const float cf = gl * 2.f * sm / (NJ_LDG(a.b.n_obs_ot + t.I[NJS_I_PATH * P + r]) * (float)a.b.batch_size_norm);
                            ca = cf * wa_ / ra; cb = cf * wb_ / rb;
                        }
                        t.F[NJS_F_CA * P + r] = ca; t.F[NJS_F_CB * P + r] = cb;
                    }
                }
                NJ_SYNCWARP();
                NJ_LANES(lane) {
                    NJ_ROWMAP(R);
                    const int r = r0 + er;
                    const bool has = t.I[NJS_I_ROW * P + r] >= 0;
                    for (int c_ = ec0; c_ < do4; c_ += LPR) {
                        float gy = 0.f, gyb = 0.f;
                        if (c_ < c.dout && has) {
                            const float x = t.XI[r * sD + c_], y = t.YY[r * sD + c_], yb = t.YBJ[r * sD + c_];
                            const float ca = t.F[NJS_F_CA * P + r], cb = t.F[NJS_F_CB * P + r];
                            if (c.loss_kind == NJODE_LOSS_STANDARD) { gy = -ca * (x - y) - cb * (yb - y); gyb = cb * (yb - y); }
                            else { gy = -ca * (x - y); gyb = cb * (yb - x); }
                        }
                        t.GOUT[(size_t)r * sO + c_] = gy;
                        t.GYBJ[r * sD + c_] = gyb;
                    }
                }
                NJ_SYNCWARP();
                nj_seg_mlp_dx<TR, REV::coop>(w, NJODE_NET_RO, true);
                NJ_LANES(lane) {
                    NJ_ROWMAP(R);
                    const int r = r0 + er;
                    for (int c_ = ec0; c_ < c.H; c_ += LPR) {
                        const float th = t.IN[(size_t)r * sI + c_];
                        float ge = t.GZ[(size_t)r * sI + c_] * (1.f - th * th);
                        if (c.residual) ge += nj_resid_bwd(t.GOUT + (size_t)r * sO, c.H, c.dout, c_);
                        t.GE[r * sH + c_] = ge;
                    }
                }
            }
            rev.glue_done();
        }
        if (any_jump) {
            NJ_SYNC();
            NJ_THREADS(tid, nt) { nj_seg_dw<REV::ovf>(c, s, t, NJODE_NET_RO, NJ_ACC(tid), gpart, tid, nt, Pt); }
            NJ_SYNC();
            // J5: encoder at X_obs, backward with g = dL/dE (from Y only)
            NJ_WARPS(wp, s.nw_b) {
                if (REV::stat && (!REV::glue || wp != 0)) { rev.serve(wp); continue; }
                NJ_SEGB_WARP_VIEW();
                NJ_LANES(lane) {
                    NJ_ROWMAP(R);
                    const int r = r0 + er, row = t.I[NJS_I_ROW * P + r];
                    for (int c_ = ec0; c_ < d4; c_ += LPR) t.IN[(size_t)r * sI + c_] = c_ < c.d ? nj_tanh(t.XI[r * sD + c_]) : 0.f;
                    for (int c_ = ec0; c_ < H4; c_ += LPR) t.GOUT[(size_t)r * sO + c_] = c_ < c.H ? t.GE[r * sH + c_] : 0.f;
                    if (ec0 == 0) { NJ_SEGB_KEY(r, row >= 0, nj_seg_event_of_jump(a, row >= 0 ? row : 0, 1u)); }
                }
                NJ_SYNCWARP();
                nj_seg_mlp_fwd<TR, REV::coop>(w, NJODE_NET_ENC, true, true);
                nj_seg_mlp_dx<TR, REV::coop>(w, NJODE_NET_ENC, false);
                rev.glue_done();
            }
            NJ_SYNC();
            NJ_THREADS(tid, nt) { nj_seg_dw<REV::ovf>(c, s, t, NJODE_NET_ENC, NJ_ACC(tid), gpart, tid, nt, Pt); }
            NJ_SYNC();
            // J6: readout at h_before, backward with g = dL/dY_bj -> gradient wrt h at the segment end
            NJ_WARPS(wp, s.nw_b) {
                if (REV::stat && (!REV::glue || wp != 0)) { rev.serve(wp); continue; }
                NJ_SEGB_WARP_VIEW();
                NJ_LANES(lane) {
                    NJ_ROWMAP(R);
                    const int r = r0 + er, row = t.I[NJS_I_ROW * P + r];
                    for (int c_ = ec0; c_ < H4; c_ += LPR) {
                        t.IN[(size_t)r * sI + c_] = c_ < c.H ? nj_tanh(t.HB[r * sH + c_]) : 0.f;
                        t.GOUT[(size_t)r * sO + c_] = c_ < c.dout ? t.GYBJ[r * sD + c_] : 0.f;     // also clears J5's H4 columns
                    }
                    if (ec0 == 0) { NJ_SEGB_KEY(r, row >= 0, nj_seg_event_of_jump(a, row >= 0 ? row : 0, 0u)); }
                }
                NJ_SYNCWARP();
                nj_seg_mlp_fwd<TR, REV::coop>(w, NJODE_NET_RO, true, true);
                nj_seg_mlp_dx<TR, REV::coop>(w, NJODE_NET_RO, true);
                NJ_LANES(lane) {
                    NJ_ROWMAP(R);
                    const int r = r0 + er;
                    if (t.I[NJS_I_ROW * P + r] >= 0) {
                        for (int c_ = ec0; c_ < c.H; c_ += LPR) {
                            const float th = t.IN[(size_t)r * sI + c_];
                            float gh = t.GZ[(size_t)r * sI + c_] * (1.f - th * th);
                            if (c.residual) gh += nj_resid_bwd(t.GOUT + (size_t)r * sO, c.H, c.dout, c_);
                            t.GH[r * sH + c_] = gh;
                        }
                    }
                }
                rev.glue_done();
            }
            NJ_SYNC();
            NJ_THREADS(tid, nt) { nj_seg_dw<REV::ovf>(c, s, t, NJODE_NET_RO, NJ_ACC(tid), gpart, tid, nt, Pt); }
            NJ_SYNC();
        }
        // ================= Euler steps, reversed =================
        if (REV::stat) { NJ_SYNC(); rev.run(maxlen); }   // (GH / TX / tau of warp 0's prelude and jump reversal are visible to every warp)
        for (int j = REV::stat ? -1 : maxlen - 1; j >= 0; --j) {
            NJ_WARPS(wp, s.nw_b) {
                if (REV::stat && (!REV::glue || wp != 0)) { rev.serve(wp); continue; }
                NJ_SEGB_WARP_VIEW();
                NJ_LANES(lane) {
                    NJ_ROWMAP(R);
                    const int r = r0 + er;
                    const bool active = j < t.I[NJS_I_LEN * P + r];
                    const int k = t.I[NJS_I_S0 * P + r] + j;
                    const float tau = t.F[NJS_F_TAU * P + r];
                    const float dt = active ? NJ_LDG(a.b.step_dt + k) : 0.f;
                    const float* hh = nullptr;
                    if (active) hh = a.scratch ? a.scratch + ((size_t)cta * a.b.S * P + (size_t)j * P + r) * sH
                                               : a.h_hist + ((size_t)k * a.b.B + t.I[NJS_I_PATH * P + r]) * c.H;
                    for (int c_ = ec0; c_ < inf4; c_ += LPR) {
                        float v = 0.f;
                        if (c_ < c.d) v = t.TX[r * sD + c_];
                        else if (c_ < c.d + c.H) v = nj_tanh(hh ? hh[c_ - c.d] : 0.f);
                        else if (c_ < c.inf) {
                            const float tcur = active ? NJ_LDG(a.b.step_t + k) : 0.f;
                            if (c_ == c.d + c.H) v = tau;
                            else if (c_ == c.d + c.H + 1) v = tcur - tau;
                            else v = tau + (tcur - tau);
                        }
                        t.IN[(size_t)r * sI + c_] = v;
                    }
                    for (int c_ = ec0; c_ < H4; c_ += LPR) t.GOUT[(size_t)r * sO + c_] = c_ < c.H ? dt * t.GH[r * sH + c_] : 0.f;
                    if (ec0 == 0) { NJ_SEGB_KEY(r, true, (unsigned)k); }
                }
                NJ_SYNCWARP();
                if (a.act_hist) {
                    // the forward pass saved the hidden activations of this step: read them instead of recomputing two layers
                    nj_seg_act_load<R>(a, 0, w.A0, s.sA, t.I + r0, P, t.I + NJS_I_PATH * P + r0, j);
                    if (a.act_nh > 1) nj_seg_act_load<R>(a, 1, w.A0 + wa, s.sA, t.I + r0, P, t.I + NJS_I_PATH * P + r0, j);
                    NJ_SYNCWARP();
                } else nj_seg_mlp_fwd<TR, REV::coop>(w, NJODE_NET_ODE, true, true);
                nj_seg_mlp_dx<TR, REV::coop>(w, NJODE_NET_ODE, true);
                NJ_LANES(lane) {
                    NJ_ROWMAP(R);
                    const int r = r0 + er;
                    if (j < t.I[NJS_I_LEN * P + r]) {
                        for (int c_ = ec0; c_ < c.H; c_ += LPR) {
                            const float th = t.IN[(size_t)r * sI + c.d + c_];
                            t.GH[r * sH + c_] += t.GZ[(size_t)r * sI + c.d + c_] * (1.f - th * th);
                        }
                    }
                }
                rev.glue_done();
            }
            NJ_SYNC();
            NJ_THREADS(tid, nt) { nj_seg_dw<REV::ovf>(c, s, t, NJODE_NET_ODE, NJ_ACC(tid), gpart, tid, nt, Pt); }
            NJ_SYNC();
        }
        // ================= the start encoder, reversed =================
        NJ_WARPS(wp, s.nw_b) {
                if (REV::stat && (!REV::glue || wp != 0)) { rev.serve(wp); continue; }
            NJ_SEGB_WARP_VIEW();
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const int r = r0 + er;
                const bool valid = t.I[NJS_I_PATH * P + r] >= 0;
                for (int c_ = ec0; c_ < d4; c_ += LPR) t.IN[(size_t)r * sI + c_] = t.TX[r * sD + c_];
                for (int c_ = ec0; c_ < H4; c_ += LPR) t.GOUT[(size_t)r * sO + c_] = (c_ < c.H && valid) ? t.GH[r * sH + c_] : 0.f;
                if (ec0 == 0) { NJ_SEGB_KEY(r, valid, nj_seg_event_of_start(a, t.I[NJS_I_START * P + r])); }
            }
            NJ_SYNCWARP();
            nj_seg_mlp_fwd<TR, REV::coop>(w, NJODE_NET_ENC, true, true);
            nj_seg_mlp_dx<TR, REV::coop>(w, NJODE_NET_ENC, false);
            rev.glue_done();
        }
        NJ_SYNC();
        NJ_THREADS(tid, nt) { nj_seg_dw<REV::ovf>(c, s, t, NJODE_NET_ENC, NJ_ACC(tid), gpart, tid, nt, Pt); }
        NJ_SYNC();
    }
}

#if !defined(NJODE_HOST_SIM)
// the dW helper warps (threads >= 32 * nw_b) of a backward CTA: they own no rows, so they skip the warp-local phases
// of nj_seg_bwd_tile and only mirror its barriers and its CTA-wide dW phases.  (On the host simulation the NJ_THREADS
// loops of nj_seg_bwd_tile already run the helper thread ids.)  Kept out of nj_seg_bwd_tile so that the row warps run
// exactly the code they ran without helpers.
template <int TR, bool OVF>
__device__ __forceinline__ void nj_seg_bwd_tile_helper(const NjCfg& c, const NjSeg& s, const NjArgs& a, const NjSegB& t,
                                                       float* acc, int cta) {
    constexpr int R = 4 * TR;
    const int P = s.P_b, nt = s.nt_b, Pt = R * s.nw_b, tid = threadIdx.x;
    float* gpart = a.partials + (size_t)cta * c.img_floats;
    NJ_SYNC();                                               // unit descriptors are in shared memory
    int maxlen = 0, any_jump = 0;
    for (int r = 0; r < Pt; ++r) {
        maxlen = t.I[NJS_I_LEN * P + r] > maxlen ? t.I[NJS_I_LEN * P + r] : maxlen;
        any_jump |= (t.I[NJS_I_ROW * P + r] >= 0);
    }
    if (any_jump) {
        NJ_SYNC(); nj_seg_dw<OVF>(c, s, t, NJODE_NET_RO, acc, gpart, tid, nt, Pt); NJ_SYNC();
        NJ_SYNC(); nj_seg_dw<OVF>(c, s, t, NJODE_NET_ENC, acc, gpart, tid, nt, Pt); NJ_SYNC();
        NJ_SYNC(); nj_seg_dw<OVF>(c, s, t, NJODE_NET_RO, acc, gpart, tid, nt, Pt); NJ_SYNC();
    }
    for (int j = maxlen - 1; j >= 0; --j) {
        NJ_SYNC(); nj_seg_dw<OVF>(c, s, t, NJODE_NET_ODE, acc, gpart, tid, nt, Pt); NJ_SYNC();
    }
    NJ_SYNC(); nj_seg_dw<OVF>(c, s, t, NJODE_NET_ENC, acc, gpart, tid, nt, Pt); NJ_SYNC();
}
#endif

// HELP: the launch has dW helper warps (a second instantiation of the kernel, so that launches without helpers run
// exactly the code -- and the register allocation -- they had before helpers existed)
template <bool HELP, bool OVF = true>
NJ_HD void nj_seg_cta_backward(const NjCfg& c, const NjSeg& s, const NjArgs& a, float* smem, int cta) {
    const int nt = s.nt_b;
    float* simg = smem + s.b_img;
    nj_stage_image(simg, a.image, c.img_floats, nt);
    nj_zero(smem + s.b_IN, s.b_smem_floats - s.b_IN, nt);
    nj_zero(a.partials + (size_t)cta * c.img_floats, c.img_floats, nt);      // overflow tiles add into it
    NJ_SYNC();
    NjSegB t;
    nj_segb_bind(t, s, smem);
    NJ_ACC_DECL(nt);
    int* ctl = t.I + NJS_I_COUNT * s.P_b;
    for (;;) {
        NJ_THREADS(tid, nt) { if (tid == 0) ctl[0] = nj_atomic_inc(a.counter); }
        NJ_SYNC();
        const int tile = ctl[0];
        NJ_SYNC();
        if (tile >= s.n_tiles_b) break;
        int ub, ue;
        const int tr = nj_seg_tile_lookup(s.b_ncls, s.b_t0, s.b_u0, s.b_u1, s.b_tr, 4 * s.nw_b, tile, ub, ue);
#if !defined(NJODE_HOST_SIM)
        if (HELP && (int)(threadIdx.x >> 5) >= s.nw_b) {
            if (tr == 2) nj_seg_bwd_tile_helper<2, OVF>(c, s, a, t, nj_acc_base, cta);
            else nj_seg_bwd_tile_helper<1, OVF>(c, s, a, t, nj_acc_base, cta);
            continue;
        }
#endif
        if (tr == 2) nj_seg_bwd_tile<2, NjSegWarpRevT<OVF>>(c, s, a, smem, t, nj_acc_base, cta, ub, ue);
        else nj_seg_bwd_tile<1, NjSegWarpRevT<OVF>>(c, s, a, smem, t, nj_acc_base, cta, ub, ue);
    }
    float* gpart = a.partials + (size_t)cta * c.img_floats;
    NJ_THREADS(tid, nt) { nj_seg_dw_flush(c, s, NJ_ACC(tid), gpart, tid, nt); }
}
