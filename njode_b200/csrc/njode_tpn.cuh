// njode_tpn.cuh -- Euler steps of SMALL whole-path batches with one thread per neuron and role-specialised warps.
//
// The reference's PhysioNet batch is 50 records x 3 000 Euler steps (NJODE/parallel_train.py:656, NJODE/models.py:431-447):
// 50 CTAs, each marching through 3 000 dependent steps.  What bounds such a launch is the latency of ONE step, and the
// K-split lane layout of the weight-stationary kernels (njode_path.cuh: 13 warps, 8 lanes per output, three shuffle
// rounds, four CTA barriers per step) spends ~900 cycles per layer on ~70 instructions per warp (ncu source page,
// profiles/r2e_*: 35 % barrier, 29 % short-scoreboard stalls).  Here a layer is ONE pass of K FFMA per thread:
//   F threads (2 warps): thread o holds row o of the three ODE weight matrices in registers and evaluates output o of
//       a layer for the R rows of the tile (activation rows are broadcast reads of shared memory);
//   T threads (3 warps, backward): thread k holds COLUMN k of the three matrices and evaluates the input gradient k;
//   D threads (2 warps, backward): own the 4x4 dW tiles of the ODE network (registers, for the whole launch);
//   the glue warp runs the jumps, the start encoder and the bookkeeping with the warp GEMMs of njode_path.cuh.
// The backward is a three-stage software pipeline over the steps between two jumps: in the same three phases F rebuilds
// the activations of step e - 1, T reverses step e, D accumulates the dW of step e + 1 (operand buffers exist three
// times); F and T meet at a named barrier after each phase, the whole CTA once per step.
//
// Dimension classes (compile-time register tiles): see NjTpnA / NjTpnB below; other networks keep the kernels of
// njode_path.cuh.  Same dual compilation as the other kernel sources (device / sequential host simulation).
#pragma once
#include "njode_path.cuh"

#define NJN_G 32                    // glue warp
#define NJN_F 64                    // thread o = output o
#define NJN_T 96                    // thread k = input k
#define NJN_D 64                    // dW tile owners
#define NJN_F0 NJN_G
#define NJN_T0 (NJN_G + NJN_F)
#define NJN_D0 (NJN_G + NJN_F + NJN_T)
#define NJN_NT_FWD (NJN_G + NJN_F)
#define NJN_NT_BWD (NJN_G + NJN_F + NJN_T + NJN_D)
#define NJN_DSLOTS 10
#define NJN_DACC (NJN_DSLOTS * 20)
enum { NJN_ALL = 0, NJN_ROLE_G, NJN_ROLE_F, NJN_ROLE_T, NJN_ROLE_D };

// KC0 / KCH: float4 chunks of the ODE network's input row / of a hidden layer, HC: of the hidden state; R: rows per tile
template <int KC0_, int KCH_, int HC_, int R_> struct NjTpnDims {
    static constexpr int KC0 = KC0_, KCH = KCH_, HC = HC_, R = R_;
};
// A: the demo networks (input d + H + 2 <= 16, hidden layers <= 52, H <= 12);  B: the PhysioNet-shaped ones (84 / 52 / 44)
#define NJN_A_KC0 4
#define NJN_A_KCH 13
#define NJN_A_HC 3
#define NJN_B_KC0 21
#define NJN_B_KCH 13
#define NJN_B_HC 11

#if defined(NJODE_HOST_SIM)
#define NJN_ROLE(ROLEID, lo, n, IDX) for (int IDX = 0; IDX < (n); ++IDX)
#define NJN_SYNC_F() ((void)0)
#define NJN_SYNC_STEP() ((void)0)
#define NJN_SYNC_FT() ((void)0)
#else
#define NJN_ROLE(ROLEID, lo, n, IDX) if (ROLE == ROLEID) for (int IDX = (int)threadIdx.x - (lo), _nj_xe = IDX + 1; IDX < _nj_xe; ++IDX)
#define NJN_SYNC_F() do { if (ROLE == NJN_ROLE_F) asm volatile("bar.sync 1, 64;" ::: "memory"); } while (0)
// end of a forward step: the F warps and the glue warp (which prepared the next step's scalars meanwhile)
#define NJN_SYNC_STEP() do { if (ROLE == NJN_ROLE_F || ROLE == NJN_ROLE_G) asm volatile("bar.sync 3, 96;" ::: "memory"); } while (0)
#define NJN_SYNC_FT() do { if (ROLE == NJN_ROLE_F || ROLE == NJN_ROLE_T) asm volatile("bar.sync 1, 160;" ::: "memory"); } while (0)
#endif

// 4-byte asynchronous copy global -> shared (prefetch of next step's h / time / step size: a register prefetch shares its
// scoreboard with the loads of the current step and stalled their first use -- ncu source page, profiles/r2l_*)
#if defined(NJODE_HOST_SIM)
NJ_HD void nj_cp_async4(float* dst, const float* src) { *dst = *src; }
NJ_HD void nj_cp_wait() {}
#else
__device__ __forceinline__ void nj_cp_async4(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void nj_cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

template <class D> struct NjTpnF {
    float w0[4 * D::KC0], w1[4 * D::KCH], w2[4 * D::KCH];
    float b0, b1, b2;
};
template <class D> struct NjTpnT {
    float c2[4 * D::HC], c1[4 * D::KCH], c0[4 * D::KCH];
};

template <class D>
NJ_HD void nj_tpn_f_load(const NjCfg& c, const float* simg, int o, NjTpnF<D>& f) {
    const NjNet& N = c.net[NJODE_NET_ODE];
#pragma unroll
    for (int j = 0; j < 4 * D::KC0; ++j) f.w0[j] = (o < N.dim[1] && j < N.dim[0]) ? simg[N.w_img[0] + o * N.ks[0] + j] : 0.f;
#pragma unroll
    for (int j = 0; j < 4 * D::KCH; ++j) f.w1[j] = (o < N.dim[2] && j < N.dim[1]) ? simg[N.w_img[1] + o * N.ks[1] + j] : 0.f;
#pragma unroll
    for (int j = 0; j < 4 * D::KCH; ++j) f.w2[j] = (o < N.dim[3] && j < N.dim[2]) ? simg[N.w_img[2] + o * N.ks[2] + j] : 0.f;
    f.b0 = (o < N.dim[1] && N.b_src[0] >= 0) ? simg[N.b_img[0] + o] : 0.f;
    f.b1 = (o < N.dim[2] && N.b_src[1] >= 0) ? simg[N.b_img[1] + o] : 0.f;
    f.b2 = (o < N.dim[3] && N.b_src[2] >= 0) ? simg[N.b_img[2] + o] : 0.f;
}

template <class D>
NJ_HD void nj_tpn_t_load(const NjCfg& c, const float* simg, int k, NjTpnT<D>& t) {
    const NjNet& N = c.net[NJODE_NET_ODE];
#pragma unroll
    for (int o = 0; o < 4 * D::HC; ++o) t.c2[o] = (o < N.dim[3] && k < N.dim[2]) ? simg[N.w_img[2] + o * N.ks[2] + k] : 0.f;
#pragma unroll
    for (int o = 0; o < 4 * D::KCH; ++o) t.c1[o] = (o < N.dim[2] && k < N.dim[1]) ? simg[N.w_img[1] + o * N.ks[1] + k] : 0.f;
#pragma unroll
    for (int o = 0; o < 4 * D::KCH; ++o) t.c0[o] = (o < N.dim[1] && k < N.dim[0]) ? simg[N.w_img[0] + o * N.ks[0] + k] : 0.f;
}

// acc[r] = sum_j w[j] * x[r][j] over KC float4 chunks.  Four partial sums per row, eight for single-row tiles (even / odd
// chunks): one thread alone on its scheduler slot needs that many independent FFMA chains to cover the FFMA latency
template <int KC, int R>
NJ_HD void nj_tpn_dot(const float* w, const float* x, int x_s, float* acc) {
    constexpr int NP = R == 1 ? 8 : 4;
    float p[R][NP];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int i = 0; i < NP; ++i) p[r][i] = 0.f;
    const nj_sp xp = nj_sp_of(x);
#pragma unroll
    for (int q = 0; q < KC; ++q) {
        const int b = NP == 8 ? 4 * (q & 1) : 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const nj_f4 v = nj_sp_ld4(NJ_SP_ADD(xp, r * x_s + 4 * q));
            p[r][b] = fmaf(w[4 * q], v.x, p[r][b]); p[r][b + 1] = fmaf(w[4 * q + 1], v.y, p[r][b + 1]);
            p[r][b + 2] = fmaf(w[4 * q + 2], v.z, p[r][b + 2]); p[r][b + 3] = fmaf(w[4 * q + 3], v.w, p[r][b + 3]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (NP == 8) acc[r] = ((p[r][0] + p[r][4]) + (p[r][1] + p[r][5])) + ((p[r][2] + p[r][6]) + (p[r][3] + p[r][7]));
        else acc[r] = (p[r][0] + p[r][1]) + (p[r][2] + p[r][3]);
    }
}

// hidden layer l of the ODE network, output o, all rows: out[r][o] = dropout(act(w . in[r] + b))
template <int KC, int R>
NJ_HD void nj_tpn_hidden(const NjCfg& c, int l, const float* w, float bias, int o, const float* in, int in_s, float* out, int out_s,
                         const int* rk, const int* lkeys = nullptr) {
    const NjNet& N = c.net[NJODE_NET_ODE];
    float acc[R];
    nj_tpn_dot<KC, R>(w, in, in_s, acc);
    if (o >= (((N.dim[l + 1] + 3) >> 2) << 2)) return;
    const bool pad = o >= N.dim[l + 1];          // columns up to the next multiple of 4 are read by the dW tiles: keep them 0
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float v = nj_act(acc[r] + bias, N.act[l]);
        if (c.has_drop) {
            // (lkeys: the layer keys of the rows, hashed ahead of time by the glue warp -- the same for every output)
            const unsigned lk = lkeys ? (unsigned)lkeys[r] : nj_layer_key((unsigned)rk[r], (unsigned)(NJODE_NET_ODE * 16 + l + 1));
            v = nj_keep(lk, (unsigned)o, c.thr) ? v * c.keep_scale : nj_u2f(NJ_DROPPED);
        }
        out[(size_t)r * out_s + o] = pad ? 0.f : v;
    }
}

// one input row of the ODE network, column c_ (ODEFunc.forward, NJODE/models.py:188-199): [tanh(last_X), tanh(h), tau, t - tau(, t)]
NJ_HD float nj_tpn_time_col(const NjCfg& c, int c_, float tau, float tcur) {
    if (c_ == c.d + c.H) return tau;
    if (c_ == c.d + c.H + 1) return tcur - tau;
    return tau + (tcur - tau);
}

// ================================================================================================
// cooperative layers for the glue warp: the jump networks (readout, encoder, GRU) live in the shared-memory parameter image;
// a lone warp needs ~2 us per layer for them (K-split lanes, shuffles), which made the ~75 jumps of a PhysioNet record
// 45 % of the forward and a third of the backward kernel (ncu source page, profiles/r2l_*).  The glue warp keeps its
// bookkeeping code; wherever it evaluates a layer it posts the layer descriptor to a mailbox and the 64 F threads -- idle
// during a jump -- compute it, one output (forward) or one float4 group of inputs (input gradient) per thread.
// ================================================================================================
struct NjCoopMB { int op, o_store, R, pad; NjWL L; NjWD D; };       // op: 0 done, 1 forward layer, 2 input gradient

template <int R>
NJ_HD void nj_coop_fwd_rows(const NjWL& L, int o_store, int o) {
    for (int oo = o; oo < o_store; oo += NJN_F) {
        float p[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) { p[r][0] = 0.f; p[r][1] = 0.f; p[r][2] = 0.f; p[r][3] = 0.f; }
        const nj_sp wp = nj_sp_of(L.W + (size_t)oo * L.w_s), ap = nj_sp_of(L.in);
#pragma unroll 4
        for (int k4 = 0; k4 < L.K4; ++k4) {
            const nj_f4 w = nj_sp_ld4(NJ_SP_ADD(wp, 4 * k4));
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const nj_f4 v = nj_sp_ld4(NJ_SP_ADD(ap, r * L.in_s + 4 * k4));
                p[r][0] = fmaf(w.x, v.x, p[r][0]); p[r][1] = fmaf(w.y, v.y, p[r][1]);
                p[r][2] = fmaf(w.z, v.z, p[r][2]); p[r][3] = fmaf(w.w, v.w, p[r][3]);
            }
        }
        const float bias = L.bias ? L.bias[oo] : 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float v = nj_act(((p[r][0] + p[r][1]) + (p[r][2] + p[r][3])) + bias, L.act);
            if (L.drop) v = nj_keep(nj_layer_key((unsigned)L.rk[r], L.tag), (unsigned)oo, L.thr) ? v * L.keep_scale : nj_u2f(NJ_DROPPED);
            L.out[(size_t)r * L.out_s + oo] = v;
        }
    }
}

// gin[r][4kg..] = (sum_o g[r][o] W[o][4kg..]) * act'(aprev) * dropout factor  (nj_pg_dx, njode_path.cuh, without the K split)
template <int R>
NJ_HD void nj_coop_dx_rows(const NjWD& L, int o) {
    for (int kg = o; kg < L.K4in; kg += NJN_F) {
        float acc[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }
        const nj_sp wp = nj_sp_of(L.W + 4 * kg), gp = nj_sp_of(L.g);
        const int ws = L.w_s;
        for (int o4 = 0; o4 < L.O4; ++o4) {
            const nj_sp q = NJ_SP_ADD(wp, 4 * o4 * ws);
            const nj_f4 w0 = nj_sp_ld4(q), w1 = nj_sp_ld4(NJ_SP_ADD(q, ws)), w2 = nj_sp_ld4(NJ_SP_ADD(q, 2 * ws)), w3 = nj_sp_ld4(NJ_SP_ADD(q, 3 * ws));
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const nj_f4 gv = nj_sp_ld4(NJ_SP_ADD(gp, r * L.g_s + 4 * o4));
                acc[r][0] = fmaf(gv.x, w0.x, fmaf(gv.y, w1.x, fmaf(gv.z, w2.x, fmaf(gv.w, w3.x, acc[r][0]))));
                acc[r][1] = fmaf(gv.x, w0.y, fmaf(gv.y, w1.y, fmaf(gv.z, w2.y, fmaf(gv.w, w3.y, acc[r][1]))));
                acc[r][2] = fmaf(gv.x, w0.z, fmaf(gv.y, w1.z, fmaf(gv.z, w2.z, fmaf(gv.w, w3.z, acc[r][2]))));
                acc[r][3] = fmaf(gv.x, w0.w, fmaf(gv.y, w1.w, fmaf(gv.z, w2.w, fmaf(gv.w, w3.w, acc[r][3]))));
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float v[4] = {acc[r][0], acc[r][1], acc[r][2], acc[r][3]};
            if (L.aprev) {
                const nj_f4 av = nj_sp_ld4(nj_sp_of(L.aprev + (size_t)r * L.a_s + 4 * kg));
                const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float a_ = a4[c];
                    if (L.drop) {
                        if (nj_f2u(a_) == NJ_DROPPED) v[c] = 0.f;
                        else { a_ *= L.one_minus_p; v[c] *= L.keep_scale; }
                    }
                    if (L.act_prev == NJODE_ACT_TANH) v[c] *= (1.f - a_ * a_);
                    else if (L.act_prev == NJODE_ACT_RELU) v[c] = a_ > 0.f ? v[c] : 0.f;
                }
            }
            nj_f4 ov; ov.x = v[0]; ov.y = v[1]; ov.z = v[2]; ov.w = v[3];
            nj_st4(L.gin + (size_t)r * L.gin_s + 4 * kg, ov);
        }
    }
}

NJ_HD void nj_coop_run(const NjCoopMB* mb, int op, int o) {
    if (op == 1) { if (mb->R == 1) nj_coop_fwd_rows<1>(mb->L, mb->o_store, o); else nj_coop_fwd_rows<4>(mb->L, mb->o_store, o); }
    else { if (mb->R == 1) nj_coop_dx_rows<1>(mb->D, o); else nj_coop_dx_rows<4>(mb->D, o); }
}

#if defined(NJODE_HOST_SIM)
NJ_HD void nj_coop_post_fwd(void* mailbox, const NjWL& L, int o_store) {
    NjCoopMB* mb = static_cast<NjCoopMB*>(mailbox);
    mb->L = L; mb->o_store = o_store;
    for (int o = 0; o < NJN_F; ++o) nj_coop_run(mb, 1, o);
}
NJ_HD void nj_coop_post_dx(void* mailbox, const NjWD& D) {
    NjCoopMB* mb = static_cast<NjCoopMB*>(mailbox);
    mb->D = D;
    for (int o = 0; o < NJN_F; ++o) nj_coop_run(mb, 2, o);
}
#define NJN_COOP_DONE(mb) ((void)0)
#define NJN_COOP_SERVE(mb, o) ((void)0)
#else
// glue warp + F threads = 96 threads on named barrier 2
#define NJN_COOP_BAR() asm volatile("bar.sync 2, 96;" ::: "memory")
NJ_HD void nj_coop_post_fwd(void* mailbox, const NjWL& L, int o_store) {
    NjCoopMB* mb = static_cast<NjCoopMB*>(mailbox);
    if ((threadIdx.x & 31) == 0) { mb->L = L; mb->o_store = o_store; mb->op = 1; }
    __syncwarp();
    NJN_COOP_BAR();
    NJN_COOP_BAR();
}
NJ_HD void nj_coop_post_dx(void* mailbox, const NjWD& D) {
    NjCoopMB* mb = static_cast<NjCoopMB*>(mailbox);
    if ((threadIdx.x & 31) == 0) { mb->D = D; mb->op = 2; }
    __syncwarp();
    NJN_COOP_BAR();
    NJN_COOP_BAR();
}
__device__ __forceinline__ void nj_coop_done(NjCoopMB* mb) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mb->op = 0;
    __syncwarp();
    NJN_COOP_BAR();
}
__device__ __forceinline__ void nj_coop_serve(const NjCoopMB* mb, int o) {
    for (;;) {
        NJN_COOP_BAR();
        const int op = *reinterpret_cast<const volatile int*>(&mb->op);
        if (!op) break;
        nj_coop_run(mb, op, o);
        NJN_COOP_BAR();
    }
}
#define NJN_COOP_DONE(mb) nj_coop_done(mb)
#define NJN_COOP_SERVE(mb, o) nj_coop_serve(mb, o)
#endif

// a section of glue code (warp 0) whose layers the F threads serve
#if defined(NJODE_HOST_SIM)
#define NJN_GLUE(mb, stmt) do { stmt; } while (0)
#else
#define NJN_GLUE(mb, stmt)                                                                        \
    do {                                                                                          \
        if (ROLE == NJN_ROLE_G) { stmt; NJN_COOP_DONE(mb); }                                      \
        else if (ROLE == NJN_ROLE_F) NJN_COOP_SERVE(mb, (int)threadIdx.x - NJN_F0);               \
    } while (0)
#endif

// ================================================================================================
// forward
// ================================================================================================
// Forward, per-step state in the region of the tile: the input rows exist twice (step parity: phase 3 and the glue warp fill
// the rows of step k + 1 while nobody reads them), so do the step sizes (F slot CA) and the dropout layer keys of the two
// hidden layers (region AUX: [parity][layer][4 rows]).
#define NJN_FWD_IN(f, s, par) ((f).w.IN + ((par) ? (s).f_IN2 - (s).f_IN : 0))
#define NJN_FWD_LK(reg, s, par, l) (reinterpret_cast<int*>((reg) + (s).f_AUX) + ((par) * 2 + (l)) * 4)

// F thread o: the whole input rows of step k from the state (first step after a jump / the start / a record)
template <class D>
NJ_HD void nj_tpn_fwd_build(const NjCfg& c, const NjPath& s, const NjArgs& a, NjPathFwd<(D::R >= 4 ? 4 : D::R), 1, true>& f, float* reg, int o, int k) {
    constexpr int R = D::R, RS = NJP_RS;
    const int inf4 = ((c.inf + 3) >> 2) << 2;
    const float tcur = NJ_LDG(a.b.step_t + k), dt = NJ_LDG(a.b.step_dt + k);
    float* in0 = NJN_FWD_IN(f, s, k & 1);
    float* in1 = NJN_FWD_IN(f, s, (k + 1) & 1);
    for (int c_ = o; c_ < inf4; c_ += NJN_F) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = f.I[NJP_I_PATH * RS + r];
            float v = 0.f;
            if (c_ < c.d) v = f.TX[r * s.sD + c_];
            else if (c_ < c.d + c.H) {
                const float h = f.HS[r * s.sH + c_ - c.d];
                if (p >= 0 && a.h_hist) a.h_hist[((size_t)k * a.b.B + p) * c.H + c_ - c.d] = h;
                v = nj_tanh(h);
            } else if (c_ < c.inf) v = nj_tpn_time_col(c, c_, f.F[NJP_F_TAU * RS + r], tcur);
            in0[(size_t)r * s.sI + c_] = v;
            // what stays the same until the next jump (tanh(last_X), tau) also goes into the rows of the other parity
            if (c_ < c.d || c_ == c.d + c.H || c_ >= c.inf) in1[(size_t)r * s.sI + c_] = v;
        }
    }
    if (o == NJN_F - 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const unsigned rk = nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(f.I[NJP_I_PATH * RS + r] + a.b.path_id_offset), (unsigned)k);
            NJN_FWD_LK(reg, s, k & 1, 0)[r] = (int)nj_layer_key(rk, (unsigned)(NJODE_NET_ODE * 16 + 1));
            NJN_FWD_LK(reg, s, k & 1, 1)[r] = (int)nj_layer_key(rk, (unsigned)(NJODE_NET_ODE * 16 + 2));
            f.F[NJP_F_CA * RS + (k & 1) * 4 + r] = dt;            // (the loss-coefficient slots are free in the forward pass)
        }
    }
}

// glue warp, while the F warps run step k: what step k + 1 needs besides tanh(h) -- time columns, step size, layer keys
template <class D>
NJ_HD void nj_tpn_fwd_aux(const NjCfg& c, const NjPath& s, const NjArgs& a, NjPathFwd<(D::R >= 4 ? 4 : D::R), 1, true>& f, float* reg, int lane, int k1) {
    constexpr int R = D::R, RS = NJP_RS;
    if (lane >= R) return;
    const int r = lane;
    const float tnext = NJ_LDG(a.b.step_t + k1), dnext = NJ_LDG(a.b.step_dt + k1);
    const unsigned rk = nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(f.I[NJP_I_PATH * RS + r] + a.b.path_id_offset), (unsigned)k1);
    NJN_FWD_LK(reg, s, k1 & 1, 0)[r] = (int)nj_layer_key(rk, (unsigned)(NJODE_NET_ODE * 16 + 1));
    NJN_FWD_LK(reg, s, k1 & 1, 1)[r] = (int)nj_layer_key(rk, (unsigned)(NJODE_NET_ODE * 16 + 2));
    f.F[NJP_F_CA * RS + (k1 & 1) * 4 + r] = dnext;
    float* in1 = NJN_FWD_IN(f, s, k1 & 1);
    const float tau = f.F[NJP_F_TAU * RS + r];
    for (int c_ = c.d + c.H + 1; c_ < c.inf; ++c_) in1[(size_t)r * s.sI + c_] = nj_tpn_time_col(c, c_, tau, tnext);
}

// the three phases of Euler step k for F thread o.  next: step k + 1 follows without a jump in between -- phase 3 then
// also writes tanh(h) into the input rows of step k + 1 and h into the history
template <class D>
NJ_HD void nj_tpn_fwd_p1(const NjCfg& c, const NjPath& s, NjPathFwd<(D::R >= 4 ? 4 : D::R), 1, true>& f, float* reg, NjTpnF<D>& q, int o, int k) {
    nj_tpn_hidden<D::KC0, D::R>(c, 0, q.w0, q.b0, o, NJN_FWD_IN(f, s, k & 1), s.sI, f.w.A0, s.sA, nullptr, NJN_FWD_LK(reg, s, k & 1, 0));
}
template <class D>
NJ_HD void nj_tpn_fwd_p2(const NjCfg& c, const NjPath& s, NjPathFwd<(D::R >= 4 ? 4 : D::R), 1, true>& f, float* reg, NjTpnF<D>& q, int o, int k) {
    nj_tpn_hidden<D::KCH, D::R>(c, 1, q.w1, q.b1, o, f.w.A0, s.sA, f.w.A1, s.sA, nullptr, NJN_FWD_LK(reg, s, k & 1, 1));
}
template <class D>
NJ_HD void nj_tpn_fwd_p3(const NjCfg& c, const NjPath& s, const NjArgs& a, NjPathFwd<(D::R >= 4 ? 4 : D::R), 1, true>& f, NjTpnF<D>& q, int o,
                         int k, bool next) {
    constexpr int R = D::R, RS = NJP_RS;
    float acc[R];
    nj_tpn_dot<D::KCH, R>(q.w2, f.w.A1, s.sA, acc);
    if (o < c.H) {
        float* in1 = NJN_FWD_IN(f, s, (k + 1) & 1);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float h = fmaf(f.F[NJP_F_CA * RS + (k & 1) * 4 + r], acc[r] + q.b2, f.HS[r * s.sH + o]);
            f.HS[r * s.sH + o] = h;
            if (next) {
                const int p = f.I[NJP_I_PATH * RS + r];
                if (p >= 0 && a.h_hist) a.h_hist[((size_t)(k + 1) * a.b.B + p) * c.H + o] = h;
                in1[(size_t)r * s.sI + c.d + o] = nj_tanh(h);
            }
        }
    }
}

#if defined(NJODE_HOST_SIM)
#define NJN_FREGS_DECL(D) std::vector<NjTpnF<D>> njn_f_store(NJN_F); NjTpnF<D>* njn_f = njn_f_store.data()
#define NJN_FREGS(x) (njn_f[x])
#define NJN_TREGS_DECL(D) std::vector<NjTpnT<D>> njn_t_store(NJN_T); NjTpnT<D>* njn_t = njn_t_store.data()
#define NJN_TREGS(x) (njn_t[x])
#define NJN_DACC_DECL() std::vector<float> njn_d_store((size_t)NJN_D * NJN_DACC, 0.f); float* njn_d = njn_d_store.data()
#define NJN_DACC_OF(x) (njn_d + (size_t)(x) * NJN_DACC)
#else
#define NJN_FREGS_DECL(D) NjTpnF<D> njn_f_store; NjTpnF<D>* njn_f = &njn_f_store
#define NJN_FREGS(x) (*njn_f)
#define NJN_TREGS_DECL(D) NjTpnT<D> njn_t_store; NjTpnT<D>* njn_t = &njn_t_store
#define NJN_TREGS(x) (*njn_t)
#define NJN_DACC_DECL() float njn_d_store[NJN_DACC]; _Pragma("unroll") for (int _i = 0; _i < NJN_DACC; ++_i) njn_d_store[_i] = 0.f; float* njn_d = njn_d_store
#define NJN_DACC_OF(x) (njn_d)
#endif

template <class D, int ROLE>
NJ_HD void nj_tpn_fwd_body(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem) {
    constexpr int R = D::R, RG = (R >= 4 ? 4 : R);
    const float* simg = smem - c.net[NJODE_NET_ENC].w_img[0];      // image offsets of the jump networks land in the staged part
    float* reg = smem + s.f_warp0;
    int* slot = reinterpret_cast<int*>(reg + s.f_I) + NJP_I_COUNT * NJP_RS;
    NjPathFwd<RG, 1, true> f(c, s, a, reg, simg);
    NjCoopMB* mb = reinterpret_cast<NjCoopMB*>(reg + s.f_MB);
    NJ_THREADS(tid, NJN_NT_FWD) { if (tid == 0) { mb->R = R; mb->op = 0; } }
    f.w.coop = mb;
    NJN_FREGS_DECL(D);
    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_f_load<D>(c, a.image, o, NJN_FREGS(o)); }
    const bool rec = a.b.E > 0;
    const int S = a.b.S;
    for (;;) {
        NJ_THREADS(tid, NJN_NT_FWD) { if (tid == 0) *slot = nj_atomic_inc(a.counter); }
        NJ_SYNC();
        const int wt = *slot;
        NJ_SYNC();
        if (wt >= s.n_tiles_f) break;
        const int ub = wt * R, ue = ub + R < a.b.n_units ? ub + R : a.b.n_units;
        NJN_GLUE(mb, f.begin(ub, ue));
        NJ_SYNC();
        int k = 0, gi = 0;
        for (;;) {
            const int nk = f.next_jump(gi);
            const int kend = nk < S ? nk : S;
            if (rec) {
                for (; k < kend; ++k) {
                    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_build<D>(c, s, a, f, reg, o, k); }
                    NJN_SYNC_F();
                    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_p1<D>(c, s, f, reg, NJN_FREGS(o), o, k); }
                    NJN_SYNC_F();
                    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_p2<D>(c, s, f, reg, NJN_FREGS(o), o, k); }
                    NJN_SYNC_F();
                    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_p3<D>(c, s, a, f, NJN_FREGS(o), o, k, false); }
                    NJ_SYNC();
                    NJN_GLUE(mb, f.record(NJ_LDG(a.b.step_event + k), NJ_EVENT_PATH_RO_BASE + (unsigned)k));
                    NJ_SYNC();
                }
            } else if (k < kend) {
                NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_build<D>(c, s, a, f, reg, o, k); }
                NJN_SYNC_F();
                for (; k < kend; ++k) {
                    const bool next = k + 1 < kend;
                    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_p1<D>(c, s, f, reg, NJN_FREGS(o), o, k); }
                    NJN_SYNC_F();
                    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_p2<D>(c, s, f, reg, NJN_FREGS(o), o, k); }
                    NJN_SYNC_F();
                    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_p3<D>(c, s, a, f, NJN_FREGS(o), o, k, next); }
                    if (next && (ROLE == NJN_ALL || ROLE == NJN_ROLE_G)) {
                        NJ_WARPS(wp, 1) { if (wp == 0) { NJ_LANES(lane) { nj_tpn_fwd_aux<D>(c, s, a, f, reg, lane, k + 1); } } }
                    }
                    NJN_SYNC_STEP();
                }
            }
            NJ_SYNC();                            // the state after the run is visible to the glue warp
            if (nk > S) break;
            const bool any = f.any_jumps_at(nk);
            NJ_SYNC();                            // every thread has read the cursors before the glue warp advances them
            NJN_GLUE(mb, { if (any) f.jump(nk); if (rec) f.record(NJ_LDG(a.b.jump_event + gi), NJ_EVENT_JUMP_BASE + 3u * (unsigned)gi + 2u); });
            if (rec) ++gi;
            NJ_SYNC();
        }
        if (ROLE == NJN_ALL || ROLE == NJN_ROLE_G) { NJ_WARPS(wp, 1) { if (wp == 0) f.finish(); } }
        NJ_SYNC();
    }
}

template <class D>
NJ_HD void nj_tpn_cta_forward(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem) {
    // (only the jump networks' part of the parameter image: the ODE network goes from global memory straight into registers)
    const int ode_floats = c.net[NJODE_NET_ENC].w_img[0];
    nj_stage_image(smem, a.image + ode_floats, c.img_floats - ode_floats, NJN_NT_FWD);
    nj_zero(smem + s.f_warp0, s.f_region, NJN_NT_FWD);
    NJ_SYNC();
#if defined(NJODE_HOST_SIM)
    nj_tpn_fwd_body<D, NJN_ALL>(c, s, a, smem);
#else
    if (threadIdx.x < NJN_F0) nj_tpn_fwd_body<D, NJN_ROLE_G>(c, s, a, smem);
    else nj_tpn_fwd_body<D, NJN_ROLE_F>(c, s, a, smem);
#endif
}

// ================================================================================================
// backward
// ================================================================================================
NJ_HD void nj_tpn_set(NjPathB& t, int off) { t.IN += off; t.A += off; t.G += off; t.GOUT += off; }
#define NJN_PRE_HDR 24
// backward: the glue warp fetches the next step's scalars and hashes its row keys (1) or the F threads 0 and 63 do (0).
// Measured on B200 (PhysioNet batch of 50, backward kernel): 9.57 ms without the glue warp's help, 10.30 ms with it (and
// 10.4 ms when it also precomputes the layer keys): the backward's eight warps share four schedulers, a busy glue warp
// competes with a T warp.  The three-warp forward gains 21 % from the same move (4.92 -> 3.90 ms).
#ifndef NJN_BWD_AUX
#define NJN_BWD_AUX 0
#endif
#define NJN_BWD_LK(smem, s, par, l) (reinterpret_cast<int*>((smem) + (s).b_PRE + 4) + ((par) * 2 + (l)) * 4)

// F, phase 1: the input rows of step e into its operand set (h from the history; the value was loaded one iteration ahead
// when `have`), then the load for step e - 1 is issued (`pre`)
template <class D>
NJ_HD void nj_tpn_bwd_build(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem, const NjPathB& t0, const NjPathB& te, int o,
                            int e, bool have, bool pre) {
    constexpr int R = D::R;
    const int P = s.P_b, inf4 = ((c.inf + 3) >> 2) << 2;
    // prefetch region: [0..1] time, [2..3] step size of the step (by parity), [4..19] dropout layer keys [parity][layer][4],
    // then h [2][P][sH].  The scalars and keys of step e were written by the glue warp during the previous pipeline
    // iteration (`have`), else (first step of a run) by this function.
    float* pre_f = smem + s.b_PRE;
    float* HP = pre_f + NJN_PRE_HDR;
    if (have) nj_cp_wait();                      // this thread's copies of the previous iteration (h of step e)
    const float tcur = have ? pre_f[e & 1] : NJ_LDG(a.b.step_t + e);
    if (o == 0 && !have) pre_f[2 + (e & 1)] = NJ_LDG(a.b.step_dt + e);
    if (!NJN_BWD_AUX && o == 0 && pre) {         // (without the glue warp's help: thread 0 fetches the scalars of step e - 1)
        nj_cp_async4(pre_f + ((e - 1) & 1), a.b.step_t + e - 1); nj_cp_async4(pre_f + 2 + ((e - 1) & 1), a.b.step_dt + e - 1);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {                // (the planner admits at most 2 * NJN_F input columns)
        const int c_ = o + NJN_F * i;
        if (c_ >= inf4) break;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = t0.I[NJB_I_PATH * P + r];
            float v = 0.f;
            if (c_ < c.d) v = t0.TX[r * s.sD + c_];
            else if (c_ < c.d + c.H) {
                float h = 0.f;
                if (p >= 0) {
                    const float* hh = a.h_hist + ((size_t)e * a.b.B + p) * c.H + c_ - c.d;
                    h = have ? HP[((e & 1) * P + r) * s.sH + c_ - c.d] : NJ_LDG(hh);
                    if (pre) nj_cp_async4(HP + (((e - 1) & 1) * P + r) * s.sH + c_ - c.d, hh - (size_t)a.b.B * c.H);
                }
                v = nj_tanh(h);
            } else if (c_ < c.inf) v = nj_tpn_time_col(c, c_, t0.F[NJP_F_TAU * P + r], tcur);
            te.IN[(size_t)r * s.sI + c_] = v;
        }
    }
    if (o == NJN_F - 1 && (!have || !NJN_BWD_AUX)) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const unsigned rk = nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(t0.I[NJB_I_PATH * P + r] + a.b.path_id_offset), (unsigned)e);
            if (NJN_BWD_AUX) NJN_BWD_LK(smem, s, e & 1, 0)[r] = (int)rk;
            else t0.I[NJB_I_RK * P + r] = (int)rk;       // (every F thread derives the layer keys itself, in parallel)
        }
    }
}

// glue warp, while F rebuilds step e: time, step size and dropout layer keys of step e - 1 (the next one F rebuilds)
template <class D>
NJ_HD void nj_tpn_bwd_aux(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem, const NjPathB& t0, int lane, int e1,
                          bool staged, bool more) {
    constexpr int R = D::R;
    const int P = s.P_b;
    float* pre_f = smem + s.b_PRE;
    if (lane == 0) {
        // the two scalars travel through a staging pair ([20..21] of the region) filled by cp.async one iteration ahead
        float tv, dv;
        if (staged) { nj_cp_wait(); tv = pre_f[20]; dv = pre_f[21]; }
        else { tv = NJ_LDG(a.b.step_t + e1); dv = NJ_LDG(a.b.step_dt + e1); }
        pre_f[e1 & 1] = tv; pre_f[2 + (e1 & 1)] = dv;
        if (more) { nj_cp_async4(pre_f + 20, a.b.step_t + e1 - 1); nj_cp_async4(pre_f + 21, a.b.step_dt + e1 - 1); }
    }
    // row keys of step e1, by parity (the F threads derive their layer keys from them in their epilogues)
    if (lane < R)
        NJN_BWD_LK(smem, s, e1 & 1, 0)[lane] = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(t0.I[NJB_I_PATH * P + lane] + a.b.path_id_offset), (unsigned)e1);
}

// T: gradient of hidden layer l's pre-activation from the partial sum over the next layer's outputs
NJ_HD float nj_tpn_hidden_grad(const NjCfg& c, int l, float v, float a_) {
    const NjNet& N = c.net[NJODE_NET_ODE];
    if (c.has_drop) {
        if (nj_f2u(a_) == NJ_DROPPED) return 0.f;
        a_ *= c.one_minus_p; v *= c.keep_scale;
    }
    if (N.act[l] == NJODE_ACT_TANH) v *= (1.f - a_ * a_);
    else if (N.act[l] == NJODE_ACT_RELU) v = a_ > 0.f ? v : 0.f;
    return v;
}

template <class D, class S, class TB>
NJ_HD void nj_tpn_bwd_t1(const NjCfg& c, const S& s, const TB& te, const NjTpnT<D>& q, int k) {
    constexpr int R = D::R;
    const int wa = s.P_b * s.sA;
    float acc[R];
    nj_tpn_dot<D::HC, R>(q.c2, te.GOUT, s.sO, acc);
    const int O = c.net[NJODE_NET_ODE].dim[2];
    if (k >= (((O + 3) >> 2) << 2)) return;
#pragma unroll
    for (int r = 0; r < R; ++r) te.G[wa + (size_t)r * s.sA + k] = k < O ? nj_tpn_hidden_grad(c, 1, acc[r], te.A[wa + (size_t)r * s.sA + k]) : 0.f;
}
template <class D, class S, class TB>
NJ_HD void nj_tpn_bwd_t2(const NjCfg& c, const S& s, const TB& te, const NjTpnT<D>& q, int k) {
    constexpr int R = D::R;
    const int wa = s.P_b * s.sA;
    float acc[R];
    nj_tpn_dot<D::KCH, R>(q.c1, te.G + wa, s.sA, acc);
    const int O = c.net[NJODE_NET_ODE].dim[1];
    if (k >= (((O + 3) >> 2) << 2)) return;
#pragma unroll
    for (int r = 0; r < R; ++r) te.G[(size_t)r * s.sA + k] = k < O ? nj_tpn_hidden_grad(c, 0, acc[r], te.A[(size_t)r * s.sA + k]) : 0.f;
}
// phase 3: the input gradient of step e into the adjoint of h (and of last_X, masked model); then the output gradient
// dt * gh of the step reversed next goes into ITS operand set `tn` (dtn: its step size; tn null: no step follows)
template <class D>
NJ_HD void nj_tpn_bwd_t3(const NjCfg& c, const NjPath& s, const NjPathB& t0, const NjPathB* te, const NjPathB* tn, float dtn,
                         const NjTpnT<D>& q, int k) {
    constexpr int R = D::R;
    const int P = s.P_b;
    float acc[R];
    if (te) nj_tpn_dot<D::KCH, R>(q.c0, te->G, s.sA, acc);
    const bool hcol = k >= c.d && k < c.d + c.H;
    const int H4 = ((c.H + 3) >> 2) << 2;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool valid = t0.I[NJB_I_PATH * P + r] >= 0;
        if (te && valid) {
            if (hcol) {
                const float th = te->IN[(size_t)r * s.sI + k];
                t0.GH[r * s.sH + k - c.d] += acc[r] * (1.f - th * th);
            } else if (c.masked && k < c.d) {             // last_X = Y[i_obs] is differentiable (NJODE/models.py:483-484)
                const float tx = te->IN[(size_t)r * s.sI + k];
                t0.GX[r * s.sD + k] += acc[r] * (1.f - tx * tx);
            }
        }
        if (tn && hcol) tn->GOUT[(size_t)r * s.sO + k - c.d] = dtn * t0.GH[r * s.sH + k - c.d];
        else if (tn && k >= c.d + c.H && k < c.d + H4) tn->GOUT[(size_t)r * s.sO + k - c.d] = 0.f;   // padding read by the dW tiles
    }
}

// D thread x owns the tiles T = slot * NJN_D + x of the ODE network.  Their operand addresses are decoded once per launch
// into a table in shared memory: [T][0] offset of g (floats from the operand set's IN buffer), [T][1] offset of a | layer << 24
// (-1: no such tile) -- decoding per step cost the D warps 1 100 instructions per step and made them the critical path
NJ_HD bool nj_tpn_ode_tile(const NjCfg& c, const NjPath& s, int T, int& l, int& og, int& kg) { return nj_path_tile_decode(c, s, NJODE_NET_ODE, T, l, og, kg); }
NJ_HD bool nj_tpn_ode_tile(const NjCfg& c, const NjSeg& s, int T, int& l, int& og, int& kg) { return nj_seg_tile_decode(c, s, NJODE_NET_ODE, T, l, og, kg); }

template <class S>
NJ_HD void nj_tpn_dw_table(const NjCfg& c, const S& s, float* smem, int x) {
    const int ode_tiles = s.tile_base[NJODE_NET_RO][0];        // the ODE network's tiles come first
    const NjNet& N = c.net[NJODE_NET_ODE];
    int* td = reinterpret_cast<int*>(smem + s.b_TD);
    const int P = s.P_b;
    for (int slot = 0; slot < NJN_DSLOTS; ++slot) {
        const int T = slot * NJN_D + x;
        int l, og, kg, e0 = -1, e1 = 0;
        if (T < ode_tiles && nj_tpn_ode_tile(c, s, T, l, og, kg)) {
            e0 = (l == N.n - 1 ? s.b_GOUT : s.b_G + l * P * s.sA) - s.b_IN + 4 * og;
            e1 = ((l == 0 ? s.b_IN : s.b_A + (l - 1) * P * s.sA) - s.b_IN + 4 * kg) | (l << 24);
        }
        td[2 * T] = e0; td[2 * T + 1] = e1;
    }
}

template <int R, class S>
NJ_HD void nj_tpn_dw(const NjCfg& c, const S& s, const float* smem, const float* set_in, float* acc, int x) {
    const int* td = reinterpret_cast<const int*>(smem + s.b_TD);
    const int last = c.net[NJODE_NET_ODE].n - 1;
#pragma unroll
    for (int slot = 0; slot < NJN_DSLOTS; ++slot) {
        const int T = slot * NJN_D + x;
        const int e0 = td[2 * T], e1 = td[2 * T + 1];
        if (e0 < 0) continue;
        const int l = e1 >> 24;
        const int g_s = l == last ? s.sO : s.sA, a_s = l == 0 ? s.sI : s.sA;
        const nj_sp gq = nj_sp_of(set_in + e0), aq = nj_sp_of(set_in + (e1 & 0xFFFFFF));
        float* q = acc + slot * 20;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const nj_f4 gv = nj_sp_ld4(NJ_SP_ADD(gq, r * g_s));
            const nj_f4 v = nj_sp_ld4(NJ_SP_ADD(aq, r * a_s));
            q[0] = fmaf(gv.x, v.x, q[0]); q[1] = fmaf(gv.x, v.y, q[1]); q[2] = fmaf(gv.x, v.z, q[2]); q[3] = fmaf(gv.x, v.w, q[3]);
            q[4] = fmaf(gv.y, v.x, q[4]); q[5] = fmaf(gv.y, v.y, q[5]); q[6] = fmaf(gv.y, v.z, q[6]); q[7] = fmaf(gv.y, v.w, q[7]);
            q[8] = fmaf(gv.z, v.x, q[8]); q[9] = fmaf(gv.z, v.y, q[9]); q[10] = fmaf(gv.z, v.z, q[10]); q[11] = fmaf(gv.z, v.w, q[11]);
            q[12] = fmaf(gv.w, v.x, q[12]); q[13] = fmaf(gv.w, v.y, q[13]); q[14] = fmaf(gv.w, v.z, q[14]); q[15] = fmaf(gv.w, v.w, q[15]);
            q[16] += gv.x; q[17] += gv.y; q[18] += gv.z; q[19] += gv.w;
        }
    }
}
template <class S>
NJ_HD void nj_tpn_dw_flush(const NjCfg& c, const S& s, const float* acc, float* gpart, int x) {
    const int ode_tiles = s.tile_base[NJODE_NET_RO][0];
#pragma unroll
    for (int slot = 0; slot < NJN_DSLOTS; ++slot) {
        const int T = slot * NJN_D + x;
        if (T >= ode_tiles) break;
        int l, og, kg;
        if (nj_tpn_ode_tile(c, s, T, l, og, kg)) nj_seg_tile_store(c, NJODE_NET_ODE, l, og, kg, acc + slot * 20, gpart, false);
    }
}

template <class D, int ROLE>
NJ_HD void nj_tpn_bwd_body(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem, int cta) {
    constexpr int R = D::R, RG = (R >= 4 ? 4 : R);
    const int nt = NJN_NT_BWD, P = s.P_b;
    const bool G = ROLE == NJN_ALL || ROLE == NJN_ROLE_G;
    // the jump networks' dW tiles are read-modify-written at every jump: into a gradient image in SHARED memory (the
    // global partial image cost ~3 us of L2 round trips per dW phase, four or five phases per jump), copied out once
    // (the ODE network comes first in the image and has its own accumulators: the shared copy starts behind it)
    const int ode_floats = c.net[NJODE_NET_ENC].w_img[0];
    float* gout = a.partials + (size_t)cta * c.img_floats;
    float* gpart = smem + s.b_GIMG - ode_floats;
    NjPathB t;
    nj_pathb_bind(t, s, smem);
    NjPathBwd<RG, 1, true> B(c, s, a, t, smem);
    NjCoopMB* mb = reinterpret_cast<NjCoopMB*>(smem + s.b_MB);
    NJ_THREADS(tid, nt) { if (tid == 0) { mb->R = R; mb->op = 0; } }
    B.coop = mb;
    NJN_FREGS_DECL(D);
    NJN_TREGS_DECL(D);
    NJN_DACC_DECL();
    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_f_load<D>(c, smem, o, NJN_FREGS(o)); }
    NJN_ROLE(NJN_ROLE_T, NJN_T0, NJN_T, k) { nj_tpn_t_load<D>(c, smem, k, NJN_TREGS(k)); }
    NJN_ROLE(NJN_ROLE_D, NJN_D0, NJN_D, x) { nj_tpn_dw_table(c, s, smem, x); }      // (read by its writer only)
    int* ctl = t.I + NJB_I_COUNT * P;
    for (;;) {
        NJ_THREADS(tid, nt) { if (tid == 0) ctl[0] = nj_atomic_inc(a.counter); }
        NJ_SYNC();
        const int tile = ctl[0];
        NJ_SYNC();
        if (tile >= s.n_tiles_b) break;
        const int ub = tile * R, ue = ub + R < a.b.n_units ? ub + R : a.b.n_units;
        NJ_THREADS(tid, nt) {
            if (tid < R) {
                const int u = ub + tid;
                if (u < ue) {
                    const int32_t* dsc = a.b.unit_desc + (size_t)u * 6;
                    t.I[NJB_I_PATH * P + tid] = dsc[0]; t.I[NJB_I_C0 * P + tid] = dsc[3]; t.I[NJB_I_CUR * P + tid] = dsc[4];
                } else { t.I[NJB_I_PATH * P + tid] = -1; t.I[NJB_I_C0 * P + tid] = 0; t.I[NJB_I_CUR * P + tid] = 0; }
                t.I[NJB_I_ACT * P + tid] = 0;
                B.set_prev(tid);
            }
        }
        NJ_SYNC();
        if (G) {
            NJ_WARPS(wp, 1) {
                if (wp == 0) {
                    NJ_LANES(lane) {
                        NJ_ROWMAP(R);
                        const int p = t.I[NJB_I_PATH * P + er];
                        const float* ght = (p >= 0 && a.grad_hT) ? a.grad_hT + (size_t)p * c.H : nullptr;
                        for (int c_ = ec0; c_ < c.H; c_ += LPR) t.GH[er * s.sH + c_] = ght ? NJ_LDG(ght + c_) : 0.f;
                        for (int c_ = ec0; c_ < c.d; c_ += LPR) t.GX[er * s.sD + c_] = 0.f;
                        B.load_state(0, lane);
                    }
                    NJ_SYNCWARP();
                }
            }
        }
        NJ_SYNC();
        int nk = nj_pathb_next(t, P, R);
        for (int k = a.b.S; ; ) {
            if (nk == k) {
                // (the pipeline is empty between runs: the jump works on operand set 0, its dW phases go through the partial image)
                NJN_GLUE(mb, B.jump_p1(0, 0, k));
                NJ_SYNC();
                NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t, NJODE_NET_RO, gpart, tid, nt, R, t.MSK, R); }
                NJ_SYNC();
                NJN_GLUE(mb, B.jump_p2(0, 0));
                NJ_SYNC();
                NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t, c.use_rnn ? NJODE_NET_GRU_HH : NJODE_NET_ENC, gpart, tid, nt, R, t.MSK, R); }
                NJ_SYNC();
                if (c.use_rnn) {
                    NJN_GLUE(mb, B.jump_p2b(0, 0));
                    NJ_SYNC();
                    NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t, NJODE_NET_GRU_IH, gpart, tid, nt, R, t.MSK, R); }
                    NJ_SYNC();
                }
                NJN_GLUE(mb, B.jump_p3(0, 0));
                NJ_SYNC();
                NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t, NJODE_NET_RO, gpart, tid, nt, R, t.MSK, R); }
                NJ_SYNC();
                nk = nj_pathb_next(t, P, R);
            }
            if (k == 0) break;
            // ---- the run of steps e0 = k - 1 ... lo between this jump and the previous one: n + 2 pipeline iterations ----
            const int lo = nk > 0 ? nk : 0, n = k - lo, e0 = k - 1;
            for (int j = 0; j <= n + 1; ++j) {
                const int eF = e0 - j, eT = eF + 1, eD = eF + 2;
                const bool vF = j < n, vT = j >= 1 && j <= n, vD = j >= 2;
                // operand set of step e: e mod 3 (views built here: an indexed array of views would live in local memory)
                NjPathB tF = t, tT = t, tD = t;
                nj_tpn_set(tF, ((eF + 3) % 3) * s.b_copy); nj_tpn_set(tT, ((eT + 3) % 3) * s.b_copy); nj_tpn_set(tD, ((eD + 3) % 3) * s.b_copy);
                // phase 1
                NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { if (vF) nj_tpn_bwd_build<D>(c, s, a, smem, t, tF, o, eF, j > 0, j + 1 < n); }
                NJN_ROLE(NJN_ROLE_T, NJN_T0, NJN_T, x) { if (vT) nj_tpn_bwd_t1<D>(c, s, tT, NJN_TREGS(x), x); }
                NJN_ROLE(NJN_ROLE_D, NJN_D0, NJN_D, x) { if (vD) nj_tpn_dw<R>(c, s, smem, tD.IN, NJN_DACC_OF(x), x); }
                NJN_SYNC_FT();
                // phase 2
                NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) {
                    if (vF) nj_tpn_hidden<D::KC0, R>(c, 0, NJN_FREGS(o).w0, NJN_FREGS(o).b0, o, tF.IN, s.sI, tF.A, s.sA, NJN_BWD_AUX ? NJN_BWD_LK(smem, s, eF & 1, 0) : t.I + NJB_I_RK * P);
                }
                NJN_ROLE(NJN_ROLE_T, NJN_T0, NJN_T, x) { if (vT) nj_tpn_bwd_t2<D>(c, s, tT, NJN_TREGS(x), x); }
                NJN_SYNC_FT();
                // phase 3
                NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) {
                    if (vF) nj_tpn_hidden<D::KCH, R>(c, 1, NJN_FREGS(o).w1, NJN_FREGS(o).b1, o, tF.A, s.sA, tF.A + P * s.sA, s.sA, NJN_BWD_AUX ? NJN_BWD_LK(smem, s, eF & 1, 0) : t.I + NJB_I_RK * P);
                    if (!NJN_BWD_AUX && o == 0) nj_cp_wait();          // scalars of the next step: visible to everyone after the barrier
                }
                NJN_ROLE(NJN_ROLE_T, NJN_T0, NJN_T, x) {
                    if (vT || vF) nj_tpn_bwd_t3<D>(c, s, t, vT ? &tT : nullptr, vF ? &tF : nullptr, smem[s.b_PRE + 2 + (eF & 1)], NJN_TREGS(x), x);
                }
                // the glue warp, idle during a run: scalars and dropout keys of the step F rebuilds next
                if (NJN_BWD_AUX && j + 1 < n && G) { NJ_WARPS(wp, 1) { if (wp == 0) { NJ_LANES(lane) { nj_tpn_bwd_aux<D>(c, s, a, smem, t, lane, eF - 1, j > 0, j + 2 < n); } } } }
                NJ_SYNC();
            }
            k = lo;
        }
        NJN_GLUE(mb, B.start_local(0));
        NJ_SYNC();
        NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t, NJODE_NET_ENC, gpart, tid, nt, R, nullptr, R); }
        NJ_SYNC();
    }
    NJN_ROLE(NJN_ROLE_D, NJN_D0, NJN_D, x) { nj_tpn_dw_flush(c, s, NJN_DACC_OF(x), gout, x); }
    NJ_THREADS(tid, nt) { for (int i = ode_floats + tid; i < c.img_floats; i += nt) gout[i] = gpart[i]; }
}

template <class D>
NJ_HD void nj_tpn_cta_backward(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem, int cta) {
    const int nt = NJN_NT_BWD;
    nj_stage_image(smem, a.image, c.img_floats, nt);
    nj_zero(smem + s.b_IN, s.b_smem_floats - s.b_IN, nt);
    nj_zero(a.partials + (size_t)cta * c.img_floats, c.net[NJODE_NET_ENC].w_img[0], nt);     // (rows no dW tile covers stay 0)
    NJ_SYNC();
#if defined(NJODE_HOST_SIM)
    nj_tpn_bwd_body<D, NJN_ALL>(c, s, a, smem, cta);
#else
    if (threadIdx.x < NJN_F0) nj_tpn_bwd_body<D, NJN_ROLE_G>(c, s, a, smem, cta);
    else if (threadIdx.x < NJN_T0) nj_tpn_bwd_body<D, NJN_ROLE_F>(c, s, a, smem, cta);
    else if (threadIdx.x < NJN_D0) nj_tpn_bwd_body<D, NJN_ROLE_T>(c, s, a, smem, cta);
    else nj_tpn_bwd_body<D, NJN_ROLE_D>(c, s, a, smem, cta);
#endif
}

// ================================================================================================
// SEGMENT units (non-masked training call) of small batches: the reference's own batch of 200 paths is ~2 200 segments of
// ~10 Euler steps, 15 per SM -- the 12-warp tiles of njode_seg.cuh then run one latency-bound warp per scheduler and the
// launch lasts as long as its longest segment (~70 steps).  Same roles as above on tiles of 4 segments whose rows step
// through their OWN Euler steps (row r is at step s0[r] + j and rests once j reaches its length); the start encoder and the
// jump that ends the segments stay with the glue warp (njode_seg.cuh code).
// ================================================================================================
#define NJN_SEG_R 4
// cooperative layers for the glue sections of segment tiles: measured neutral to slightly slower on B200 (bs_demo_200: forward
// 0.21 -> 0.22 ms, backward 0.41 -> 0.44 ms) -- with 4 rows the glue warp's own warp GEMM has seven independent accumulators
// per lane and is as fast as 64 threads with one dot product each plus two named barriers per layer.  Kept switchable.
#define NJN_SEG_COOP false

NJ_HD void nj_tpn_set(NjSegB& t, int off) { t.IN += off; t.A += off; t.G += off; t.GOUT += off; }

// forward per-row scalars in the F slots of the region: step size by parity (the DT slot), prefetched time / step size (CA slot)
#define NJN_SEG_DTV(f, par, r) (f).F[NJS_F_DT * 16 + (par) * 4 + (r)]
#define NJN_SEG_TS(f, par, r) (f).F[NJS_F_CA * 16 + (par) * 4 + (r)]
#define NJN_SEG_DS(f, par, r) (f).F[NJS_F_CA * 16 + 8 + (par) * 4 + (r)]

template <class D>
NJ_HD void nj_segtpn_fwd_build(const NjCfg& c, const NjSeg& s, const NjArgs& a, NjSegFwd<1, NJN_SEG_COOP>& f, int o, int j) {
    constexpr int R = NJN_SEG_R, RS = 16;
    const int inf4 = ((c.inf + 3) >> 2) << 2;
    for (int c_ = o; c_ < inf4; c_ += NJN_F) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = f.I[NJS_I_PATH * RS + r], k = f.I[NJS_I_S0 * RS + r] + j;
            const bool active = j < f.I[NJS_I_LEN * RS + r];
            float v = 0.f;
            if (c_ < c.d) v = f.TX[r * s.sD + c_];
            else if (c_ < c.d + c.H) {
                const float h = f.HS[r * s.sH + c_ - c.d];
                if (active && a.h_hist) a.h_hist[((size_t)k * a.b.B + p) * c.H + c_ - c.d] = h;
                v = nj_tanh(h);
            } else if (c_ < c.inf) v = nj_tpn_time_col(c, c_, f.F[NJS_F_TAU * RS + r], active ? NJ_LDG(a.b.step_t + k) : 0.f);
            f.w.IN[(size_t)r * s.sI + c_] = v;
        }
    }
    if (o == NJN_F - 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int k = f.I[NJS_I_S0 * RS + r] + j;
            f.w.RK[r] = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(f.I[NJS_I_PATH * RS + r] + a.b.path_id_offset), (unsigned)k);
            NJN_SEG_DTV(f, j & 1, r) = j < f.I[NJS_I_LEN * RS + r] ? NJ_LDG(a.b.step_dt + k) : 0.f;
        }
    }
}
template <class D>
NJ_HD void nj_segtpn_fwd_p1(const NjCfg& c, const NjSeg& s, const NjArgs& a, NjSegFwd<1, NJN_SEG_COOP>& f, NjTpnF<D>& q, int o, int j, bool next) {
    constexpr int R = NJN_SEG_R, RS = 16;
    if (next && o == NJN_F - 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (j + 1 < f.I[NJS_I_LEN * RS + r]) {
                const int k1 = f.I[NJS_I_S0 * RS + r] + j + 1;
                nj_cp_async4(&NJN_SEG_TS(f, (j + 1) & 1, r), a.b.step_t + k1);
                nj_cp_async4(&NJN_SEG_DS(f, (j + 1) & 1, r), a.b.step_dt + k1);
            }
        }
    }
    nj_tpn_hidden<D::KC0, R>(c, 0, q.w0, q.b0, o, f.w.IN, s.sI, f.w.A0, s.sA, f.w.RK);
}
template <class D>
NJ_HD void nj_segtpn_fwd_p2(const NjCfg& c, const NjSeg& s, NjSegFwd<1, NJN_SEG_COOP>& f, NjTpnF<D>& q, int o) {
    nj_tpn_hidden<D::KCH, NJN_SEG_R>(c, 1, q.w1, q.b1, o, f.w.A0, s.sA, f.w.A1, s.sA, f.w.RK);
}
template <class D>
NJ_HD void nj_segtpn_fwd_p3(const NjCfg& c, const NjSeg& s, const NjArgs& a, NjSegFwd<1, NJN_SEG_COOP>& f, NjTpnF<D>& q, int o, int j, bool next) {
    constexpr int R = NJN_SEG_R, RS = 16;
    float acc[R];
    nj_tpn_dot<D::KCH, R>(q.w2, f.w.A1, s.sA, acc);
    if (o < c.H) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float h = fmaf(NJN_SEG_DTV(f, j & 1, r), acc[r] + q.b2, f.HS[r * s.sH + o]);
            f.HS[r * s.sH + o] = h;
            if (next) {
                if (j + 1 < f.I[NJS_I_LEN * RS + r] && a.h_hist)
                    a.h_hist[((size_t)(f.I[NJS_I_S0 * RS + r] + j + 1) * a.b.B + f.I[NJS_I_PATH * RS + r]) * c.H + o] = h;
                f.w.IN[(size_t)r * s.sI + c.d + o] = nj_tanh(h);
            }
        }
    }
    if (next && o == NJN_F - 1) {
        nj_cp_wait();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool an = j + 1 < f.I[NJS_I_LEN * RS + r];
            const int k1 = f.I[NJS_I_S0 * RS + r] + j + 1;
            const float tau = f.F[NJS_F_TAU * RS + r], tnext = an ? NJN_SEG_TS(f, (j + 1) & 1, r) : 0.f;
            for (int c_ = c.d + c.H + 1; c_ < c.inf; ++c_) f.w.IN[(size_t)r * s.sI + c_] = nj_tpn_time_col(c, c_, tau, tnext);
            f.w.RK[r] = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(f.I[NJS_I_PATH * RS + r] + a.b.path_id_offset), (unsigned)k1);
            NJN_SEG_DTV(f, (j + 1) & 1, r) = an ? NJN_SEG_DS(f, (j + 1) & 1, r) : 0.f;
        }
    }
}

template <class D, int ROLE>
NJ_HD void nj_segtpn_fwd_body(const NjCfg& c, const NjSeg& s, const NjArgs& a, float* smem) {
    constexpr int R = NJN_SEG_R, RS = 16;
    float* simg = smem + s.f_img;
    float* reg = smem + s.f_warp0;
    NjSegFwd<1, NJN_SEG_COOP> f(c, s, a, reg, simg);
    NjCoopMB* mb = reinterpret_cast<NjCoopMB*>(reg + s.f_MB);
    NJ_THREADS(tid, NJN_NT_FWD) { if (tid == 0) { mb->R = NJN_SEG_R; mb->op = 0; } }
    f.w.coop = mb;
    int* slot = f.I + NJS_I_COUNT * RS;
    NJN_FREGS_DECL(D);
    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_f_load<D>(c, simg, o, NJN_FREGS(o)); }
    for (;;) {
        NJ_THREADS(tid, NJN_NT_FWD) { if (tid == 0) *slot = nj_atomic_inc(a.counter); }
        NJ_SYNC();
        const int wt = *slot;
        NJ_SYNC();
        if (wt >= s.n_tiles_f) break;
        int ub, ue;
        nj_seg_tile_lookup(s.f_ncls, s.f_t0, s.f_u0, s.f_u1, s.f_tr, 4, wt, ub, ue);
        if (NJN_SEG_COOP) NJN_GLUE(mb, f.begin(ub, ue));
        else if (ROLE == NJN_ALL || ROLE == NJN_ROLE_G) { NJ_WARPS(wp, 1) { if (wp == 0) f.begin(ub, ue); } }
        NJ_SYNC();
        int maxlen = 0;
        for (int r = 0; r < R; ++r) maxlen = f.I[NJS_I_LEN * RS + r] > maxlen ? f.I[NJS_I_LEN * RS + r] : maxlen;
        if (maxlen > 0) {
            NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_segtpn_fwd_build<D>(c, s, a, f, o, 0); }
            NJN_SYNC_F();
            for (int j = 0; j < maxlen; ++j) {
                const bool next = j + 1 < maxlen;
                NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_segtpn_fwd_p1<D>(c, s, a, f, NJN_FREGS(o), o, j, next); }
                NJN_SYNC_F();
                NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_segtpn_fwd_p2<D>(c, s, f, NJN_FREGS(o), o); }
                NJN_SYNC_F();
                NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_segtpn_fwd_p3<D>(c, s, a, f, NJN_FREGS(o), o, j, next); }
                NJN_SYNC_F();
            }
        }
        NJ_SYNC();
        if (NJN_SEG_COOP) NJN_GLUE(mb, { f.maxlen = maxlen; f.finish(); });
        else if (ROLE == NJN_ALL || ROLE == NJN_ROLE_G) { NJ_WARPS(wp, 1) { if (wp == 0) { f.maxlen = maxlen; f.finish(); } } }
        NJ_SYNC();
    }
}

template <class D>
NJ_HD void nj_segtpn_cta_forward(const NjCfg& c, const NjSeg& s, const NjArgs& a, float* smem) {
    nj_stage_image(smem + s.f_img, a.image, c.img_floats, NJN_NT_FWD);
    nj_zero(smem + s.f_warp0, s.f_region, NJN_NT_FWD);
    NJ_SYNC();
#if defined(NJODE_HOST_SIM)
    nj_segtpn_fwd_body<D, NJN_ALL>(c, s, a, smem);
#else
    if (threadIdx.x < NJN_F0) nj_segtpn_fwd_body<D, NJN_ROLE_G>(c, s, a, smem);
    else nj_segtpn_fwd_body<D, NJN_ROLE_F>(c, s, a, smem);
#endif
}

// ---- backward: the reversed Euler steps of a tile as the REV functor of nj_seg_bwd_tile ----
// prefetch slots (s.b_PRE): time [2][4], step size [2][4], then h [2][P][sH]
template <class D>
NJ_HD void nj_segtpn_bwd_build(const NjCfg& c, const NjSeg& s, const NjArgs& a, float* smem, const NjSegB& t0, const NjSegB& te, int o,
                               int e, bool have, bool pre, int cta) {
    constexpr int R = NJN_SEG_R;
    const int P = s.P_b, inf4 = ((c.inf + 3) >> 2) << 2;
    float* pre_f = smem + s.b_PRE;
    float* HP = pre_f + 16;
    if (have) nj_cp_wait();
    if (o == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int len = t0.I[NJS_I_LEN * P + r], k = t0.I[NJS_I_S0 * P + r] + e;
            if (!have) pre_f[8 + (e & 1) * 4 + r] = e < len ? NJ_LDG(a.b.step_dt + k) : 0.f;
            if (pre) {
                if (e - 1 < len) {
                    nj_cp_async4(pre_f + ((e - 1) & 1) * 4 + r, a.b.step_t + k - 1);
                    nj_cp_async4(pre_f + 8 + ((e - 1) & 1) * 4 + r, a.b.step_dt + k - 1);
                } else pre_f[8 + ((e - 1) & 1) * 4 + r] = 0.f;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c_ = o + NJN_F * i;
        if (c_ >= inf4) break;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = t0.I[NJS_I_PATH * P + r], len = t0.I[NJS_I_LEN * P + r], k = t0.I[NJS_I_S0 * P + r] + e;
            const bool active = e < len;
            float v = 0.f;
            if (c_ < c.d) v = t0.TX[r * s.sD + c_];
            else if (c_ < c.d + c.H) {
                const int cc = c_ - c.d;
                // h before step e: the history of the forward pass, or the recomputed chain of this CTA (recompute mode)
                const float* hh = a.scratch ? a.scratch + ((size_t)cta * a.b.S * P + (size_t)e * P + r) * s.sH + cc
                                            : a.h_hist + ((size_t)k * a.b.B + p) * c.H + cc;
                float h = 0.f;
                if (active) h = have ? HP[((e & 1) * P + r) * s.sH + cc] : NJ_LDG(hh);
                if (pre && e - 1 < len)
                    nj_cp_async4(HP + (((e - 1) & 1) * P + r) * s.sH + cc, a.scratch ? hh - (size_t)P * s.sH : hh - (size_t)a.b.B * c.H);
                v = nj_tanh(h);
            } else if (c_ < c.inf) {
                const float tcur = active ? (have ? pre_f[(e & 1) * 4 + r] : NJ_LDG(a.b.step_t + k)) : 0.f;
                v = nj_tpn_time_col(c, c_, t0.F[NJS_F_TAU * P + r], tcur);
            }
            te.IN[(size_t)r * s.sI + c_] = v;
        }
    }
    if (o == NJN_F - 1) {
#pragma unroll
        for (int r = 0; r < R; ++r)
            t0.I[NJS_I_RK * P + r] = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(t0.I[NJS_I_PATH * P + r] + a.b.path_id_offset),
                                                     (unsigned)(t0.I[NJS_I_S0 * P + r] + e));
    }
}

template <class D>
NJ_HD void nj_segtpn_bwd_t3(const NjCfg& c, const NjSeg& s, const float* smem, const NjSegB& t0, const NjSegB* te, const NjSegB* tn, int eT, int eF,
                            const NjTpnT<D>& q, int k) {
    constexpr int R = NJN_SEG_R;
    const int P = s.P_b;
    float acc[R];
    if (te) nj_tpn_dot<D::KCH, R>(q.c0, te->G, s.sA, acc);
    const bool hcol = k >= c.d && k < c.d + c.H;
    const int H4 = ((c.H + 3) >> 2) << 2;
    const float* ds = smem + s.b_PRE + 8 + (eF & 1) * 4;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (te && hcol && eT < t0.I[NJS_I_LEN * P + r]) {
            const float th = te->IN[(size_t)r * s.sI + k];
            t0.GH[r * s.sH + k - c.d] += acc[r] * (1.f - th * th);
        }
        if (tn && hcol) tn->GOUT[(size_t)r * s.sO + k - c.d] = ds[r] * t0.GH[r * s.sH + k - c.d];      // (step size 0: the row rests)
        else if (tn && k >= c.d + c.H && k < c.d + H4) tn->GOUT[(size_t)r * s.sO + k - c.d] = 0.f;
    }
}

template <class D, int ROLE>
struct NjSegTpnRev {
    static constexpr bool stat = true;
    static constexpr bool glue = (ROLE == NJN_ALL || ROLE == NJN_ROLE_G);
    static constexpr bool ovf = true;              // every jump-network dW tile goes through the gradient image
    static constexpr bool coop = NJN_SEG_COOP;     // the layers of the glue sections: served by the F threads / the glue warp's own GEMMs
    const NjCfg& c; const NjSeg& s; const NjArgs& a; float* smem; const NjSegB& t;
    NjTpnF<D>* njn_f; NjTpnT<D>* njn_t; float* njn_d; int cta; float* gimg; NjCoopMB* mb;

    NJ_HD float* gpart(float*) const { return gimg; }
    NJ_HD void* mailbox() const { return mb; }
    // a warp-local section of nj_seg_bwd_tile: the glue warp runs it and releases the servers, the F warps serve its layers
    NJ_HD void serve(int) const {
#if !defined(NJODE_HOST_SIM)
        if (coop && ROLE == NJN_ROLE_F) NJN_COOP_SERVE(mb, (int)threadIdx.x - NJN_F0);
#endif
    }
    NJ_HD void glue_done() const {
#if !defined(NJODE_HOST_SIM)
        if (coop && ROLE == NJN_ROLE_G) NJN_COOP_DONE(mb);
#endif
    }

    // steps n - 1 ... 0 of the tile, n + 2 pipeline iterations
    NJ_HD void run(int n) const {
        constexpr int R = NJN_SEG_R;
        const int P = s.P_b;
        for (int it = 0; it <= n + 1; ++it) {
            const int eF = n - 1 - it, eT = eF + 1, eD = eF + 2;
            const bool vF = it < n, vT = it >= 1 && it <= n, vD = it >= 2;
            NjSegB tF = t, tT = t, tD = t;
            nj_tpn_set(tF, ((eF + 3) % 3) * s.b_copy); nj_tpn_set(tT, ((eT + 3) % 3) * s.b_copy); nj_tpn_set(tD, ((eD + 3) % 3) * s.b_copy);
            NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { if (vF) nj_segtpn_bwd_build<D>(c, s, a, smem, t, tF, o, eF, it > 0, it + 1 < n, cta); }
            NJN_ROLE(NJN_ROLE_T, NJN_T0, NJN_T, x) { if (vT) nj_tpn_bwd_t1<D>(c, s, tT, NJN_TREGS(x), x); }
            NJN_ROLE(NJN_ROLE_D, NJN_D0, NJN_D, x) { if (vD) nj_tpn_dw<R>(c, s, smem, tD.IN, NJN_DACC_OF(x), x); }
            NJN_SYNC_FT();
            NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) {
                if (vF) nj_tpn_hidden<D::KC0, R>(c, 0, NJN_FREGS(o).w0, NJN_FREGS(o).b0, o, tF.IN, s.sI, tF.A, s.sA, t.I + NJS_I_RK * P);
            }
            NJN_ROLE(NJN_ROLE_T, NJN_T0, NJN_T, x) { if (vT) nj_tpn_bwd_t2<D>(c, s, tT, NJN_TREGS(x), x); }
            NJN_SYNC_FT();
            NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) {
                if (vF) nj_tpn_hidden<D::KCH, R>(c, 1, NJN_FREGS(o).w1, NJN_FREGS(o).b1, o, tF.A, s.sA, tF.A + P * s.sA, s.sA, t.I + NJS_I_RK * P);
                if (o == 0) nj_cp_wait();
            }
            NJN_ROLE(NJN_ROLE_T, NJN_T0, NJN_T, x) {
                if (vT || vF) nj_segtpn_bwd_t3<D>(c, s, smem, t, vT ? &tT : nullptr, vF ? &tF : nullptr, eT, eF, NJN_TREGS(x), x);
            }
            NJ_SYNC();
        }
    }
};

template <class D, int ROLE>
NJ_HD void nj_segtpn_bwd_body(const NjCfg& c, const NjSeg& s, const NjArgs& a, float* smem, int cta) {
    const int nt = NJN_NT_BWD;
    const int ode_floats = c.net[NJODE_NET_ENC].w_img[0];
    float* gout = a.partials + (size_t)cta * c.img_floats;
    NjSegB t;
    nj_segb_bind(t, s, smem);
    NJN_FREGS_DECL(D);
    NJN_TREGS_DECL(D);
    NJN_DACC_DECL();
    float* simg = smem + s.b_img;
    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_f_load<D>(c, simg, o, NJN_FREGS(o)); }
    NJN_ROLE(NJN_ROLE_T, NJN_T0, NJN_T, k) { nj_tpn_t_load<D>(c, simg, k, NJN_TREGS(k)); }
    NJN_ROLE(NJN_ROLE_D, NJN_D0, NJN_D, x) { nj_tpn_dw_table(c, s, smem, x); }
    NjCoopMB* mb = reinterpret_cast<NjCoopMB*>(smem + s.b_MB);
    NJ_THREADS(tid, nt) { if (tid == 0) { mb->R = NJN_SEG_R; mb->op = 0; } }
    const NjSegTpnRev<D, ROLE> rev{c, s, a, smem, t, njn_f, njn_t, njn_d, cta, smem + s.b_GIMG - ode_floats, mb};
    int* ctl = t.I + NJS_I_COUNT * s.P_b;
    for (;;) {
        NJ_THREADS(tid, nt) { if (tid == 0) ctl[0] = nj_atomic_inc(a.counter); }
        NJ_SYNC();
        const int tile = ctl[0];
        NJ_SYNC();
        if (tile >= s.n_tiles_b) break;
        int ub, ue;
        nj_seg_tile_lookup(s.b_ncls, s.b_t0, s.b_u0, s.b_u1, s.b_tr, 4 * s.nw_b, tile, ub, ue);
        nj_seg_bwd_tile<1, NjSegTpnRev<D, ROLE>>(c, s, a, smem, t, nullptr, cta, ub, ue, rev);
    }
    NJN_ROLE(NJN_ROLE_D, NJN_D0, NJN_D, x) { nj_tpn_dw_flush(c, s, NJN_DACC_OF(x), gout, x); }
    const float* gimg = smem + s.b_GIMG - ode_floats;
    NJ_THREADS(tid, nt) { for (int i = ode_floats + tid; i < c.img_floats; i += nt) gout[i] = gimg[i]; }
}

template <class D>
NJ_HD void nj_segtpn_cta_backward(const NjCfg& c, const NjSeg& s, const NjArgs& a, float* smem, int cta) {
    const int nt = NJN_NT_BWD;
    nj_stage_image(smem + s.b_img, a.image, c.img_floats, nt);
    nj_zero(smem + s.b_IN, s.b_smem_floats - s.b_IN, nt);
    nj_zero(a.partials + (size_t)cta * c.img_floats, c.net[NJODE_NET_ENC].w_img[0], nt);
    NJ_SYNC();
#if defined(NJODE_HOST_SIM)
    nj_segtpn_bwd_body<D, NJN_ALL>(c, s, a, smem, cta);
#else
    if (threadIdx.x < NJN_F0) nj_segtpn_bwd_body<D, NJN_ROLE_G>(c, s, a, smem, cta);
    else if (threadIdx.x < NJN_T0) nj_segtpn_bwd_body<D, NJN_ROLE_F>(c, s, a, smem, cta);
    else if (threadIdx.x < NJN_D0) nj_segtpn_bwd_body<D, NJN_ROLE_T>(c, s, a, smem, cta);
    else nj_segtpn_bwd_body<D, NJN_ROLE_D>(c, s, a, smem, cta);
#endif
}
