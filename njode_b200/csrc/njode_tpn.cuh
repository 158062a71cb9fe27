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
#define NJN_SYNC_FT() ((void)0)
#else
#define NJN_ROLE(ROLEID, lo, n, IDX) if (ROLE == ROLEID) for (int IDX = (int)threadIdx.x - (lo), _nj_xe = IDX + 1; IDX < _nj_xe; ++IDX)
#define NJN_SYNC_F() do { if (ROLE == NJN_ROLE_F) asm volatile("bar.sync 1, 64;" ::: "memory"); } while (0)
#define NJN_SYNC_FT() do { if (ROLE == NJN_ROLE_F || ROLE == NJN_ROLE_T) asm volatile("bar.sync 1, 160;" ::: "memory"); } while (0)
#endif

template <class D> struct NjTpnF {
    float w0[4 * D::KC0], w1[4 * D::KCH], w2[4 * D::KCH];
    float b0, b1, b2;
    float hpre[2 * D::R];            // h of the NEXT step to rebuild (backward), loaded one pipeline iteration ahead
    float tpre, dpre;                // time / step size of the next step
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
#pragma unroll
    for (int j = 0; j < 2 * D::R; ++j) f.hpre[j] = 0.f;
    f.tpre = 0.f; f.dpre = 0.f;
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

// acc[r] = sum_j w[j] * x[r][j] over KC float4 chunks; four partial sums per row (the chain of dependent FFMA is KC long)
template <int KC, int R>
NJ_HD void nj_tpn_dot(const float* w, const float* x, int x_s, float* acc) {
    float p[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r) { p[r][0] = 0.f; p[r][1] = 0.f; p[r][2] = 0.f; p[r][3] = 0.f; }
    const nj_sp xp = nj_sp_of(x);
#pragma unroll
    for (int q = 0; q < KC; ++q) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const nj_f4 v = nj_sp_ld4(NJ_SP_ADD(xp, r * x_s + 4 * q));
            p[r][0] = fmaf(w[4 * q], v.x, p[r][0]); p[r][1] = fmaf(w[4 * q + 1], v.y, p[r][1]);
            p[r][2] = fmaf(w[4 * q + 2], v.z, p[r][2]); p[r][3] = fmaf(w[4 * q + 3], v.w, p[r][3]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = (p[r][0] + p[r][1]) + (p[r][2] + p[r][3]);
}

// hidden layer l of the ODE network, output o, all rows: out[r][o] = dropout(act(w . in[r] + b))
template <int KC, int R>
NJ_HD void nj_tpn_hidden(const NjCfg& c, int l, const float* w, float bias, int o, const float* in, int in_s, float* out, int out_s,
                         const int* rk) {
    const NjNet& N = c.net[NJODE_NET_ODE];
    float acc[R];
    nj_tpn_dot<KC, R>(w, in, in_s, acc);
    if (o >= (((N.dim[l + 1] + 3) >> 2) << 2)) return;
    const bool pad = o >= N.dim[l + 1];          // columns up to the next multiple of 4 are read by the dW tiles: keep them 0
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float v = nj_act(acc[r] + bias, N.act[l]);
        if (c.has_drop) {
            const unsigned lk = nj_layer_key((unsigned)rk[r], (unsigned)(NJODE_NET_ODE * 16 + l + 1));
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
// forward
// ================================================================================================
// F thread o: the whole input rows of step k from the state (first step after a jump / the start / a record)
template <class D>
NJ_HD void nj_tpn_fwd_build(const NjCfg& c, const NjPath& s, const NjArgs& a, NjPathFwd<(D::R >= 4 ? 4 : D::R), 1>& f, int o, int k) {
    constexpr int R = D::R, RS = NJP_RS;
    const int inf4 = ((c.inf + 3) >> 2) << 2;
    const float tcur = NJ_LDG(a.b.step_t + k), dt = NJ_LDG(a.b.step_dt + k);
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
            f.w.IN[(size_t)r * s.sI + c_] = v;
        }
    }
    if (o == NJN_F - 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            f.w.RK[r] = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(f.I[NJP_I_PATH * RS + r] + a.b.path_id_offset), (unsigned)k);
            f.F[NJP_F_CA * RS + (k & 1) * 4 + r] = dt;            // (the loss-coefficient slots are free in the forward pass)
        }
    }
}

// the three phases of Euler step k for F thread o.  next: step k + 1 follows without a jump in between -- phase 3 then
// also writes what changes in the input rows (tanh(h), the time columns), the history and the dropout keys of step k + 1
template <class D>
NJ_HD void nj_tpn_fwd_p1(const NjCfg& c, const NjPath& s, const NjArgs& a, NjPathFwd<(D::R >= 4 ? 4 : D::R), 1>& f, NjTpnF<D>& q, int o,
                         int k, bool next) {
    if (next && o == NJN_F - 1) { q.tpre = NJ_LDG(a.b.step_t + k + 1); q.dpre = NJ_LDG(a.b.step_dt + k + 1); }
    nj_tpn_hidden<D::KC0, D::R>(c, 0, q.w0, q.b0, o, f.w.IN, s.sI, f.w.A0, s.sA, f.w.RK);
}
template <class D>
NJ_HD void nj_tpn_fwd_p2(const NjCfg& c, const NjPath& s, NjPathFwd<(D::R >= 4 ? 4 : D::R), 1>& f, NjTpnF<D>& q, int o) {
    nj_tpn_hidden<D::KCH, D::R>(c, 1, q.w1, q.b1, o, f.w.A0, s.sA, f.w.A1, s.sA, f.w.RK);
}
template <class D>
NJ_HD void nj_tpn_fwd_p3(const NjCfg& c, const NjPath& s, const NjArgs& a, NjPathFwd<(D::R >= 4 ? 4 : D::R), 1>& f, NjTpnF<D>& q, int o,
                         int k, bool next) {
    constexpr int R = D::R, RS = NJP_RS;
    float acc[R];
    nj_tpn_dot<D::KCH, R>(q.w2, f.w.A1, s.sA, acc);
    if (o < c.H) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float h = fmaf(f.F[NJP_F_CA * RS + (k & 1) * 4 + r], acc[r] + q.b2, f.HS[r * s.sH + o]);
            f.HS[r * s.sH + o] = h;
            if (next) {
                const int p = f.I[NJP_I_PATH * RS + r];
                if (p >= 0 && a.h_hist) a.h_hist[((size_t)(k + 1) * a.b.B + p) * c.H + o] = h;
                f.w.IN[(size_t)r * s.sI + c.d + o] = nj_tanh(h);
            }
        }
    }
    if (next && o == NJN_F - 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float tau = f.F[NJP_F_TAU * RS + r];
            for (int c_ = c.d + c.H + 1; c_ < c.inf; ++c_) f.w.IN[(size_t)r * s.sI + c_] = nj_tpn_time_col(c, c_, tau, q.tpre);
            f.w.RK[r] = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(f.I[NJP_I_PATH * RS + r] + a.b.path_id_offset), (unsigned)(k + 1));
            f.F[NJP_F_CA * RS + ((k + 1) & 1) * 4 + r] = q.dpre;
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
    const float* simg = smem;
    float* reg = smem + s.f_warp0;
    int* slot = reinterpret_cast<int*>(reg + s.f_I) + NJP_I_COUNT * NJP_RS;
    NjPathFwd<RG, 1> f(c, s, a, reg, simg);
    NJN_FREGS_DECL(D);
    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_f_load<D>(c, simg, o, NJN_FREGS(o)); }
    const bool rec = a.b.E > 0;
    const int S = a.b.S;
    for (;;) {
        NJ_THREADS(tid, NJN_NT_FWD) { if (tid == 0) *slot = nj_atomic_inc(a.counter); }
        NJ_SYNC();
        const int wt = *slot;
        NJ_SYNC();
        if (wt >= s.n_tiles_f) break;
        const int ub = wt * R, ue = ub + R < a.b.n_units ? ub + R : a.b.n_units;
        if (ROLE == NJN_ALL || ROLE == NJN_ROLE_G) { NJ_WARPS(wp, 1) { if (wp == 0) f.begin(ub, ue); } }
        NJ_SYNC();
        int k = 0, gi = 0;
        for (;;) {
            const int nk = f.next_jump(gi);
            const int kend = nk < S ? nk : S;
            if (rec) {
                for (; k < kend; ++k) {
                    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_build<D>(c, s, a, f, o, k); }
                    NJN_SYNC_F();
                    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_p1<D>(c, s, a, f, NJN_FREGS(o), o, k, false); }
                    NJN_SYNC_F();
                    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_p2<D>(c, s, f, NJN_FREGS(o), o); }
                    NJN_SYNC_F();
                    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_p3<D>(c, s, a, f, NJN_FREGS(o), o, k, false); }
                    NJ_SYNC();
                    if (ROLE == NJN_ALL || ROLE == NJN_ROLE_G) {
                        NJ_WARPS(wp, 1) { if (wp == 0) f.record(NJ_LDG(a.b.step_event + k), NJ_EVENT_PATH_RO_BASE + (unsigned)k); }
                    }
                    NJ_SYNC();
                }
            } else if (k < kend) {
                NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_build<D>(c, s, a, f, o, k); }
                NJN_SYNC_F();
                for (; k < kend; ++k) {
                    const bool next = k + 1 < kend;
                    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_p1<D>(c, s, a, f, NJN_FREGS(o), o, k, next); }
                    NJN_SYNC_F();
                    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_p2<D>(c, s, f, NJN_FREGS(o), o); }
                    NJN_SYNC_F();
                    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_fwd_p3<D>(c, s, a, f, NJN_FREGS(o), o, k, next); }
                    NJN_SYNC_F();
                }
            }
            NJ_SYNC();                            // the state after the run is visible to the glue warp
            if (nk > S) break;
            const bool any = f.any_jumps_at(nk);
            NJ_SYNC();                            // every thread has read the cursors before the glue warp advances them
            if (ROLE == NJN_ALL || ROLE == NJN_ROLE_G) {
                NJ_WARPS(wp, 1) {
                    if (wp == 0) {
                        if (any) f.jump(nk);
                        if (rec) f.record(NJ_LDG(a.b.jump_event + gi), NJ_EVENT_JUMP_BASE + 3u * (unsigned)gi + 2u);
                    }
                }
            }
            if (rec) ++gi;
            NJ_SYNC();
        }
        if (ROLE == NJN_ALL || ROLE == NJN_ROLE_G) { NJ_WARPS(wp, 1) { if (wp == 0) f.finish(); } }
        NJ_SYNC();
    }
}

template <class D>
NJ_HD void nj_tpn_cta_forward(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem) {
    nj_stage_image(smem, a.image, c.img_floats, NJN_NT_FWD);
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

// F, phase 1: the input rows of step e into its operand set (h from the history; the value was loaded one iteration ahead
// when `have`), then the load for step e - 1 is issued (`pre`)
template <class D>
NJ_HD void nj_tpn_bwd_build(const NjCfg& c, const NjPath& s, const NjArgs& a, const NjPathB& t0, const NjPathB& te, NjTpnF<D>& q, int o,
                            int e, bool have, bool pre) {
    constexpr int R = D::R;
    const int P = s.P_b, inf4 = ((c.inf + 3) >> 2) << 2;
    // (every thread keeps the time of the step: the time columns belong to whichever threads their indices fall on)
    float tcur = q.tpre;
    if (!have) tcur = NJ_LDG(a.b.step_t + e);
    if (pre) q.tpre = NJ_LDG(a.b.step_t + e - 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) {                // (the planner admits at most 2 * NJN_F input columns; constant indices into hpre)
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
                    h = have ? q.hpre[2 * r + i] : NJ_LDG(hh);
                    if (pre) q.hpre[2 * r + i] = NJ_LDG(hh - (size_t)a.b.B * c.H);
                }
                v = nj_tanh(h);
            } else if (c_ < c.inf) v = nj_tpn_time_col(c, c_, t0.F[NJP_F_TAU * P + r], tcur);
            te.IN[(size_t)r * s.sI + c_] = v;
        }
    }
    if (o == NJN_F - 1) {
#pragma unroll
        for (int r = 0; r < R; ++r)
            t0.I[NJB_I_RK * P + r] = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(t0.I[NJB_I_PATH * P + r] + a.b.path_id_offset), (unsigned)e);
    }
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

template <class D>
NJ_HD void nj_tpn_bwd_t1(const NjCfg& c, const NjPath& s, const NjPathB& te, const NjTpnT<D>& q, int k) {
    constexpr int R = D::R;
    const int wa = s.P_b * s.sA;
    float acc[R];
    nj_tpn_dot<D::HC, R>(q.c2, te.GOUT, s.sO, acc);
    const int O = c.net[NJODE_NET_ODE].dim[2];
    if (k >= (((O + 3) >> 2) << 2)) return;
#pragma unroll
    for (int r = 0; r < R; ++r) te.G[wa + (size_t)r * s.sA + k] = k < O ? nj_tpn_hidden_grad(c, 1, acc[r], te.A[wa + (size_t)r * s.sA + k]) : 0.f;
}
template <class D>
NJ_HD void nj_tpn_bwd_t2(const NjCfg& c, const NjPath& s, const NjPathB& te, const NjTpnT<D>& q, int k) {
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

// D thread x: dW of the ODE network for the rows of the step held in operand set `te`
NJ_HD void nj_tpn_dw(const NjCfg& c, const NjPath& s, const NjPathB& te, float* acc, int x, int Pt) {
    const int ode_tiles = s.tile_base[NJODE_NET_RO][0];        // the ODE network's tiles come first
#pragma unroll
    for (int slot = 0; slot < NJN_DSLOTS; ++slot) {
        const int T = slot * NJN_D + x;
        if (T >= ode_tiles) break;
        int l, og, kg;
        if (!nj_path_tile_decode(c, s, NJODE_NET_ODE, T, l, og, kg)) continue;
        nj_path_dw_rows(c, s, te, NJODE_NET_ODE, l, og, kg, Pt, nullptr, 1, acc + slot * 20);
    }
}
NJ_HD void nj_tpn_dw_flush(const NjCfg& c, const NjPath& s, const float* acc, float* gpart, int x) {
    const int ode_tiles = s.tile_base[NJODE_NET_RO][0];
#pragma unroll
    for (int slot = 0; slot < NJN_DSLOTS; ++slot) {
        const int T = slot * NJN_D + x;
        if (T >= ode_tiles) break;
        int l, og, kg;
        if (nj_path_tile_decode(c, s, NJODE_NET_ODE, T, l, og, kg)) nj_seg_tile_store(c, NJODE_NET_ODE, l, og, kg, acc + slot * 20, gpart, false);
    }
}

template <class D, int ROLE>
NJ_HD void nj_tpn_bwd_body(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem, int cta) {
    constexpr int R = D::R, RG = (R >= 4 ? 4 : R);
    const int nt = NJN_NT_BWD, P = s.P_b;
    const bool G = ROLE == NJN_ALL || ROLE == NJN_ROLE_G;
    float* gpart = a.partials + (size_t)cta * c.img_floats;
    NjPathB t;
    nj_pathb_bind(t, s, smem);
    const NjPathBwd<RG, 1> B(c, s, a, t, smem);
    NJN_FREGS_DECL(D);
    NJN_TREGS_DECL(D);
    NJN_DACC_DECL();
    NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { nj_tpn_f_load<D>(c, smem, o, NJN_FREGS(o)); }
    NJN_ROLE(NJN_ROLE_T, NJN_T0, NJN_T, k) { nj_tpn_t_load<D>(c, smem, k, NJN_TREGS(k)); }
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
                if (G) { NJ_WARPS(wp, 1) { if (wp == 0) B.jump_p1(0, 0, k); } }
                NJ_SYNC();
                NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t, NJODE_NET_RO, gpart, tid, nt, R, t.MSK, R); }
                NJ_SYNC();
                if (G) { NJ_WARPS(wp, 1) { if (wp == 0) B.jump_p2(0, 0); } }
                NJ_SYNC();
                NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t, c.use_rnn ? NJODE_NET_GRU_HH : NJODE_NET_ENC, gpart, tid, nt, R, t.MSK, R); }
                NJ_SYNC();
                if (c.use_rnn) {
                    if (G) { NJ_WARPS(wp, 1) { if (wp == 0) B.jump_p2b(0, 0); } }
                    NJ_SYNC();
                    NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t, NJODE_NET_GRU_IH, gpart, tid, nt, R, t.MSK, R); }
                    NJ_SYNC();
                }
                if (G) { NJ_WARPS(wp, 1) { if (wp == 0) B.jump_p3(0, 0); } }
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
                float dtn = 0.f;
                // phase 1
                NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) { if (vF) nj_tpn_bwd_build<D>(c, s, a, t, tF, NJN_FREGS(o), o, eF, j > 0, j + 1 < n); }
#if !defined(NJODE_HOST_SIM)
                if (ROLE == NJN_ROLE_T && vF) dtn = NJ_LDG(a.b.step_dt + eF);      // used in phase 3
#else
                if (vF) dtn = a.b.step_dt[eF];
#endif
                NJN_ROLE(NJN_ROLE_T, NJN_T0, NJN_T, x) { if (vT) nj_tpn_bwd_t1<D>(c, s, tT, NJN_TREGS(x), x); }
                NJN_ROLE(NJN_ROLE_D, NJN_D0, NJN_D, x) { if (vD) nj_tpn_dw(c, s, tD, NJN_DACC_OF(x), x, R); }
                NJN_SYNC_FT();
                // phase 2
                NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) {
                    if (vF) nj_tpn_hidden<D::KC0, R>(c, 0, NJN_FREGS(o).w0, NJN_FREGS(o).b0, o, tF.IN, s.sI, tF.A, s.sA, t.I + NJB_I_RK * P);
                }
                NJN_ROLE(NJN_ROLE_T, NJN_T0, NJN_T, x) { if (vT) nj_tpn_bwd_t2<D>(c, s, tT, NJN_TREGS(x), x); }
                NJN_SYNC_FT();
                // phase 3
                NJN_ROLE(NJN_ROLE_F, NJN_F0, NJN_F, o) {
                    if (vF) nj_tpn_hidden<D::KCH, R>(c, 1, NJN_FREGS(o).w1, NJN_FREGS(o).b1, o, tF.A, s.sA, tF.A + P * s.sA, s.sA, t.I + NJB_I_RK * P);
                }
                NJN_ROLE(NJN_ROLE_T, NJN_T0, NJN_T, x) {
                    if (vT || vF) nj_tpn_bwd_t3<D>(c, s, t, vT ? &tT : nullptr, vF ? &tF : nullptr, dtn, NJN_TREGS(x), x);
                }
                NJ_SYNC();
            }
            k = lo;
        }
        if (G) { NJ_WARPS(wp, 1) { if (wp == 0) B.start_local(0); } }
        NJ_SYNC();
        NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t, NJODE_NET_ENC, gpart, tid, nt, R, nullptr, R); }
        NJ_SYNC();
    }
    NJN_ROLE(NJN_ROLE_D, NJN_D0, NJN_D, x) { nj_tpn_dw_flush(c, s, NJN_DACC_OF(x), gpart, x); }
}

template <class D>
NJ_HD void nj_tpn_cta_backward(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem, int cta) {
    const int nt = NJN_NT_BWD;
    nj_stage_image(smem, a.image, c.img_floats, nt);
    nj_zero(smem + s.b_IN, s.b_smem_floats - s.b_IN, nt);
    nj_zero(a.partials + (size_t)cta * c.img_floats, c.img_floats, nt);
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
