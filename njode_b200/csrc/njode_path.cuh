// njode_path.cuh -- whole-path units (masked model, GRU jump, return_path) on the warp GEMMs of njode_seg.cuh.
//
// A whole path carries a true recurrence through its jumps (NJODE/models.py:464-467 masked imputation,
// 460-461 GRU cell, 483-484 last_X = Y), so it cannot be cut into independent segments.  The generic kernels
// (njode_core.cuh) march such units CTA-cooperatively with ~12 __syncthreads per Euler step; a PhysioNet-shaped
// batch (3 700 dependent steps) was bound by barrier latency at 0.2-2.5 % of the FMA roofline.  Here:
//   * forward: WARP-AUTONOMOUS.  A warp owns R = RG*TR paths from the start encoder to hT: Euler steps, jumps
//     (readout, imputation, encoder / GRU, readout, loss row), path recording -- only __syncwarp after the
//     parameter image is in shared memory.  Rows of a warp jump at different steps: the warp marches to the next
//     step at which any of its rows jumps (kept in a register), runs the jump networks over all R rows and commits
//     the result for the jumping rows only.
//   * small batches (the reference's PhysioNet batch is 50): RG < 4 row groups.  The 4/RG lanes that would hold
//     further rows split the reduction dimension of every layer GEMM instead and meet in two shuffles, so ONE
//     path per warp (RG = 1) runs a layer in a quarter of the dependent FMA chain.
//   * backward: a CTA owns P = R*nw rows in lockstep over the batch-global steps, two CTA barriers per Euler step
//     for the dW phase (thread-owned 4x4 register tiles, helper warps as in njode_seg.cuh).  A jump is reversed
//     warp-locally by the warps that own jumping rows; its three dW phases (readout, encoder / GRU, readout) run
//     over the jumping rows only (per-warp row masks).
// h at every Euler step comes from the forward's h_hist: an O(S*B*H) buffer, 30 MB for the PhysioNet batch -- the
// segment path (non-masked training) is the one that recomputes from its checkpoints.
//
// Same dual-compilation scheme as njode_seg.cuh (-DNJODE_HOST_SIM: sequential host simulation for the CPU tests).
#pragma once
#include "njode_seg.cuh"

#define NJP_RS 8                    // row-slot stride of the per-warp scalar arrays (forward: R <= 8)
enum { NJP_I_PATH = 0, NJP_I_CUR, NJP_I_END, NJP_I_NEXTK, NJP_I_ROW, NJP_I_ACT, NJP_I_RK, NJP_I_COUNT };
enum { NJP_F_TAU = 0, NJP_F_CA, NJP_F_CB, NJP_F_COUNT };
#define NJP_NEVER 0x7FFFFFFF

struct NjPath {
    int ok;
    int rg_f, tr_f, nw_f;           // forward: a warp owns rg*tr rows, nw warps per CTA
    int rg_b, tr_b, nw_b, nt_b;     // backward: nw_b row warps of rg*tr rows + helper warps up to nt_b threads
    int sI, sA, sO, sH, sD, s3, nA;
    int f_region, f_IN, f_A0, f_A1, f_OUT, f_HS, f_EE, f_LX, f_TX, f_XI, f_YBJ, f_YY, f_MM, f_GI, f_F, f_I;
    int f_warp0, f_smem_floats, n_tiles_f;
    int b_IN, b_A, b_G, b_GOUT, b_GZ, b_OUT, b_GH, b_HB, b_EE, b_GE, b_XI, b_LX, b_TX, b_YBJ, b_YY, b_GYBJ, b_GX, b_MM,
        b_GI, b_GHH, b_F, b_I;
    int b_smem_floats, P_b, n_tiles_b;
    int tile_base[NJODE_NUM_NETS][NJODE_MAX_LINEAR];    // first dW tile of (net, layer); order ODE, RO, ENC, GRU_HH, GRU_IH
    int tiles_total, nt_slots;
    // weight-stationary Euler steps (small batches): one CTA of nw_s warps per tile of rg*tr rows, ODE weights in registers
    int stat, nw_s, b_PART;
    // pipelined backward: dW of the ODE network on helper warps, operand buffers (IN, A, G, GOUT) twice, b_copy floats apart
    int pipe, b_copy;
    // thread-per-neuron kernels of small batches (njode_tpn.cuh): dimension class (1: demo networks, 2: PhysioNet-shaped),
    // operand buffers three times
    int tpn, tpn_fwd, b_TD, b_PRE, f_MB, b_MB, b_GIMG, f_IN2, f_AUX;      // tpn_fwd: the forward takes them too (else warp kernels)
};

// ------------------------------------------------------------------------------------------------
// warp GEMMs with RG row groups.  lane = (rg, og), rg = lane >> 3, og = lane & 7.  Row of the lane's i-th accumulator
// row: (rg mod RG) + RG*i; the KSN = 4/RG lanes that share rows split the float4 chunks of the reduction dimension
// (chunk q belongs to split q mod KSN) and add their partial sums by shuffles (xor 16, then xor 8).  The host
// simulation has no shuffles: there the split-0 lane evaluates all KSN partial sums itself, in the same order.
// ------------------------------------------------------------------------------------------------
template <int RG> struct NjRG {
    static constexpr int KSN = 4 / RG;
    NJ_HD static int row(int rg) { return RG == 4 ? rg : (RG == 2 ? (rg & 1) : 0); }
    NJ_HD static int ks(int rg) { return RG == 4 ? 0 : (RG == 2 ? (rg >> 1) : rg); }
};

template <int RG>
NJ_HD float nj_ks_sum(float v) {
#if !defined(NJODE_HOST_SIM)
    if (RG <= 2) v += __shfl_xor_sync(0xFFFFFFFFu, v, 16);
    if (RG == 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, 8);
#endif
    return v;
}
// host simulation: combines the KSN partial sums of one accumulator in the order of the device's shuffle tree
template <int RG>
NJ_HD float nj_ks_combine(const float* p, int stride) {
    if (RG == 4) return p[0];
    if (RG == 2) return p[0] + p[stride];
    return (p[0] + p[2 * stride]) + (p[stride] + p[3 * stride]);
}

template <int RG, int TR, int TO>
NJ_HD void nj_pg_fwd_partial(const NjWL& L, int rr, int og, int ks, float (&acc)[TR][TO]) {
    constexpr int KSN = 4 / RG;
#pragma unroll
    for (int j = 0; j < TO; ++j) {
        const float b = (L.bias && ks == 0) ? L.bias[L.o_base + og + 8 * j] : 0.f;
#pragma unroll
        for (int i = 0; i < TR; ++i) acc[i][j] = b;
    }
    nj_sp ap[TR], wp[TO];
#pragma unroll
    for (int i = 0; i < TR; ++i) ap[i] = nj_sp_of(L.in + (size_t)(rr + RG * i) * L.in_s);
#pragma unroll
    for (int j = 0; j < TO; ++j) wp[j] = nj_sp_of(L.W + (size_t)(L.o_base + og + 8 * j) * L.w_s);
#pragma unroll 2
    for (int k4 = ks; k4 < L.K4; k4 += KSN) {
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
}

// out[r][o] = act(b[o] + sum_k in[r][k] W[o][k]) (* dropout); outputs o = o_base + og + 8j
template <int RG, int TR, int TO>
NJ_HD void nj_pg_fwd(const NjWL& L, int lane) {
    const int rg = lane >> 3, og = lane & 7;
    const int rr = NjRG<RG>::row(rg), ks = NjRG<RG>::ks(rg);
    float acc[TR][TO];
#if defined(NJODE_HOST_SIM)
    if (ks != 0) return;
    {
        constexpr int KSN = 4 / RG;
        float part[KSN][TR][TO];
        for (int q = 0; q < KSN; ++q) nj_pg_fwd_partial<RG, TR, TO>(L, rr, og, q, part[q]);
        for (int i = 0; i < TR; ++i)
            for (int j = 0; j < TO; ++j) acc[i][j] = nj_ks_combine<RG>(&part[0][i][j], TR * TO);
    }
#else
    nj_pg_fwd_partial<RG, TR, TO>(L, rr, og, ks, acc);
    if (RG < 4) {
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int j = 0; j < TO; ++j) acc[i][j] = nj_ks_sum<RG>(acc[i][j]);
        if (ks != 0) return;
    }
#endif
    const unsigned obase16 = (unsigned)(L.o_base + og);
#pragma unroll
    for (int i = 0; i < TR; ++i) {
        const int r = rr + RG * i;
        float* orow = L.out + (size_t)r * L.out_s + L.o_base + og;
        if (L.drop) {
            const unsigned lk = nj_layer_key((unsigned)L.rk[r], L.tag);
#pragma unroll
            for (int j = 0; j < TO; j += 2) {
                const unsigned o = obase16 + 8u * j;               // neurons o and o ^ 8 share a hash word (nj_keep)
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

template <int RG, int TR>
NJ_HDN void nj_pg_layer_fwd(NjWL L, int to0, int to_last, int nch) {
    NJ_ASSUME_SHARED(L.in); NJ_ASSUME_SHARED(L.W); NJ_ASSUME_SHARED(L.out); NJ_ASSUME_SHARED(L.rk);
    if (L.bias) NJ_ASSUME_SHARED(L.bias);
    for (int ch = 0; ch < nch; ++ch) {
        L.o_base = ch * 8 * to0;
        const int to = ch < nch - 1 ? to0 : to_last;
        NJ_LANES(lane) {
            switch (to) {
                case 1: nj_pg_fwd<RG, TR, 1>(L, lane); break;
                case 2: nj_pg_fwd<RG, TR, 2>(L, lane); break;
                case 3: nj_pg_fwd<RG, TR, 3>(L, lane); break;
                case 4: nj_pg_fwd<RG, TR, 4>(L, lane); break;
                case 5: nj_pg_fwd<RG, TR, 5>(L, lane); break;
                case 6: nj_pg_fwd<RG, TR, 6>(L, lane); break;
                case 7: nj_pg_fwd<RG, TR, 7>(L, lane); break;
                default: nj_pg_fwd<RG, TR, 8>(L, lane); break;
            }
        }
    }
}

template <int RG, int TR, int TK>
NJ_HD void nj_pg_dx_partial(const NjWD& L, int rr, int kq, int ks, float (&acc)[TR][TK][4]) {
    constexpr int KSN = 4 / RG;
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
    for (int i = 0; i < TR; ++i) gp[i] = nj_sp_of(L.g + (size_t)(rr + RG * i) * L.g_s);
    const int ws = L.w_s;
    for (int o4 = ks; o4 < L.O4; o4 += KSN) {
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
}

// gin[r][k] = (sum_o g[r][o] W[o][k]) * act'(aprev[r][k]) * dropout factor; k-groups (float4) kg = kg_base + kq + 8*jk
template <int RG, int TR, int TK>
NJ_HD void nj_pg_dx(const NjWD& L, int lane) {
    const int rg = lane >> 3, kq = lane & 7;
    const int rr = NjRG<RG>::row(rg), ks = NjRG<RG>::ks(rg);
    float acc[TR][TK][4];
#if defined(NJODE_HOST_SIM)
    if (ks != 0) return;
    {
        constexpr int KSN = 4 / RG;
        float part[KSN][TR][TK][4];
        for (int q = 0; q < KSN; ++q) nj_pg_dx_partial<RG, TR, TK>(L, rr, kq, q, part[q]);
        for (int i = 0; i < TR; ++i)
            for (int jk = 0; jk < TK; ++jk)
                for (int c = 0; c < 4; ++c) acc[i][jk][c] = nj_ks_combine<RG>(&part[0][i][jk][c], TR * TK * 4);
    }
#else
    nj_pg_dx_partial<RG, TR, TK>(L, rr, kq, ks, acc);
    if (RG < 4) {
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int jk = 0; jk < TK; ++jk)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[i][jk][c] = nj_ks_sum<RG>(acc[i][jk][c]);
        if (ks != 0) return;
    }
#endif
#pragma unroll
    for (int i = 0; i < TR; ++i) {
        const int r = rr + RG * i;
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

template <int RG, int TR>
NJ_HDN void nj_pg_layer_dx(NjWD D) {
    NJ_ASSUME_SHARED(D.g); NJ_ASSUME_SHARED(D.W); NJ_ASSUME_SHARED(D.gin);
    if (D.aprev) NJ_ASSUME_SHARED(D.aprev);
    for (int kb = 0; kb < D.K4in; kb += 16) {
        D.kg_base = kb;
        const bool two = (D.K4in - kb) > 8;
        NJ_LANES(lane) {
            if (two) nj_pg_dx<RG, TR, 2>(D, lane); else nj_pg_dx<RG, TR, 1>(D, lane);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// per-warp view of the MLP buffers
// ------------------------------------------------------------------------------------------------
struct NjPW {
    const NjCfg* c;
    const float* wimg;
    float *IN, *A0, *A1, *OUT, *G0, *GOUT, *GZ;
    int sI, sA, sO;
    int a_buf_stride, g_buf_stride;
    int* RK;
    void* coop;                     // mailbox of the cooperative layer service (njode_tpn.cuh) or null: the warp's own GEMM
};

template <int RG, int TR, bool COOP = false>
NJ_HD void nj_path_mlp_fwd(const NjPW& w, int netid, bool keep_all, bool skip_last) {
    const NjCfg& c = *w.c;
    const NjNet& N = c.net[netid];
    const float* in = w.IN; int in_s = w.sI;
    for (int l = 0; l < N.n; ++l) {
        const bool last = (l == N.n - 1);
        if (last && skip_last) break;
        NjWL L;
        L.in = in; L.in_s = in_s; L.K4 = (N.dim[l] + 3) >> 2;
        L.W = w.wimg + N.w_img[l]; L.w_s = N.ks[l];
        L.bias = N.b_src[l] >= 0 ? w.wimg + N.b_img[l] : nullptr;
        if (last) { L.out = w.OUT; L.out_s = w.sO; }
        else { L.out = keep_all ? w.A0 + (size_t)l * w.a_buf_stride : ((l & 1) ? w.A1 : w.A0); L.out_s = w.sA; }
        L.act = last ? NJODE_ACT_NONE : N.act[l];
        L.drop = (!last) && c.has_drop; L.thr = c.thr; L.keep_scale = c.keep_scale;
        L.rk = w.RK; L.tag = (unsigned)(netid * 16 + l + 1);
        L.o_base = 0;
        if (COOP) nj_coop_post_fwd(w.coop, L, 8 * (N.to[l] * (N.nch[l] - 1) + N.tol[l]));
        else nj_pg_layer_fwd<RG, TR>(L, N.to[l], N.tol[l], N.nch[l]);
        NJ_SYNCWARP();
        in = L.out; in_s = L.out_s;
    }
}

template <int RG, int TR, bool COOP = false>
NJ_HD void nj_path_mlp_dx(const NjPW& w, int netid, bool need_in_grad) {
    const NjCfg& c = *w.c;
    const NjNet& N = c.net[netid];
    for (int l = N.n - 1; l >= 0; --l) {
        if (l == 0 && !need_in_grad) break;
        NjWD D;
        if (l == N.n - 1) { D.g = w.GOUT; D.g_s = w.sO; } else { D.g = w.G0 + (size_t)l * w.g_buf_stride; D.g_s = w.sA; }
        D.O4 = (N.dim[l + 1] + 3) >> 2;
        D.W = w.wimg + N.w_img[l]; D.w_s = N.ks[l];
        D.K4in = (N.dim[l] + 3) >> 2;
        if (l > 0) {
            D.gin = w.G0 + (size_t)(l - 1) * w.g_buf_stride; D.gin_s = w.sA;
            D.aprev = w.A0 + (size_t)(l - 1) * w.a_buf_stride; D.a_s = w.sA; D.act_prev = N.act[l - 1];
        } else {
            D.gin = w.GZ; D.gin_s = w.sI; D.aprev = nullptr; D.a_s = 0; D.act_prev = NJODE_ACT_NONE;
        }
        D.drop = c.has_drop; D.keep_scale = c.keep_scale; D.one_minus_p = c.one_minus_p;
        D.kg_base = 0;
        if (COOP) nj_coop_post_dx(w.coop, D);
        else nj_pg_layer_dx<RG, TR>(D);
        NJ_SYNCWARP();
    }
}

NJ_HD unsigned nj_path_jump_key(const NjCfg& c, const NjArgs& a, int p, int jump, unsigned which) {
    return nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(p + a.b.path_id_offset), NJ_EVENT_JUMP_BASE + 3u * (unsigned)jump + which);
}

// ================================================================================================
// forward: one warp = R = RG*TR whole paths
// ================================================================================================
template <int RG, int TR, bool COOP = false>
struct NjPathFwd {
    static constexpr int R = RG * TR;
    static constexpr int RS = NJP_RS;
    const NjCfg& c; const NjPath& s; const NjArgs& a;
    NjPW w;
    float *HS, *EE, *LX, *TX, *XI, *YBJ, *YY, *MM, *GI, *F;
    int* I;
    int d4, H4, inf4, ein4;

    NJ_HD NjPathFwd(const NjCfg& c_, const NjPath& s_, const NjArgs& a_, float* reg, const float* wimg) : c(c_), s(s_), a(a_) {
        w.c = &c; w.wimg = wimg;
        w.IN = reg + s.f_IN; w.A0 = reg + s.f_A0; w.A1 = reg + s.f_A1; w.OUT = reg + s.f_OUT;
        w.G0 = w.GOUT = w.GZ = nullptr; w.a_buf_stride = 0; w.g_buf_stride = 0;
        w.sI = s.sI; w.sA = s.sA; w.sO = s.sO; w.coop = nullptr;
        HS = reg + s.f_HS; EE = reg + s.f_EE; LX = reg + s.f_LX; TX = reg + s.f_TX; XI = reg + s.f_XI; YBJ = reg + s.f_YBJ; YY = reg + s.f_YY;
        MM = reg + s.f_MM; GI = reg + s.f_GI; F = reg + s.f_F;
        I = reinterpret_cast<int*>(reg + s.f_I);
        w.RK = I + NJP_I_RK * RS;
        d4 = ((c.d + 3) >> 2) << 2; H4 = ((c.H + 3) >> 2) << 2; inf4 = ((c.inf + 3) >> 2) << 2; ein4 = ((c.enc_in + 3) >> 2) << 2;
    }

    // cursor -> (next jump row, step count at which it happens)
    NJ_HD void set_next(int r) {
        const int cur = I[NJP_I_CUR * RS + r];
        if (cur < I[NJP_I_END * RS + r]) {
            const int row = NJ_LDG(a.b.path_rows + cur);
            I[NJP_I_ROW * RS + r] = row;
            I[NJP_I_NEXTK * RS + r] = NJ_LDG(a.b.jump_step + NJ_LDG(a.b.row_jump + row));
        } else { I[NJP_I_ROW * RS + r] = -1; I[NJP_I_NEXTK * RS + r] = NJP_NEVER; }
    }

    // path_h / path_y record e: h and readout(h) of every row (return_path, NJODE/models.py:440-444, 491-494)
    NJ_HD void record(int e, unsigned event_key) {
        const int sI = s.sI, sO = s.sO, sH = s.sH;
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int p = I[NJP_I_PATH * RS + er];
            for (int c_ = ec0; c_ < H4; c_ += LPR) {
                const float h = c_ < c.H ? HS[er * sH + c_] : 0.f;
                if (p >= 0 && c_ < c.H) a.path_h[((size_t)e * a.b.B + p) * c.H + c_] = h;
                w.IN[(size_t)er * sI + c_] = c_ < c.H ? nj_tanh(h) : 0.f;
            }
            if (ec0 == 0) w.RK[er] = p >= 0 ? (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(p + a.b.path_id_offset), event_key) : 0;
        }
        NJ_SYNCWARP();
        nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_RO, false, false);
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int p = I[NJP_I_PATH * RS + er];
            if (p >= 0)
                for (int c_ = ec0; c_ < c.dout; c_ += LPR) {
                    float y = w.OUT[er * sO + c_];
                    if (c.residual) y += nj_resid(HS + er * sH, c.H, c.dout, c_);
                    a.path_y[((size_t)e * a.b.B + p) * c.dout + c_] = y;
                }
        }
        NJ_SYNCWARP();
    }

    NJ_HD void euler_step(int k) {
        const int sI = s.sI, sO = s.sO, sH = s.sH, sD = s.sD;
        const float tcur = NJ_LDG(a.b.step_t + k), dt = NJ_LDG(a.b.step_dt + k);
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int p = I[NJP_I_PATH * RS + er];
            const float tau = F[NJP_F_TAU * RS + er];
            float* hh = (p >= 0 && a.h_hist) ? a.h_hist + ((size_t)k * a.b.B + p) * c.H : nullptr;
            for (int c_ = ec0; c_ < inf4; c_ += LPR) {
                float v = 0.f;
                if (c_ < c.d) v = TX[er * sD + c_];
                else if (c_ < c.d + c.H) {
                    const float h = HS[er * sH + c_ - c.d];
                    if (hh) hh[c_ - c.d] = h;
                    v = nj_tanh(h);
                } else if (c_ < c.inf) {
                    if (c_ == c.d + c.H) v = tau;
                    else if (c_ == c.d + c.H + 1) v = tcur - tau;
                    else v = tau + (tcur - tau);
                }
                w.IN[(size_t)er * sI + c_] = v;
            }
            if (ec0 == 0) w.RK[er] = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(p + a.b.path_id_offset), (unsigned)k);
        }
        NJ_SYNCWARP();
        nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_ODE, false, false);
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            for (int c_ = ec0; c_ < c.H; c_ += LPR)
                HS[er * sH + c_] = fmaf(dt, w.OUT[er * sO + c_], HS[er * sH + c_]);
        }
        NJ_SYNCWARP();
    }

    // the jump of NJODE/models.py:449-489 for the rows whose next jump happens after nk Euler steps
    NJ_HD void jump(int nk) {
        const int sI = s.sI, sO = s.sO, sH = s.sH, sD = s.sD, s3 = s.s3, H = c.H;
        // (a) Y_bj = readout(h before the jump)
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const bool act = I[NJP_I_NEXTK * RS + er] == nk;
            const int row = I[NJP_I_ROW * RS + er], p = I[NJP_I_PATH * RS + er];
            for (int c_ = ec0; c_ < H4; c_ += LPR) {
                const float h = c_ < H ? HS[er * sH + c_] : 0.f;
                if (act && c_ < H && a.h_before) a.h_before[(size_t)row * H + c_] = h;
                w.IN[(size_t)er * sI + c_] = c_ < H ? nj_tanh(h) : 0.f;
            }
            if (ec0 == 0) {
                I[NJP_I_ACT * RS + er] = act ? 1 : 0;
                w.RK[er] = act ? (int)nj_path_jump_key(c, a, p, NJ_LDG(a.b.row_jump + row), 0u) : 0;
            }
        }
        NJ_SYNCWARP();
        nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_RO, false, false);
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            for (int c_ = ec0; c_ < c.dout; c_ += LPR) {
                float y = w.OUT[er * sO + c_];
                if (c.residual) y += nj_resid(HS + er * sH, H, c.dout, c_);
                YBJ[er * sD + c_] = y;
            }
        }
        NJ_SYNCWARP();
        if (c.use_rnn) {
            // (b') h[i_obs] = GRUCell(tanh(X_obs), tanh(h[i_obs]))  (NJODE/models.py:202-217, 460-461; gates r, z, n)
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const bool act = I[NJP_I_ACT * RS + er] != 0;
                const int row = I[NJP_I_ROW * RS + er];
                for (int c_ = ec0; c_ < d4; c_ += LPR) {
                    const float x = (act && c_ < c.d) ? NJ_LDG(a.b.X + (size_t)row * c.d + c_) : 0.f;
                    XI[er * sD + c_] = x;
                    w.IN[(size_t)er * sI + c_] = c_ < c.d ? nj_tanh(x) : 0.f;
                }
            }
            NJ_SYNCWARP();
            nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_GRU_IH, false, false);
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                for (int c_ = ec0; c_ < 3 * H; c_ += LPR) GI[er * s3 + c_] = w.OUT[er * sO + c_];
                for (int c_ = ec0; c_ < H4; c_ += LPR) w.IN[(size_t)er * sI + c_] = c_ < H ? nj_tanh(HS[er * sH + c_]) : 0.f;
            }
            NJ_SYNCWARP();
            nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_GRU_HH, false, false);
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const bool act = I[NJP_I_ACT * RS + er] != 0;
                for (int c_ = ec0; c_ < H4; c_ += LPR) {
                    float e = 0.f;
                    if (c_ < H) {
                        const float* gi = GI + er * s3;
                        const float* o = w.OUT + er * sO;
                        const float hh = w.IN[(size_t)er * sI + c_];
                        const float r = nj_sigmoid(gi[c_] + o[c_]);
                        const float z = nj_sigmoid(gi[H + c_] + o[H + c_]);
                        const float n = nj_tanh(fmaf(r, o[2 * H + c_], gi[2 * H + c_]));
                        e = act ? fmaf(z, hh - n, n) : HS[er * sH + c_];
                    }
                    if (c_ < H) EE[er * sH + c_] = e;
                }
            }
            NJ_SYNCWARP();
        } else {
            // (b) encoder at the (imputed) observation
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const bool act = I[NJP_I_ACT * RS + er] != 0;
                const int row = I[NJP_I_ROW * RS + er], p = I[NJP_I_PATH * RS + er];
                for (int c_ = ec0; c_ < ein4; c_ += LPR) {
                    float v = 0.f;
                    if (c_ < c.d) {
                        float x = act ? NJ_LDG(a.b.X + (size_t)row * c.d + c_) : 0.f;
                        if (c.masked) {
                            const float m = act ? NJ_LDG(a.b.M + (size_t)row * c.d + c_) : 0.f;
                            x = x * m + (1.f - m) * YBJ[er * sD + c_];
                            MM[er * sD + c_] = m;
                        }
                        XI[er * sD + c_] = x;
                        v = nj_tanh(x);
                    } else if (c.masked && c_ < 2 * c.d) {
                        v = act ? NJ_LDG(a.b.M + (size_t)row * c.d + c_ - c.d) : 0.f;
                    }
                    w.IN[(size_t)er * sI + c_] = v;
                }
                if (ec0 == 0) w.RK[er] = act ? (int)nj_path_jump_key(c, a, p, NJ_LDG(a.b.row_jump + row), 1u) : 0;
            }
            NJ_SYNCWARP();
            nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_ENC, false, false);
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const bool act = I[NJP_I_ACT * RS + er] != 0;
                for (int c_ = ec0; c_ < H; c_ += LPR) {
                    float e = w.OUT[er * sO + c_];
                    if (c.residual) e += nj_resid(XI + er * sD, c.d, H, c_);
                    EE[er * sH + c_] = act ? e : HS[er * sH + c_];
                }
            }
            NJ_SYNCWARP();
        }
        // (c) commit h, Y = readout(h after the jump)
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const bool act = I[NJP_I_ACT * RS + er] != 0;
            const int row = I[NJP_I_ROW * RS + er], p = I[NJP_I_PATH * RS + er];
            for (int c_ = ec0; c_ < H4; c_ += LPR) {
                float e = 0.f;
                if (c_ < H) { e = EE[er * sH + c_]; HS[er * sH + c_] = e; }
                w.IN[(size_t)er * sI + c_] = c_ < H ? nj_tanh(e) : 0.f;
            }
            if (ec0 == 0) w.RK[er] = act ? (int)nj_path_jump_key(c, a, p, NJ_LDG(a.b.row_jump + row), 2u) : 0;
        }
        NJ_SYNCWARP();
        nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_RO, false, false);
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const bool act = I[NJP_I_ACT * RS + er] != 0;
            const int row = I[NJP_I_ROW * RS + er];
            if (act)
                for (int c_ = ec0; c_ < c.dout; c_ += LPR) {
                    float y = w.OUT[er * sO + c_];
                    if (c.residual) y += nj_resid(HS + er * sH, H, c.dout, c_);
                    YY[er * sD + c_] = y;
                    if (a.y_after) a.y_after[(size_t)row * c.dout + c_] = y;
                    // last_X = Y[i_obs] (masked, differentiable) or X_obs  (NJODE/models.py:481-486)
                    const float lx = c.masked ? y : NJ_LDG(a.b.X + (size_t)row * c.d + c_);
                    LX[er * sD + c_] = lx; TX[er * sD + c_] = nj_tanh(lx);
                }
        }
        NJ_SYNCWARP();
        // (d) loss row, tau, cursor
        NJ_LANES(lane) {
            if (lane < R && I[NJP_I_ACT * RS + lane]) {
                const int r = lane, row = I[NJP_I_ROW * RS + r];
                if (a.get_loss) {
                    float sa = 0.f, sb = 0.f;
                    NJ_UNROLL4
                    for (int c_ = 0; c_ < c.dout; ++c_) {      // (unrolled: the independent loads of four features go out together)
                        const float x = NJ_LDG(a.b.X + (size_t)row * c.d + c_);
                        const float m = c.masked ? NJ_LDG(a.b.M + (size_t)row * c.d + c_) : 1.f;
                        const float y = YY[r * sD + c_], yb = YBJ[r * sD + c_];
                        const float da = x - y, db = (c.loss_kind == NJODE_LOSS_STANDARD) ? (yb - y) : (yb - x);
                        sa = fmaf(m * da, da, sa); sb = fmaf(m * db, db, sb);
                    }
                    const float ra = sqrtf(sa + 1e-10f), rb = sqrtf(sb + 1e-10f);
                    const float sm = (c.loss_kind == NJODE_LOSS_STANDARD) ? (2.f * c.w * ra + 2.f * (1.f - c.w) * rb)
                                                                          : (c.w * ra + (1.f - c.w) * rb);
                    a.row_loss[row] = sm * sm / NJ_LDG(a.b.n_obs_ot + I[NJP_I_PATH * RS + r]);
                }
                F[NJP_F_TAU * RS + r] = NJ_LDG(a.b.jump_tau + NJ_LDG(a.b.row_jump + row));
                I[NJP_I_CUR * RS + r] += 1;
                set_next(r);
            }
        }
        NJ_SYNCWARP();
    }

    // unit descriptors, start encoder, first record
    NJ_HD void begin(int u0, int u1) {
        const int sI = s.sI, sO = s.sO, sH = s.sH, sD = s.sD;
        const bool rec = a.b.E > 0;
        NJ_LANES(lane) {
            if (lane < R) {
                const int u = u0 + lane;
                if (u < u1) {
                    const int32_t* dsc = a.b.unit_desc + (size_t)u * 6;
                    I[NJP_I_PATH * RS + lane] = dsc[0]; I[NJP_I_CUR * RS + lane] = dsc[3]; I[NJP_I_END * RS + lane] = dsc[4];
                } else { I[NJP_I_PATH * RS + lane] = -1; I[NJP_I_CUR * RS + lane] = 0; I[NJP_I_END * RS + lane] = 0; }
                I[NJP_I_ACT * RS + lane] = 0;
                set_next(lane);
            }
        }
        NJ_SYNCWARP();
        // ---- start: h = encoder(start_X, mask = 0)  (NJODE/models.py:411-419) ----
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int p = I[NJP_I_PATH * RS + er];
            for (int c_ = ec0; c_ < ein4; c_ += LPR) {
                float v = 0.f;
                if (c_ < c.d) {
                    const float x = p >= 0 ? NJ_LDG(a.b.start_X + (size_t)p * c.d + c_) : 0.f;
                    XI[er * sD + c_] = x; LX[er * sD + c_] = x;
                    v = nj_tanh(x);
                    TX[er * sD + c_] = v;
                }
                w.IN[(size_t)er * sI + c_] = v;
            }
            if (ec0 == 0) {
                F[NJP_F_TAU * RS + er] = 0.f;
                w.RK[er] = p >= 0 ? (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(p + a.b.path_id_offset), NJ_EVENT_INIT) : 0;
            }
        }
        NJ_SYNCWARP();
        nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_ENC, false, false);
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            for (int c_ = ec0; c_ < c.H; c_ += LPR) {
                float e = w.OUT[er * sO + c_];
                if (c.residual) e += nj_resid(XI + er * sD, c.d, c.H, c_);
                HS[er * sH + c_] = e;
            }
        }
        NJ_SYNCWARP();
        if (rec) record(0, NJ_EVENT_INIT);
    }
    // step count at which the next jump of the tile happens (NJP_NEVER: none left); gi: global jump cursor (return_path)
    NJ_HD int next_jump(int gi) const {
        int nk = NJP_NEVER;
        if (a.b.E > 0) { if (gi < a.b.K) nk = NJ_LDG(a.b.jump_step + gi); }
        else for (int r = 0; r < R; ++r) nk = I[NJP_I_NEXTK * RS + r] < nk ? I[NJP_I_NEXTK * RS + r] : nk;
        return nk;
    }
    NJ_HD bool any_jumps_at(int nk) const {
        bool any = false;
        for (int r = 0; r < R; ++r) any |= (I[NJP_I_NEXTK * RS + r] == nk);
        return any;
    }
    NJ_HD void finish() {
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int p = I[NJP_I_PATH * RS + er];
            if (p >= 0)
                for (int c_ = ec0; c_ < c.H; c_ += LPR) a.hT[(size_t)p * c.H + c_] = HS[er * s.sH + c_];
        }
        NJ_SYNCWARP();
    }

    NJ_HD void run(int u0, int u1) {
        const bool rec = a.b.E > 0;
        begin(u0, u1);
        const int S = a.b.S;
        int k = 0, gi = 0;
        for (;;) {
            const int nk = next_jump(gi);
            const int kend = nk < S ? nk : S;
            for (; k < kend; ++k) {
                euler_step(k);
                if (rec) record(NJ_LDG(a.b.step_event + k), NJ_EVENT_PATH_RO_BASE + (unsigned)k);
            }
            if (nk > S) break;
            if (any_jumps_at(nk)) jump(nk);
            if (rec) { record(NJ_LDG(a.b.jump_event + gi), NJ_EVENT_JUMP_BASE + 3u * (unsigned)gi + 2u); ++gi; }
        }
        finish();
    }
};

template <int RG, int TR>
NJ_HD void nj_path_cta_forward(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem) {
    float* simg = smem;
    nj_stage_image(simg, a.image, c.img_floats, s.nw_f * 32);
    nj_zero(smem + s.f_warp0, s.nw_f * s.f_region, s.nw_f * 32);
    NJ_SYNC();
    constexpr int R = RG * TR;
    NJ_WARPS(wp, s.nw_f) {
        float* reg = smem + s.f_warp0 + (size_t)wp * s.f_region;
        int* slot = reinterpret_cast<int*>(reg + s.f_I) + NJP_I_COUNT * NJP_RS;
        for (;;) {
            NJ_LANES(lane) { if (lane == 0) *slot = nj_atomic_inc(a.counter); }
            NJ_SYNCWARP();
            const int wt = *slot;
            NJ_SYNCWARP();
            if (wt >= s.n_tiles_f) break;
            const int ub = wt * R, ue = ub + R < a.b.n_units ? ub + R : a.b.n_units;
            NjPathFwd<RG, TR> f(c, s, a, reg, simg);
            f.run(ub, ue);
        }
    }
}

// ================================================================================================
// backward: one CTA = P rows (nw_b row warps of R rows + dW helper warps), lockstep over the batch-global steps
// ================================================================================================
enum { NJB_I_PATH = 0, NJB_I_CUR, NJB_I_C0, NJB_I_PREVK, NJB_I_ROW, NJB_I_ACT, NJB_I_RK, NJB_I_COUNT };

struct NjPathB {                    // CTA-level views
    float *IN, *A, *G, *GOUT, *GZ, *OUT, *GH, *HB, *EE, *GE, *XI, *LX, *TX, *YBJ, *YY, *GYBJ, *GX, *MM, *GI, *GHH, *F;
    int* I;
    int* MSK;                       // [nw_b] bit i: row i of the warp takes part in the current jump
};

NJ_HD void nj_pathb_bind(NjPathB& t, const NjPath& s, float* smem) {
    t.IN = smem + s.b_IN; t.A = smem + s.b_A; t.G = smem + s.b_G; t.GOUT = smem + s.b_GOUT; t.GZ = smem + s.b_GZ;
    t.OUT = smem + s.b_OUT; t.GH = smem + s.b_GH; t.HB = smem + s.b_HB; t.EE = smem + s.b_EE; t.GE = smem + s.b_GE;
    t.XI = smem + s.b_XI; t.LX = smem + s.b_LX; t.TX = smem + s.b_TX; t.YBJ = smem + s.b_YBJ; t.YY = smem + s.b_YY;
    t.GYBJ = smem + s.b_GYBJ; t.GX = smem + s.b_GX; t.MM = smem + s.b_MM; t.GI = smem + s.b_GI; t.GHH = smem + s.b_GHH;
    t.F = smem + s.b_F; t.I = reinterpret_cast<int*>(smem + s.b_I);
    t.MSK = t.I + NJB_I_COUNT * s.P_b + 4;
}

// (net, layer, og, kg) of dW tile T; false when T belongs to another network
NJ_HD bool nj_path_tile_decode(const NjCfg& c, const NjPath& s, int netid, int T, int& l, int& og, int& kg) {
    const NjNet& N = c.net[netid];
    if (N.n == 0 || T < s.tile_base[netid][0]) return false;
    l = 0;
    for (int ll = N.n - 1; ll > 0; --ll) if (T >= s.tile_base[netid][ll]) { l = ll; break; }
    const int K4 = (N.dim[l] + 3) >> 2, O4 = (N.dim[l + 1] + 3) >> 2;
    const int tl = T - s.tile_base[netid][l];
    if (tl >= K4 * O4) return false;
    kg = tl % K4; og = tl / K4;
    return true;
}

// q[0..15] += sum_r g[r][4og..] (x) a[r][4kg..], q[16..19] += sum_r g[r][4og..]; rows = all Pt rows (msk == nullptr) or
// the rows flagged in the per-warp masks, in ascending order (deterministic)
NJ_HD void nj_path_dw_rows(const NjCfg& c, const NjPath& s, const NjPathB& t, int netid, int l, int og, int kg, int Pt,
                           const int* msk, int R, float* q) {
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
#define NJP_DW_ROW(GP, AP)                                                                                          \
    {                                                                                                               \
        const nj_f4 gv = nj_sp_ld4(GP);                                                                             \
        const nj_f4 x = nj_sp_ld4(AP);                                                                              \
        r00 = fmaf(gv.x, x.x, r00); r01 = fmaf(gv.x, x.y, r01); r02 = fmaf(gv.x, x.z, r02); r03 = fmaf(gv.x, x.w, r03); \
        r10 = fmaf(gv.y, x.x, r10); r11 = fmaf(gv.y, x.y, r11); r12 = fmaf(gv.y, x.z, r12); r13 = fmaf(gv.y, x.w, r13); \
        r20 = fmaf(gv.z, x.x, r20); r21 = fmaf(gv.z, x.y, r21); r22 = fmaf(gv.z, x.z, r22); r23 = fmaf(gv.z, x.w, r23); \
        r30 = fmaf(gv.w, x.x, r30); r31 = fmaf(gv.w, x.y, r31); r32 = fmaf(gv.w, x.z, r32); r33 = fmaf(gv.w, x.w, r33); \
        b0 += gv.x; b1 += gv.y; b2 += gv.z; b3 += gv.w;                                                             \
    }
    if (!msk) {
#pragma unroll 4
        for (int r = 0; r < Pt; ++r) {
            NJP_DW_ROW(gq, aq);
            gq = NJ_SP_ADD(gq, g_s); aq = NJ_SP_ADD(aq, a_s);
        }
    } else {
        for (int wq = 0, r0 = 0; r0 < Pt; ++wq, r0 += R) {
            int m = msk[wq];
            for (int i = 0; m; ++i, m >>= 1)
                if (m & 1) NJP_DW_ROW(NJ_SP_ADD(gq, (r0 + i) * g_s), NJ_SP_ADD(aq, (r0 + i) * a_s));
        }
    }
#undef NJP_DW_ROW
    q[0] = r00; q[1] = r01; q[2] = r02; q[3] = r03; q[4] = r10; q[5] = r11; q[6] = r12; q[7] = r13;
    q[8] = r20; q[9] = r21; q[10] = r22; q[11] = r23; q[12] = r30; q[13] = r31; q[14] = r32; q[15] = r33;
    q[16] = b0; q[17] = b1; q[18] = b2; q[19] = b3;
}

// tiles beyond the register slots: accumulated from zero and added into this CTA's partial image (L2 resident) right
// away.  Kept out of line: inlined into the 40-accumulator dW phase, ptxas (12.9, -O3) produced a wrong partial-image
// address for these tiles on sm_100a (compute-sanitizer: out-of-bounds LDG in the read-modify-write; a printf next to it
// made the problem disappear), while the same source is correct in the host simulation.
NJ_HDN void nj_path_dw_overflow(const NjCfg* cp, const NjPath* sp, const NjPathB* tp, int netid, float* gpart, int tid, int nt,
                                int Pt, const int* msk, int R) {
    const NjCfg& c = *cp; const NjPath& s = *sp; const NjPathB& t = *tp;
    for (int T = NJ_SEG_NT_MAX * nt + tid; T < s.tiles_total; T += nt) {
        int l, og, kg;
        if (!nj_path_tile_decode(c, s, netid, T, l, og, kg)) continue;
        float q[20];
#pragma unroll
        for (int i = 0; i < 20; ++i) q[i] = 0.f;
        nj_path_dw_rows(c, s, t, netid, l, og, kg, Pt, msk, R, q);
        nj_seg_tile_store(c, netid, l, og, kg, q, gpart, true);
    }
}

// dW phase of one network: thread-owned tiles, the first NJ_SEG_NT_MAX * nt of them in registers for the whole launch
NJ_HD void nj_path_dw(const NjCfg& c, const NjPath& s, const NjPathB& t, int netid, float* acc, float* gpart,
                      int tid, int nt, int Pt, const int* msk, int R) {
#pragma unroll
    for (int slot = 0; slot < NJ_SEG_NT_MAX; ++slot) {
        if (slot >= s.nt_slots) break;
        int l, og, kg;
        if (!nj_path_tile_decode(c, s, netid, slot * nt + tid, l, og, kg)) continue;
        nj_path_dw_rows(c, s, t, netid, l, og, kg, Pt, msk, R, acc + slot * 20);
    }
    if (s.tiles_total > NJ_SEG_NT_MAX * nt) nj_path_dw_overflow(&c, &s, &t, netid, gpart, tid, nt, Pt, msk, R);
}

NJ_HD void nj_path_dw_flush(const NjCfg& c, const NjPath& s, const float* acc, float* gpart, int tid, int nt) {
#pragma unroll
    for (int slot = 0; slot < NJ_SEG_NT_MAX; ++slot) {
        if (slot >= s.nt_slots) break;
        const int T = slot * nt + tid;
        if (T >= s.tiles_total) continue;
        for (int netid = 0; netid < NJODE_NUM_NETS; ++netid) {
            int l, og, kg;
            if (nj_path_tile_decode(c, s, netid, T, l, og, kg)) { nj_seg_tile_store(c, netid, l, og, kg, acc + slot * 20, gpart, false); break; }
        }
    }
}

// CTA-wide: the latest step count at which a row of the tile still has a jump to reverse (-1: none)
NJ_HD int nj_pathb_next(const NjPathB& t, int P, int Pt) {
    int m = -1;
    for (int r = 0; r < Pt; ++r) m = t.I[NJB_I_PREVK * P + r] > m ? t.I[NJB_I_PREVK * P + r] : m;
    return m;
}

template <int RG, int TR, bool COOP = false>
struct NjPathBwd {
    static constexpr int R = RG * TR;
    const NjCfg& c; const NjPath& s; const NjArgs& a; const NjPathB& t;
    const float* simg;
    int P, d4, H4, inf4, ein4, do4, wa;
    void* coop = nullptr;

    NJ_HD NjPathBwd(const NjCfg& c_, const NjPath& s_, const NjArgs& a_, const NjPathB& t_, const float* simg_)
        : c(c_), s(s_), a(a_), t(t_), simg(simg_) {
        P = s.P_b;
        d4 = ((c.d + 3) >> 2) << 2; H4 = ((c.H + 3) >> 2) << 2; inf4 = ((c.inf + 3) >> 2) << 2;
        ein4 = ((c.enc_in + 3) >> 2) << 2; do4 = ((c.dout + 3) >> 2) << 2;
        wa = P * s.sA;
    }
    NJ_HD NjPW view(int r0) const {
        NjPW w;
        w.c = &c; w.wimg = simg;
        w.IN = t.IN + (size_t)r0 * s.sI; w.A0 = t.A + (size_t)r0 * s.sA; w.A1 = nullptr; w.OUT = t.OUT + (size_t)r0 * s.sO;
        w.G0 = t.G + (size_t)r0 * s.sA; w.GOUT = t.GOUT + (size_t)r0 * s.sO; w.GZ = t.GZ + (size_t)r0 * s.sI;
        w.sI = s.sI; w.sA = s.sA; w.sO = s.sO;
        w.a_buf_stride = wa; w.g_buf_stride = wa; w.RK = t.I + NJB_I_RK * P + r0; w.coop = coop;
        return w;
    }
    NJ_HD void set_key(int r, bool valid, unsigned ev) const {
        t.I[NJB_I_RK * P + r] = valid ? (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(t.I[NJB_I_PATH * P + r] + a.b.path_id_offset), ev) : 0;
    }
    // cursor -> (row of the latest jump not yet reversed, its step count)
    NJ_HD void set_prev(int r) const {
        const int cur = t.I[NJB_I_CUR * P + r];
        if (cur > t.I[NJB_I_C0 * P + r]) {
            const int row = NJ_LDG(a.b.path_rows + cur - 1);
            t.I[NJB_I_ROW * P + r] = row;
            t.I[NJB_I_PREVK * P + r] = NJ_LDG(a.b.jump_step + NJ_LDG(a.b.row_jump + row));
        } else { t.I[NJB_I_ROW * P + r] = -1; t.I[NJB_I_PREVK * P + r] = -1; }
    }
    // (last_X, tanh(last_X), tau) valid between the jump of row NJB_I_ROW (or the start) and the next jump
    NJ_HD void load_state(int r0, int lane) const {
        NJ_ROWMAP(R);
        const int r = r0 + er, p = t.I[NJB_I_PATH * P + r], prev = t.I[NJB_I_ROW * P + r];
        for (int c_ = ec0; c_ < d4; c_ += LPR) {
            float x = 0.f;
            if (p >= 0 && c_ < c.d) {
                if (prev < 0) x = NJ_LDG(a.b.start_X + (size_t)p * c.d + c_);
                else if (c.masked) x = a.y_after[(size_t)prev * c.dout + c_];
                else x = NJ_LDG(a.b.X + (size_t)prev * c.d + c_);
            }
            t.LX[r * s.sD + c_] = x; t.TX[r * s.sD + c_] = c_ < c.d ? nj_tanh(x) : 0.f;
        }
        if (ec0 == 0) t.F[NJP_F_TAU * P + r] = (p >= 0 && prev >= 0) ? NJ_LDG(a.b.jump_tau + NJ_LDG(a.b.row_jump + prev)) : 0.f;
    }

    // ---- reverse of Euler step k, warp-local part (rows r0 .. r0 + R) ----
    NJ_HD void step_local(int r0, int k) const {
        const int sI = s.sI, sO = s.sO, sH = s.sH, sD = s.sD;
        const NjPW w = view(r0);
        const float tcur = NJ_LDG(a.b.step_t + k), dt = NJ_LDG(a.b.step_dt + k);
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int r = r0 + er, p = t.I[NJB_I_PATH * P + r];
            const float tau = t.F[NJP_F_TAU * P + r];
            const float* hh = p >= 0 ? a.h_hist + ((size_t)k * a.b.B + p) * c.H : nullptr;
            for (int c_ = ec0; c_ < inf4; c_ += LPR) {
                float v = 0.f;
                if (c_ < c.d) v = t.TX[r * sD + c_];
                else if (c_ < c.d + c.H) v = nj_tanh(hh ? hh[c_ - c.d] : 0.f);
                else if (c_ < c.inf) {
                    if (c_ == c.d + c.H) v = tau;
                    else if (c_ == c.d + c.H + 1) v = tcur - tau;
                    else v = tau + (tcur - tau);
                }
                t.IN[(size_t)r * sI + c_] = v;
            }
            for (int c_ = ec0; c_ < H4; c_ += LPR) t.GOUT[(size_t)r * sO + c_] = c_ < c.H ? dt * t.GH[r * sH + c_] : 0.f;
            if (ec0 == 0) set_key(r, true, (unsigned)k);
        }
        NJ_SYNCWARP();
        nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_ODE, true, true);
        nj_path_mlp_dx<RG, TR, COOP>(w, NJODE_NET_ODE, true);
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int r = r0 + er;
            if (t.I[NJB_I_PATH * P + r] >= 0) {
                for (int c_ = ec0; c_ < c.H; c_ += LPR) {
                    const float th = t.IN[(size_t)r * sI + c.d + c_];
                    t.GH[r * sH + c_] += t.GZ[(size_t)r * sI + c.d + c_] * (1.f - th * th);
                }
                if (c.masked)            // last_X = Y[i_obs] is differentiable (NJODE/models.py:483-484)
                    for (int c_ = ec0; c_ < c.d; c_ += LPR) {
                        const float tx = t.IN[(size_t)r * sI + c_];
                        t.GX[r * sD + c_] += t.GZ[(size_t)r * sI + c_] * (1.f - tx * tx);
                    }
            }
        }
        NJ_SYNCWARP();
    }

    // ---- jump reversal, phase 1 (warp-local): recompute Y_bj, E, Y; loss gradients; readout backward at E ----
    NJ_HD void jump_p1(int r0, int wp, int k) const {
        const int sI = s.sI, sO = s.sO, sH = s.sH, sD = s.sD, s3 = s.s3, H = c.H;
        const NjPW w = view(r0);
        int msk = 0;
        for (int i = 0; i < R; ++i) msk |= (t.I[NJB_I_PREVK * P + r0 + i] == k) ? (1 << i) : 0;
        NJ_LANES(lane) { if (lane == 0) t.MSK[wp] = msk; }
        if (!msk) { NJ_SYNCWARP(); return; }
        const float gl = NJ_LDG(a.grad_loss);
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int r = r0 + er, row = t.I[NJB_I_ROW * P + r];
            const bool act = (msk >> er) & 1;
            for (int c_ = ec0; c_ < H4; c_ += LPR) {
                const float h = (act && c_ < H) ? a.h_before[(size_t)row * H + c_] : 0.f;
                if (c_ < H) t.HB[r * sH + c_] = h;
                t.IN[(size_t)r * sI + c_] = c_ < H ? nj_tanh(h) : 0.f;
            }
            if (ec0 == 0) {
                t.I[NJB_I_ACT * P + r] = act ? 1 : 0;
                set_key(r, act, NJ_EVENT_JUMP_BASE + 3u * (unsigned)(act ? NJ_LDG(a.b.row_jump + row) : 0) + 0u);
            }
        }
        NJ_SYNCWARP();
        nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_RO, true, false);
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int r = r0 + er;
            for (int c_ = ec0; c_ < c.dout; c_ += LPR) {
                float y = t.OUT[r * sO + c_];
                if (c.residual) y += nj_resid(t.HB + r * sH, H, c.dout, c_);
                t.YBJ[r * sD + c_] = y;
            }
        }
        NJ_SYNCWARP();
        if (c.use_rnn) {
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const int r = r0 + er, row = t.I[NJB_I_ROW * P + r];
                const bool act = (msk >> er) & 1;
                for (int c_ = ec0; c_ < d4; c_ += LPR) {
                    const float x = (act && c_ < c.d) ? NJ_LDG(a.b.X + (size_t)row * c.d + c_) : 0.f;
                    t.XI[r * sD + c_] = x;
                    t.IN[(size_t)r * sI + c_] = c_ < c.d ? nj_tanh(x) : 0.f;
                }
            }
            NJ_SYNCWARP();
            nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_GRU_IH, true, false);
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const int r = r0 + er;
                for (int c_ = ec0; c_ < 3 * H; c_ += LPR) t.GI[r * s3 + c_] = t.OUT[r * sO + c_];
                for (int c_ = ec0; c_ < H4; c_ += LPR) t.IN[(size_t)r * sI + c_] = c_ < H ? nj_tanh(t.HB[r * sH + c_]) : 0.f;
            }
            NJ_SYNCWARP();
            nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_GRU_HH, true, false);
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const int r = r0 + er, row = t.I[NJB_I_ROW * P + r];
                const bool act = (msk >> er) & 1;
                for (int c_ = ec0; c_ < H4; c_ += LPR) {
                    float e = 0.f;
                    if (c_ < H) {
                        float* gi = t.GI + r * s3;
                        float* gh = t.GHH + r * s3;
                        const float* o = t.OUT + r * sO;
                        const float hh = t.IN[(size_t)r * sI + c_];
                        const float rr_ = nj_sigmoid(gi[c_] + o[c_]);
                        const float z = nj_sigmoid(gi[H + c_] + o[H + c_]);
                        const float ghn = o[2 * H + c_];
                        const float n = nj_tanh(fmaf(rr_, ghn, gi[2 * H + c_]));
                        gi[c_] = rr_; gi[H + c_] = z; gi[2 * H + c_] = n;
                        gh[c_] = hh; gh[2 * H + c_] = ghn;
                        e = fmaf(z, hh - n, n);
                        t.EE[r * sH + c_] = e;
                    }
                    t.IN[(size_t)r * sI + c_] = c_ < H ? nj_tanh(e) : 0.f;
                }
                if (ec0 == 0) set_key(r, act, NJ_EVENT_JUMP_BASE + 3u * (unsigned)(act ? NJ_LDG(a.b.row_jump + row) : 0) + 2u);
            }
            NJ_SYNCWARP();
        } else {
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const int r = r0 + er, row = t.I[NJB_I_ROW * P + r];
                const bool act = (msk >> er) & 1;
                for (int c_ = ec0; c_ < ein4; c_ += LPR) {
                    float v = 0.f;
                    if (c_ < c.d) {
                        float x = act ? NJ_LDG(a.b.X + (size_t)row * c.d + c_) : 0.f;
                        if (c.masked) {
                            const float m = act ? NJ_LDG(a.b.M + (size_t)row * c.d + c_) : 0.f;
                            x = x * m + (1.f - m) * t.YBJ[r * sD + c_];
                            t.MM[r * sD + c_] = m;
                        }
                        t.XI[r * sD + c_] = x;
                        v = nj_tanh(x);
                    } else if (c.masked && c_ < 2 * c.d) v = act ? NJ_LDG(a.b.M + (size_t)row * c.d + c_ - c.d) : 0.f;
                    t.IN[(size_t)r * sI + c_] = v;
                }
                if (ec0 == 0) set_key(r, act, NJ_EVENT_JUMP_BASE + 3u * (unsigned)(act ? NJ_LDG(a.b.row_jump + row) : 0) + 1u);
            }
            NJ_SYNCWARP();
            nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_ENC, true, false);
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const int r = r0 + er, row = t.I[NJB_I_ROW * P + r];
                const bool act = (msk >> er) & 1;
                for (int c_ = ec0; c_ < H4; c_ += LPR) {
                    float e = 0.f;
                    if (c_ < H) {
                        e = t.OUT[r * sO + c_];
                        if (c.residual) e += nj_resid(t.XI + r * sD, c.d, H, c_);
                        t.EE[r * sH + c_] = e;
                    }
                    t.IN[(size_t)r * sI + c_] = c_ < H ? nj_tanh(e) : 0.f;
                }
                if (ec0 == 0) set_key(r, act, NJ_EVENT_JUMP_BASE + 3u * (unsigned)(act ? NJ_LDG(a.b.row_jump + row) : 0) + 2u);
            }
            NJ_SYNCWARP();
        }
        nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_RO, true, false);
        // loss derivative (compute_loss / compute_loss_2, NJODE/models.py:71-126)
        NJ_LANES(lane) {
            if (lane < R) {
                const int r = r0 + lane, row = t.I[NJB_I_ROW * P + r];
                float ca = 0.f, cb = 0.f;
                if ((msk >> lane) & 1) {
                    float sa = 0.f, sb = 0.f;
                    NJ_UNROLL4
                    for (int c_ = 0; c_ < c.dout; ++c_) {      // (unrolled: the independent loads of four features go out together)
                        float y = t.OUT[r * sO + c_];
                        if (c.residual) y += nj_resid(t.EE + r * sH, H, c.dout, c_);
                        t.YY[r * sD + c_] = y;
                        const float x = NJ_LDG(a.b.X + (size_t)row * c.d + c_), yb = t.YBJ[r * sD + c_];
                        const float m = c.masked ? NJ_LDG(a.b.M + (size_t)row * c.d + c_) : 1.f;
                        const float da = x - y, db = (c.loss_kind == NJODE_LOSS_STANDARD) ? (yb - y) : (yb - x);
                        sa = fmaf(m * da, da, sa); sb = fmaf(m * db, db, sb);
                    }
                    const float ra = sqrtf(sa + 1e-10f), rb = sqrtf(sb + 1e-10f);
                    const float wa_ = (c.loss_kind == NJODE_LOSS_STANDARD) ? 2.f * c.w : c.w;
                    const float wb_ = (c.loss_kind == NJODE_LOSS_STANDARD) ? 2.f * (1.f - c.w) : (1.f - c.w);
                    const float sm = wa_ * ra + wb_ * rb;
                    const float nobs = a.b.n_obs_ot ? NJ_LDG(a.b.n_obs_ot + t.I[NJB_I_PATH * P + r]) : 1.f;
                    const float cf = a.b.n_obs_ot ? gl * 2.f * sm / (nobs * (float)a.b.batch_size_norm) : 0.f;
                    ca = cf * wa_ / ra; cb = cf * wb_ / rb;
                }
                t.F[NJP_F_CA * P + r] = ca; t.F[NJP_F_CB * P + r] = cb;
            }
        }
        NJ_SYNCWARP();
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int r = r0 + er, row = t.I[NJB_I_ROW * P + r];
            const bool act = (msk >> er) & 1;
            for (int c_ = ec0; c_ < do4; c_ += LPR) {
                float gy = 0.f, gyb = 0.f;
                if (c_ < c.dout && act) {
                    const float x = NJ_LDG(a.b.X + (size_t)row * c.d + c_), y = t.YY[r * sD + c_], yb = t.YBJ[r * sD + c_];
                    const float m = c.masked ? NJ_LDG(a.b.M + (size_t)row * c.d + c_) : 1.f;
                    const float ca = t.F[NJP_F_CA * P + r], cb = t.F[NJP_F_CB * P + r];
                    if (c.loss_kind == NJODE_LOSS_STANDARD) { gy = -ca * m * (x - y) - cb * m * (yb - y); gyb = cb * m * (yb - y); }
                    else { gy = -ca * m * (x - y); gyb = cb * m * (yb - x); }
                    if (c.masked) gy += t.GX[r * sD + c_];           // last_X = Y[i_obs]
                }
                t.GOUT[(size_t)r * sO + c_] = gy;
                t.GYBJ[r * sD + c_] = gyb;
            }
        }
        NJ_SYNCWARP();
        nj_path_mlp_dx<RG, TR, COOP>(w, NJODE_NET_RO, true);
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int r = r0 + er;
            const bool act = (msk >> er) & 1;
            for (int c_ = ec0; c_ < H; c_ += LPR) {
                const float th = t.IN[(size_t)r * sI + c_];
                float ge = t.GZ[(size_t)r * sI + c_] * (1.f - th * th);
                if (c.residual) ge += nj_resid_bwd(t.GOUT + (size_t)r * sO, H, c.dout, c_);
                t.GE[r * sH + c_] = act ? ge + t.GH[r * sH + c_] : 0.f;
            }
        }
        NJ_SYNCWARP();
    }

    // ---- phase 2 (warp-local): encoder (or GRU hidden map) at the observation, backward with g = dL/dE ----
    NJ_HD void jump_p2(int r0, int wp) const {
        const int sI = s.sI, sO = s.sO, sH = s.sH, sD = s.sD, s3 = s.s3, H = c.H;
        const int msk = t.MSK[wp];
        if (!msk) return;
        const NjPW w = view(r0);
        if (c.use_rnn) {
            // backward of h' = (1 - z) n + z tanh(h_old), gates from (gi, gh)
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const int r = r0 + er;
                const bool act = (msk >> er) & 1;
                for (int c_ = ec0; c_ < H4; c_ += LPR) {
                    if (c_ < H) {
                        float* gi = t.GI + r * s3;
                        const float* gh = t.GHH + r * s3;
                        float* go = t.GOUT + (size_t)r * sO;
                        const float ge = act ? t.GE[r * sH + c_] : 0.f;
                        const float rr_ = gi[c_], z = gi[H + c_], n = gi[2 * H + c_], hh = gh[c_], ghn = gh[2 * H + c_];
                        const float dpn = ge * (1.f - z) * (1.f - n * n);
                        const float dpr = dpn * ghn * rr_ * (1.f - rr_);
                        const float dpz = ge * (hh - n) * z * (1.f - z);
                        go[c_] = dpr; go[H + c_] = dpz; go[2 * H + c_] = dpn * rr_;
                        gi[c_] = dpr; gi[H + c_] = dpz; gi[2 * H + c_] = dpn;
                        t.EE[r * sH + c_] = ge * z;
                        t.IN[(size_t)r * sI + c_] = hh;
                    } else t.IN[(size_t)r * sI + c_] = 0.f;
                }
            }
            NJ_SYNCWARP();
            nj_path_mlp_dx<RG, TR, COOP>(w, NJODE_NET_GRU_HH, true);
        } else {
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const int r = r0 + er, row = t.I[NJB_I_ROW * P + r];
                const bool act = (msk >> er) & 1;
                for (int c_ = ec0; c_ < ein4; c_ += LPR) {
                    float v = 0.f;
                    if (c_ < c.d) v = nj_tanh(t.XI[r * sD + c_]);
                    else if (c.masked && c_ < 2 * c.d) v = t.MM[r * sD + c_ - c.d];
                    t.IN[(size_t)r * sI + c_] = v;
                }
                for (int c_ = ec0; c_ < H4; c_ += LPR) t.GOUT[(size_t)r * sO + c_] = (c_ < H && act) ? t.GE[r * sH + c_] : 0.f;
                if (ec0 == 0) set_key(r, act, NJ_EVENT_JUMP_BASE + 3u * (unsigned)(act ? NJ_LDG(a.b.row_jump + row) : 0) + 1u);
            }
            NJ_SYNCWARP();
            nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_ENC, true, true);
            nj_path_mlp_dx<RG, TR, COOP>(w, NJODE_NET_ENC, c.masked != 0);
            if (c.masked) {
                // imputation X*M + (1 - M)*Y_bj: the gradient wrt the encoder input reaches Y_bj where M = 0
                NJ_LANES(lane) {
                    NJ_ROWMAP(R);
                    const int r = r0 + er;
                    if ((msk >> er) & 1)
                        for (int c_ = ec0; c_ < c.d; c_ += LPR) {
                            const float tx = t.IN[(size_t)r * sI + c_];
                            float gx = t.GZ[(size_t)r * sI + c_] * (1.f - tx * tx);
                            if (c.residual) gx += nj_resid_bwd(t.GOUT + (size_t)r * sO, c.d, H, c_);
                            t.GYBJ[r * sD + c_] += (1.f - t.MM[r * sD + c_]) * gx;
                        }
                }
                NJ_SYNCWARP();
            }
        }
    }

    // ---- GRU only, phase 2b (warp-local): operands of the input map's dW ----
    NJ_HD void jump_p2b(int r0, int wp) const {
        const int sI = s.sI, sO = s.sO, sH = s.sH, sD = s.sD, s3 = s.s3, H = c.H;
        const int msk = t.MSK[wp];
        if (!msk) return;
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int r = r0 + er;
            for (int c_ = ec0; c_ < H; c_ += LPR) {
                const float hh = t.GHH[r * s3 + c_];
                t.EE[r * sH + c_] = (t.EE[r * sH + c_] + t.GZ[(size_t)r * sI + c_]) * (1.f - hh * hh);
            }
        }
        NJ_SYNCWARP();
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int r = r0 + er;
            for (int c_ = ec0; c_ < 3 * H; c_ += LPR) t.GOUT[(size_t)r * sO + c_] = t.GI[r * s3 + c_];
            for (int c_ = ec0; c_ < d4; c_ += LPR) t.IN[(size_t)r * sI + c_] = c_ < c.d ? nj_tanh(t.XI[r * sD + c_]) : 0.f;
        }
        NJ_SYNCWARP();
    }

    // ---- phase 3 (warp-local): readout at h_before, backward with g = dL/dY_bj -> adjoint before the jump; cursor ----
    NJ_HD void jump_p3(int r0, int wp) const {
        const int sI = s.sI, sO = s.sO, sH = s.sH, sD = s.sD, H = c.H;
        const int msk = t.MSK[wp];
        if (!msk) return;
        const NjPW w = view(r0);
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int r = r0 + er, row = t.I[NJB_I_ROW * P + r];
            const bool act = (msk >> er) & 1;
            const int zc = c.use_rnn ? 3 * H : H4;        // clears what the previous phase left in GOUT
            for (int c_ = ec0; c_ < H4; c_ += LPR) t.IN[(size_t)r * sI + c_] = c_ < H ? nj_tanh(t.HB[r * sH + c_]) : 0.f;
            for (int c_ = ec0; c_ < zc || c_ < do4; c_ += LPR) t.GOUT[(size_t)r * sO + c_] = (c_ < c.dout && act) ? t.GYBJ[r * sD + c_] : 0.f;
            if (ec0 == 0) set_key(r, act, NJ_EVENT_JUMP_BASE + 3u * (unsigned)(act ? NJ_LDG(a.b.row_jump + row) : 0) + 0u);
        }
        NJ_SYNCWARP();
        nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_RO, true, true);
        nj_path_mlp_dx<RG, TR, COOP>(w, NJODE_NET_RO, true);
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int r = r0 + er;
            if ((msk >> er) & 1) {
                for (int c_ = ec0; c_ < H; c_ += LPR) {
                    const float th = t.IN[(size_t)r * sI + c_];
                    float gh = t.GZ[(size_t)r * sI + c_] * (1.f - th * th);
                    if (c.residual) gh += nj_resid_bwd(t.GOUT + (size_t)r * sO, H, c.dout, c_);
                    if (c.use_rnn) gh += t.EE[r * sH + c_];
                    t.GH[r * sH + c_] = gh;
                }
                for (int c_ = ec0; c_ < c.d; c_ += LPR) t.GX[r * sD + c_] = 0.f;
            }
        }
        NJ_SYNCWARP();
        NJ_LANES(lane) {
            if (lane < R && ((msk >> lane) & 1)) {
                t.I[NJB_I_CUR * P + r0 + lane] -= 1;
                set_prev(r0 + lane);
            }
        }
        NJ_SYNCWARP();
        NJ_LANES(lane) { load_state(r0, lane); }
        NJ_SYNCWARP();
    }

    // ---- start encoder reversed (warp-local) ----
    NJ_HD void start_local(int r0) const {
        const int sI = s.sI, sO = s.sO, sH = s.sH, sD = s.sD;
        const NjPW w = view(r0);
        NJ_LANES(lane) {
            NJ_ROWMAP(R);
            const int r = r0 + er;
            const bool valid = t.I[NJB_I_PATH * P + r] >= 0;
            for (int c_ = ec0; c_ < ein4; c_ += LPR) t.IN[(size_t)r * sI + c_] = c_ < c.d ? t.TX[r * sD + c_] : 0.f;
            for (int c_ = ec0; c_ < H4; c_ += LPR) t.GOUT[(size_t)r * sO + c_] = (c_ < c.H && valid) ? t.GH[r * sH + c_] : 0.f;
            if (ec0 == 0) set_key(r, valid, NJ_EVENT_INIT);
        }
        NJ_SYNCWARP();
        nj_path_mlp_fwd<RG, TR, COOP>(w, NJODE_NET_ENC, true, true);
        nj_path_mlp_dx<RG, TR, COOP>(w, NJODE_NET_ENC, false);
    }
};

// the barrier protocol of one tile, shared by the row warps and the helper warps:
//   per reversed jump step: p1 | dW RO | p2 | dW ENC or GRU_HH | [GRU: p2b | dW GRU_IH |] p3 | dW RO |
//   per reversed Euler step: local | dW ODE |;  start encoder: local | dW ENC |       ("|" = CTA barrier)
template <int RG, int TR, bool ROWS>
NJ_HD void nj_path_bwd_tile(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem, const NjPathB& t, float* nj_acc_base,
                            int cta, int u0, int u1) {
    constexpr int R = RG * TR;
    const int P = s.P_b, nt = s.nt_b, Pt = R * s.nw_b;
    float* gpart = a.partials + (size_t)cta * c.img_floats;
    const NjPathBwd<RG, TR> B(c, s, a, t, smem);
    if (ROWS) {
        NJ_THREADS(tid, nt) {
            if (tid < Pt) {
                const int u = u0 + tid;
                if (u < u1) {
                    const int32_t* dsc = a.b.unit_desc + (size_t)u * 6;
                    t.I[NJB_I_PATH * P + tid] = dsc[0]; t.I[NJB_I_C0 * P + tid] = dsc[3]; t.I[NJB_I_CUR * P + tid] = dsc[4];
                } else { t.I[NJB_I_PATH * P + tid] = -1; t.I[NJB_I_C0 * P + tid] = 0; t.I[NJB_I_CUR * P + tid] = 0; }
                t.I[NJB_I_ACT * P + tid] = 0;
                B.set_prev(tid);
            }
        }
    }
    NJ_SYNC();
    if (ROWS) {
        NJ_WARPS(wp, s.nw_b) {
            const int r0 = wp * R;
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const int r = r0 + er, p = t.I[NJB_I_PATH * P + r];
                const float* ght = (p >= 0 && a.grad_hT) ? a.grad_hT + (size_t)p * c.H : nullptr;
                for (int c_ = ec0; c_ < c.H; c_ += LPR) t.GH[r * s.sH + c_] = ght ? NJ_LDG(ght + c_) : 0.f;
                for (int c_ = ec0; c_ < c.d; c_ += LPR) t.GX[r * s.sD + c_] = 0.f;
                B.load_state(r0, lane);
            }
            NJ_SYNCWARP();
        }
    }
    int nk = nj_pathb_next(t, P, Pt);
    for (int k = a.b.S; ; --k) {
        if (nk == k) {
            if (ROWS) { NJ_WARPS(wp, s.nw_b) { B.jump_p1(wp * R, wp, k); } }
            NJ_SYNC();
            NJ_THREADS(tid, nt) { nj_path_dw(c, s, t, NJODE_NET_RO, NJ_ACC(tid), gpart, tid, nt, Pt, t.MSK, R); }
            NJ_SYNC();
            if (ROWS) { NJ_WARPS(wp, s.nw_b) { B.jump_p2(wp * R, wp); } }
            NJ_SYNC();
            NJ_THREADS(tid, nt) { nj_path_dw(c, s, t, c.use_rnn ? NJODE_NET_GRU_HH : NJODE_NET_ENC, NJ_ACC(tid), gpart, tid, nt, Pt, t.MSK, R); }
            NJ_SYNC();
            if (c.use_rnn) {
                if (ROWS) { NJ_WARPS(wp, s.nw_b) { B.jump_p2b(wp * R, wp); } }
                NJ_SYNC();
                NJ_THREADS(tid, nt) { nj_path_dw(c, s, t, NJODE_NET_GRU_IH, NJ_ACC(tid), gpart, tid, nt, Pt, t.MSK, R); }
                NJ_SYNC();
            }
            if (ROWS) { NJ_WARPS(wp, s.nw_b) { B.jump_p3(wp * R, wp); } }
            NJ_SYNC();
            NJ_THREADS(tid, nt) { nj_path_dw(c, s, t, NJODE_NET_RO, NJ_ACC(tid), gpart, tid, nt, Pt, t.MSK, R); }
            NJ_SYNC();
            nk = nj_pathb_next(t, P, Pt);
        }
        if (k == 0) break;
        if (ROWS) { NJ_WARPS(wp, s.nw_b) { B.step_local(wp * R, k - 1); } }
        NJ_SYNC();
        NJ_THREADS(tid, nt) { nj_path_dw(c, s, t, NJODE_NET_ODE, NJ_ACC(tid), gpart, tid, nt, Pt, nullptr, R); }
        NJ_SYNC();
    }
    if (ROWS) { NJ_WARPS(wp, s.nw_b) { B.start_local(wp * R); } }
    NJ_SYNC();
    NJ_THREADS(tid, nt) { nj_path_dw(c, s, t, NJODE_NET_ENC, NJ_ACC(tid), gpart, tid, nt, Pt, nullptr, R); }
    NJ_SYNC();
}

// ------------------------------------------------------------------------------------------------
// pipelined variant: the dW phase of Euler step k runs on the HELPER warps while the row warps already reverse step
// k - 1.  The operand buffers of a step (IN, A, G, GOUT) exist twice; the ODE network's gradient tiles live in the helper
// warps' registers only (up to NJP_HSLOTS 4x4 tiles per helper thread), so the row warps never wait for a dW phase: one CTA
// barrier per Euler step, at which "row warps finished step k into buffer b" meets "helpers finished the dW of step k + 1
// from buffer 1 - b".  A jump drains the pipeline and runs its (rare) three dW phases through the partial image.
// (ncu, PhysioNet-shaped batch of 2 000 before this: 9.5 barrier stalls per issued instruction, issue slots 19 % busy --
// seven row warps and five helpers took turns.)
// ------------------------------------------------------------------------------------------------
#define NJP_HSLOTS 5
#define NJP_HACC (NJP_HSLOTS * 20)
NJ_HDN void nj_stat_dw(const NjCfg* cp, const NjPath* sp, const NjPathB* tp, int netid, float* gpart, int tid, int nt, int Pt,
                       const int* msk, int R);          // every dW tile of one network through the partial image (below)

// helper thread `hid` of `nth`: dW of the ODE network for the Pt rows of the step held in the operand buffers of `t`
NJ_HD void nj_path_dw_helper(const NjCfg& c, const NjPath& s, const NjPathB& t, float* acc, int hid, int nth, int Pt) {
    const int ode_tiles = s.tile_base[NJODE_NET_RO][0];        // the ODE network's tiles come first
#pragma unroll
    for (int slot = 0; slot < NJP_HSLOTS; ++slot) {
        const int T = slot * nth + hid;
        if (T >= ode_tiles) break;
        int l, og, kg;
        if (!nj_path_tile_decode(c, s, NJODE_NET_ODE, T, l, og, kg)) continue;
        nj_path_dw_rows(c, s, t, NJODE_NET_ODE, l, og, kg, Pt, nullptr, 1, acc + slot * 20);
    }
}
NJ_HD void nj_path_helper_flush(const NjCfg& c, const NjPath& s, const float* acc, float* gpart, int hid, int nth) {
    const int ode_tiles = s.tile_base[NJODE_NET_RO][0];
#pragma unroll
    for (int slot = 0; slot < NJP_HSLOTS; ++slot) {
        const int T = slot * nth + hid;
        if (T >= ode_tiles) break;
        int l, og, kg;
        if (nj_path_tile_decode(c, s, NJODE_NET_ODE, T, l, og, kg)) nj_seg_tile_store(c, NJODE_NET_ODE, l, og, kg, acc + slot * 20, gpart, false);
    }
}

#if defined(NJODE_HOST_SIM)
#define NJP_HACC_DECL(nt) std::vector<float> njp_hacc_store((size_t)(nt) * NJP_HACC, 0.f); float* njp_hacc = njp_hacc_store.data()
#define NJP_HACC_OF(tid) (njp_hacc + (size_t)(tid) * NJP_HACC)
#else
#define NJP_HACC_DECL(nt) float njp_hacc_store[NJP_HACC]; _Pragma("unroll") for (int _i = 0; _i < NJP_HACC; ++_i) njp_hacc_store[_i] = 0.f; float* njp_hacc = njp_hacc_store
#define NJP_HACC_OF(tid) (njp_hacc)
#endif

// ROWS: the calling warp owns rows (device: warps < nw_b; host simulation: always, the NJ_THREADS loops cover the helper ids)
template <int RG, int TR, bool ROWS>
NJ_HD void nj_path_bwd_tile_pipe(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem, const NjPathB& t0, const NjPathB& t1,
                                 float* njp_hacc, int cta, int u0, int u1) {
    constexpr int R = RG * TR;
    const int P = s.P_b, nt = s.nt_b, Pt = R * s.nw_b, nrow = 32 * s.nw_b, nth = nt - nrow;
    float* gpart = a.partials + (size_t)cta * c.img_floats;
    const NjPathBwd<RG, TR> B0(c, s, a, t0, smem), B1(c, s, a, t1, smem);
    const NjPathB& t = t0;
#if defined(NJODE_HOST_SIM)
    const bool helper_here = true;
#else
    const bool helper_here = !ROWS;
#endif
    if (ROWS) {
        NJ_THREADS(tid, nt) {
            if (tid < Pt) {
                const int u = u0 + tid;
                if (u < u1) {
                    const int32_t* dsc = a.b.unit_desc + (size_t)u * 6;
                    t.I[NJB_I_PATH * P + tid] = dsc[0]; t.I[NJB_I_C0 * P + tid] = dsc[3]; t.I[NJB_I_CUR * P + tid] = dsc[4];
                } else { t.I[NJB_I_PATH * P + tid] = -1; t.I[NJB_I_C0 * P + tid] = 0; t.I[NJB_I_CUR * P + tid] = 0; }
                t.I[NJB_I_ACT * P + tid] = 0;
                B0.set_prev(tid);
            }
        }
    }
    NJ_SYNC();
    if (ROWS) {
        NJ_WARPS(wp, s.nw_b) {
            const int r0 = wp * R;
            NJ_LANES(lane) {
                NJ_ROWMAP(R);
                const int r = r0 + er, p = t.I[NJB_I_PATH * P + r];
                const float* ght = (p >= 0 && a.grad_hT) ? a.grad_hT + (size_t)p * c.H : nullptr;
                for (int c_ = ec0; c_ < c.H; c_ += LPR) t.GH[r * s.sH + c_] = ght ? NJ_LDG(ght + c_) : 0.f;
                for (int c_ = ec0; c_ < c.d; c_ += LPR) t.GX[r * s.sD + c_] = 0.f;
                B0.load_state(r0, lane);
            }
            NJ_SYNCWARP();
        }
    }
    int nk = nj_pathb_next(t, P, Pt);
    int par = 0;                 // buffer the row warps write next
    bool pending = false;        // the helpers still owe the dW of the step in buffer par ^ 1
    for (int k = a.b.S; ; --k) {
        if (nk == k) {
            // drain, then the jump on buffer 0 with its dW phases through the partial image (all threads)
            if (pending && helper_here) {
                NJ_THREADS(tid, nt) { if (tid >= nrow) nj_path_dw_helper(c, s, par ? t0 : t1, NJP_HACC_OF(tid), tid - nrow, nth, Pt); }
            }
            pending = false;
            NJ_SYNC();
            if (ROWS) { NJ_WARPS(wp, s.nw_b) { B0.jump_p1(wp * R, wp, k); } }
            NJ_SYNC();
            NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t0, NJODE_NET_RO, gpart, tid, nt, Pt, t.MSK, R); }
            NJ_SYNC();
            if (ROWS) { NJ_WARPS(wp, s.nw_b) { B0.jump_p2(wp * R, wp); } }
            NJ_SYNC();
            NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t0, c.use_rnn ? NJODE_NET_GRU_HH : NJODE_NET_ENC, gpart, tid, nt, Pt, t.MSK, R); }
            NJ_SYNC();
            if (c.use_rnn) {
                if (ROWS) { NJ_WARPS(wp, s.nw_b) { B0.jump_p2b(wp * R, wp); } }
                NJ_SYNC();
                NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t0, NJODE_NET_GRU_IH, gpart, tid, nt, Pt, t.MSK, R); }
                NJ_SYNC();
            }
            if (ROWS) { NJ_WARPS(wp, s.nw_b) { B0.jump_p3(wp * R, wp); } }
            NJ_SYNC();
            NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t0, NJODE_NET_RO, gpart, tid, nt, Pt, t.MSK, R); }
            NJ_SYNC();
            nk = nj_pathb_next(t, P, Pt);
        }
        if (k == 0) break;
        if (ROWS) { NJ_WARPS(wp, s.nw_b) { if (par) B1.step_local(wp * R, k - 1); else B0.step_local(wp * R, k - 1); } }
        if (pending && helper_here) {
            NJ_THREADS(tid, nt) { if (tid >= nrow) nj_path_dw_helper(c, s, par ? t0 : t1, NJP_HACC_OF(tid), tid - nrow, nth, Pt); }
        }
        NJ_SYNC();
        pending = true; par ^= 1;
    }
    if (pending && helper_here) {
        NJ_THREADS(tid, nt) { if (tid >= nrow) nj_path_dw_helper(c, s, par ? t0 : t1, NJP_HACC_OF(tid), tid - nrow, nth, Pt); }
    }
    NJ_SYNC();
    if (ROWS) { NJ_WARPS(wp, s.nw_b) { B0.start_local(wp * R); } }
    NJ_SYNC();
    NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t0, NJODE_NET_ENC, gpart, tid, nt, Pt, nullptr, R); }
    NJ_SYNC();
}

template <int RG, int TR>
NJ_HD void nj_path_cta_backward_pipe(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem, int cta) {
    const int nt = s.nt_b;
    constexpr int R = RG * TR;
    nj_stage_image(smem, a.image, c.img_floats, nt);
    nj_zero(smem + s.b_IN, s.b_smem_floats - s.b_IN, nt);
    float* gpart = a.partials + (size_t)cta * c.img_floats;
    nj_zero(gpart, c.img_floats, nt);
    NJ_SYNC();
    NjPathB t0, t1;
    nj_pathb_bind(t0, s, smem);
    t1 = t0;
    t1.IN = t0.IN + s.b_copy; t1.A = t0.A + s.b_copy; t1.G = t0.G + s.b_copy; t1.GOUT = t0.GOUT + s.b_copy;
    int* ctl = t0.I + NJB_I_COUNT * s.P_b;
    const int nrow = 32 * s.nw_b, rows = R * s.nw_b;
#if defined(NJODE_HOST_SIM)
    NJP_HACC_DECL(nt);
    for (;;) {
        ctl[0] = nj_atomic_inc(a.counter);
        const int tile = ctl[0];
        if (tile >= s.n_tiles_b) break;
        const int ub = tile * rows, ue = ub + rows < a.b.n_units ? ub + rows : a.b.n_units;
        nj_path_bwd_tile_pipe<RG, TR, true>(c, s, a, smem, t0, t1, njp_hacc, cta, ub, ue);
    }
    NJ_THREADS(tid, nt) { if (tid >= nrow) nj_path_helper_flush(c, s, NJP_HACC_OF(tid), gpart, tid - nrow, nt - nrow); }
#else
    // two copies of the tile loop: the helpers' 100 accumulator registers are live in their branch only, so the row
    // warps' GEMM code does not compete with them for registers
    if ((int)threadIdx.x >= nrow) {
        NJP_HACC_DECL(nt);
        for (;;) {
            __syncthreads();
            const int tile = ctl[0];
            __syncthreads();
            if (tile >= s.n_tiles_b) break;
            const int ub = tile * rows, ue = ub + rows < a.b.n_units ? ub + rows : a.b.n_units;
            nj_path_bwd_tile_pipe<RG, TR, false>(c, s, a, smem, t0, t1, njp_hacc, cta, ub, ue);
        }
        nj_path_helper_flush(c, s, njp_hacc, gpart, (int)threadIdx.x - nrow, nt - nrow);
    } else {
        for (;;) {
            if (threadIdx.x == 0) ctl[0] = nj_atomic_inc(a.counter);
            __syncthreads();
            const int tile = ctl[0];
            __syncthreads();
            if (tile >= s.n_tiles_b) break;
            const int ub = tile * rows, ue = ub + rows < a.b.n_units ? ub + rows : a.b.n_units;
            nj_path_bwd_tile_pipe<RG, TR, true>(c, s, a, smem, t0, t1, nullptr, cta, ub, ue);
        }
    }
#endif
}

template <int RG, int TR>
NJ_HD void nj_path_cta_backward(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem, int cta) {
    const int nt = s.nt_b;
    constexpr int R = RG * TR;
    nj_stage_image(smem, a.image, c.img_floats, nt);
    nj_zero(smem + s.b_IN, s.b_smem_floats - s.b_IN, nt);
    nj_zero(a.partials + (size_t)cta * c.img_floats, c.img_floats, nt);
    NJ_SYNC();
    NjPathB t;
    nj_pathb_bind(t, s, smem);
    NJ_ACC_DECL(nt);
    int* ctl = t.I + NJB_I_COUNT * s.P_b;
    for (;;) {
        NJ_THREADS(tid, nt) { if (tid == 0) ctl[0] = nj_atomic_inc(a.counter); }
        NJ_SYNC();
        const int tile = ctl[0];
        NJ_SYNC();
        if (tile >= s.n_tiles_b) break;
        const int rows = R * s.nw_b;
        const int ub = tile * rows, ue = ub + rows < a.b.n_units ? ub + rows : a.b.n_units;
#if !defined(NJODE_HOST_SIM)
        if ((int)(threadIdx.x >> 5) >= s.nw_b) { nj_path_bwd_tile<RG, TR, false>(c, s, a, smem, t, nj_acc_base, cta, ub, ue); continue; }
#endif
        nj_path_bwd_tile<RG, TR, true>(c, s, a, smem, t, nj_acc_base, cta, ub, ue);
    }
    float* gpart = a.partials + (size_t)cta * c.img_floats;
    NJ_THREADS(tid, nt) { nj_path_dw_flush(c, s, NJ_ACC(tid), gpart, tid, nt); }
}

// ================================================================================================
// weight-stationary Euler steps: small whole-path batches (the reference's PhysioNet batch is 50 records)
//
// With at most a handful of paths per SM the warp GEMMs above leave the machine idle: one warp per path exposes every
// instruction latency of a 3 700-step dependent chain.  Here ONE CTA of NW = ceil(widest layer / 4) warps owns the R <= 8
// paths of a tile and all of its threads cooperate on every path:
//   * the ODE network's weights live in REGISTERS for the whole launch: thread (warp w, lane = (oq, ksl)) holds the slice
//     ksl (of 8) of the weight row of output o = 4 w + oq of every layer -- at most 12 floats per layer;
//   * a layer of a step = 3 broadcast LDS.128 of the activation slice + 12 FFMA + 3 shuffles (the 8 slices of an output
//     meet), one CTA barrier; no weight ever moves after the launch has started;
//   * backward: dW[o][slice] += g[o] * a[slice] is thread-local (the gradient tile of a thread mirrors its weight tile,
//     also in registers for the whole launch); the input gradient g . W is reduced over the 4 outputs of a warp by
//     shuffles and over the warps through a partial buffer in shared memory, summed by its consumer after the barrier.
// Jumps, the start encoder and path records stay with warp 0 on the warp GEMMs above (the other warps wait).
// ================================================================================================
#define NJT_MAXL 3                  // Linear layers of the ODE network
#define NJT_SLICE 12                // floats of one k-slice: 3 float4 chunks, 8 slices -> layer inputs up to 96 wide
#define NJT_PARTW 96

// layer 0 (input width up to 96): slices of 12 floats; the other layers (hidden widths up to 64): slices of 8
struct NjStatRegs {
    float w0[12], w1[8], w2[8];
    float dw0[12], dw1[8], dw2[8];
    float b[NJT_MAXL], db[NJT_MAXL];
};
template <int L> struct NjStatL { static constexpr int Q = L == 0 ? 3 : 2; };     // float4 chunks per slice
template <int L> NJ_HD float* nj_stat_w(NjStatRegs& r) { return L == 0 ? r.w0 : (L == 1 ? r.w1 : r.w2); }
template <int L> NJ_HD const float* nj_stat_w(const NjStatRegs& r) { return L == 0 ? r.w0 : (L == 1 ? r.w1 : r.w2); }
template <int L> NJ_HD float* nj_stat_dw(NjStatRegs& r) { return L == 0 ? r.dw0 : (L == 1 ? r.dw1 : r.dw2); }

struct NjStatGeo { int n, K4[NJT_MAXL], KS[NJT_MAXL], O[NJT_MAXL]; };

NJ_HD NjStatGeo nj_stat_geo(const NjCfg& c) {
    NjStatGeo g;
    const NjNet& N = c.net[NJODE_NET_ODE];
    g.n = N.n;
    for (int l = 0; l < NJT_MAXL; ++l) {
        g.K4[l] = l < N.n ? (N.dim[l] + 3) >> 2 : 0;
        g.KS[l] = (g.K4[l] + 7) >> 3;                  // <= 3 for layer 0, <= 2 for the others (planner)
        g.O[l] = l < N.n ? N.dim[l + 1] : 0;
    }
    return g;
}

#if defined(NJODE_HOST_SIM)
#define NJT_REGS_DECL(nt) std::vector<NjStatRegs> njt_regs_store(nt); NjStatRegs* njt_regs = njt_regs_store.data()
#define NJT_REGS(tid) (njt_regs[tid])
#else
#define NJT_REGS_DECL(nt) NjStatRegs njt_regs_store; NjStatRegs* njt_regs = &njt_regs_store
#define NJT_REGS(tid) (*njt_regs)
#endif

// weight slices of this thread from the parameter image in shared memory (rows beyond a layer's outputs: zero)
template <int L>
NJ_HD void nj_stat_load_layer(const NjCfg& c, const NjStatGeo& g, const float* simg, int tid, NjStatRegs& R_) {
    const NjNet& N = c.net[NJODE_NET_ODE];
    const int o = 4 * (tid >> 5) + ((tid & 31) >> 3), ksl = tid & 7;
    float* w = nj_stat_w<L>(R_);
    float* dw = nj_stat_dw<L>(R_);
#pragma unroll
    for (int j = 0; j < 4 * NjStatL<L>::Q; ++j) {
        const int q = j >> 2, c4 = ksl * g.KS[L] + q;
        float v = 0.f;
        if (L < g.n && o < g.O[L] && q < g.KS[L] && c4 < g.K4[L]) v = simg[N.w_img[L] + o * N.ks[L] + 4 * c4 + (j & 3)];
        w[j] = v; dw[j] = 0.f;
    }
    R_.b[L] = (L < g.n && o < g.O[L] && N.b_src[L] >= 0) ? simg[N.b_img[L] + o] : 0.f;
    R_.db[L] = 0.f;
}
NJ_HD void nj_stat_load(const NjCfg& c, const NjStatGeo& g, const float* simg, int tid, NjStatRegs& R_) {
    nj_stat_load_layer<0>(c, g, simg, tid, R_);
    nj_stat_load_layer<1>(c, g, simg, tid, R_);
    nj_stat_load_layer<2>(c, g, simg, tid, R_);
}

// partial dot product of one weight slice with the matching activation slice of a row
template <int Q>
NJ_HD float nj_stat_dot(const float* w, const float* in_row, int K4, int KS, int ksl) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int c4 = ksl * KS + q;
        if (q < KS && c4 < K4) {
            const nj_f4 a = nj_sp_ld4(nj_sp_of(in_row + 4 * c4));
            s0 = fmaf(a.x, w[4 * q], s0); s1 = fmaf(a.y, w[4 * q + 1], s1);
            s2 = fmaf(a.z, w[4 * q + 2], s2); s3 = fmaf(a.w, w[4 * q + 3], s3);
        }
    }
    return (s0 + s1) + (s2 + s3);
}

// sum over the 8 k-slices of an output (lanes ksl = 0..7 of one oq group); the total is valid in the ksl == 0 lane
template <int L>
NJ_HD float nj_stat_sum8(NjStatRegs* regs, int tid, const float* in_row, int K4, int KS) {
#if defined(NJODE_HOST_SIM)
    float p[8];
    for (int sl = 0; sl < 8; ++sl) p[sl] = nj_stat_dot<NjStatL<L>::Q>(nj_stat_w<L>(regs[tid + sl]), in_row, K4, KS, sl);
    return ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
#else
    (void)tid;
    float v = nj_stat_dot<NjStatL<L>::Q>(nj_stat_w<L>(*regs), in_row, K4, KS, threadIdx.x & 7);
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 4);
    return v;
#endif
}

// one layer of the forward step over the R rows of the tile; hidden layers: out = act(. + b) with dropout marks,
// last layer (hs != nullptr): hs[r][o] += dt * (. + b).  The layer index is a template parameter: the register arrays of
// NjStatRegs must only ever be indexed with compile-time constants, or they end up in local memory.
template <int l>
NJ_HD void nj_stat_layer_fwd(const NjCfg& c, const NjStatGeo& g, NjStatRegs* regs, int tid, int R,
                             const float* in, int in_s, float* out, int out_s, float* hs, int hs_s, const float* dtv, const int* rk) {
    const int lane = tid & 31, o = 4 * (tid >> 5) + (lane >> 3), ksl = lane & 7;
#if defined(NJODE_HOST_SIM)
    if (ksl != 0) return;
#endif
    const NjNet& N = c.net[NJODE_NET_ODE];
    const bool lead = ksl == 0 && o < g.O[l];
#if defined(NJODE_HOST_SIM)
    const float bias = regs[tid].b[l];
#else
    const float bias = regs->b[l];
#endif
    for (int r = 0; r < R; ++r) {
        const float sum = nj_stat_sum8<l>(regs, tid, in + (size_t)r * in_s, g.K4[l], g.KS[l]);
        if (lead) {
            if (hs) hs[r * hs_s + o] = fmaf(dtv[r], sum + bias, hs[r * hs_s + o]);     // dtv[r] = 0: the row rests
            else {
                float v = nj_act(sum + bias, N.act[l]);
                if (c.has_drop) {
                    const unsigned lk = nj_layer_key((unsigned)rk[r], (unsigned)(NJODE_NET_ODE * 16 + l + 1));
                    v = nj_keep(lk, (unsigned)o, c.thr) ? v * c.keep_scale : nj_u2f(NJ_DROPPED);
                }
                out[(size_t)r * out_s + o] = v;
            }
        }
    }
}

// the ODE network's input rows of step k for the R rows of the tile: [tanh(last_X), tanh(h), tau, t - tau(, t)]
// (ODEFunc.forward, NJODE/models.py:188-199); hsrc: h of the step (forward: HS, also written to h_hist; backward: h_hist)
template <bool BWD>
NJ_HD void nj_stat_build_in(const NjCfg& c, const NjArgs& a, int tid, int nt, int R, int k, const int* path, const float* TX, int sD,
                            const float* tau, float* HS, int sH, float* IN, int sI, int* rk, const float* GH, float* GOUT, int sO,
                            float* dtv) {
    const int inf4 = ((c.inf + 3) >> 2) << 2;
    const float tcur = NJ_LDG(a.b.step_t + k), dt = NJ_LDG(a.b.step_dt + k);
    const int c_ = tid & 127, groups = nt >> 7;        // 128 threads per row; a trailing partial group stays idle
    for (int r = tid >> 7; r < R && (tid >> 7) < groups; r += groups) {
        const int p = path[r];
        if (c_ < inf4) {
            float v = 0.f;
            if (c_ < c.d) v = TX[r * sD + c_];
            else if (c_ < c.d + c.H) {
                float h;
                float* hh = (p >= 0 && a.h_hist) ? a.h_hist + ((size_t)k * a.b.B + p) * c.H + (c_ - c.d) : nullptr;
                if (BWD) h = hh ? *hh : 0.f;
                else { h = HS[r * sH + c_ - c.d]; if (hh) *hh = h; }
                v = nj_tanh(h);
            } else if (c_ < c.inf) {
                const float t_ = tau[r];
                if (c_ == c.d + c.H) v = t_;
                else if (c_ == c.d + c.H + 1) v = tcur - t_;
                else v = t_ + (tcur - t_);
            }
            IN[(size_t)r * sI + c_] = v;
        }
        if (BWD && c_ < c.H) GOUT[(size_t)r * sO + c_] = dt * GH[r * sH + c_];
        if (c_ == 127) {
            rk[r] = (int)nj_row_key(c.seed_lo, c.seed_hi, (unsigned)(p + a.b.path_id_offset), (unsigned)k);
            if (dtv) dtv[r] = dt;
        }
    }
}

// forward Euler step k of the tile, all threads of the CTA (n + 1 barriers)
template <int RG, int TR>
NJ_HD void nj_stat_fwd_step(const NjCfg& c, const NjPath& s, const NjArgs& a, const NjStatGeo& g, NjStatRegs* njt_regs,
                            NjPathFwd<RG, TR>& f, int nt, int k) {
    constexpr int R = RG * TR, RS = NJP_RS;
    float* dtv = f.F + NJP_F_CA * RS;              // (the loss-coefficient slots are free in the forward pass)
    NJ_THREADS(tid, nt) {
        nj_stat_build_in<false>(c, a, tid, nt, R, k, f.I + NJP_I_PATH * RS, f.TX, s.sD, f.F + NJP_F_TAU * RS, f.HS, s.sH,
                                f.w.IN, s.sI, f.w.RK, nullptr, nullptr, 0, dtv);
    }
    NJ_SYNC();
    const int* rk = f.w.RK;
    NJ_THREADS(tid, nt) { nj_stat_layer_fwd<0>(c, g, njt_regs, tid, R, f.w.IN, s.sI, f.w.A0, s.sA, nullptr, 0, nullptr, rk); }
    NJ_SYNC();
    if (g.n == 2) {
        NJ_THREADS(tid, nt) { nj_stat_layer_fwd<1>(c, g, njt_regs, tid, R, f.w.A0, s.sA, nullptr, 0, f.HS, s.sH, dtv, rk); }
        NJ_SYNC();
    } else {
        NJ_THREADS(tid, nt) { nj_stat_layer_fwd<1>(c, g, njt_regs, tid, R, f.w.A0, s.sA, f.w.A1, s.sA, nullptr, 0, nullptr, rk); }
        NJ_SYNC();
        NJ_THREADS(tid, nt) { nj_stat_layer_fwd<2>(c, g, njt_regs, tid, R, f.w.A1, s.sA, nullptr, 0, f.HS, s.sH, dtv, rk); }
        NJ_SYNC();
    }
}

template <int RG, int TR>
NJ_HD void nj_stat_cta_forward(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem) {
    constexpr int R = RG * TR;
    const int nt = 32 * s.nw_s;
    float* simg = smem;
    nj_stage_image(simg, a.image, c.img_floats, nt);
    nj_zero(smem + s.f_warp0, s.f_region, nt);
    NJ_SYNC();
    const NjStatGeo g = nj_stat_geo(c);
    NJT_REGS_DECL(nt);
    NJ_THREADS(tid, nt) { nj_stat_load(c, g, simg, tid, NJT_REGS(tid)); }
    float* reg = smem + s.f_warp0;
    int* slot = reinterpret_cast<int*>(reg + s.f_I) + NJP_I_COUNT * NJP_RS;
    NjPathFwd<RG, TR> f(c, s, a, reg, simg);
    const bool rec = a.b.E > 0;
    const int S = a.b.S;
    for (;;) {
        NJ_THREADS(tid, nt) { if (tid == 0) *slot = nj_atomic_inc(a.counter); }
        NJ_SYNC();
        const int wt = *slot;
        NJ_SYNC();
        if (wt >= s.n_tiles_f) break;
        const int ub = wt * R, ue = ub + R < a.b.n_units ? ub + R : a.b.n_units;
        NJ_WARPS(wp, 1) { if (wp == 0) f.begin(ub, ue); }
        NJ_SYNC();
        int k = 0, gi = 0;
        for (;;) {
            const int nk = f.next_jump(gi);
            const int kend = nk < S ? nk : S;
            for (; k < kend; ++k) {
                nj_stat_fwd_step<RG, TR>(c, s, a, g, njt_regs, f, nt, k);
                if (rec) {
                    NJ_WARPS(wp, 1) { if (wp == 0) f.record(NJ_LDG(a.b.step_event + k), NJ_EVENT_PATH_RO_BASE + (unsigned)k); }
                    NJ_SYNC();
                }
            }
            if (nk > S) break;
            const bool any = f.any_jumps_at(nk);
            NJ_SYNC();                            // every warp has read the cursors before warp 0 advances them
            NJ_WARPS(wp, 1) {
                if (wp == 0) {
                    if (any) f.jump(nk);
                    if (rec) f.record(NJ_LDG(a.b.jump_event + gi), NJ_EVENT_JUMP_BASE + 3u * (unsigned)gi + 2u);
                }
            }
            if (rec) ++gi;
            NJ_SYNC();
        }
        NJ_WARPS(wp, 1) { if (wp == 0) f.finish(); }
        NJ_SYNC();
    }
}

// ---- backward ----
// sum over the 4 outputs of a warp (lanes oq = 0..3 with equal ksl); the total is valid in the oq == 0 lanes
NJ_HD float nj_stat_sum4(float v) {
#if !defined(NJODE_HOST_SIM)
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 8);
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 16);
#endif
    return v;
}

// reverse of layer l for the R rows: with g[r][o] = dL/d(pre-activation output o),
//   dW[o][slice] += g a[slice], db[o] += g               (thread-local registers)
//   part[r][warp][k] = sum over the warp's 4 outputs of g[o] W[o][k]   (input-gradient partials, summed by the consumer)
template <int l>
NJ_HD void nj_stat_layer_bwd(const NjStatGeo& g, NjStatRegs* regs, int tid, int nw, int R,
                             const float* gout, int g_s, const float* in, int in_s, float* part) {
    const int lane = tid & 31, wq = tid >> 5, oq = lane >> 3, ksl = lane & 7, o = 4 * wq + oq;
    const int KS = g.KS[l], K4 = g.K4[l];
#if defined(NJODE_HOST_SIM)
    // sequential lanes: every lane updates its own dW; the oq == 0 lane of a slice sums the four partial products
    for (int r = 0; r < R; ++r) {
        const float gv = o < g.O[l] ? gout[(size_t)r * g_s + o] : 0.f;
        NjStatRegs& me = regs[tid];
        float* dw = nj_stat_dw<l>(me);
        for (int q = 0; q < NjStatL<l>::Q; ++q) {
            const int c4 = ksl * KS + q;
            if (q < KS && c4 < K4) {
                const nj_f4 a = nj_ld4(in + (size_t)r * in_s + 4 * c4);
                dw[4 * q] = fmaf(gv, a.x, dw[4 * q]); dw[4 * q + 1] = fmaf(gv, a.y, dw[4 * q + 1]);
                dw[4 * q + 2] = fmaf(gv, a.z, dw[4 * q + 2]); dw[4 * q + 3] = fmaf(gv, a.w, dw[4 * q + 3]);
            }
        }
        if (ksl == 0) me.db[l] += gv;
        if (oq == 0) {
            for (int q = 0; q < NjStatL<l>::Q; ++q) {
                const int c4 = ksl * KS + q;
                if (!(q < KS && c4 < K4)) continue;
                for (int e = 0; e < 4; ++e) {
                    float p4[4];
                    for (int oo = 0; oo < 4; ++oo) {
                        const int o2 = 4 * wq + oo;
                        const float g2 = o2 < g.O[l] ? gout[(size_t)r * g_s + o2] : 0.f;
                        p4[oo] = g2 * nj_stat_w<l>(regs[tid + 8 * oo])[4 * q + e];
                    }
                    part[((size_t)r * nw + wq) * NJT_PARTW + 4 * c4 + e] = (p4[0] + p4[1]) + (p4[2] + p4[3]);
                }
            }
        }
    }
#else
    NjStatRegs& me = *regs;
    float* dw = nj_stat_dw<l>(me);
    const float* w = nj_stat_w<l>(me);
    for (int r = 0; r < R; ++r) {
        const float gv = o < g.O[l] ? gout[(size_t)r * g_s + o] : 0.f;
#pragma unroll
        for (int q = 0; q < NjStatL<l>::Q; ++q) {
            const int c4 = ksl * KS + q;
            if (q < KS && c4 < K4) {
                const nj_f4 a = nj_sp_ld4(nj_sp_of(in + (size_t)r * in_s + 4 * c4));
                dw[4 * q] = fmaf(gv, a.x, dw[4 * q]); dw[4 * q + 1] = fmaf(gv, a.y, dw[4 * q + 1]);
                dw[4 * q + 2] = fmaf(gv, a.z, dw[4 * q + 2]); dw[4 * q + 3] = fmaf(gv, a.w, dw[4 * q + 3]);
            }
        }
        if (ksl == 0) me.db[l] += gv;
#pragma unroll
        for (int q = 0; q < NjStatL<l>::Q; ++q) {
            if (q < KS) {                     // uniform over the warp: every lane takes part in the shuffles
                nj_f4 v;
                v.x = nj_stat_sum4(gv * w[4 * q]); v.y = nj_stat_sum4(gv * w[4 * q + 1]);
                v.z = nj_stat_sum4(gv * w[4 * q + 2]); v.w = nj_stat_sum4(gv * w[4 * q + 3]);
                const int c4 = ksl * KS + q;
                if (oq == 0 && c4 < K4) nj_st4(part + ((size_t)r * nw + wq) * NJT_PARTW + 4 * c4, v);
            }
        }
    }
#endif
}

// consumer of the partials of layer l (l >= 1): g_prev[r][o'] = (sum over the warps) * act'(a[r][o']) for the hidden
// activation a = output of layer l - 1; written by the ksl == 0 lane of output o', read by its 8 lanes after a warp sync
NJ_HD void nj_stat_consume_hidden(const NjCfg& c, const NjStatGeo& g, int tid, int nw, int l, int R, const float* part,
                                  const float* act, int a_s, float* gprev, int gp_s) {
    const int lane = tid & 31, o = 4 * (tid >> 5) + (lane >> 3), ksl = lane & 7;
    const NjNet& N = c.net[NJODE_NET_ODE];
    if (ksl != 0 || o >= g.O[l - 1]) return;
    for (int r = 0; r < R; ++r) {
        float v = 0.f;
        for (int wq = 0; wq < nw; ++wq) v += part[((size_t)r * nw + wq) * NJT_PARTW + o];
        float a_ = act[(size_t)r * a_s + o];
        if (c.has_drop) {
            if (nj_f2u(a_) == NJ_DROPPED) v = 0.f;
            else { a_ *= c.one_minus_p; v *= c.keep_scale; }
        }
        if (N.act[l - 1] == NJODE_ACT_TANH) v *= (1.f - a_ * a_);
        else if (N.act[l - 1] == NJODE_ACT_RELU) v = a_ > 0.f ? v : 0.f;
        gprev[(size_t)r * gp_s + o] = v;
    }
}

// reverse Euler step k of the tile, all threads of the CTA (2n + 2 barriers)
template <int RG, int TR>
NJ_HD void nj_stat_bwd_step(const NjCfg& c, const NjPath& s, const NjArgs& a, const NjStatGeo& g, NjStatRegs* njt_regs,
                            const NjPathB& t, float* part0, float* part1, int nt, int k) {
    constexpr int R = RG * TR;
    const int P = s.P_b, nw = s.nw_s, wa = P * s.sA;
    NJ_THREADS(tid, nt) {
        nj_stat_build_in<true>(c, a, tid, nt, R, k, t.I + NJB_I_PATH * P, t.TX, s.sD, t.F + NJP_F_TAU * P, nullptr, s.sH,
                               t.IN, s.sI, t.I + NJB_I_RK * P, t.GH, t.GOUT, s.sO, nullptr);
    }
    NJ_SYNC();
    // recompute the hidden activations (kept with their dropout marks)
    int* rk = t.I + NJB_I_RK * P;
    float* A0 = t.A; float* A1 = t.A + wa; float* G0 = t.G; float* G1 = t.G + wa;
    NJ_THREADS(tid, nt) { nj_stat_layer_fwd<0>(c, g, njt_regs, tid, R, t.IN, s.sI, A0, s.sA, nullptr, 0, nullptr, rk); }
    NJ_SYNC();
    if (g.n == 3) {
        NJ_THREADS(tid, nt) { nj_stat_layer_fwd<1>(c, g, njt_regs, tid, R, A0, s.sA, A1, s.sA, nullptr, 0, nullptr, rk); }
        NJ_SYNC();
    }
    // layers in reverse: dW in registers, input-gradient partials through shared memory (two buffers in turn: a warp may
    // still sum the previous layer's partials while another one writes the next)
    float* part;
    if (g.n == 3) {
        NJ_THREADS(tid, nt) { nj_stat_layer_bwd<2>(g, njt_regs, tid, nw, R, t.GOUT, s.sO, A1, s.sA, part0); }
        NJ_SYNC();
        NJ_THREADS(tid, nt) { nj_stat_consume_hidden(c, g, tid, nw, 2, R, part0, A1, s.sA, G1, s.sA); }
        NJ_THREADS(tid, nt) { NJ_SYNCWARP(); }
        NJ_THREADS(tid, nt) { nj_stat_layer_bwd<1>(g, njt_regs, tid, nw, R, G1, s.sA, A0, s.sA, part1); }
        NJ_SYNC();
        NJ_THREADS(tid, nt) { nj_stat_consume_hidden(c, g, tid, nw, 1, R, part1, A0, s.sA, G0, s.sA); }
        NJ_THREADS(tid, nt) { NJ_SYNCWARP(); }
        NJ_THREADS(tid, nt) { nj_stat_layer_bwd<0>(g, njt_regs, tid, nw, R, G0, s.sA, t.IN, s.sI, part0); }
        NJ_SYNC();
        part = part1;
    } else {
        NJ_THREADS(tid, nt) { nj_stat_layer_bwd<1>(g, njt_regs, tid, nw, R, t.GOUT, s.sO, A0, s.sA, part0); }
        NJ_SYNC();
        NJ_THREADS(tid, nt) { nj_stat_consume_hidden(c, g, tid, nw, 1, R, part0, A0, s.sA, G0, s.sA); }
        NJ_THREADS(tid, nt) { NJ_SYNCWARP(); }
        NJ_THREADS(tid, nt) { nj_stat_layer_bwd<0>(g, njt_regs, tid, nw, R, G0, s.sA, t.IN, s.sI, part1); }
        NJ_SYNC();
        part = part0;
    }
    // the partials of layer 0 -> adjoint of h (and of last_X for the masked model)
    const float* pl = part == part0 ? part1 : part0;
    NJ_THREADS(tid, nt) {
        const int lane = tid & 31, o = 4 * (tid >> 5) + (lane >> 3), ksl = lane & 7;
        if (ksl == 0) {
            for (int r = 0; r < R; ++r) {
                if (t.I[NJB_I_PATH * P + r] < 0) continue;
                if (o < c.H) {
                    float v = 0.f;
                    for (int wq = 0; wq < nw; ++wq) v += pl[((size_t)r * nw + wq) * NJT_PARTW + c.d + o];
                    const float th = t.IN[(size_t)r * s.sI + c.d + o];
                    t.GH[r * s.sH + o] += v * (1.f - th * th);
                }
                if (c.masked && o < c.d) {
                    float v = 0.f;
                    for (int wq = 0; wq < nw; ++wq) v += pl[((size_t)r * nw + wq) * NJT_PARTW + o];
                    const float tx = t.IN[(size_t)r * s.sI + o];
                    t.GX[r * s.sD + o] += v * (1.f - tx * tx);
                }
            }
        }
    }
    NJ_SYNC();
}

// the register tiles of the ODE network's gradient -> this CTA's partial image (every element has exactly one owner)
template <int L>
NJ_HD void nj_stat_flush_layer(const NjCfg& c, const NjStatGeo& g, NjStatRegs& R_, int tid, float* gpart) {
    const NjNet& N = c.net[NJODE_NET_ODE];
    const int o = 4 * (tid >> 5) + ((tid & 31) >> 3), ksl = tid & 7;
    if (L >= g.n || o >= g.O[L]) return;
    const float* dw = nj_stat_dw<L>(R_);
#pragma unroll
    for (int j = 0; j < 4 * NjStatL<L>::Q; ++j) {
        const int q = j >> 2, c4 = ksl * g.KS[L] + q;
        if (q < g.KS[L] && c4 < g.K4[L]) gpart[N.w_img[L] + o * N.ks[L] + 4 * c4 + (j & 3)] = dw[j];
    }
    if (ksl == 0 && N.b_src[L] >= 0) gpart[N.b_img[L] + o] = R_.db[L];
}
NJ_HD void nj_stat_flush(const NjCfg& c, const NjStatGeo& g, NjStatRegs& R_, int tid, float* gpart) {
    nj_stat_flush_layer<0>(c, g, R_, tid, gpart);
    nj_stat_flush_layer<1>(c, g, R_, tid, gpart);
    nj_stat_flush_layer<2>(c, g, R_, tid, gpart);
}

// jump-network dW phase of the weight-stationary kernel: every tile goes through the partial image (no register tiles:
// the registers hold the ODE network; jumps are ~2 % of the steps)
NJ_HDN void nj_stat_dw(const NjCfg* cp, const NjPath* sp, const NjPathB* tp, int netid, float* gpart, int tid, int nt, int Pt,
                       const int* msk, int R) {
    const NjCfg& c = *cp; const NjPath& s = *sp; const NjPathB& t = *tp;
    // the tiles of one network are a contiguous range of the tile numbering: walk that range only (decoding every tile of
    // every network per phase was 6 decodes per thread and phase for 2 matching tiles)
    const int lo = s.tile_base[netid][0];
    int hi = s.tiles_total;
    for (int q = 0; q < NJODE_NUM_NETS; ++q) {
        const int b = s.tile_base[q][0];
        if (b > lo && b < hi) hi = b;
    }
    for (int T = lo + tid; T < hi; T += nt) {
        int l, og, kg;
        if (!nj_path_tile_decode(c, s, netid, T, l, og, kg)) continue;
        float q[20];
#pragma unroll
        for (int i = 0; i < 20; ++i) q[i] = 0.f;
        nj_path_dw_rows(c, s, t, netid, l, og, kg, Pt, msk, R, q);
        nj_seg_tile_store(c, netid, l, og, kg, q, gpart, true);
    }
}

template <int RG, int TR>
NJ_HD void nj_stat_cta_backward(const NjCfg& c, const NjPath& s, const NjArgs& a, float* smem, int cta) {
    constexpr int R = RG * TR;
    const int nt = 32 * s.nw_s, P = s.P_b;
    nj_stage_image(smem, a.image, c.img_floats, nt);
    nj_zero(smem + s.b_IN, s.b_smem_floats - s.b_IN, nt);
    float* gpart = a.partials + (size_t)cta * c.img_floats;
    nj_zero(gpart, c.img_floats, nt);
    NJ_SYNC();
    NjPathB t;
    nj_pathb_bind(t, s, smem);
    const NjStatGeo g = nj_stat_geo(c);
    NJT_REGS_DECL(nt);
    NJ_THREADS(tid, nt) { nj_stat_load(c, g, smem, tid, NJT_REGS(tid)); }
    float* part0 = smem + s.b_PART;
    float* part1 = part0 + (size_t)R * s.nw_s * NJT_PARTW;
    const NjPathBwd<RG, TR> B(c, s, a, t, smem);
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
        NJ_SYNC();
        int nk = nj_pathb_next(t, P, R);
        for (int k = a.b.S; ; --k) {
            if (nk == k) {
                NJ_WARPS(wp, 1) { if (wp == 0) B.jump_p1(0, 0, k); }
                NJ_SYNC();
                NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t, NJODE_NET_RO, gpart, tid, nt, R, t.MSK, R); }
                NJ_SYNC();
                NJ_WARPS(wp, 1) { if (wp == 0) B.jump_p2(0, 0); }
                NJ_SYNC();
                NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t, c.use_rnn ? NJODE_NET_GRU_HH : NJODE_NET_ENC, gpart, tid, nt, R, t.MSK, R); }
                NJ_SYNC();
                if (c.use_rnn) {
                    NJ_WARPS(wp, 1) { if (wp == 0) B.jump_p2b(0, 0); }
                    NJ_SYNC();
                    NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t, NJODE_NET_GRU_IH, gpart, tid, nt, R, t.MSK, R); }
                    NJ_SYNC();
                }
                NJ_WARPS(wp, 1) { if (wp == 0) B.jump_p3(0, 0); }
                NJ_SYNC();
                NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t, NJODE_NET_RO, gpart, tid, nt, R, t.MSK, R); }
                NJ_SYNC();
                nk = nj_pathb_next(t, P, R);
            }
            if (k == 0) break;
            nj_stat_bwd_step<RG, TR>(c, s, a, g, njt_regs, t, part0, part1, nt, k - 1);
        }
        NJ_WARPS(wp, 1) { if (wp == 0) B.start_local(0); }
        NJ_SYNC();
        NJ_THREADS(tid, nt) { nj_stat_dw(&c, &s, &t, NJODE_NET_ENC, gpart, tid, nt, R, nullptr, R); }
        NJ_SYNC();
    }
    NJ_THREADS(tid, nt) { nj_stat_flush(c, g, NJT_REGS(tid), tid, gpart); }
}

