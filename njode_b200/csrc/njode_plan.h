// njode_plan.h -- host-side launch planning: validates the model description, lays out the padded
// parameter image and the shared-memory matrices of njode_core.cuh, picks the tile size.
#pragma once
#include <string>
#include <algorithm>
#include "njode_core.cuh"
#include "njode_seg.cuh"
#include "njode_tpn.cuh"

struct NjPlanOut {
    NjCfg fwd, bwd;
    NjSeg seg;                      // seg.ok: the segment fast path (njode_seg.cuh) serves this call
    NjPath path;                    // path.ok: the warp-GEMM whole-path kernels (njode_path.cuh) serve this call
    int path_grid_f, path_grid_b;
    size_t path_smem_f_bytes, path_smem_b_bytes;
    int seg_grid_f, seg_grid_b;
    size_t seg_smem_f_bytes, seg_smem_b_bytes;
    int n_tiles;
    int grid_fwd, grid_bwd;
    size_t smem_fwd_bytes, smem_bwd_bytes;
    size_t ws_image_off, ws_rowloss_off, ws_counter_off, ws_partials_off, ws_bytes;
    size_t act_bytes; int act_nh, act_wp;      // saved hidden activations of the ODE network (segment warp kernels), see NjArgs
    size_t ws_scratch_off, scratch_bytes;      // segment backward in recompute mode: h chains of the tiles in flight
};

// compact: the generic kernels (njode_core.cuh) only touch rows < 4 * og of a weight block; the 8*to*nch row padding
// is for the warp GEMMs of the segment kernels.  Wide-ish nets (2x100) that cannot take the segment path anyway get
// the compact image so that it fits shared memory.
static inline bool nj_fill_nets(const njode_model_t& m, NjCfg& c, bool compact, std::string& err) {
    int off = 0;
    const int nnets = m.use_rnn ? NJODE_NUM_NETS : 3;
    for (int n = 0; n < nnets; ++n) {
        const njode_mlp_t& s = m.net[n];
        NjNet& N = c.net[n];
        if (s.n_linear < 1 || s.n_linear > NJODE_MAX_LINEAR) { err = "n_linear out of range"; return false; }
        N.n = s.n_linear;
        for (int l = 0; l <= s.n_linear; ++l) {
            N.dim[l] = s.dims[l];
            if (s.dims[l] < 1) { err = "layer width < 1"; return false; }
        }
        for (int l = 0; l < s.n_linear; ++l) {
            N.act[l] = s.act[l];
            if (l < s.n_linear - 1 && s.act[l] != NJODE_ACT_TANH && s.act[l] != NJODE_ACT_RELU) { err = "unknown activation"; return false; }
            N.ks[l] = nj_stride_host(s.dims[l]);
            N.og[l] = (s.dims[l + 1] + 3) / 4;
            const int o8 = (s.dims[l + 1] + 7) / 8;
            // warp GEMM chunks: 64 outputs (8 per lane) each, the last one 8 * tol; chunk bases stay multiples of 16
            // (hash pairs) and the image holds exactly 8 * ceil(out / 8) rows (2x100 nets: 104 rows, not 128)
            N.nch[l] = (o8 + 7) / 8;
            N.to[l] = N.nch[l] > 1 ? 8 : o8;
            N.tol[l] = o8 - 8 * (N.nch[l] - 1);
            N.rp[l] = compact ? 4 * N.og[l] : 8 * o8;
            N.w_img[l] = off; off += N.rp[l] * N.ks[l];
            N.b_img[l] = off; off += N.rp[l];
            N.w_src[l] = s.w_off[l]; N.b_src[l] = s.b_off[l];
            N.m_in[l] = nj_magic_host(s.dims[l]); N.m_out[l] = nj_magic_host(s.dims[l + 1]);
        }
    }
    c.img_floats = off;
    return true;
}

// slices for the row-loop layer GEMMs (nj_rl_* in njode_core.cuh): `width` threads side by side (outputs / input
// columns), `chunks` groups of 4 along the reduction.  Returns the number of slices (0: micro-tile path) and the chunks
// per slice.
static inline int nj_rl_pick(int nrows, int width, int chunks, int nt, int& per) {
    per = 0;
    if (nrows > NJ_RL_MAXROWS || width > nt || width < 1 || chunks < 1) return 0;
    int ns = nt / width;
    if (ns > chunks) ns = chunks;
    if (ns > 16) ns = 16;
    per = (chunks + ns - 1) / ns;
    if (per > 8) return 0;
    ns = (chunks + per - 1) / per;
    if ((size_t)ns * nrows * width > NJ_RL_SCRATCH) return 0;
    return ns;
}

static inline void nj_fill_rl(NjCfg& c) {
    const char* off = getenv("NJODE_NO_ROWLOOP");
    const bool on = !(off && atoi(off)) && c.P <= NJ_RL_MAXROWS;
    for (int n = 0; n < NJODE_NUM_NETS; ++n) {
        NjNet& N = c.net[n];
        for (int l = 0; l < N.n; ++l) {
            N.rl_f_ns[l] = N.rl_f_kc[l] = N.rl_d_t0[l] = N.rl_d_ns[l] = N.rl_d_oc[l] = 0;
            if (!on) continue;
            const int K4 = (N.dim[l] + 3) / 4;
            N.rl_f_ns[l] = nj_rl_pick(c.P, N.dim[l + 1], K4, c.nt, N.rl_f_kc[l]);
            const int dw_tiles = N.og[l] * K4;
            N.rl_d_t0[l] = (dw_tiles + 2 * N.dim[l] <= c.nt) ? dw_tiles : 0;
            N.rl_d_ns[l] = nj_rl_pick(c.P, N.dim[l], N.og[l], c.nt - N.rl_d_t0[l], N.rl_d_oc[l]);
        }
    }
}

// lays out shared memory for tile size P; returns the number of floats needed
static inline void nj_layout(NjCfg& c, int P, int nt, bool bwd, bool w_smem, int dw_smem) {
    c.P = P; c.nt = nt; c.w_smem = w_smem; c.dw_smem = bwd ? dw_smem : 0;
    nj_fill_rl(c);
    // gradient image part kept in shared memory: everything (1) or the ODE network's blocks (2; first in the image)
    c.dimg_floats = c.dw_smem == 1 ? c.img_floats : (c.dw_smem == 2 ? c.net[NJODE_NET_ENC].w_img[0] : 0);
    int o = 0;
    c.o_img = o; if (w_smem) o += c.img_floats;
    c.o_dimg = o; o += c.dimg_floats;
    c.o_IN = o; o += P * c.sIN;
    c.o_ACT = o; o += c.nACT * P * c.sACT;
    c.o_OUT = o; o += P * c.sOUT;
    c.o_H = o; o += P * c.sH;
    c.o_LX = o; o += P * c.sD;
    c.o_XI = o; o += P * c.sD;
    c.o_YBJ = o; o += P * c.sDO;
    c.o_YY = o; o += P * c.sDO;
    c.o_XH = o; o += P * c.sH;
    c.o_EE = o; o += P * c.sH;
    c.o_KS = o; o += NJ_RL_SCRATCH;
    c.o_GI = o; if (c.use_rnn) o += P * c.s3H;
    c.o_GHH = o; if (c.use_rnn) o += P * c.s3H;
    c.o_GOUT = c.o_GTMP = c.o_GA = c.o_GB = c.o_GH = c.o_GX = c.o_GYBJ = o;
    if (bwd) {
        c.o_GOUT = o; o += P * c.sOUT;
        c.o_GTMP = o; o += P * c.sOUT;
        c.o_GA = o; o += P * c.sG;
        c.o_GB = o; o += P * c.sG;
        c.o_GH = o; o += P * c.sH;
        c.o_GX = o; o += P * c.sD;
        c.o_GYBJ = o; o += P * c.sDO;
    }
    c.o_F = o; o += NJ_F_COUNT * P;
    o = (o + 3) & ~3;
    c.o_I = o; o += NJ_I_COUNT * P + NJ_CTL_COUNT;
    o = (o + 3) & ~3;
    if (bwd) c.smem_floats_bwd = o; else c.smem_floats_fwd = o;
}

static inline bool nj_make_cfg(const njode_model_t& m, NjCfg& c, bool compact, std::string& err) {
    memset(&c, 0, sizeof(c));
    if (!nj_fill_nets(m, c, compact, err)) return false;
    c.compact = compact ? 1 : 0;
    c.d = m.input_size; c.H = m.hidden_size; c.dout = m.output_size;
    c.masked = m.masked; c.curt = m.input_current_t; c.loss_kind = m.loss_kind; c.residual = m.residual;
    c.training = m.training; c.use_rnn = m.use_rnn ? 1 : 0;
    c.inf = c.net[NJODE_NET_ODE].dim[0]; c.enc_in = c.net[NJODE_NET_ENC].dim[0];
    const NjNet &O = c.net[NJODE_NET_ODE], &E = c.net[NJODE_NET_ENC], &R = c.net[NJODE_NET_RO];
    if (c.inf != c.d + c.H + 2 + (c.curt ? 1 : 0)) { err = "ode_f input width != input+hidden+2(+1)"; return false; }
    if (O.dim[O.n] != c.H) { err = "ode_f output width != hidden_size"; return false; }
    if (c.enc_in != c.d * (c.masked ? 2 : 1)) { err = "encoder input width mismatch"; return false; }
    if (E.dim[E.n] != c.H) { err = "encoder output width != hidden_size"; return false; }
    if (R.dim[0] != c.H || R.dim[R.n] != c.dout) { err = "readout widths mismatch"; return false; }
    if (c.dout != c.d) { err = "output_size must equal input_size (loss compares X with Y)"; return false; }
    if (c.use_rnn) {
        // torch.nn.GRUCell(input_size, hidden_size): weight_ih [3H, d], weight_hh [3H, H]  (NJODE/models.py:208)
        const NjNet &GI = c.net[NJODE_NET_GRU_IH], &GH = c.net[NJODE_NET_GRU_HH];
        if (GI.n != 1 || GH.n != 1 || GI.dim[0] != c.d || GH.dim[0] != c.H || GI.dim[1] != 3 * c.H || GH.dim[1] != 3 * c.H) {
            err = "use_rnn: GRU maps must be single Linear layers input->3*hidden and hidden->3*hidden"; return false;
        }
    }
    if (c.residual) {
        // FFNN.__init__ residual cases, NJODE/models.py:240-257
        if ((c.d <= c.H && c.H % c.d) || (c.d > c.H && c.d % c.H)) { err = "for residual: encoder sizes must be multiples"; return false; }
        if ((c.H <= c.dout && c.dout % c.H) || (c.H > c.dout && c.H % c.dout)) { err = "for residual: readout sizes must be multiples"; return false; }
    }
    c.m_H = nj_magic_host(c.H); c.m_d = nj_magic_host(c.d); c.m_dout = nj_magic_host(c.dout);
    c.m_inf = nj_magic_host(c.inf); c.m_dH = nj_magic_host(c.d + c.H); c.m_3H = nj_magic_host(3 * c.H);
    c.w = m.weight;
    const float p = m.dropout_p;
    c.has_drop = (m.training && p > 0.f) ? 1 : 0;
    c.one_minus_p = 1.f - p;
    c.keep_scale = (p < 1.f) ? 1.f / (1.f - p) : 0.f;
    double thr = (double)p * 65536.0;
    c.thr = thr >= 65536.0 ? 65536u : (unsigned)thr;          // 16-bit fields, see nj_keep
    c.seed_lo = (unsigned)(m.dropout_seed & 0xFFFFFFFFull); c.seed_hi = (unsigned)(m.dropout_seed >> 32);
    int maxhid = 1, maxin = 1, maxn = 1;
    for (int n = 0; n < NJODE_NUM_NETS; ++n) {
        const NjNet& N = c.net[n];
        maxn = std::max(maxn, N.n);
        for (int l = 0; l < N.n; ++l) maxin = std::max(maxin, N.dim[l]);
        for (int l = 1; l < N.n; ++l) maxhid = std::max(maxhid, N.dim[l]);
    }
    c.sIN = nj_stride_host(std::max(std::max(c.inf, c.enc_in), c.H));
    c.sACT = nj_stride_host(maxhid);
    c.nACT = std::max(1, maxn - 1);
    c.sOUT = nj_stride_host(std::max(std::max(c.H, c.dout), c.use_rnn ? 3 * c.H : 0));
    c.s3H = nj_stride_host(3 * c.H);
    c.sH = nj_stride_host(c.H); c.sD = nj_stride_host(c.d); c.sDO = nj_stride_host(c.dout);
    c.sG = nj_stride_host(maxin);
    return true;
}

// choose tile size / residency so that both kernels fit `smem_limit` bytes per CTA.
//   * 256 threads per CTA whatever the tile height: the lockstep march is latency bound (dependent Euler steps), so
//     every layer GEMM wants as many threads as it has 1x4 micro-tiles;
//   * tile height P: the smallest that still covers the units in whole waves of one CTA per SM (a whole-path batch of
//     2000 records -> P = 14, 143 CTAs; the reference's PhysioNet batch of 50 -> P = 1), at most 64;
//   * residency: parameter image (and, if it also fits, the gradient image) in shared memory is worth a smaller P.
static inline bool nj_make_plan(const njode_model_t& m, int n_units_fwd, int n_units_bwd, int N_rows,
                                int num_sms, size_t smem_limit, int force_P, bool compact, NjPlanOut& out, std::string& err) {
    NjCfg base;
    if (!nj_make_cfg(m, base, compact, err)) return false;
    int P_want = 64;
    if (n_units_fwd < 64 * num_sms) {
        const int waves = std::max(1, (n_units_fwd + 64 * num_sms - 1) / (64 * num_sms));
        P_want = std::max(1, std::min(64, (n_units_fwd + waves * num_sms - 1) / (waves * num_sms)));
    } else {
        const int tiles64 = (n_units_fwd + 63) / 64;
        const int waves = (tiles64 + num_sms - 1) / num_sms;
        P_want = std::max(1, std::min(64, (n_units_fwd + waves * num_sms - 1) / (waves * num_sms)));
    }
    if (force_P > 0) P_want = force_P;
    // residency options of one kernel at tile height P, best first; dw: 1 = whole gradient image in shared memory,
    // 2 = the ODE network's part only (the one every Euler step accumulates into), 0 = per-CTA partial in global memory
    auto place = [&](int P, bool need_w) {
        // 256 threads whatever the tile height.  (Measured on B200: 1024-thread CTAs for the small tiles are 2-3x
        // SLOWER -- the per-warp fixed cost of the ~12 barrier-separated phases of an Euler step dominates, not the
        // arithmetic, so more warps means more issue slots spent on loop/phase boilerplate.)
        const int nt = 256;
        out.fwd = base; out.bwd = base;
        bool okf = false, okb = false;
        for (int w = 1; w >= (need_w ? 1 : 0) && !okf; --w) {
            nj_layout(out.fwd, P, nt, false, w != 0, 0);
            okf = (size_t)out.fwd.smem_floats_fwd * 4 <= smem_limit;
        }
        static const int opts[4][2] = {{1, 1}, {1, 2}, {1, 0}, {0, 0}};
        const char* fdw = getenv("NJODE_FORCE_DW");          // tests: pin the gradient-image residency
        for (int k = 0; k < (need_w ? 3 : 4) && !okb; ++k) {
            if (fdw && opts[k][0] && opts[k][1] != atoi(fdw)) continue;
            nj_layout(out.bwd, P, nt, true, opts[k][0] != 0, opts[k][1]);
            okb = (size_t)out.bwd.smem_floats_bwd * 4 <= smem_limit;
        }
        return okf && okb;
    };
    bool ok = false;
    // the parameter image in shared memory is worth a tile of half the height
    const int cands[3] = {P_want, (3 * P_want) / 4, P_want / 2};
    for (int i = 0; i < (force_P > 0 ? 1 : 3) && !ok; ++i)
        if (cands[i] >= 1 && (i == 0 || cands[i] != cands[i - 1])) ok = place(cands[i], true);
    for (int P = P_want; !ok && P >= 1; P /= 2) ok = place(P, false);
    if (!ok) { err = "model too wide for the shared-memory tile kernels"; return false; }
    out.smem_fwd_bytes = (size_t)out.fwd.smem_floats_fwd * 4;
    out.smem_bwd_bytes = (size_t)out.bwd.smem_floats_bwd * 4;
    const int P_ = out.fwd.P;
    out.n_tiles = (n_units_fwd + P_ - 1) / P_;
    const int tiles_b = (n_units_bwd + P_ - 1) / P_;
    auto per_sm = [&](size_t bytes, int nt) {
        int k = (int)((smem_limit + 1024) / (bytes + 1024));
        (void)nt;
        return std::max(1, std::min(k, 2));          // 128 registers x 256 threads: two CTAs per SM
    };
    out.grid_fwd = std::max(1, std::min(out.n_tiles, num_sms * per_sm(out.smem_fwd_bytes, out.fwd.nt)));
    out.grid_bwd = std::max(1, std::min(tiles_b, num_sms * per_sm(out.smem_bwd_bytes, out.bwd.nt)));
    size_t o = 0;
    out.ws_image_off = o; o += (size_t)base.img_floats * 4; o = (o + 255) & ~(size_t)255;
    out.ws_rowloss_off = o; o += (size_t)std::max(N_rows, 1) * 4; o = (o + 255) & ~(size_t)255;
    out.ws_counter_off = o; o += 256;
    out.ws_partials_off = o; o += (size_t)out.grid_bwd * base.img_floats * 4;
    out.ws_bytes = o;
    return true;
}


// ------------------------------------------------------------------------------------------------
// segment fast path (njode_seg.cuh): eligibility + shared-memory layout
// ------------------------------------------------------------------------------------------------
// whole-path batches of at most this many paths per SM take the thread-per-neuron kernels, one path per tile (measured on B200)
// (B200, PhysioNet nets, ms per pass: thread per neuron 0.025 / 0.058 per path forward / backward, pipelined warps
// 17 + 0.0048 / 36 + 0.044 per path: the forward crosses over at ~5.7 paths per SM, the backward at ~17; profiles/r2af_*)
#ifndef NJ_TPN_MAX_WAVES
#define NJ_TPN_MAX_WAVES 16
#endif
#ifndef NJ_TPN_MAX_WAVES_FWD
#define NJ_TPN_MAX_WAVES_FWD 5
#endif
// segment batches of at most this many units per SM take the thread-per-neuron kernels (measured on B200)
#ifndef NJ_SEGTPN_MAX_UNITS_PER_SM
#define NJ_SEGTPN_MAX_UNITS_PER_SM 32
#endif

static inline int nj_seg_fwd_region(const NjCfg& c, NjSeg& s, int R) {
    (void)c;
    int o = 0;
    s.f_IN = o; o += R * s.sI;
    s.f_A0 = o; o += R * s.sA;
    s.f_A1 = o; o += R * s.sA;
    s.f_OUT = o; o += R * s.sO;
    s.f_HS = o; o += R * s.sH;
    s.f_LX = o; o += R * s.sD;
    s.f_TX = o; o += R * s.sD;
    s.f_XI = o; o += R * s.sD;
    s.f_YBJ = o; o += R * s.sD;
    // the per-row scalar slots keep the stride of the tallest tile (RS = 16 in nj_seg_forward_warp) whatever R is
    s.f_F = o; o += NJS_F_COUNT * 16;
    s.f_I = o; o += NJS_I_COUNT * 16 + 4;
    return (o + 3) & ~3;
}

static inline int nj_seg_bwd_layout(const NjCfg& c, NjSeg& s, int P, int sets = 1) {
    int o = c.img_floats;
    s.b_img = 0;
    s.b_IN = o; o += P * s.sI;
    s.b_A = o; o += s.nA * P * s.sA;
    s.b_G = o; o += s.nA * P * s.sA;
    s.b_GOUT = o; o += P * s.sO;
    s.b_copy = o - s.b_IN; o += (sets - 1) * s.b_copy;       // thread-per-neuron backward: the operand buffers three times
    s.b_GZ = o; o += P * s.sI;
    s.b_OUT = o; o += P * s.sO;
    s.b_GH = o; o += P * s.sH;
    s.b_HB = o; o += P * s.sH;
    s.b_EE = o; o += P * s.sH;
    s.b_GE = o; o += P * s.sH;
    s.b_XI = o; o += P * s.sD;
    s.b_LX = o; o += P * s.sD;
    s.b_TX = o; o += P * s.sD;
    s.b_YBJ = o; o += P * s.sD;
    s.b_YY = o; o += P * s.sD;
    s.b_GYBJ = o; o += P * s.sD;
    s.b_F = o; o += NJS_F_COUNT * P;
    s.b_I = o; o += NJS_I_COUNT * P + 4;
    return (o + 3) & ~3;
}

// classes of tile heights over the two sorted unit runs (loss units, tail units); see NjSeg
static inline int nj_seg_classes(const int run_b[2], const int run_e[2], const int n1[2], const int n2[2],
                                 const int* trs, int ntr, int rows_per_tr, int* t0, int* u0, int* u1, int* trc) {
    // trs: tile heights from lowest to tallest, e.g. {1, 2, 4}; cut points: n1 (-> trs[0]), n2 (-> trs[1]), rest
    int ncls = 0, tiles = 0;
    for (int k = 0; k < ntr; ++k) {
        for (int r = 0; r < 2; ++r) {
            const int len = run_e[r] - run_b[r];
            int cut[4] = {0, std::min(n1[r], len), std::min(std::max(n2[r], n1[r]), len), len};
            int lo, hi;
            if (ntr == 3) { lo = cut[k]; hi = cut[k + 1]; }
            else if (ntr == 2) { lo = k == 0 ? 0 : cut[1]; hi = k == 0 ? cut[1] : len; }
            else { lo = 0; hi = len; }
            if (hi <= lo) continue;
            const int rows = rows_per_tr * trs[k];
            t0[ncls] = tiles; u0[ncls] = run_b[r] + lo; u1[ncls] = run_b[r] + hi; trc[ncls] = trs[k];
            tiles += (hi - lo + rows - 1) / rows;
            ++ncls;
        }
    }
    t0[ncls] = tiles;
    return ncls;
}

static inline void nj_make_seg(const NjCfg& c, const njode_batch_t& b, int num_sms, size_t smem_limit, NjPlanOut& out) {
    NjSeg& s = out.seg;
    memset(&s, 0, sizeof(s));
    const int n_units = b.n_units;
    const char* off = getenv("NJODE_NO_SEG");
    if (off && atoi(off)) return;
    if (b.unit_kind != 1 || c.masked || c.use_rnn || b.E > 0 || n_units <= 0) return;
    int maxhid = 1, maxn = 1, maxlast = 1;
    for (int n = 0; n < 3; ++n) {
        const NjNet& N = c.net[n];
        maxn = std::max(maxn, N.n);
        for (int l = 0; l < N.n - 1; ++l) maxhid = std::max(maxhid, N.rp[l]);
        maxlast = std::max(maxlast, N.rp[N.n - 1]);
    }
    s.sI = nj_stride_act(std::max(std::max(c.inf, c.enc_in), c.H));
    s.sA = nj_stride_act(maxhid);
    s.sO = nj_stride_act(maxlast);
    s.sH = (c.H + 3) & ~3; s.sD = (std::max(c.d, c.dout) + 3) & ~3;
    s.nA = std::max(1, maxn - 1);
    // dW tiles: order ODE, RO, ENC
    static const int order[3] = {NJODE_NET_ODE, NJODE_NET_RO, NJODE_NET_ENC};
    int tiles = 0;
    for (int oi = 0; oi < 3; ++oi) {
        const NjNet& N = c.net[order[oi]];
        for (int l = 0; l < NJODE_MAX_LINEAR; ++l) {
            s.tile_base[order[oi]][l] = tiles;
            if (l < N.n) tiles += ((N.dim[l] + 3) / 4) * ((N.dim[l + 1] + 3) / 4);
        }
    }
    s.tiles_total = tiles;
    // ---- small batches: thread-per-neuron kernels (njode_tpn.cuh, nj_segtpn_*): tiles of 4 segments, one per CTA at a time ----
    {
        const NjNet& O = c.net[NJODE_NET_ODE];
        const char* fs = getenv("NJODE_SEG_TPN");             // 0: never, 1: whatever the batch size (tests)
        const int want = fs ? atoi(fs) : -1;
        auto fits = [&](int kc0, int kch, int hc) {
            return O.dim[0] <= 4 * kc0 && O.dim[1] <= 4 * kch && O.dim[2] <= 4 * kch && O.dim[3] <= 4 * hc && c.H <= 4 * hc &&
                   c.d + 4 * hc <= NJN_T && O.dim[0] <= NJN_T && c.inf <= 2 * NJN_F;
        };
        int cls = 0;
        if (want != 0 && O.n == 3) cls = fits(NJN_A_KC0, NJN_A_KCH, NJN_A_HC) ? 1 : (fits(NJN_B_KC0, NJN_B_KCH, NJN_B_HC) ? 2 : 0);
        if (cls && want < 0 && n_units > NJ_SEGTPN_MAX_UNITS_PER_SM * num_sms) cls = 0;
        if (cls && tiles - s.tile_base[NJODE_NET_RO][0] >= 0 && s.tile_base[NJODE_NET_RO][0] > NJN_D * NJN_DSLOTS) cls = 0;
        if (cls) {
            NjSeg keep = s;
            const int kc0 = cls == 1 ? NJN_A_KC0 : NJN_B_KC0, kch = cls == 1 ? NJN_A_KCH : NJN_B_KCH, hc = cls == 1 ? NJN_A_HC : NJN_B_HC;
            s.tpn = cls;
            s.sI = std::max(s.sI, nj_stride_act(4 * kc0));
            s.sA = std::max(s.sA, nj_stride_act(4 * kch));
            s.sO = std::max(s.sO, nj_stride_act(4 * hc));
            const int n_loss_ = std::max(0, std::min(b.n_loss_units, n_units));
            const int rb[2] = {0, n_loss_}, re[2] = {n_loss_, n_units};
            const int z1[2] = {0, 0}, z2[2] = {0, 0};
            const int one[1] = {1};
            s.f_region = nj_seg_fwd_region(c, s, NJN_SEG_R);
            s.f_MB = s.f_region; s.f_region += (int)((sizeof(NjCoopMB) + 15) / 16) * 4;       // mailbox of the cooperative layers
            s.f_img = 0; s.f_warp0 = c.img_floats;
            s.f_ncls = nj_seg_classes(rb, re, z1, z2, one, 1, 4, s.f_t0, s.f_u0, s.f_u1, s.f_tr);
            s.n_tiles_f = s.f_t0[s.f_ncls];
            s.nw_f = 1;
            s.f_smem_floats = c.img_floats + s.f_region;
            int fl = nj_seg_bwd_layout(c, s, NJN_SEG_R, 3);
            s.b_TD = fl; fl += 2 * NJN_DSLOTS * NJN_D;
            s.b_PRE = fl; fl += 16 + 2 * NJN_SEG_R * s.sH;
            fl = (fl + 3) & ~3;
            s.b_GIMG = fl; fl += c.img_floats - c.net[NJODE_NET_ENC].w_img[0];
            fl = (fl + 3) & ~3;
            s.b_MB = fl; fl += (int)((sizeof(NjCoopMB) + 15) / 16) * 4;
            s.b_smem_floats = (fl + 3) & ~3; s.P_b = NJN_SEG_R; s.nw_b = 1; s.nt_b = NJN_NT_BWD; s.nt_slots = 0;
            s.b_ncls = nj_seg_classes(rb, re, z1, z2, one, 1, 4, s.b_t0, s.b_u0, s.b_u1, s.b_tr);
            s.n_tiles_b = s.b_t0[s.b_ncls];
            if ((size_t)s.f_smem_floats * 4 <= smem_limit && (size_t)s.b_smem_floats * 4 <= smem_limit) {
                out.seg_smem_f_bytes = (size_t)s.f_smem_floats * 4;
                out.seg_smem_b_bytes = (size_t)s.b_smem_floats * 4;
                // forward CTAs are 3 warps with ~215 registers: three per SM, tiles handed out by an atomic counter
                out.seg_grid_f = std::max(1, std::min(s.n_tiles_f, 3 * num_sms));
                out.seg_grid_b = std::max(1, std::min(s.n_tiles_b, num_sms));
                s.ok = 1;
                return;
            }
            s = keep;                                         // does not fit: the warp kernels below
        }
    }
    const char* ftr = getenv("NJODE_FORCE_TR");
    const int force_tr = ftr ? atoi(ftr) : 0;
    const char* fnw = getenv("NJODE_FORCE_NW");
    const int force_nw = fnw ? atoi(fnw) : 0;
    const int n_loss = std::max(0, std::min(b.n_loss_units, n_units));
    const int run_b[2] = {0, n_loss}, run_e[2] = {n_loss, n_units};
    int n1[2] = {b.seg_n1[0], b.seg_n1[1]}, n2[2] = {b.seg_n2[0], b.seg_n2[1]};
    // ---- forward: per-warp regions sized for the tallest tile (16 rows) ----
    {
        s.f_region = nj_seg_fwd_region(c, s, 16);
        s.f_img = 0; s.f_warp0 = c.img_floats;
        auto warps_that_fit = [&]() {
            for (int cand = 12; cand >= 2; --cand)          // launch bounds: 384 threads
                if ((size_t)(c.img_floats + cand * s.f_region) * 4 <= smem_limit) return cand;
            return 0;
        };
        int nw = warps_that_fit();
        // big nets (2x100: the image leaves room for three 16-row regions): regions of 8 rows and tiles of at most
        // 8 rows instead -- twice the warps to hide the latency of the warp-autonomous marches
        bool low = false;
        const char* flow = getenv("NJODE_SEG_LOW");           // tests: take the 8-row layout whatever fits
        if (!force_tr && (nw < 6 || (flow && atoi(flow)))) {
            s.f_region = nj_seg_fwd_region(c, s, 8);
            const int nw8 = warps_that_fit();
            if ((nw8 >= 2 * nw || (flow && atoi(flow))) && nw8 >= 2) { nw = nw8; low = true; }
            else s.f_region = nj_seg_fwd_region(c, s, 16);
        }
        static const int trs3[3] = {1, 2, 4};
        static const int trs2f[2] = {1, 2};
        if (force_tr) { const int one[1] = {std::min(4, force_tr)}; s.f_ncls = nj_seg_classes(run_b, run_e, n1, n2, one, 1, 4, s.f_t0, s.f_u0, s.f_u1, s.f_tr); }
        else if (low) s.f_ncls = nj_seg_classes(run_b, run_e, n1, n2, trs2f, 2, 4, s.f_t0, s.f_u0, s.f_u1, s.f_tr);
        else s.f_ncls = nj_seg_classes(run_b, run_e, n1, n2, trs3, 3, 4, s.f_t0, s.f_u0, s.f_u1, s.f_tr);
        s.n_tiles_f = s.f_t0[s.f_ncls];
        if (!nw) return;
        // small batches: fewer warps per CTA so that every SM gets work
        nw = std::max(2, std::min(nw, (s.n_tiles_f + num_sms - 1) / num_sms));
        if (force_nw) nw = std::max(2, std::min(12, force_nw));
        s.nw_f = nw;
        s.f_smem_floats = c.img_floats + nw * s.f_region;

    }
    // ---- backward: CTA-level arrays of P = 8 * nw rows (tallest tile) ----
    {
        // small batches: fewer warps (rows) per CTA so that every SM gets a tile; dW tiles beyond the register
        // capacity of the smaller CTA go through the partial image in global memory (nj_seg_dw)
        int want = std::max(4, std::min(12, (n_units + 8 * num_sms - 1) / (8 * num_sms)));
        if (force_nw) want = std::max(2, std::min(12, force_nw));
        int nw = 0;
        for (int cand = want; cand >= 2; --cand) {       // launch bounds: 384 threads
            const int fl = nj_seg_bwd_layout(c, s, 8 * cand);
            if ((size_t)fl * 4 <= smem_limit) { nw = cand; s.b_smem_floats = fl; s.P_b = 8 * cand; break; }
        }
        if (!nw) return;
        s.nw_b = nw;
        // dW tiles beyond NJ_SEG_NT_MAX per thread go through the L2-resident partial image at every step.  Nets with
        // many more tiles than a CTA has register slots (2x100 nets: 2250 tiles; the image leaves room for 24 rows =
        // 3 row warps = 192 slots) get helper warps that own no rows and only join the dW phases, up to the 12 warps of
        // the launch bounds (B200, 2x100 nets: backward 11.5 -> 7.7 ms; the generic backward takes 9.6 ms).  For the
        // demo nets (702 tiles) helpers measured neutral to 6 % slower, so small batches of small nets get none.
        const char* fh = getenv("NJODE_SEG_HELPERS");
        const bool helpers = fh ? atoi(fh) != 0 : tiles > 1024;
        int wtot = nw;
        if (helpers) wtot = std::max(nw, std::min(12, (tiles + 32 * NJ_SEG_NT_MAX - 1) / (32 * NJ_SEG_NT_MAX)));
        // with spare warps in the launch the rows of the CTA are spread over twice as many row warps (4 rows each
        // instead of 8): the warp-local phases, during which the helpers wait, take half as long
        const bool thin = helpers && !force_tr && 2 * nw <= wtot;
        if (thin) { nw *= 2; s.nw_b = nw; }
        s.nt_b = 32 * wtot;
        s.nt_slots = std::min(NJ_SEG_NT_MAX, (tiles + s.nt_b - 1) / s.nt_b);
        static const int trs2[2] = {1, 2};
        if (thin) { const int one[1] = {1}; s.b_ncls = nj_seg_classes(run_b, run_e, n1, n2, one, 1, 4 * nw, s.b_t0, s.b_u0, s.b_u1, s.b_tr); }
        else if (force_tr) { const int one[1] = {std::min(2, force_tr)}; s.b_ncls = nj_seg_classes(run_b, run_e, n1, n2, one, 1, 4 * nw, s.b_t0, s.b_u0, s.b_u1, s.b_tr); }
        else s.b_ncls = nj_seg_classes(run_b, run_e, n1, n2, trs2, 2, 4 * nw, s.b_t0, s.b_u0, s.b_u1, s.b_tr);
        s.n_tiles_b = s.b_t0[s.b_ncls];
    }
    out.seg_smem_f_bytes = (size_t)s.f_smem_floats * 4;
    out.seg_smem_b_bytes = (size_t)s.b_smem_floats * 4;
    out.seg_grid_f = std::max(1, std::min((s.n_tiles_f + s.nw_f - 1) / s.nw_f, num_sms));
    out.seg_grid_b = std::max(1, std::min(s.n_tiles_b, num_sms));
    s.ok = 1;
}

// ------------------------------------------------------------------------------------------------
// whole-path units on the warp GEMMs (njode_path.cuh): eligibility, tile shapes, shared-memory layout
// ------------------------------------------------------------------------------------------------
static inline void nj_make_path(const NjCfg& c, const njode_batch_t& b, int num_sms, size_t smem_limit, NjPlanOut& out) {
    NjPath& s = out.path;
    memset(&s, 0, sizeof(s));
    const char* off = getenv("NJODE_NO_PATH");
    if (off && atoi(off)) return;
    const int n = b.n_units;
    if (b.unit_kind != 0 || n <= 0 || c.compact) return;
    int maxhid = 1, maxn = 1, maxlast = 1;
    const int nnets = c.use_rnn ? NJODE_NUM_NETS : 3;
    for (int q = 0; q < nnets; ++q) {
        const NjNet& N = c.net[q];
        maxn = std::max(maxn, N.n);
        for (int l = 0; l < N.n - 1; ++l) maxhid = std::max(maxhid, N.rp[l]);
        maxlast = std::max(maxlast, N.rp[N.n - 1]);
    }
    s.sI = nj_stride_act(std::max(std::max(c.inf, c.enc_in), std::max(c.H, c.d)));
    s.sA = nj_stride_act(maxhid);
    s.sO = nj_stride_act(maxlast);
    s.sH = (c.H + 3) & ~3; s.sD = (std::max(c.d, c.dout) + 3) & ~3; s.s3 = (3 * c.H + 3) & ~3;
    s.nA = std::max(1, maxn - 1);
    const char* frw = getenv("NJODE_PATH_R");                 // tests: rows per warp (1, 2, 4, 8)
    const int force_r = frw ? atoi(frw) : 0;
    auto shape = [](int R, int& rg, int& tr) { rg = R >= 4 ? 4 : R; tr = R >= 8 ? 2 : 1; };
    // ---- weight-stationary Euler steps: batches of at most 8 paths per SM whose ODE network fits the register tiles ----
    int stat_R = 0;
    {
        const NjNet& O = c.net[NJODE_NET_ODE];
        const char* ns = getenv("NJODE_NO_STAT");
        bool ok = !(ns && atoi(ns)) && (O.n == 2 || O.n == 3);
        int maxo = std::max(c.H, c.masked ? c.d : 0);
        for (int l = 0; l < O.n && ok; ++l) {
            if (O.dim[l] > (l == 0 ? 96 : 64)) ok = false;      // register slices: 8 x 12 floats (layer 0), 8 x 8 (the others)
            maxo = std::max(maxo, O.dim[l + 1]);
        }
        const int nw = std::max(4, (maxo + 3) / 4);
        if (nw > 13) ok = false;                              // launch bounds: 416 threads
        if (ok) {
            // one path per CTA only: with several rows per tile the 12-warp pipelined kernels are faster (B200, PhysioNet
            // nets: 300 records 136 ms here against 67 ms there, 600 records 342 against 85; profiles/r2z2_*)
            if (n <= num_sms) stat_R = 1;
            const char* fs = getenv("NJODE_FORCE_STAT");      // tests: take the stationary kernels whatever the batch size
            if (fs && atoi(fs)) {
                for (int cand = 1; cand <= 8 && !stat_R; cand *= 2)
                    if ((n + cand - 1) / cand <= num_sms) stat_R = cand;
                if (!stat_R) stat_R = 8;
            }
            if (force_r && (stat_R || (fs && atoi(fs)))) stat_R = force_r;
        }
        if (stat_R) { s.stat = 1; s.nw_s = nw; }
    }
    // ---- thread-per-neuron kernels (njode_tpn.cuh): tiles of 1 or 4 paths, one per SM, ODE network of a known dimension class ----
    int tpn_R = 0;
    bool tpn_fwd = true;
    {
        const NjNet& O = c.net[NJODE_NET_ODE];
        const char* nt_ = getenv("NJODE_NO_TPN");
        const char* ft_ = getenv("NJODE_FORCE_TPN");           // tests: whatever the batch size
        auto fits = [&](int kc0, int kch, int hc) {
            return O.dim[0] <= 4 * kc0 && O.dim[1] <= 4 * kch && O.dim[2] <= 4 * kch && O.dim[3] <= 4 * hc && c.H <= 4 * hc &&
                   c.d + 4 * hc <= NJN_T && O.dim[0] <= NJN_T && c.inf <= 2 * NJN_F;
        };
        int cls = 0;
        if (!(nt_ && atoi(nt_)) && O.n == 3) cls = fits(NJN_A_KC0, NJN_A_KCH, NJN_A_HC) ? 1 : (fits(NJN_B_KC0, NJN_B_KCH, NJN_B_HC) ? 2 : 0);
        if (cls) {
            // tiles of 4 paths exist (tests, NJODE_FORCE_TPN) but lose to the pipelined warp kernels on B200 (300 PhysioNet
            // records: 80 against 67 ms; 500 demo paths with the GRU jump: 3.9 against 3.0 ms): one path per CTA only
            // Several paths per SM are served as several one-path tiles per CTA in sequence (atomic tile counter), up to
            // NJ_TPN_MAX_WAVES tiles per CTA (env NJODE_TPN_WAVES): beyond that the 12-warp pipelined kernels win.
            const char* tw_ = getenv("NJODE_TPN_WAVES");
            const char* twf_ = getenv("NJODE_TPN_WAVES_FWD");
            // (the demo networks (class A) cross over at ~5 paths per SM in both passes: GRU jump, 500 paths 2.28 against
            // 3.02 ms, extrapolated crossover at ~730 paths)
            const int waves = tw_ ? atoi(tw_) : (cls == 1 ? NJ_TPN_MAX_WAVES_FWD : NJ_TPN_MAX_WAVES);
            // (class B forward, two CTAs per SM since only the jump networks' image is staged: 0.0185 ms per path, crossover at ~8)
            const int waves_f = twf_ ? atoi(twf_) : (tw_ ? atoi(tw_) : (cls == 1 ? NJ_TPN_MAX_WAVES_FWD : 8));
            // (calls that record the path -- evaluation, E > 0 -- run a readout on the glue warp after every step: there the
            // warp kernels win beyond one path per SM: 500 demo paths 1.87 against 1.43 ms, 100 paths 1.02 against 1.36)
            const int w_b = b.E > 0 ? 1 : std::max(1, waves), w_f = b.E > 0 ? 1 : std::max(1, waves_f);
            if (n <= w_b * num_sms) { tpn_R = 1; tpn_fwd = n <= w_f * num_sms; }
            else if (ft_ && atoi(ft_)) tpn_R = 4;
            if (force_r && (tpn_R || (ft_ && atoi(ft_)))) tpn_R = (force_r == 1 || force_r == 4) ? force_r : 0;
        }
        if (tpn_R) {
            s.tpn = cls; s.tpn_fwd = tpn_fwd ? 1 : 0; s.stat = 0; s.nw_s = 0; stat_R = 0;
            const int kc0 = cls == 1 ? NJN_A_KC0 : NJN_B_KC0, kch = cls == 1 ? NJN_A_KCH : NJN_B_KCH, hc = cls == 1 ? NJN_A_HC : NJN_B_HC;
            // the register tiles read whole classes of columns: rows at least that wide (the padding stays zero)
            s.sI = std::max(s.sI, nj_stride_act(4 * kc0));
            s.sA = std::max(s.sA, nj_stride_act(4 * kch));
            s.sO = std::max(s.sO, nj_stride_act(4 * hc));
        }
    }
    // dW tiles of the thread-owned 4x4 scheme; the stationary kernels keep the ODE network's gradient in their own
    // register tiles, so the ODE network has no tiles there
    static const int order[NJODE_NUM_NETS] = {NJODE_NET_ODE, NJODE_NET_RO, NJODE_NET_ENC, NJODE_NET_GRU_HH, NJODE_NET_GRU_IH};
    int tiles = 0;
    for (int oi = 0; oi < NJODE_NUM_NETS; ++oi) {
        const NjNet& N = c.net[order[oi]];
        for (int l = 0; l < NJODE_MAX_LINEAR; ++l) {
            if (s.stat && order[oi] == NJODE_NET_ODE) { s.tile_base[order[oi]][l] = 0x3FFFFFFF; continue; }
            s.tile_base[order[oi]][l] = tiles;
            if (l < N.n) tiles += ((N.dim[l] + 3) / 4) * ((N.dim[l + 1] + 3) / 4);
        }
    }
    s.tiles_total = tiles;
    // ---- forward: per-warp regions ----
    {
        auto region = [&](int R) {
            int o = 0;
            s.f_IN = o; o += R * s.sI;
            s.f_A0 = o; o += R * s.sA;
            s.f_A1 = o; o += R * s.sA;
            s.f_OUT = o; o += R * s.sO;
            s.f_HS = o; o += R * s.sH;
            s.f_EE = o; o += R * s.sH;
            s.f_LX = o; o += R * s.sD;
            s.f_TX = o; o += R * s.sD;
            s.f_XI = o; o += R * s.sD;
            s.f_YBJ = o; o += R * s.sD;
            s.f_YY = o; o += R * s.sD;
            s.f_MM = o; o += R * s.sD;
            s.f_GI = o; if (c.use_rnn) o += R * s.s3;
            s.f_F = o; o += NJP_F_COUNT * NJP_RS;
            s.f_I = o; o += NJP_I_COUNT * NJP_RS + 4;
            o = (o + 3) & ~3;
            s.f_MB = o; if (s.tpn && s.tpn_fwd) o += (int)((sizeof(NjCoopMB) + 15) / 16) * 4;    // mailbox of the cooperative jump layers
            s.f_IN2 = o; if (s.tpn && s.tpn_fwd) o += R * s.sI;                                // input rows of the other step parity
            s.f_AUX = o; if (s.tpn && s.tpn_fwd) o += 16;                                      // dropout layer keys [parity][layer][4]
            return (o + 3) & ~3;
        };
        // the smallest tile height whose warps still fit the machine in one wave of 12-warp CTAs: smaller tiles = more
        // warps to hide the latency of the dependent steps and, below 4 rows, a split reduction dimension
        int R = 8;
        for (int cand = 1; cand <= 8; cand *= 2)
            if ((n + cand - 1) / cand <= num_sms * 12) { R = cand; break; }
        if (force_r) R = force_r;
        if (s.stat) R = stat_R;
        if (s.tpn && s.tpn_fwd) R = tpn_R;
        shape(R, s.rg_f, s.tr_f);
        s.f_region = region(R);
        s.f_warp0 = c.img_floats;
        s.n_tiles_f = (n + R - 1) / R;
        int nw = 0;
        for (int cand = 12; cand >= 1; --cand)
            if ((size_t)(c.img_floats + cand * s.f_region) * 4 <= smem_limit) { nw = cand; break; }
        if (!nw) return;
        nw = std::max(1, std::min(nw, (s.n_tiles_f + num_sms - 1) / num_sms));
        if (s.stat || (s.tpn && s.tpn_fwd)) nw = 1;           // one tile per CTA at a time, all warps on it
        s.nw_f = nw;
        s.f_smem_floats = c.img_floats + nw * s.f_region;
        if (s.tpn && s.tpn_fwd) {
            // the thread-per-neuron forward holds the ODE network in registers: only the jump networks' part of the image is
            // staged in shared memory (PhysioNet nets: 112 -> 72 KB, so two of its 3-warp CTAs fit an SM)
            s.f_warp0 = c.img_floats - c.net[NJODE_NET_ENC].w_img[0];
            s.f_smem_floats = s.f_warp0 + s.f_region;
        }
        out.path_grid_f = std::max(1, std::min((s.n_tiles_f + nw - 1) / nw, num_sms));
        // thread-per-neuron forward CTAs are 3 warps with up to 255 registers: two fit an SM
        if (s.tpn && s.tpn_fwd) out.path_grid_f = std::max(1, std::min(s.n_tiles_f, 2 * num_sms));
        out.path_smem_f_bytes = (size_t)s.f_smem_floats * 4;
    }
    // ---- backward: CTA-level arrays of P rows ----
    {
        auto layout = [&](int P, int nw) {
            int o = c.img_floats;
            s.b_IN = o; o += P * s.sI;
            s.b_A = o; o += s.nA * P * s.sA;
            s.b_G = o; o += s.nA * P * s.sA;
            s.b_GOUT = o; o += P * s.sO;
            s.b_copy = o - s.b_IN;                            // pipelined backward: a second set of the four operand buffers
            if (s.pipe) o += s.b_copy;
            if (s.tpn) o += 2 * s.b_copy;                     // thread-per-neuron backward: three sets
            s.b_GZ = o; o += P * s.sI;
            s.b_OUT = o; o += P * s.sO;
            s.b_GH = o; o += P * s.sH;
            s.b_HB = o; o += P * s.sH;
            s.b_EE = o; o += P * s.sH;
            s.b_GE = o; o += P * s.sH;
            s.b_XI = o; o += P * s.sD;
            s.b_LX = o; o += P * s.sD;
            s.b_TX = o; o += P * s.sD;
            s.b_YBJ = o; o += P * s.sD;
            s.b_YY = o; o += P * s.sD;
            s.b_GYBJ = o; o += P * s.sD;
            s.b_GX = o; o += P * s.sD;
            s.b_MM = o; o += P * s.sD;
            s.b_GI = o; if (c.use_rnn) o += P * s.s3;
            s.b_GHH = o; if (c.use_rnn) o += P * s.s3;
            s.b_PART = o; if (s.stat) o += 2 * P * s.nw_s * NJT_PARTW;
            // thread-per-neuron backward: dW tile table [slots * owners][2], prefetch slots (time, step size, h) of the next step
            s.b_TD = o; if (s.tpn) o += 2 * NJN_DSLOTS * NJN_D;
            s.b_PRE = o; if (s.tpn) o += NJN_PRE_HDR + 2 * P * s.sH;
            o = (o + 3) & ~3;
            s.b_MB = o; if (s.tpn) o += (int)((sizeof(NjCoopMB) + 15) / 16) * 4;
            // gradient image of the jump networks (everything behind the ODE network; zeroed with the other buffers)
            s.b_GIMG = o; if (s.tpn) o += c.img_floats - c.net[NJODE_NET_ENC].w_img[0];
            s.b_F = o; o += NJP_F_COUNT * P;
            s.b_I = o; o += NJB_I_COUNT * P + 4 + nw + 4;
            return (o + 3) & ~3;
        };
        const int P_want = std::max(1, (n + num_sms - 1) / num_sms);
        int R = 8;
        for (int cand = 1; cand <= 8; cand *= 2)
            if ((P_want + cand - 1) / cand <= 12) { R = cand; break; }
        if (force_r) R = force_r;
        int nw = std::max(1, std::min(12, (P_want + R - 1) / R));
        if (s.stat) { R = stat_R; nw = 1; }
        if (s.tpn) { R = tpn_R; nw = 1; }
        // pipelined dW: enough helper warps to hold the ODE network's tiles in their registers
        const int ode_tiles = s.tile_base[NJODE_NET_RO][0];
        const int helpers_min = std::max(3, (ode_tiles + 32 * NJP_HSLOTS - 1) / (32 * NJP_HSLOTS));
        const char* np_ = getenv("NJODE_NO_PIPE");
        if (!s.stat && !s.tpn && !(np_ && atoi(np_)) && helpers_min <= 8) {
            // rows per warp so that the row warps leave room for the helpers
            int Rp = R;
            while ((P_want + Rp - 1) / Rp > 12 - helpers_min && Rp < 8 && !force_r) Rp *= 2;
            const int nwp = std::max(1, std::min(12 - helpers_min, (P_want + Rp - 1) / Rp));
            // worth it when the helpers' dW (16 FFMA per tile and row) is well hidden behind the row warps' step (five
            // layer passes over Rp rows): measured on B200, PhysioNet nets (ratio 0.3) backward 145 -> 128 ms, demo nets
            // with the GRU jump at 36 rows per CTA (ratio 0.7) 5.6 -> 6.3 ms
            double w_ode = 0;
            for (int l = 0; l < c.net[NJODE_NET_ODE].n; ++l) w_ode += (double)c.net[NJODE_NET_ODE].dim[l] * c.net[NJODE_NET_ODE].dim[l + 1];
            const double row_ffma = 5.0 * w_ode * Rp / 32.0;
            const double help_ffma = (double)ode_tiles * 16.0 * (Rp * nwp) / (32.0 * (12 - nwp));
            const char* fp_ = getenv("NJODE_FORCE_PIPE");
            if (help_ffma < 0.5 * row_ffma || (fp_ && atoi(fp_))) { s.pipe = 1; R = Rp; nw = nwp; }
        }
        int fl = 0;
        for (; nw >= 1; --nw) {
            fl = layout(R * nw, nw);
            if ((size_t)fl * 4 <= smem_limit) break;
        }
        if (nw < 1) return;
        shape(R, s.rg_b, s.tr_b);
        s.nw_b = nw; s.P_b = R * nw; s.b_smem_floats = fl;
        const int wtot = std::max(nw, std::min(12, (tiles + 32 * NJ_SEG_NT_MAX - 1) / (32 * NJ_SEG_NT_MAX)));
        s.nt_b = 32 * wtot;
        s.nt_slots = std::min(NJ_SEG_NT_MAX, (tiles + s.nt_b - 1) / s.nt_b);
        if (s.stat) { s.nt_b = 32 * s.nw_s; s.nt_slots = 0; }
        if (s.pipe) { s.nt_b = 32 * 12; s.nt_slots = 0; }
        if (s.tpn) {
            s.nt_b = NJN_NT_BWD; s.nt_slots = 0;
            if (ode_tiles > NJN_D * NJN_DSLOTS) return;
        }
        s.n_tiles_b = (n + s.P_b - 1) / s.P_b;
        out.path_grid_b = std::max(1, std::min(s.n_tiles_b, num_sms));
        out.path_smem_b_bytes = (size_t)s.b_smem_floats * 4;
    }
    s.ok = 1;
}

// the whole launch plan of one (model, batch) pair: the segment fast path when it serves the call (padded parameter
// image), else the generic kernels on the compact image
static inline bool nj_plan_all(const njode_model_t& m, const njode_batch_t& b, int num_sms, size_t smem_limit, int force_P,
                               NjPlanOut& out, std::string& err) {
    memset(&out.path, 0, sizeof(out.path));
    if (!nj_make_plan(m, b.n_units, b.n_units, b.N, num_sms, smem_limit, force_P, false, out, err)) return false;
    out.act_bytes = 0; out.act_nh = 0; out.act_wp = 0;
    nj_make_seg(out.fwd, b, num_sms, smem_limit, out);
    if (out.seg.ok) {
        const NjNet& O = out.fwd.net[NJODE_NET_ODE];
        if (!out.seg.tpn && O.n >= 2 && O.n <= 3 && b.S > 0) {
            int wmax = 0;
            for (int l = 1; l < O.n; ++l) wmax = std::max(wmax, O.dim[l]);
            out.act_nh = O.n - 1; out.act_wp = ((wmax + 3) / 4) * 4;
            out.act_bytes = (size_t)b.S * (size_t)b.B * out.act_nh * out.act_wp * sizeof(float);
        }
        return true;
    }
    nj_make_path(out.fwd, b, num_sms, smem_limit, out);
    if (out.path.ok) return true;
    if (!nj_make_plan(m, b.n_units, b.n_units, b.N, num_sms, smem_limit, force_P, true, out, err)) return false;
    memset(&out.seg, 0, sizeof(out.seg));
    memset(&out.path, 0, sizeof(out.path));
    return true;
}
