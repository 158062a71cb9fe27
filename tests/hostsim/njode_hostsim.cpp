// TEST INFRASTRUCTURE ONLY -- sequential host simulation of njode_b200/csrc/njode_core.cuh.
// The kernel source is written as barrier-separated phases over shared arrays, so compiling it with
// -DNJODE_HOST_SIM runs every "thread" of a CTA in turn between barriers.  This lets the CPU test
// suite exercise the exact index/stride/cursor/gradient logic of the CUDA kernels without a GPU.
// It exports the same C ABI as libnjode_b200.so but takes HOST pointers.  The product package never
// loads this library (njode_b200/_ext.py only ever opens libnjode_b200.so and refuses to run
// without a CUDA device).
#define NJODE_HOST_SIM 1
#include <stdlib.h>
#include <stdio.h>
#include <vector>
#include "../../njode_b200/csrc/njode_plan.h"

static thread_local std::string g_err;
extern "C" const char* njode_last_error(void) { return g_err.c_str(); }
extern "C" int njode_abi_version(void) { return NJODE_ABI_VERSION; }

// simulated SM count: 4 keeps several tiles per CTA in play; NJODE_SIM_SMS=148 reproduces the launch plans of a B200
static int sim_sms() { const char* e = getenv("NJODE_SIM_SMS"); return e ? atoi(e) : 4; }
#define kSimSMs sim_sms()
static const size_t kSimSmem = 227 * 1024;

static int plan_for(const njode_model_t* m, const njode_batch_t* b, NjPlanOut& out) {
    std::string err;
    const char* fp = getenv("NJODE_FORCE_TILE");
    if (!nj_plan_all(*m, *b, kSimSMs, kSimSmem, fp ? atoi(fp) : 0, out, err)) { g_err = err; return -3; }
    if (getenv("NJODE_DEBUG_PLAN"))
        fprintf(stderr, "[plan] units %d kind %d E %d -> seg %d (stat %d) path %d stat %d tpn %d nw_s %d pipe %d (fwd rg %d tr %d nw %d | bwd rg %d tr %d nw %d nt %d P %d slots %d tiles %d)\n",
                b->n_units, b->unit_kind, b->E, out.seg.ok, out.seg.tpn, out.path.ok, out.path.stat, out.path.tpn, out.path.nw_s, out.path.pipe, out.path.rg_f, out.path.tr_f, out.path.nw_f, out.path.rg_b,
                out.path.tr_b, out.path.nw_b, out.path.nt_b, out.path.P_b, out.path.nt_slots, out.path.tiles_total);
    const size_t cap = (size_t)kSimSMs * 2;
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

// test hook (host simulation only): which kernel family the planner picks for (model, batch) --
// bit 0 segment units, 1 their weight-stationary kernels, 2 whole-path warp kernels, 3 their weight-stationary kernels,
// 4 pipelined path backward, 5 thread-per-neuron backward, 6 thread-per-neuron forward too
extern "C" int njode_hostsim_plan_kind(const njode_model_t* model, const njode_batch_t* bs) {
    NjPlanOut o;
    if (int rc = plan_for(model, bs, o)) return rc;
    return (o.seg.ok ? 1 : 0) | (o.seg.ok && o.seg.tpn ? 2 : 0) | (o.path.ok ? 4 : 0) | (o.path.ok && o.path.stat ? 8 : 0) |
           (o.path.ok && o.path.pipe ? 16 : 0) | (o.path.ok && o.path.tpn ? 32 : 0) | (o.path.ok && o.path.tpn && o.path.tpn_fwd ? 64 : 0);
}

extern "C" int njode_plan(const njode_model_t* model, const njode_batch_t* bs, int, njode_plan_t* p) {
    NjPlanOut o;
    if (int rc = plan_for(model, bs, o)) return rc;
    p->tile_paths = o.fwd.P; p->threads = o.fwd.nt; p->grid_fwd = o.grid_fwd; p->grid_bwd = o.grid_bwd;
    p->weights_in_smem = o.fwd.w_smem; p->grads_in_smem = o.bwd.dw_smem;
    p->smem_fwd_bytes = (int64_t)o.smem_fwd_bytes; p->smem_bwd_bytes = (int64_t)o.smem_bwd_bytes;
    p->image_floats = o.fwd.img_floats; p->workspace_bytes = (int64_t)o.ws_bytes;
    p->recompute_bytes = (int64_t)o.scratch_bytes;
    p->act_bytes = (int64_t)o.act_bytes;
    return 0;
}

static void pack(const NjCfg& c, const float* params, float* image) {
    for (int i = 0; i < c.img_floats; ++i) image[i] = 0.f;
    for (int n = 0; n < NJODE_NUM_NETS; ++n) {
        const NjNet& N = c.net[n];
        for (int l = 0; l < N.n; ++l) {
            const int K = N.dim[l], O = N.dim[l + 1];
            for (int o = 0; o < O; ++o) {
                for (int k = 0; k < K; ++k) image[N.w_img[l] + o * N.ks[l] + k] = params[N.w_src[l] + (long long)o * K + k];
                if (N.b_src[l] >= 0) image[N.b_img[l] + o] = params[N.b_src[l] + o];
            }
        }
    }
}

static void fill_args(NjArgs& a, const njode_batch_t* b, const NjPlanOut& pl, char* ws) {
    memset(&a, 0, sizeof(a));
    a.b = *b;
    a.image = reinterpret_cast<const float*>(ws + pl.ws_image_off);
    a.row_loss = reinterpret_cast<float*>(ws + pl.ws_rowloss_off);
    a.partials = reinterpret_cast<float*>(ws + pl.ws_partials_off);
    a.counter = reinterpret_cast<int*>(ws + pl.ws_counter_off);
}

extern "C" int njode_forward(const njode_model_t* model, const njode_batch_t* batch, const float* params,
                             float* hT, float* loss, float* path_h, float* path_y,
                             const njode_saved_t* saved, void* workspace, void*) {
    NjPlanOut pl;
    if (int rc = plan_for(model, batch, pl)) return rc;
    NjArgs a;
    fill_args(a, batch, pl, (char*)workspace);
    a.hT = hT; a.path_h = path_h; a.path_y = path_y;
    a.h_hist = saved ? saved->h_hist : nullptr;
    a.h_before = saved ? saved->h_before : nullptr;
    a.y_after = saved ? saved->y_after : nullptr;
    a.act_hist = (saved && pl.act_bytes > 0) ? saved->act_hist : nullptr; a.act_nh = pl.act_nh; a.act_wp = pl.act_wp;
    a.get_loss = loss ? 1 : 0;
    a.n_tiles = pl.n_tiles;
    pack(pl.fwd, params, const_cast<float*>(a.image));
    for (int i = 0; i < batch->N; ++i) a.row_loss[i] = 0.f;
    if (pl.seg.ok) {
        std::vector<float> smem(pl.seg.f_smem_floats);
        for (int cta = 0; cta < pl.seg_grid_f; ++cta) {
            std::fill(smem.begin(), smem.end(), NAN);
            if (cta == 0) *a.counter = 0;
            if (pl.seg.tpn == 1) nj_segtpn_cta_forward<NjTpnDims<NJN_A_KC0, NJN_A_KCH, NJN_A_HC, 4>>(pl.fwd, pl.seg, a, smem.data());
            else if (pl.seg.tpn == 2) nj_segtpn_cta_forward<NjTpnDims<NJN_B_KC0, NJN_B_KCH, NJN_B_HC, 4>>(pl.fwd, pl.seg, a, smem.data());
            else nj_seg_cta_forward(pl.fwd, pl.seg, a, smem.data());
        }
    } else if (pl.path.ok) {
        std::vector<float> smem(pl.path.f_smem_floats);
        for (int cta = 0; cta < pl.path_grid_f; ++cta) {
            std::fill(smem.begin(), smem.end(), NAN);
            if (cta == 0) *a.counter = 0;
            if (pl.path.tpn && pl.path.tpn_fwd) {
                const int R = pl.path.rg_f * pl.path.tr_f;
                if (pl.path.tpn == 1 && R == 1) nj_tpn_cta_forward<NjTpnDims<NJN_A_KC0, NJN_A_KCH, NJN_A_HC, 1>>(pl.fwd, pl.path, a, smem.data());
                else if (pl.path.tpn == 1) nj_tpn_cta_forward<NjTpnDims<NJN_A_KC0, NJN_A_KCH, NJN_A_HC, 4>>(pl.fwd, pl.path, a, smem.data());
                else if (R == 1) nj_tpn_cta_forward<NjTpnDims<NJN_B_KC0, NJN_B_KCH, NJN_B_HC, 1>>(pl.fwd, pl.path, a, smem.data());
                else nj_tpn_cta_forward<NjTpnDims<NJN_B_KC0, NJN_B_KCH, NJN_B_HC, 4>>(pl.fwd, pl.path, a, smem.data());
            } else if (pl.path.stat) {
                if (pl.path.rg_f == 1) nj_stat_cta_forward<1, 1>(pl.fwd, pl.path, a, smem.data());
                else if (pl.path.rg_f == 2) nj_stat_cta_forward<2, 1>(pl.fwd, pl.path, a, smem.data());
                else if (pl.path.tr_f == 1) nj_stat_cta_forward<4, 1>(pl.fwd, pl.path, a, smem.data());
                else nj_stat_cta_forward<4, 2>(pl.fwd, pl.path, a, smem.data());
            }
            else if (pl.path.rg_f == 1) nj_path_cta_forward<1, 1>(pl.fwd, pl.path, a, smem.data());
            else if (pl.path.rg_f == 2) nj_path_cta_forward<2, 1>(pl.fwd, pl.path, a, smem.data());
            else if (pl.path.tr_f == 1) nj_path_cta_forward<4, 1>(pl.fwd, pl.path, a, smem.data());
            else nj_path_cta_forward<4, 2>(pl.fwd, pl.path, a, smem.data());
        }
    } else {
        std::vector<float> smem(pl.fwd.smem_floats_fwd);
        for (int cta = 0; cta < pl.grid_fwd && batch->n_units > 0; ++cta) {
            std::fill(smem.begin(), smem.end(), NAN);       // uninitialised shared memory must never matter
            nj_cta_forward(pl.fwd, a, smem.data(), cta, pl.grid_fwd);
        }
    }
    if (loss) {
        double s = 0.0;
        for (int i = 0; i < batch->N; ++i) s += a.row_loss[i];
        loss[0] = (float)(s * (double)(1.f / (float)batch->batch_size_norm));
    }
    return 0;
}

extern "C" int njode_backward(const njode_model_t* model, const njode_batch_t* batch, const float* params,
                              const njode_saved_t* saved, const float* grad_loss, const float* grad_hT,
                              float* grads, void* workspace, void*) {
    NjPlanOut pl;
    if (int rc = plan_for(model, batch, pl)) return rc;
    NjArgs a;
    fill_args(a, batch, pl, (char*)workspace);
    a.h_hist = saved->h_hist; a.h_before = saved->h_before; a.y_after = saved->y_after;
    a.act_hist = pl.act_bytes > 0 ? saved->act_hist : nullptr; a.act_nh = pl.act_nh; a.act_wp = pl.act_wp;
    if (!saved->h_hist && !saved->h_before) {
        if (!(pl.seg.ok && pl.scratch_bytes > 0)) { g_err = "backward needs the history of the forward pass"; return -1; }
        a.scratch = reinterpret_cast<float*>((char*)workspace + pl.ws_scratch_off);
    }
    a.grad_loss = grad_loss; a.grad_hT = grad_hT; a.get_loss = 1; a.n_tiles = pl.n_tiles;
    pack(pl.bwd, params, const_cast<float*>(a.image));
    int nparts = 0;
    if (pl.seg.ok) {
        std::vector<float> smem(pl.seg.b_smem_floats);
        for (int cta = 0; cta < pl.seg_grid_b; ++cta) {
            std::fill(smem.begin(), smem.end(), NAN);
            if (cta == 0) *a.counter = 0;
            if (pl.seg.tpn == 1) nj_segtpn_cta_backward<NjTpnDims<NJN_A_KC0, NJN_A_KCH, NJN_A_HC, 4>>(pl.bwd, pl.seg, a, smem.data(), cta);
            else if (pl.seg.tpn == 2) nj_segtpn_cta_backward<NjTpnDims<NJN_B_KC0, NJN_B_KCH, NJN_B_HC, 4>>(pl.bwd, pl.seg, a, smem.data(), cta);
            else nj_seg_cta_backward<false>(pl.bwd, pl.seg, a, smem.data(), cta);
        }
        nparts = pl.seg_grid_b;
    } else if (pl.path.ok) {
        std::vector<float> smem(pl.path.b_smem_floats);
        for (int cta = 0; cta < pl.path_grid_b; ++cta) {
            std::fill(smem.begin(), smem.end(), NAN);
            if (cta == 0) *a.counter = 0;
            if (pl.path.tpn) {
                const int R = pl.path.rg_b * pl.path.tr_b;
                if (pl.path.tpn == 1 && R == 1) nj_tpn_cta_backward<NjTpnDims<NJN_A_KC0, NJN_A_KCH, NJN_A_HC, 1>>(pl.bwd, pl.path, a, smem.data(), cta);
                else if (pl.path.tpn == 1) nj_tpn_cta_backward<NjTpnDims<NJN_A_KC0, NJN_A_KCH, NJN_A_HC, 4>>(pl.bwd, pl.path, a, smem.data(), cta);
                else if (R == 1) nj_tpn_cta_backward<NjTpnDims<NJN_B_KC0, NJN_B_KCH, NJN_B_HC, 1>>(pl.bwd, pl.path, a, smem.data(), cta);
                else nj_tpn_cta_backward<NjTpnDims<NJN_B_KC0, NJN_B_KCH, NJN_B_HC, 4>>(pl.bwd, pl.path, a, smem.data(), cta);
            } else if (pl.path.stat) {
                if (pl.path.rg_b == 1) nj_stat_cta_backward<1, 1>(pl.bwd, pl.path, a, smem.data(), cta);
                else if (pl.path.rg_b == 2) nj_stat_cta_backward<2, 1>(pl.bwd, pl.path, a, smem.data(), cta);
                else if (pl.path.tr_b == 1) nj_stat_cta_backward<4, 1>(pl.bwd, pl.path, a, smem.data(), cta);
                else nj_stat_cta_backward<4, 2>(pl.bwd, pl.path, a, smem.data(), cta);
            }
            else if (pl.path.pipe) {
                if (pl.path.rg_b == 1) nj_path_cta_backward_pipe<1, 1>(pl.bwd, pl.path, a, smem.data(), cta);
                else if (pl.path.rg_b == 2) nj_path_cta_backward_pipe<2, 1>(pl.bwd, pl.path, a, smem.data(), cta);
                else if (pl.path.tr_b == 1) nj_path_cta_backward_pipe<4, 1>(pl.bwd, pl.path, a, smem.data(), cta);
                else nj_path_cta_backward_pipe<4, 2>(pl.bwd, pl.path, a, smem.data(), cta);
            }
            else if (pl.path.rg_b == 1) nj_path_cta_backward<1, 1>(pl.bwd, pl.path, a, smem.data(), cta);
            else if (pl.path.rg_b == 2) nj_path_cta_backward<2, 1>(pl.bwd, pl.path, a, smem.data(), cta);
            else if (pl.path.tr_b == 1) nj_path_cta_backward<4, 1>(pl.bwd, pl.path, a, smem.data(), cta);
            else nj_path_cta_backward<4, 2>(pl.bwd, pl.path, a, smem.data(), cta);
        }
        nparts = pl.path_grid_b;
    } else {
        std::vector<float> smem(pl.bwd.smem_floats_bwd);
        for (int cta = 0; cta < pl.grid_bwd && batch->n_units > 0; ++cta) {
            std::fill(smem.begin(), smem.end(), NAN);
            nj_cta_backward(pl.bwd, a, smem.data(), cta, pl.grid_bwd);
            nparts = pl.grid_bwd;
        }
    }
    const NjCfg& c = pl.bwd;
    for (int n = 0; n < NJODE_NUM_NETS; ++n) {
        const NjNet& N = c.net[n];
        for (int l = 0; l < N.n; ++l) {
            const int K = N.dim[l], O = N.dim[l + 1];
            for (int o = 0; o < O; ++o) {
                for (int k = 0; k < K; ++k) {
                    float s = 0.f;
                    for (int p = 0; p < nparts; ++p) s += a.partials[(size_t)p * c.img_floats + N.w_img[l] + o * N.ks[l] + k];
                    grads[N.w_src[l] + (long long)o * K + k] = s;
                }
                if (N.b_src[l] >= 0) {
                    float s = 0.f;
                    for (int p = 0; p < nparts; ++p) s += a.partials[(size_t)p * c.img_floats + N.b_img[l] + o];
                    grads[N.b_src[l] + o] = s;
                }
            }
        }
    }
    return 0;
}
