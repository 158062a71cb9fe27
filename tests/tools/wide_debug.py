#!/usr/bin/env python
"""TEST INFRASTRUCTURE (developer tool, may import oracle/): stage-wise comparison of the tcgen05 forward (njode_wide_forward) with the fp32 kernels on the same
batch: h_hist (every Euler step), h_before, y_after, hT, loss.  usage: wide_debug.py CASE"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # tests/tools/ -> repo root
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
import oracle.njode_oracle as orc
from njode_b200 import models

CASES = {
    # name: (d, H, widths, n_layers, B, steps, obs_perc, dropout)
    "tiny": (16, 64, 64, 1, 4, 4, 0.3, 0.0),
    "tiny2": (16, 64, 64, 2, 40, 6, 0.3, 0.0),
    "mid": (4, 128, 192, 2, 300, 12, 0.25, 0.0),
    "cfg5": (16, 256, 256, 4, 300, 20, 0.2, 0.0),
    "cfg5_drop": (16, 256, 256, 4, 300, 20, 0.2, 0.1),
    "cfg5_big": (16, 256, 256, 4, 4096, 100, 0.1, 0.1),
    "cfg5_huge": (16, 256, 256, 4, 16384, 200, 0.1, 0.1),
}


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def main(name):
    d, H, W, L, B, steps, op, drop = CASES[name]
    nn = [[W, "tanh"]] * L
    cfg = cases.demo_cfg(input_size=d, output_size=d, hidden_size=H, ode_nn=nn, enc_nn=nn, readout_nn=nn, dropout_rate=drop)
    sd = orc.init_state_dict(orc.Config(**cfg), seed=3)
    m = models.NJODE(**cfg); m.load_state_dict(sd); m.to("cuda:0")
    m.train() if drop else m.eval()
    batch = cases.grid_batch(B, d, steps, op, seed=11)
    dt = 1.0 / steps
    pb = m.prepare_batch(batch["times"], batch["time_ptr"], batch["X"].cuda(), batch["obs_idx"], dt, 1.0,
                         batch["start_X"].cuda(), batch["n_obs_ot"])
    m._ensure_flat()
    mt = m._model_struct(12345 if drop else 0)
    r = pb.runner
    print("case", name, "B", B, "N", pb.N, "S", pb.sched.S, "units", pb.n_units, "wide_supported", r.wide_supported(mt), flush=True)
    hT1, loss1, _, _, s1 = r.forward(mt, pb, m._flat, H, d, True, True)
    torch.cuda.synchronize()
    print("fp32 loss", float(loss1), flush=True)
    hT2, loss2, _, _, s2 = r.forward_wide(mt, pb, m._flat, H, d, True, True, fp32_backward=True)
    torch.cuda.synchronize()
    print("tc   loss", float(loss2), "rel", abs(float(loss2) - float(loss1)) / abs(float(loss1)), flush=True)
    S = pb.sched.S
    hh1, hh2 = s1[0].view(S, B, H), s2[0].view(S, B, H)
    print("hT rel", rel(hT2, hT1))
    print("h_before rel", rel(s2[1], s1[1]), " y_after rel", rel(s2[2], s1[2]))
    print("h_hist rel (all)", rel(hh2, hh1))
    for s in list(range(min(S, 6))) + [S - 1]:
        print("  step", s, "h_hist rel", rel(hh2[s], hh1[s]), "nan", int(torch.isnan(hh2[s]).sum()))
    if rel(hh2, hh1) > 0.05:
        e = (hh2 - hh1).abs()
        s_, b_, c_ = np.unravel_index(int(e.argmax()), e.shape)
        print("worst at step", s_, "path", b_, "col", c_, float(hh2[s_, b_, c_]), float(hh1[s_, b_, c_]))
        print("err by column block of 8 (step 0):", [round(float(e[0][:, i:i + 8].max()), 4) for i in range(0, H, 8)][:32])
        print("err by path (step 0):", [round(float(e[0][i].max()), 4) for i in range(min(B, 16))])
    # gradients: tcgen05 backward vs the fp32 backward (both on this batch)
    gl = torch.ones((), device="cuda")
    g1 = r.backward(mt, pb, m._flat, s1, gl, None)
    hT3, loss3, _, _, s3 = r.forward_wide(mt, pb, m._flat, H, d, True, True)
    torch.cuda.synchronize()
    print("tc(train) loss", float(loss3), flush=True)
    g2 = r.backward_wide(mt, pb, m._flat, s3[1], gl, None)
    torch.cuda.synchronize()
    print("grad rel (all)", rel(g2, g1), "nan", int(torch.isnan(g2).sum()), flush=True)
    for (nme, p_), (o, n_, shp) in zip(m.named_parameters(), m._flat_layout):
        a_, b_ = g2[o:o + n_], g1[o:o + n_]
        print("   %-28s rel %.4f  max|ref| %.3e" % (nme, rel(a_, b_), float(b_.abs().max())))
    if name.endswith("big") or name.endswith("huge"):
        import ctypes as C
        prof = torch.zeros(1024, dtype=torch.int64, device="cuda")
        r.lib.dll.njode_wide_set_profile(C.c_void_p(prof.data_ptr()))
        r.forward_wide(mt, pb, m._flat, H, d, True, False)
        torch.cuda.synchronize()
        r.lib.dll.njode_wide_set_profile(None)
        pr = prof.cpu().numpy().reshape(256, 4)
        t0 = pr[0, 0]
        print("layer-GEMM stamps of CTA 0 (cycles): a_ready seen | MMAs issued | acc seen | epilogue done ; deltas")
        for i in range(1, 16):
            a0, a1, a2, a3 = pr[i]
            print("  %3d: mma_start %7d  issue %5d  acc_wait %6d  epilogue %6d  -> next start %6d" % (
                i, a0 - t0, a1 - a0, a2 - a0, a3 - a2, pr[i + 1, 0] - a3))
        r.lib.dll.njode_set_timing(1)
        for _ in range(3):
            r.forward_wide(mt, pb, m._flat, H, d, True, True)
        torch.cuda.synchronize()
        a, b_, c_ = C.c_float(), C.c_float(), C.c_float()
        r.lib.dll.njode_wide_get_timing(C.byref(a), C.byref(b_), C.byref(c_))
        F_ode = 2 * ((d + H + 2) * W + (L - 1) * W * W + W * H)
        print("timing ms: enc %.3f ode %.3f ro %.3f ; ode TFLOP/s %.1f" % (a.value, b_.value, c_.value, B * S * F_ode / (b_.value * 1e-3) / 1e12))
        for _ in range(2):
            hT3, loss3, _, _, s3 = r.forward_wide(mt, pb, m._flat, H, d, True, True)
            g2 = r.backward_wide(mt, pb, m._flat, s3[1], gl, None)
        torch.cuda.synchronize()
        r.lib.dll.njode_wide_get_timing(C.byref(a), C.byref(b_), C.byref(c_))
        ch, dw = C.c_float(), C.c_float()
        r.lib.dll.njode_wide_get_timing_bwd(C.byref(ch), C.byref(dw))
        tot = a.value + b_.value + c_.value + ch.value + dw.value
        print("train: fwd enc %.3f ode %.3f ro %.3f | bwd chains %.3f dW %.3f ms ; total %.3f ms = %.1f M path-steps/s, %.1f TFLOP/s (3x fwd flops)" % (
            a.value, b_.value, c_.value, ch.value, dw.value, tot, B * S / tot / 1e3, 3 * B * S * F_ode / (tot * 1e-3) / 1e12))
        f, bb = C.c_float(), C.c_float()
        for _ in range(2):
            r.forward(mt, pb, m._flat, H, d, True, True)
        torch.cuda.synchronize()
        r.lib.dll.njode_get_timing(C.byref(f), C.byref(bb))
        r.backward(mt, pb, m._flat, s1, gl, None)
        torch.cuda.synchronize()
        r.lib.dll.njode_get_timing(C.byref(f), C.byref(bb))
        print("fp32 fwd kernel ms %.3f bwd kernel ms %.3f" % (f.value, bb.value))


if __name__ == "__main__":
    main(sys.argv[1])
