"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

CPU restatement (functional PyTorch, fp32 or fp64) of the NJ-ODE hot path of the
reference, written from the behaviour of /root/reference/NJODE/models.py.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module; the product package
``njode_b200`` never does (its ops fail loudly without the CUDA library).

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the *real* reference
(imported from /root/reference in the build container) and this restatement on
the same seeded inputs; ``tests/test_oracle_golden.py`` checks this file against
the committed outputs of the real reference (loss, hT, path_t, path_h, path_y
and all parameter gradients).

Every function cites the reference lines it restates (paths relative to
/root/reference).
"""
import math

import numpy as np
import torch

# ----------------------------------------------------------------------------------------------
# counter-based dropout masks (shared definition with the CUDA kernels, see
# njode_b200/csrc/njode_common.cuh::philox4x32_10 and dropout_keep).  The reference draws its
# masks from aten::bernoulli_ (NJODE/models.py:160,164 -> torch.nn.Dropout); that stream cannot
# be reproduced on the device, so train-mode parity is defined as: *given the same keep-masks*,
# same numbers.  The oracle therefore regenerates the device's masks here.
# ----------------------------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


# tanh of the restated computation.  ``torch.tanh`` = the reference.  tests/parity_util.py swaps in ``tanh_kernel_formula``
# (fp32 evaluation only) to MEASURE how much fp32 noise a tanh of the kernels' accuracy class adds to each output.
TANH = torch.tanh


def tanh_kernel_formula(x):
    """1 - 2 / (2^(2 x log2 e) + 1), every operation rounded to the dtype of x: the formula of nj_tanh
    (njode_b200/csrc/njode_hash.cuh); CUDA's own tanhf uses the same expression for |x| > 0.55"""
    return 1 - 2 / (torch.exp2(x * 2.8853900817779268) + 1)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox-4x32-10 on numpy uint32 arrays (Salmon et al. 2011).  Returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint32).copy()
    c1 = np.asarray(c1, dtype=np.uint32).copy()
    c2 = np.asarray(c2, dtype=np.uint32).copy()
    c3 = np.asarray(c3, dtype=np.uint32).copy()
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask32 = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PHILOX_M0 * c0.astype(np.uint64)
            p1 = _PHILOX_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask32).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n1 = lo1
            n2 = hi0 ^ c3 ^ k1
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32(k0 + _PHILOX_W0)
            k1 = np.uint32(k1 + _PHILOX_W1)
    return c0, c1, c2, c3


def _fmix32(h):
    h = np.asarray(h, dtype=np.uint32)
    with np.errstate(over="ignore"):
        h = h ^ (h >> np.uint32(16)); h = h * np.uint32(0x85EBCA6B)
        h = h ^ (h >> np.uint32(13)); h = h * np.uint32(0xC2B2AE35)
        h = h ^ (h >> np.uint32(16))
    return h


def dropout_keep_mask(seed, path_ids, event_id, net_id, layer_id, width, p):
    """keep-mask [len(path_ids), width] (float32 0/1) of one MLP hidden layer; restates
    nj_row_key / nj_layer_key / nj_keep of njode_b200/csrc/njode_core.cuh (murmur3 finaliser chain
    keyed by seed, path, event, net, layer, neuron pair); keep iff 16-bit field >= floor(p * 2^16)."""
    path_ids = np.asarray(path_ids, dtype=np.uint32)
    seed_lo, seed_hi = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    thr16 = np.uint32(min(int(np.float32(p).astype(np.float64) * 65536.0), 65536))
    with np.errstate(over="ignore"):
        rk = _fmix32(_fmix32(path_ids ^ seed_lo) + np.uint32(event_id & 0xFFFFFFFF) * np.uint32(0x9E3779B9)) ^ seed_hi
        tag = np.uint32(net_id * 16 + layer_id + 1)
        lk = _fmix32(rk + tag * np.uint32(0x85EBCA77))
        neuron = np.arange(width, dtype=np.uint32)[None, :]
        # one 32-bit word per neuron pair (o, o ^ 8): 16-bit fields (nj_keep in njode_core.cuh)
        widx = (neuron & np.uint32(7)) | ((neuron >> np.uint32(4)) << np.uint32(3))
        word = _fmix32(lk[:, None] + widx * np.uint32(0xC2B2AE3D))
        field = np.where(((neuron >> np.uint32(3)) & np.uint32(1)) == 1, word >> np.uint32(16), word & np.uint32(0xFFFF))
    return (field >= thr16).astype(np.float32)


NET_ODE, NET_ENC, NET_RO = 0, 1, 2
# event ids used to key dropout masks for the MLP calls that are not Euler steps.  Euler step k of
# the batch-global schedule uses event id k.  Jump number i uses three ids (see below) offset by
# EVENT_JUMP_BASE; the initial encoder call uses EVENT_INIT.
EVENT_JUMP_BASE = 1 << 30
EVENT_PATH_RO_BASE = 1 << 29      # readout after Euler step k when return_path=True: id = base + k
EVENT_INIT = (1 << 31) - 1


def jump_event_id(i, which):
    """which: 0 = readout before jump (Y_bj), 1 = encoder, 2 = readout after jump (Y)."""
    return EVENT_JUMP_BASE + 3 * i + which


# ----------------------------------------------------------------------------------------------
# networks
# ----------------------------------------------------------------------------------------------
def _act(name):
    # NJODE/models.py:134-137 (nonlinears)
    return {"tanh": lambda v: TANH(v), "relu": torch.relu}[name]


def linear_indices(nn_desc):
    """positions of the Linear modules inside the reference's nn.Sequential
    (NJODE/models.py:153-166: Linear, then [act, Dropout, Linear] per hidden layer)."""
    if nn_desc is None:
        return [0]
    return [3 * i for i in range(len(nn_desc) + 1)]


def mlp(x, sd, prefix, nn_desc, bias, dropout=None):
    """get_ffnn forward (NJODE/models.py:140-166).  ``dropout`` = None (eval) or a callable
    (layer_index, width) -> keep-mask tensor [rows, width] already scaled by 1/(1-p)."""
    idx = linear_indices(nn_desc)
    out = x
    for li, pos in enumerate(idx):
        w = sd["%s.%d.weight" % (prefix, pos)]
        b = sd["%s.%d.bias" % (prefix, pos)] if bias else None
        out = torch.nn.functional.linear(out, w, b)
        if li < len(idx) - 1:
            out = _act(nn_desc[li][1])(out)
            if dropout is not None:
                out = out * dropout(li, out.shape[1])
    return out


def ffnn(x, sd, prefix, nn_desc, bias, residual, in_size, out_size, mask=None, dropout=None):
    """FFNN.forward (NJODE/models.py:261-276) incl. the residual cases set up in
    NJODE/models.py:240-259."""
    if mask is not None:
        out = mlp(torch.cat((TANH(x), mask), 1), sd, prefix, nn_desc, bias, dropout)
    else:
        out = mlp(TANH(x), sd, prefix, nn_desc, bias, dropout)
    if not residual:
        return out
    if in_size <= out_size:
        if out_size % in_size != 0:
            raise ValueError("for residual: output_size needs to be multiple of input_size")
        return x.repeat(1, out_size // in_size) + out
    if in_size % out_size != 0:
        raise ValueError("for residual: input_size needs to be multiple of output_size")
    mult = in_size // out_size
    ident = torch.mean(torch.stack(x.chunk(mult, dim=1)), dim=0)
    return ident + out


def gru_cell(x, h, sd, prefix, bias):
    """torch.nn.GRUCell as used by GRUCell.forward (NJODE/models.py:202-217): gates in (r, z, n) order,
    r = sig(W_ir x + b_ir + W_hr h + b_hr), z likewise, n = tanh(W_in x + b_in + r (W_hn h + b_hn)),
    h' = (1 - z) n + z h.  The caller passes x = tanh(X_obs), h = tanh(h[i_obs])."""
    gi = torch.nn.functional.linear(x, sd[prefix + ".weight_ih"], sd[prefix + ".bias_ih"] if bias else None)
    gh = torch.nn.functional.linear(h, sd[prefix + ".weight_hh"], sd[prefix + ".bias_hh"] if bias else None)
    i_r, i_z, i_n = gi.chunk(3, dim=1)
    h_r, h_z, h_n = gh.chunk(3, dim=1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = TANH(i_n + r * h_n)
    return (1 - z) * n + z * h


def loss_term(which, X_obs, Y_obs, Y_obs_bj, n_obs_ot, batch_size, weight, M_obs=None, eps=1e-10):
    """compute_loss / compute_loss_2 (NJODE/models.py:71-126)."""
    m = 1.0 if M_obs is None else M_obs
    if which == "standard":
        a = torch.sqrt(torch.sum(m * (X_obs - Y_obs) ** 2, dim=1) + eps)
        b = torch.sqrt(torch.sum(m * (Y_obs_bj - Y_obs) ** 2, dim=1) + eps)
        inner = (2 * weight * a + 2 * (1 - weight) * b) ** 2
    elif which == "easy":
        a = torch.sqrt(torch.sum(m * (X_obs - Y_obs) ** 2, dim=1) + eps)
        b = torch.sqrt(torch.sum(m * (Y_obs_bj - X_obs) ** 2, dim=1) + eps)
        inner = (weight * a + (1 - weight) * b) ** 2
    else:
        raise AssertionError(which)
    return torch.sum(inner / n_obs_ot) / batch_size


# ----------------------------------------------------------------------------------------------
# the forward pass
# ----------------------------------------------------------------------------------------------
class Config:
    """constructor arguments of the reference NJODE (NJODE/models.py:284-341)."""

    def __init__(self, input_size, hidden_size, output_size, ode_nn, readout_nn, enc_nn,
                 use_rnn=False, bias=True, dropout_rate=0.0, solver="euler", weight=0.5,
                 weight_decay=1.0, **options):
        o = options.get("options", {})
        self.input_size, self.hidden_size, self.output_size = input_size, hidden_size, output_size
        self.ode_nn, self.readout_nn, self.enc_nn = ode_nn, readout_nn, enc_nn
        self.use_rnn, self.bias, self.dropout_rate = use_rnn, bias, dropout_rate
        self.weight = weight
        self.which_loss = o.get("which_loss", "standard")
        self.residual = o.get("residual_enc_dec", True)
        self.input_current_t = o.get("input_current_t", False)
        self.masked = o.get("masked", False)
        assert self.which_loss in ("standard", "easy")


def forward(cfg, sd, times, time_ptr, X, obs_idx, delta_t, T, start_X, n_obs_ot,
            return_path=False, get_loss=True, until_T=False, M=None, dropout_seed=None,
            weight=None):
    """NJODE.forward (NJODE/models.py:379-518).

    ``sd``: dict name -> tensor with the reference's state_dict keys.  All tensors share one dtype
    (fp32 = the reference's arithmetic; fp64 = high-precision evaluation of the same function; the
    scalars the reference rounds to fp32 -- delta_t_, current_time, obs_time -- are rounded here
    too so both precisions evaluate the same function).
    ``dropout_seed``: None = eval mode; int = train mode with the device's counter-based masks;
    "native" = train mode with torch's own dropout stream (as the reference draws it).
    """
    dt_ = start_X.dtype
    p = cfg.dropout_rate
    B = start_X.shape[0]
    path_ids = np.arange(B)
    weight = cfg.weight if weight is None else weight

    def dropper(net, event, rows):
        if dropout_seed is None or p == 0.0:
            return None
        if dropout_seed == "native":
            # the reference's own masks: torch.nn.Dropout -> aten::bernoulli_ (NJODE/models.py:160,164);
            # used by bench.py's CPU baseline so the port runs the reference's op sequence
            return lambda layer, width: torch.nn.functional.dropout(
                torch.ones(len(rows), width, dtype=dt_), p, True)

        def f(layer, width):
            keep = dropout_keep_mask(dropout_seed, rows, event, net, layer, width, p)
            return torch.as_tensor(keep, dtype=dt_) * float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))
        return f

    def enc(x, mask, event, rows):
        return ffnn(x, sd, "encoder_map.ffnn", cfg.enc_nn, cfg.bias, cfg.residual,
                    cfg.input_size, cfg.hidden_size, mask=mask, dropout=dropper(NET_ENC, event, rows))

    def ro(h, event, rows):
        return ffnn(h, sd, "readout_map.ffnn", cfg.readout_nn, cfg.bias, cfg.residual,
                    cfg.hidden_size, cfg.output_size, dropout=dropper(NET_RO, event, rows))

    def ode_f(x, h, tau, tdiff, event):
        # ODEFunc.forward (NJODE/models.py:188-199)
        parts = [TANH(x), TANH(h), tau, tdiff]
        if cfg.input_current_t:
            parts.append(tau + tdiff)
        return mlp(torch.cat(parts, dim=1), sd, "ode_f.f", cfg.ode_nn, cfg.bias,
                   dropper(NET_ODE, event, path_ids))

    # NJODE/models.py:411-419
    if cfg.masked:
        h = enc(start_X, torch.zeros_like(start_X), EVENT_INIT, path_ids)
    else:
        h = enc(start_X, None, EVENT_INIT, path_ids)
    last_X = start_X
    tau = torch.zeros(B, 1, dtype=dt_)
    current_time = 0.0
    loss = 0
    step_no = 0
    if return_path:
        path_t, path_h = [0], [h]
        path_y = [ro(h, EVENT_INIT, path_ids)]
    assert len(times) + 1 == len(time_ptr)           # NJODE/models.py:428

    def euler_to(h, current_time, target, step_no):
        # NJODE/models.py:432-445 / 498-511 ; ode_step 369-377
        while current_time < (target - 1e-10 * delta_t):
            if current_time < target - delta_t:
                d = delta_t
            else:
                d = target - current_time
            t32 = float(np.float32(current_time))
            d32 = float(np.float32(d))
            tdiff = torch.full((B, 1), t32, dtype=dt_) - tau
            h = h + d32 * ode_f(last_X, h, tau, tdiff, step_no)
            current_time = current_time + d
            step_no += 1
            if return_path:
                path_t.append(current_time)
                path_h.append(h)
                path_y.append(ro(h, EVENT_PATH_RO_BASE + step_no - 1, path_ids))
        return h, current_time, step_no

    for i, obs_time in enumerate(times):
        h, current_time, step_no = euler_to(h, current_time, obs_time, step_no)
        start, end = int(time_ptr[i]), int(time_ptr[i + 1])     # NJODE/models.py:449-456
        X_obs = X[start:end]
        i_obs = obs_idx[start:end]
        rows = i_obs.numpy()
        M_obs = M[start:end] if cfg.masked else None
        Y_bj = ro(h, jump_event_id(i, 0), path_ids)              # NJODE/models.py:459
        temp = h.clone()                                           # NJODE/models.py:463-470
        if cfg.use_rnn:                                            # NJODE/models.py:460-461, 213-217
            temp[i_obs] = gru_cell(TANH(X_obs), TANH(h[i_obs]), sd, "obs_c.gru_d", cfg.bias)
        elif cfg.masked:
            X_imp = X_obs * M_obs + (torch.ones_like(M_obs) - M_obs) * Y_bj[i_obs]
            temp[i_obs] = enc(X_imp, M_obs, jump_event_id(i, 1), rows)
        else:
            temp[i_obs] = enc(X_obs, None, jump_event_id(i, 1), rows)
        h = temp
        Y = ro(h, jump_event_id(i, 2), path_ids)                 # NJODE/models.py:471
        if get_loss:                                               # NJODE/models.py:473-477
            loss = loss + loss_term(cfg.which_loss, X_obs, Y[i_obs], Y_bj[i_obs],
                                    n_obs_ot[i_obs], B, weight, M_obs)
        temp_X = last_X.clone()                                    # NJODE/models.py:481-489
        temp_tau = tau.clone()
        temp_X[i_obs] = Y[i_obs] if cfg.masked else X_obs
        temp_tau[i_obs] = float(np.float32(np.float64(obs_time)))
        last_X, tau = temp_X, temp_tau
        if return_path:                                            # NJODE/models.py:491-494
            path_t.append(obs_time)
            path_h.append(h)
            path_y.append(Y)

    if until_T:                                                    # NJODE/models.py:497-511
        h, current_time, step_no = euler_to(h, current_time, T, step_no)

    if return_path:
        return h, loss, np.array(path_t), torch.stack(path_h), torch.stack(path_y)
    return h, loss


def init_state_dict(cfg, seed=0, dtype=torch.float32):
    """random parameters with the reference's shapes/keys (Xavier-uniform weights, small random
    biases so the bias path is exercised; NJODE/models.py:21-26 uses zero biases)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def add(prefix, in_f, out_f, nn_desc):
        sizes = [in_f] + ([] if nn_desc is None else [l[0] for l in nn_desc]) + [out_f]
        for pos, (a, b) in zip(linear_indices(nn_desc), zip(sizes[:-1], sizes[1:])):
            bound = math.sqrt(6.0 / (a + b))
            sd["%s.%d.weight" % (prefix, pos)] = ((torch.rand(b, a, generator=g) * 2 - 1) * bound).to(dtype)
            if cfg.bias:
                sd["%s.%d.bias" % (prefix, pos)] = ((torch.rand(b, generator=g) * 2 - 1) * 0.1).to(dtype)

    add_t = 3 if cfg.input_current_t else 2
    add("ode_f.f", cfg.input_size + cfg.hidden_size + add_t, cfg.hidden_size, cfg.ode_nn)
    add("encoder_map.ffnn", cfg.input_size * (2 if cfg.masked else 1), cfg.hidden_size, cfg.enc_nn)
    add("readout_map.ffnn", cfg.hidden_size, cfg.output_size, cfg.readout_nn)
    if cfg.use_rnn:
        # torch.nn.GRUCell default init: U(-1/sqrt(H), 1/sqrt(H)) for all four tensors (init_weights only
        # touches nn.Linear, NJODE/models.py:21-26)
        k = 1.0 / math.sqrt(cfg.hidden_size)
        H3 = 3 * cfg.hidden_size
        sd["obs_c.gru_d.weight_ih"] = ((torch.rand(H3, cfg.input_size, generator=g) * 2 - 1) * k).to(dtype)
        sd["obs_c.gru_d.weight_hh"] = ((torch.rand(H3, cfg.hidden_size, generator=g) * 2 - 1) * k).to(dtype)
        if cfg.bias:
            sd["obs_c.gru_d.bias_ih"] = ((torch.rand(H3, generator=g) * 2 - 1) * k).to(dtype)
            sd["obs_c.gru_d.bias_hh"] = ((torch.rand(H3, generator=g) * 2 - 1) * k).to(dtype)
    return sd


def loss_and_grads(cfg, sd, batch, delta_t, T, dtype=torch.float32, dropout_seed=None,
                   until_T=False, grad_hT=None, weight=None):
    """convenience: forward + autograd backward; returns (hT, loss, {name: grad})."""
    sd = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    M = batch.get("M")
    hT, loss = forward(cfg, sd, batch["times"], batch["time_ptr"], batch["X"].to(dtype),
                       batch["obs_idx"], delta_t, T, batch["start_X"].to(dtype), batch["n_obs_ot"],
                       until_T=until_T, M=None if M is None else M.to(dtype),
                       dropout_seed=dropout_seed, weight=weight)
    obj = loss
    if grad_hT is not None:
        obj = obj + (hT * grad_hT.to(dtype)).sum()
    grads = torch.autograd.grad(obj, list(sd.values()), allow_unused=True)
    g = {k: (torch.zeros_like(v) if gi is None else gi) for (k, v), gi in zip(sd.items(), grads)}
    return hT.detach(), loss.detach(), g
