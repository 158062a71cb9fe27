"""Drop-in module for the reference's ``NJODE/models.py``: same classes, constructor arguments,
``forward`` signature, ``state_dict`` keys and checkpoint format -- but ``NJODE.forward`` and its
backward run in the hand-written sm_100a kernels of ``libnjode_b200.so`` (njode_b200/csrc) instead
of ~50k ATen calls per batch.  Reference citations are relative to /root/reference.

What stays Python here: argument parsing, the parameter tree (ordinary ``nn.Linear`` modules, so
Adam state / ``load_state_dict`` / shipped ``checkpt.tar`` files work unchanged) and the host
schedule (njode_b200/schedule.py).  There is no CPU execution path for ``NJODE.forward``: the
module must live on a CUDA device and raises otherwise.
"""
import os

import numpy as np
import torch

from . import _ext


# ----------------------------------------------------------------------------------------------
# helpers with the reference's names
# ----------------------------------------------------------------------------------------------
def init_weights(m, bias=0.0):
    """NJODE/models.py:21-26"""
    if type(m) == torch.nn.Linear:
        torch.nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            m.bias.data.fill_(bias)


def save_checkpoint(model, optimizer, path, epoch):
    """NJODE/models.py:29-45 (same file name and dict keys)"""
    os.makedirs(path, exist_ok=True)
    torch.save({"epoch": epoch, "weight": model.weight,
                "model_state_dict": model.state_dict(),
                "optimizer_state_dict": optimizer.state_dict()},
               os.path.join(path, "checkpt.tar"))


def get_ckpt_model(ckpt_path, model, optimizer, device):
    """NJODE/models.py:48-67; ``weights_only=False`` because the file holds a python float and the
    optimizer state (the reference relied on torch<2.6 defaults)."""
    ckpt_path = os.path.join(ckpt_path, "checkpt.tar")
    if not os.path.exists(ckpt_path):
        raise Exception("Checkpoint " + ckpt_path + " does not exist.")
    checkpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    optimizer.load_state_dict(checkpt["optimizer_state_dict"])
    model.load_state_dict(checkpt["model_state_dict"])
    model.epoch = checkpt["epoch"]
    model.weight = checkpt["weight"]
    model.to(device)


def compute_loss(X_obs, Y_obs, Y_obs_bj, n_obs_ot, batch_size, eps=1e-10, weight=0.5, M_obs=None):
    """NJODE/models.py:71-106 on tensors (API helper; inside NJODE.forward the loss is accumulated
    by the kernels)."""
    m = 1.0 if M_obs is None else M_obs
    inner = (2 * weight * torch.sqrt(torch.sum(m * (X_obs - Y_obs) ** 2, dim=1) + eps) +
             2 * (1 - weight) * torch.sqrt(torch.sum(m * (Y_obs_bj - Y_obs) ** 2, dim=1) + eps)) ** 2
    return torch.sum(inner / n_obs_ot) / batch_size


def compute_loss_2(X_obs, Y_obs, Y_obs_bj, n_obs_ot, batch_size, eps=1e-10, weight=0.5, M_obs=None):
    """NJODE/models.py:109-126"""
    m = 1.0 if M_obs is None else M_obs
    inner = (weight * torch.sqrt(torch.sum(m * (X_obs - Y_obs) ** 2, dim=1) + eps) +
             (1 - weight) * torch.sqrt(torch.sum(m * (Y_obs_bj - X_obs) ** 2, dim=1) + eps)) ** 2
    return torch.sum(inner / n_obs_ot) / batch_size


LOSS_FUN_DICT = {"standard": compute_loss, "easy": compute_loss_2}


def _mean_square_diff(x, y):
    """default ``diff_fun`` of NJODE.evaluate (NJODE/models.py:523)"""
    return np.mean((x - y) ** 2)
nonlinears = {"tanh": torch.nn.Tanh, "relu": torch.nn.ReLU}


def get_ffnn(input_size, output_size, nn_desc, dropout_rate, bias):
    """NJODE/models.py:140-166: Linear, then (activation, Dropout, Linear) per hidden layer -- the
    Sequential indices (0, 3, 6, ...) are part of the checkpoint format."""
    if nn_desc is None:
        layers = [torch.nn.Linear(input_size, output_size, bias=bias)]
    else:
        layers = [torch.nn.Linear(input_size, nn_desc[0][0], bias=bias)]
        for i in range(len(nn_desc) - 1):
            layers += [nonlinears[nn_desc[i][1]](), torch.nn.Dropout(p=dropout_rate),
                       torch.nn.Linear(nn_desc[i][0], nn_desc[i + 1][0], bias=bias)]
        layers += [nonlinears[nn_desc[-1][1]](), torch.nn.Dropout(p=dropout_rate),
                   torch.nn.Linear(nn_desc[-1][0], output_size, bias=bias)]
    return torch.nn.Sequential(*layers)


class ODEFunc(torch.nn.Module):
    """f_theta (NJODE/models.py:170-199).  Holds the parameters; a direct call evaluates the MLP with
    ATen on the parameters' device (API compatibility only -- NJODE.forward never takes this path)."""

    def __init__(self, input_size, hidden_size, ode_nn, dropout_rate=0.0, bias=True,
                 input_current_t=False):
        super().__init__()
        self.input_current_t = input_current_t
        add = 3 if input_current_t else 2
        self.f = get_ffnn(input_size=input_size + hidden_size + add, output_size=hidden_size,
                          nn_desc=ode_nn, dropout_rate=dropout_rate, bias=bias)

    def forward(self, x, h, tau, tdiff):
        parts = [torch.tanh(x), torch.tanh(h), tau, tdiff]
        if self.input_current_t:
            parts.append(tau + tdiff)
        return self.f(torch.cat(parts, dim=1))


class GRUCell(torch.nn.Module):
    """rho_theta (NJODE/models.py:202-217); only reachable with use_rnn=True."""

    def __init__(self, input_size, hidden_size, bias=True):
        super().__init__()
        self.gru_d = torch.nn.GRUCell(input_size, hidden_size, bias=bias)
        self.input_size = input_size

    def forward(self, h, X_obs, i_obs):
        temp = h.clone()
        temp[i_obs] = self.gru_d(torch.tanh(X_obs), torch.tanh(h[i_obs]))
        return temp


class FFNN(torch.nn.Module):
    """encoder / readout network with tanh on the input and the residual cases of
    NJODE/models.py:220-276 (API compatibility, see ODEFunc)."""

    def __init__(self, input_size, output_size, nn_desc, dropout_rate=0.0, bias=True,
                 residual=False, masked=False):
        super().__init__()
        in_size = 2 * input_size if masked else input_size
        self.masked = masked
        self.ffnn = get_ffnn(input_size=in_size, output_size=output_size, nn_desc=nn_desc,
                             dropout_rate=dropout_rate, bias=bias)
        self.case = 0
        if residual:
            if input_size <= output_size:
                if output_size % input_size != 0:
                    raise ValueError("for residual: output_size needs to be multiple of input_size")
                self.case, self.mult = 1, output_size // input_size
            else:
                if input_size % output_size != 0:
                    raise ValueError("for residual: input_size needs to be multiple of output_size")
                self.case, self.mult = 2, input_size // output_size

    def forward(self, nn_input, mask=None):
        if self.masked:
            assert mask is not None
            out = self.ffnn(torch.cat((torch.tanh(nn_input), mask), 1))
        else:
            out = self.ffnn(torch.tanh(nn_input))
        if self.case == 1:
            return nn_input.repeat(1, self.mult) + out
        if self.case == 2:
            return torch.mean(torch.stack(nn_input.chunk(self.mult, dim=1)), dim=0) + out
        return out


# ----------------------------------------------------------------------------------------------
# autograd bridge
# ----------------------------------------------------------------------------------------------
class _NJODEFunction(torch.autograd.Function):
    """forward = njode_forward, backward = njode_backward (include/njode_b200.h)"""

    @staticmethod
    def forward(ctx, module, runner, pb, model_t, get_loss, need_grad, *params):
        flat = module._flat
        H, dout = module.hidden_size, module.output_size
        if module._use_tensor_cores(runner, pb, model_t):
            hT, loss, path_h, path_y, saved = runner.forward_wide(
                model_t, pb, flat, H, dout, get_loss, need_grad, fp32_backward=module.tensor_core_backward == "fp32")
        else:
            hT, loss, path_h, path_y, saved = runner.forward(model_t, pb, flat, H, dout, get_loss, need_grad,
                                                             recompute=module.recompute)
        ctx.module, ctx.runner, ctx.pb, ctx.model_t, ctx.saved = module, runner, pb, model_t, saved
        # the backward re-reads the parameters: remember which buffer / which in-place versions the forward saw
        ctx.flat_version = module._flat_version
        ctx.param_versions = tuple(p._version for p in params)
        ctx.params = params
        ctx.set_materialize_grads(False)      # an unused output (hT in loss.backward()) arrives as None, not zeros
        ctx.mark_non_differentiable(*[t for t in (path_h, path_y) if t is not None])
        outs = (hT, loss if loss is not None else hT.new_zeros(()))
        ctx.n_extra = 0
        if path_h is not None:
            outs = outs + (path_h, path_y)
        return outs

    @staticmethod
    def backward(ctx, g_hT, g_loss, *unused):
        module, runner = ctx.module, ctx.runner
        if ctx.saved is None:
            raise RuntimeError("NJODE.forward ran without gradient bookkeeping")
        if (ctx.flat_version != module._flat_version
                or tuple(p._version for p in ctx.params) != ctx.param_versions):
            # same contract as autograd's saved-tensor version check: the kernels would otherwise combine the NEW weights
            # with activations saved under the old ones
            raise RuntimeError("njode_b200: a parameter of the model was modified (optimizer.step(), load_state_dict(), "
                               ".to()) between NJODE.forward and the backward of its outputs")
        flat = module._flat
        dev = flat.device
        if g_loss is None and g_hT is None:
            return (None,) * (6 + len(ctx.params))
        g_loss = torch.zeros((), device=dev) if g_loss is None else g_loss.to(dev, torch.float32).contiguous()
        if g_hT is not None:
            g_hT = g_hT.to(dev, torch.float32).contiguous()
        if len(ctx.saved) and isinstance(ctx.saved[0], str):          # ("wide", blob): operand tiles of the tensor-core forward
            grads = runner.backward_wide(ctx.model_t, ctx.pb, flat, ctx.saved[1], g_loss, g_hT)
        else:
            grads = runner.backward(ctx.model_t, ctx.pb, flat, ctx.saved, g_loss, g_hT)
        if module._grad_sync is not None:
            module._grad_sync(grads)
        views = tuple(grads[o:o + n].view(s) for (o, n, s) in module._flat_layout)
        return (None,) * 6 + views


class NJODE(torch.nn.Module):
    """NJ-ODE model (NJODE/models.py:280-584) on the B200 kernels."""

    def __init__(self, input_size, hidden_size, output_size, ode_nn, readout_nn, enc_nn, use_rnn,
                 bias=True, dropout_rate=0, solver="euler", weight=0.5, weight_decay=1.,
                 **options):
        super().__init__()
        self.epoch = 1
        self.weight = weight
        self.weight_decay = weight_decay
        self.use_rnn = use_rnn
        options1 = options["options"]                      # required, NJODE/models.py:321
        self.which_loss = options1.get("which_loss", "standard")
        assert self.which_loss in LOSS_FUN_DICT
        self.residual_enc_dec = options1.get("residual_enc_dec", True)
        self.input_current_t = options1.get("input_current_t", False)
        self.masked = options1.get("masked", False)
        self.ode_f = ODEFunc(input_size, hidden_size, ode_nn, dropout_rate, bias,
                             input_current_t=self.input_current_t)
        self.encoder_map = FFNN(input_size, hidden_size, enc_nn, dropout_rate, bias,
                                masked=self.masked, residual=self.residual_enc_dec)
        self.readout_map = FFNN(hidden_size, output_size, readout_nn, dropout_rate, bias,
                                residual=self.residual_enc_dec)
        if self.use_rnn:
            self.obs_c = GRUCell(input_size, hidden_size, bias=bias)
        self.solver = solver
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.output_size = output_size
        self.bias = bias
        self.dropout_rate = float(dropout_rate)
        self._nn_desc = (ode_nn, enc_nn, readout_nn)
        self.apply(init_weights)
        # kernel-side state (not part of state_dict)
        self._flat = None
        self._flat_layout = None
        self._flat_version = 0
        self._grad_sync = None           # set by njode_b200.dist for data-parallel training
        self.output_device = "cpu"       # where loss / path_h / path_y are returned (reference: CPU)
        self.batch_size_norm = None      # global batch size under data parallelism
        self.path_id_offset = 0
        self.last_h2d_bytes = 0
        # "auto": the tcgen05 tensor-core kernels (bf16 operands, fp32 accumulation and state) serve the
        # non-masked loss/training call when every MLP is a real dense contraction (all hidden widths and
        # hidden_size >= 128); "off": always the fp32 FMA kernels; "on": whenever the model qualifies.
        self.tensor_cores = os.environ.get("NJODE_TENSOR_CORES", "auto")
        # backward of a tensor-core forward: "tcgen05" (chain + dW kernels on the spilled operand tiles) or
        # "fp32" (the FMA backward kernels re-reading h_hist)
        self.tensor_core_backward = os.environ.get("NJODE_TENSOR_CORE_BACKWARD", "tcgen05")
        self.last_forward_path = None
        # backward of the segment kernels: "off" = re-read the [S, B, H] history the forward pass wrote; "on" = save nothing
        # and recompute every segment from its checkpoint at the observation time (memory independent of S * B);
        # "auto" = recompute once the history would exceed _ext.RECOMPUTE_AUTO_BYTES
        self.recompute = os.environ.get("NJODE_RECOMPUTE", "auto")

    def _use_tensor_cores(self, runner, pb, model_t):
        mode = self.tensor_cores
        use = False
        if mode != "off" and pb.fwd.unit_kind == 1 and not pb.return_path and hasattr(runner, "wide_supported"):
            if runner.wide_supported(model_t):
                widths = [w for desc in self._nn_desc if desc is not None for (w, _) in desc]
                use = mode == "on" or (self.hidden_size >= 128 and len(widths) > 0 and min(widths) >= 128)
        self.last_forward_path = "tcgen05" if use else "fp32"
        return use

    # -- reference API ------------------------------------------------------------------------
    def weight_decay_step(self):
        """NJODE/models.py:364-367"""
        inc = (self.weight - 0.5)
        self.weight = 0.5 + inc * self.weight_decay
        return self.weight

    # -- parameter flattening -------------------------------------------------------------------
    def _kernel_params(self):
        """per kernel network (ODE, ENC, RO[, GRU_IH, GRU_HH]) the list of its affine layers as
        (weight, bias or None) parameter pairs, nn.Linear layout [out, in]"""
        nets = (self.ode_f.f, self.encoder_map.ffnn, self.readout_map.ffnn)
        out = []
        for seq in nets:
            out.append([(m.weight, m.bias) for m in seq if isinstance(m, torch.nn.Linear)])
        if self.use_rnn:
            g = self.obs_c.gru_d          # torch.nn.GRUCell: weight_ih [3H, d], weight_hh [3H, H], gates (r, z, n)
            out.append([(g.weight_ih, g.bias_ih if g.bias else None)])
            out.append([(g.weight_hh, g.bias_hh if g.bias else None)])
        return out

    def _ensure_flat(self):
        """all MLP parameters alias one contiguous fp32 buffer (re-established after .to(device))"""
        plist = []
        for lin in self._kernel_params():
            for w, b in lin:
                plist.append(w)
                if b is not None:
                    plist.append(b)
        ok = self._flat is not None and self._flat.device == plist[0].device
        if ok:
            base = self._flat.data_ptr()
            for p, (o, n, s) in zip(plist, self._flat_layout):
                if p.data_ptr() != base + 4 * o or p.dtype != torch.float32:
                    ok = False
                    break
        if not ok:
            layout, off = [], 0
            for p in plist:
                layout.append((off, p.numel(), tuple(p.shape)))
                off += p.numel()
            flat = torch.empty(off, dtype=torch.float32, device=plist[0].device)
            for p, (o, n, s) in zip(plist, layout):
                flat[o:o + n].copy_(p.data.reshape(-1))
                p.data = flat[o:o + n].view(s)
            self._flat, self._flat_layout = flat, layout
            self._flat_version += 1
        return plist

    def _model_struct(self, seed):
        ode_nn, enc_nn, readout_nn = self._nn_desc
        mt = _ext.ModelT()
        mt.input_size, mt.hidden_size, mt.output_size = self.input_size, self.hidden_size, self.output_size
        mt.masked = int(bool(self.masked))
        mt.input_current_t = int(bool(self.input_current_t))
        mt.loss_kind = _ext.LOSS_CODES[self.which_loss]
        mt.residual = int(bool(self.residual_enc_dec))
        mt.training = int(self.training)
        mt.weight = float(self.weight)
        mt.dropout_p = self.dropout_rate
        mt.dropout_seed = int(seed)
        mt.use_rnn = int(bool(self.use_rnn))
        it = iter(self._flat_layout)
        for n, (lin, desc) in enumerate(zip(self._kernel_params(), (ode_nn, enc_nn, readout_nn, None, None))):
            net = mt.net[n]
            if len(lin) > _ext.MAX_LINEAR:
                raise ValueError("at most %d Linear layers per network are supported" % _ext.MAX_LINEAR)
            net.n_linear = len(lin)
            net.dims[0] = lin[0][0].shape[1]
            for l, (w, b) in enumerate(lin):
                net.dims[l + 1] = w.shape[0]
                net.act[l] = _ext.ACT_CODES[desc[l][1]] if l < len(lin) - 1 else 0
                net.w_off[l] = next(it)[0]
                net.b_off[l] = next(it)[0] if b is not None else -1
        mt.n_params = self._flat.numel()
        return mt

    # -- the hot path ---------------------------------------------------------------------------
    def prepare_batch(self, times, time_ptr, X, obs_idx, delta_t, T, start_X, n_obs_ot,
                      return_path=False, get_loss=True, until_T=False, M=None):
        """host half of ``forward``: builds the Euler schedule and the work units, stages every input
        in one pinned block and issues one host->device copy.  The result can be fed to
        ``forward_prepared`` any number of times (inputs then stay resident in HBM)."""
        if self.solver != "euler":
            raise ValueError("Unknown solver '{}'.".format(self.solver))      # NJODE/models.py:374
        assert len(times) + 1 == len(time_ptr)                                  # NJODE/models.py:428
        if self.masked:
            assert M is not None                                                # NJODE/models.py:263
        if get_loss and n_obs_ot is None:
            raise ValueError("get_loss=True needs n_obs_ot")
        self._ensure_flat()
        runner = _ext.cuda_runner(self._flat.device)      # raises off CUDA: there is no CPU path
        # the encoder jump forgets the old hidden state -> (path, segment) units; the masked imputation and the
        # GRU jump (use_rnn) carry h through the jump -> whole-path units
        segments = (not self.masked) and (not return_path) and (not self.use_rnn)
        pb = runner.prepare(times, time_ptr, X, obs_idx, delta_t, T, start_X,
                            n_obs_ot if get_loss else None, M if self.masked else None, until_T,
                            return_path, segments, self.input_size,
                            batch_size_norm=self.batch_size_norm, path_id_offset=self.path_id_offset)
        pb.get_loss, pb.return_path, pb.runner = bool(get_loss), bool(return_path), runner
        self.last_h2d_bytes = pb.h2d_bytes
        return pb

    def forward_prepared(self, pb):
        """device half of ``forward`` (njode_forward; njode_backward through autograd)"""
        plist = self._ensure_flat()
        seed = 0
        if self.training and self.dropout_rate > 0:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        mt = self._model_struct(seed)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in plist)
        outs = _NJODEFunction.apply(self, pb.runner, pb, mt, pb.get_loss, need_grad, *plist)
        hT, loss = outs[0], outs[1]
        if pb.get_loss:
            if self.output_device == "cpu":
                loss = loss.cpu()
        else:
            loss = 0
        if pb.return_path:
            path_h, path_y = outs[2], outs[3]
            if self.output_device == "cpu":
                path_h, path_y = path_h.cpu(), path_y.cpu()
            return hT, loss, pb.sched.path_t, path_h, path_y
        return hT, loss

    def forward(self, times, time_ptr, X, obs_idx, delta_t, T, start_X, n_obs_ot,
                return_path=False, get_loss=True, until_T=False, M=None):
        """NJODE.forward (NJODE/models.py:379-518): same arguments, same returns
        ``(h, loss)`` or ``(h, loss, path_t, path_h, path_y)``; ``loss`` is the python int 0 when
        ``get_loss=False``."""
        return self.forward_prepared(self.prepare_batch(
            times, time_ptr, X, obs_idx, delta_t, T, start_X, n_obs_ot, return_path=return_path,
            get_loss=get_loss, until_T=until_T, M=M))

    def _pred_path(self, times, time_ptr, X, obs_idx, delta_t, T, start_X, M, on_device=False):
        """forward(return_path=True, get_loss=False, until_T=True) as evaluate / get_pred call it; only the predictions
        path_y go to the host (path_h, ten times larger for the demo nets, is not used by either caller).  Returns
        (path_t, path_y, prepared batch)."""
        keep = self.output_device
        self.output_device = "cuda"
        try:
            with torch.no_grad():
                pb = self.prepare_batch(times, time_ptr, X, obs_idx, delta_t, T, start_X, None, return_path=True,
                                        get_loss=False, until_T=True, M=M)
                _, _, path_t, _, path_y = self.forward_prepared(pb)
        finally:
            self.output_device = keep
        return path_t, (path_y.cpu() if (keep == "cpu" and not on_device) else path_y), pb

    def evaluate(self, times, time_ptr, X, obs_idx, delta_t, T, start_X, n_obs_ot, stockmodel,
                 cond_exp_fun_kwargs=None, diff_fun=_mean_square_diff, return_paths=False, M=None):
        """NJODE/models.py:521-562.  With the default ``diff_fun`` and a stock model that has the device kernel
        (njode_b200.stock_model: BlackScholes / Heston / HestonWOFeller / OrnsteinUhlenbeck) the analytic conditional
        expectation path and the mean square difference are computed on the device (one scalar comes back); any other
        combination takes the reference's NumPy route."""
        self.eval()
        dev_ok = (diff_fun is _mean_square_diff and not return_paths
                  and next(self.parameters()).device.type == "cuda"
                  and getattr(stockmodel, "supports_cond_exp_device", lambda d: False)(self.output_size))
        path_t, path_y, pb = self._pred_path(times, time_ptr, X, obs_idx, delta_t, T, start_X, M, on_device=dev_ok)
        if dev_ok:
            true_y = stockmodel.compute_cond_exp_device(pb, self.output_size)
            return float(((path_y.double() - true_y.double()) ** 2).mean())
        _, true_path_t, true_path_y = stockmodel.compute_cond_exp(
            times, time_ptr, X.detach().cpu().numpy(), obs_idx.detach().cpu().numpy(), delta_t, T,
            start_X.detach().cpu().numpy(), n_obs_ot.detach().cpu().numpy(), return_path=True,
            get_loss=False)
        eval_loss = diff_fun(path_y.detach().cpu().numpy(), true_path_y)
        if return_paths:
            return eval_loss, path_t, true_path_t, path_y, true_path_y
        return eval_loss

    def get_pred(self, times, time_ptr, X, obs_idx, delta_t, T, start_X, M=None):
        """NJODE/models.py:564-584"""
        self.eval()
        path_t, path_y, _ = self._pred_path(times, time_ptr, X, obs_idx, delta_t, T, start_X, M)
        return {"pred": path_y, "pred_t": path_t}
