"""Drop-in for the part of the reference's ``NJODE/stock_model.py`` that sits on the hot path:

* ``generate_paths`` of BlackScholes / OrnsteinUhlenbeck / Heston / HestonWOFeller
  (NJODE/stock_model.py:356-375, 397-418, 181-221, 288-335) runs as ONE CUDA kernel
  (``njode_sde_generate`` in njode_b200/csrc/njode_sde.cu, Philox-4x32-10, one subsequence per global
  path id) instead of a Python loop over paths x steps.  ``generate_paths()`` keeps the reference's
  return contract ``(paths float64 numpy [nb_paths, dim, nb_steps+1], dt)``;
  ``generate_paths_device()`` returns device tensors (paths, observed, nb_obs) without a host copy.
* ``next_cond_exp`` / ``compute_cond_exp`` / ``get_optimal_loss`` (NJODE/stock_model.py:50-158 and the
  per-model formulas) stay host NumPy: they are the cheap analytic oracle of the training loop
  (SURVEY.md §8 a15), O(B x steps) element-wise work.

There is no CPU generator: without the CUDA library / a CUDA device ``generate_paths`` raises.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _ext

SDE_CODES = {"BlackScholes": 0, "OrnsteinUhlenbeck": 1, "Heston": 2, "HestonWOFeller": 3}


def _bind(lib):
    d = lib.dll
    if getattr(d, "_sde_bound", False):
        return d
    d.njode_sde_generate.argtypes = [C.POINTER(_ext.SdeT), C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    d.njode_sde_generate.restype = C.c_int
    d.njode_collate.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    d.njode_collate.restype = C.c_int
    d.njode_cond_exp.argtypes = [C.POINTER(_ext.SdeT), C.POINTER(_ext.BatchT), C.c_void_p, C.c_void_p]
    d.njode_cond_exp.restype = C.c_int
    d._sde_bound = True
    return d


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def compute_loss(X_obs, Y_obs, Y_obs_bj, n_obs_ot, batch_size, eps=1e-10, weight=0.5):
    """NJODE/stock_model.py:471-481"""
    inner = (2 * weight * np.sqrt(np.sum((X_obs - Y_obs) ** 2, axis=1) + eps) +
             2 * (1 - weight) * np.sqrt(np.sum((Y_obs_bj - Y_obs) ** 2, axis=1) + eps)) ** 2
    return np.sum(inner / n_obs_ot) / batch_size


def _gen_kw(kwargs):
    """generator-only knobs of this implementation; every other extra kwarg is ignored like in the reference"""
    return {k: kwargs[k] for k in ("seed", "first_path", "device") if k in kwargs}


class StockModel:
    """NJODE/stock_model.py:15-158"""
    model_name = None

    def __init__(self, drift, volatility, S0, nb_paths, nb_steps, maturity, sine_coeff, **kwargs):
        self.drift = drift
        self.volatility = volatility
        self.S0 = S0
        self.nb_paths = nb_paths
        self.nb_steps = nb_steps
        self.maturity = maturity
        self.dimensions = np.size(S0)
        self.sine_coeff = sine_coeff
        if sine_coeff is None:
            self.periodic_coeff = lambda t: 1
        else:
            self.periodic_coeff = lambda t: (1 + np.sin(sine_coeff * t))
        # generator-only knobs of this implementation (not in the reference's constructor)
        self.seed = int(kwargs.get("seed", 0))
        self.first_path = int(kwargs.get("first_path", 0))
        self.device = kwargs.get("device", "cuda")

    # ---- generators --------------------------------------------------------------------------
    def _sde_struct(self, obs_perc=0.0):
        p = _ext.SdeT()
        p.model = SDE_CODES[self.model_name]
        p.dimension = int(self.dimensions)
        p.nb_steps = int(self.nb_steps)
        p.return_vol = int(bool(getattr(self, "retur_vol", False)))
        p.drift = float(self.drift) if self.drift is not None else 0.0
        p.volatility = float(self.volatility)
        p.mean = float(getattr(self, "mean", 0.0))
        p.speed = float(getattr(self, "speed", 0.0))
        p.correlation = float(getattr(self, "correlation", 0.0))
        p.v0 = float(getattr(self, "v0", 0.0))
        p.maturity = float(self.maturity)
        p.sine_coeff = float("nan") if self.sine_coeff is None else float(self.sine_coeff)
        p.obs_perc = float(obs_perc)
        p.t0 = 0.0
        p.seed = int(self.seed) & 0xFFFFFFFFFFFFFFFF
        return p

    def generate_paths_device(self, start_X=None, obs_perc=None, nb_paths=None, first_path=None, device=None):
        """-> (paths f64 [n, out_dim, steps+1], observed i32 [n, steps+1] or None, nb_obs i32 [n] or None, dt),
        all on the device.  ``observed``/``nb_obs`` are produced when ``obs_perc`` is given
        (NJODE/data_utils.py:79-81: every column incl. column 0 is Bernoulli(obs_perc), nb_obs counts columns >= 1)."""
        device = torch.device(device or self.device)
        if device.type != "cuda" or not torch.cuda.is_available():
            raise _ext.NjodeError("njode_b200.stock_model: the generators run on CUDA devices only (no CPU fallback)")
        d = _bind(_ext.cuda_lib())
        n = int(self.nb_paths if nb_paths is None else nb_paths)
        first = int(self.first_path if first_path is None else first_path)
        dim = int(self.dimensions)
        out_dim = dim * (2 if getattr(self, "retur_vol", False) else 1)
        with torch.cuda.device(device):
            if start_X is not None:
                s0 = torch.as_tensor(np.asarray(start_X, dtype=np.float64).reshape(n, dim)).to(device)
                per_path = 1
            else:
                s0 = torch.as_tensor(np.broadcast_to(np.asarray(self.S0, dtype=np.float64).reshape(-1), (dim,)).copy()).to(device)
                per_path = 0
            paths = torch.empty(n, out_dim, self.nb_steps + 1, dtype=torch.float64, device=device)
            observed = nb_obs = None
            if obs_perc is not None:
                observed = torch.empty(n, self.nb_steps + 1, dtype=torch.int32, device=device)
                nb_obs = torch.empty(n, dtype=torch.int32, device=device)
            p = self._sde_struct(0.0 if obs_perc is None else obs_perc)
            rc = d.njode_sde_generate(C.byref(p), first, n, C.c_void_p(s0.data_ptr()), per_path,
                                      C.c_void_p(paths.data_ptr()),
                                      None if observed is None else C.c_void_p(observed.data_ptr()),
                                      None if nb_obs is None else C.c_void_p(nb_obs.data_ptr()), _stream(device))
            _ext.cuda_lib().check(rc, "njode_sde_generate")
        return paths, observed, nb_obs, self.maturity / self.nb_steps

    def generate_paths(self, start_X=None):
        """reference contract: (np.float64 [nb_paths, data_dim, nb_steps+1], dt)"""
        paths, _, _, dt = self.generate_paths_device(start_X=start_X)
        return paths.cpu().numpy(), dt

    # ---- analytic conditional expectation ---------------------------------------------------------
    def supports_cond_exp_device(self, d):
        """True when ``compute_cond_exp_device`` serves batches of data dimension ``d`` (not with func_appl_X features,
        not for the regime-switching Combined model, whose conditional expectation restarts per regime)"""
        if self.model_name not in SDE_CODES:
            return False
        return int(self.dimensions) * (2 if getattr(self, "retur_vol", False) else 1) == int(d)

    def compute_cond_exp_device(self, pb, d):
        """path_y of ``compute_cond_exp`` as a device tensor [E, B, d] (njode_cond_exp) for the prepared batch ``pb``
        of a forward(return_path=True, until_T=True) call -- same records, same order as the model's path_y, so
        NJODE.evaluate (NJODE/models.py:551-558) can take the mean square difference without leaving the device."""
        lib = _ext.cuda_lib()
        dll = _bind(lib)
        if not pb.return_path or pb.dev.type != "cuda":
            raise _ext.NjodeError("compute_cond_exp_device needs the prepared batch of a return_path call on a CUDA device")
        sde = self._sde_struct()
        if int(sde.dimension) * (2 if sde.return_vol else 1) != int(d):
            raise ValueError("data dimension of the batch does not match the stock model")
        out = torch.empty(pb.sched.E, pb.B, d, dtype=torch.float32, device=pb.dev)
        with torch.cuda.device(pb.dev):
            rc = dll.njode_cond_exp(C.byref(sde), C.byref(pb.fwd), C.c_void_p(out.data_ptr()), _stream(pb.dev))
        lib.check(rc, "njode_cond_exp")
        return out

    def _decay(self, step_t, step_dt, d):
        """the analytic conditional expectation of every model moves each coordinate towards a fixed point at an exponential
        rate: y(t + s) = fp + (y(t) - fp) exp(rate(t) s).  Returns (rate * step_dt as [S, d], fp as [d]) for Euler steps
        starting at ``step_t`` of length ``step_dt``; subclasses define it from their ``next_cond_exp`` formula."""
        raise ValueError("not implemented yet")          # same error as the reference's abstract next_cond_exp

    def _coeff(self, t):
        """periodic_coeff on an array of times (NJODE/stock_model.py:29-32)"""
        t = np.asarray(t, dtype=np.float64)
        return np.ones_like(t) if self.sine_coeff is None else 1 + np.sin(self.sine_coeff * t)

    def next_cond_exp(self, y, delta_t, current_t):
        """E[X_{t + delta_t} | X_t = y] (NJODE/stock_model.py:178-179, 277-286, 353-354, 393-395)"""
        la, fp = self._decay(np.array([current_t], dtype=np.float64), np.array([delta_t], dtype=np.float64), np.shape(y)[1])
        return fp + (y - fp) * np.exp(la[0])

    def compute_cond_exp(self, times, time_ptr, X, obs_idx, delta_t, T, start_X, n_obs_ot,
                         return_path=True, get_loss=False, weight=0.5, start_time=None, **kwargs):
        """StockModel.compute_cond_exp (NJODE/stock_model.py:50-151): the analytic counterpart of NJODE.forward --
        between observations every path follows its model's conditional expectation, at an observation the observed
        paths are reset to the observed value.  Returns ``loss`` or ``(loss, path_t, path_y [E, B, d])`` like the
        reference.

        Evaluated in closed form over the event list (schedule.cond_exp_schedule) instead of stepping: with
        P[k] = sum of the first k log-factors, the value of a path at a record is
        fp + (x_a - fp) exp(P[k_record] - P[k_a]) where (x_a, k_a) is the path's latest observation (or the start)
        -- one gather and one exp for the whole [E, B, d] block.  One deviation from the reference: its tail loop calls
        ``next_cond_exp`` without the current time (stock_model.py:139) and raises TypeError whenever time remains after
        the last observation; here the tail uses the current time like every other step."""
        from . import schedule as _sched
        X, start_X = np.asarray(X), np.asarray(start_X)
        obs_idx, time_ptr = np.asarray(obs_idx, dtype=np.int64), np.asarray(time_ptr, dtype=np.int64)
        B, d = start_X.shape
        cs = _sched.cond_exp_schedule(times, delta_t, T, start_time)
        la, fp = self._decay(cs.step_t, cs.step_dt, d)
        P = np.zeros((len(cs.step_dt) + 1, d))
        np.cumsum(la, axis=0, out=P[1:])
        # observation rows of the processed observation times, in processing order
        J = len(cs.jump_index)
        lo = time_ptr[cs.jump_index] if J else np.zeros(0, dtype=np.int64)
        cnt = (time_ptr[cs.jump_index + 1] - lo) if J else lo
        first = np.cumsum(cnt) - cnt
        row_j = np.repeat(np.arange(J), cnt)
        rows = np.arange(int(cnt.sum())) - first[row_j] + lo[row_j]
        row_p = obs_idx[rows]
        # anchor[j, b]: position (into rows) of path b's latest observation among the first j observation times, -1 = start
        anchor = np.full((J + 1, B), -1, dtype=np.int64)
        anchor[row_j + 1, row_p] = np.arange(len(rows))
        np.maximum.accumulate(anchor, axis=0, out=anchor)

        def value_at(pos, paths, k):
            """conditional expectation after k Euler steps of the given paths, anchored at rows[pos] (or the start)"""
            has = pos >= 0
            safe = np.where(has, pos, 0)
            xa = np.where(has[..., None], X[rows[safe]] if len(rows) else 0.0, start_X[paths]).astype(np.float64)
            ka = np.where(has, cs.jump_step[row_j[safe]] if len(rows) else 0, 0)
            k = np.broadcast_to(k, ka.shape)
            y = fp + (xa - fp) * np.exp(P[k] - P[ka])
            return np.where((k == ka)[..., None], xa, y)          # an observation is reproduced exactly

        loss = 0
        if get_loss and len(rows):
            # compute_loss (NJODE/stock_model.py:471-481) with Y = X at the observed rows and Y_bj = value before the reset
            ybj = value_at(anchor[row_j, row_p], row_p, cs.jump_step[row_j])
            eps = 1e-10
            inner = (2 * weight * np.sqrt(eps) +
                     2 * (1 - weight) * np.sqrt(np.sum((ybj - X[rows]) ** 2, axis=1) + eps)) ** 2
            loss = np.sum(inner / np.asarray(n_obs_ot)[row_p]) / B
        if not return_path:
            return loss
        path_y = value_at(anchor[cs.rec_jumps], np.broadcast_to(np.arange(B), (len(cs.rec_t), B)), cs.rec_step[:, None])
        return loss, cs.rec_t, path_y

    def get_optimal_loss(self, times, time_ptr, X, obs_idx, delta_t, T, start_X, n_obs_ot, weight=0.5):
        """NJODE/stock_model.py:153-158: the loss of the true conditional expectation"""
        return self.compute_cond_exp(times, time_ptr, X, obs_idx, delta_t, T, start_X, n_obs_ot,
                                     return_path=False, get_loss=True, weight=weight)


class Heston(StockModel):
    """NJODE/stock_model.py:161-221"""
    model_name = "Heston"

    def __init__(self, drift, volatility, mean, speed, correlation, nb_paths, nb_steps, S0, maturity,
                 sine_coeff=None, **kwargs):
        super().__init__(drift=drift, volatility=volatility, nb_paths=nb_paths, nb_steps=nb_steps, S0=S0,
                         maturity=maturity, sine_coeff=sine_coeff, **_gen_kw(kwargs))
        self.mean = mean
        self.speed = speed
        self.correlation = correlation

    def _decay(self, step_t, step_dt, d):
        """NJODE/stock_model.py:178-179: y exp(drift c(t) s)"""
        return np.repeat((self.drift * self._coeff(step_t) * step_dt)[:, None], d, axis=1), np.zeros(d)


class HestonWOFeller(StockModel):
    """NJODE/stock_model.py:250-335"""
    model_name = "HestonWOFeller"

    def __init__(self, drift, volatility, mean, speed, correlation, nb_paths, nb_steps, S0, maturity,
                 scheme='euler', return_vol=False, v0=None, sine_coeff=None, **kwargs):
        super().__init__(drift=drift, volatility=volatility, nb_paths=nb_paths, nb_steps=nb_steps, S0=S0,
                         maturity=maturity, sine_coeff=sine_coeff, **_gen_kw(kwargs))
        self.mean = mean
        self.speed = speed
        self.correlation = correlation
        if scheme != 'euler':
            raise ValueError('unknown sampling scheme')
        self.scheme = scheme
        self.retur_vol = return_vol
        self.v0 = self.mean if v0 is None else v0

    def _decay(self, step_t, step_dt, d):
        """NJODE/stock_model.py:277-286: spot coordinates y exp(drift c(t) s); with return_vol the second half of the
        coordinates is the variance, mean reverting at ``speed`` (no periodic coefficient)"""
        la = np.repeat((self.drift * self._coeff(step_t) * step_dt)[:, None], d, axis=1)
        fp = np.zeros(d)
        if self.retur_vol:
            la[:, d // 2:] = (-self.speed * np.asarray(step_dt, dtype=np.float64))[:, None]
            fp[d // 2:] = self.mean
        return la, fp


class BlackScholes(StockModel):
    """NJODE/stock_model.py:340-375"""
    model_name = "BlackScholes"

    def __init__(self, drift, volatility, nb_paths, nb_steps, S0, maturity, sine_coeff=None, **kwargs):
        super().__init__(drift=drift, volatility=volatility, nb_paths=nb_paths, nb_steps=nb_steps, S0=S0,
                         maturity=maturity, sine_coeff=sine_coeff, **_gen_kw(kwargs))

    def _decay(self, step_t, step_dt, d):
        """NJODE/stock_model.py:353-354: y exp(drift c(t) s)"""
        return np.repeat((self.drift * self._coeff(step_t) * step_dt)[:, None], d, axis=1), np.zeros(d)


class OrnsteinUhlenbeck(StockModel):
    """NJODE/stock_model.py:378-418"""
    model_name = "OrnsteinUhlenbeck"

    def __init__(self, volatility, nb_paths, nb_steps, S0, mean, speed, maturity, sine_coeff=None, **kwargs):
        super().__init__(volatility=volatility, nb_paths=nb_paths, drift=None, nb_steps=nb_steps, S0=S0,
                         maturity=maturity, sine_coeff=sine_coeff, **_gen_kw(kwargs))
        self.mean = mean
        self.speed = speed

    def _decay(self, step_t, step_dt, d):
        """NJODE/stock_model.py:393-395: mean reversion, y e + mean (1 - e) with e = exp(-speed c(t) s)"""
        return np.repeat((-self.speed * self._coeff(step_t) * step_dt)[:, None], d, axis=1), np.full(d, float(self.mean))


class Combined(StockModel):
    """NJODE/stock_model.py:421-468"""

    def __init__(self, stock_model_names, hyperparam_dicts, **kwargs):
        self.stock_model_names = stock_model_names
        self.hyperparam_dicts = hyperparam_dicts

    def compute_cond_exp(self, times, time_ptr, X, obs_idx, delta_t, T, start_X, n_obs_ot,
                         return_path=True, get_loss=False, weight=0.5, **kwargs):
        """NJODE/stock_model.py:426-460: one regime after the other; regime i covers (end of regime i-1, + its own
        maturity], starts from the last conditional expectation of the previous regime and only looks at the observation
        times after its start (``start_time``)."""
        loss, horizon, ts, ys = 0, 0., [], []
        for name, hp in zip(self.stock_model_names, self.hyperparam_dicts):
            regime = STOCK_MODELS[name](**hp)
            horizon += hp['maturity']
            part = regime.compute_cond_exp(
                times, time_ptr, X, obs_idx, delta_t, horizon, ys[-1][-1] if ys else start_X, n_obs_ot,
                return_path=True, get_loss=get_loss, weight=weight, start_time=ts[-1][-1] if ts else None)
            loss += part[0]
            ts.append(part[1])
            ys.append(part[2])
        if return_path:
            return loss, np.concatenate(ts), np.concatenate(ys, axis=0)
        return loss


STOCK_MODELS = {                           # NJODE/stock_model.py:486-495
    "BlackScholes": BlackScholes,
    "Heston": Heston,
    "OrnsteinUhlenbeck": OrnsteinUhlenbeck,
    "HestonWOFeller": HestonWOFeller,
    "combined": Combined,
    "sine_BlackScholes": BlackScholes,
    "sine_Heston": Heston,
    "sine_OrnsteinUhlenbeck": OrnsteinUhlenbeck,
}


# ------------------------------------------------------------------------------------------------
# device-resident dataset + on-device collate (NJODE/data_utils.py:59-108, 278-316)
# ------------------------------------------------------------------------------------------------
class DeviceDataset:
    """paths / observation mask generated on the device and kept there; ``collate(sel)`` builds one
    batch of the reference's collate contract with ``njode_collate`` (rows ordered time ascending,
    batch position ascending) and returns the same dict as ``custom_collate_fn`` -- ``X`` / ``start_X``
    stay on the device, the small index arrays come back to the host because the reference contract
    holds them there (NJODE/train.py:493-507)."""

    def __init__(self, stock_model_name, hyperparam_dict, seed=0, first_path=0, device="cuda"):
        hp = dict(hyperparam_dict)
        self.obs_perc = hp['obs_perc']
        self.hyperparam_dict = hp
        self.model = STOCK_MODELS[stock_model_name](**hp, seed=seed, first_path=first_path, device=device)
        self.paths, self.observed, self.nb_obs, self.dt = self.model.generate_paths_device(obs_perc=self.obs_perc)
        self.device = self.paths.device
        self.n, self.dim, n1 = self.paths.shape
        self.nb_steps = n1 - 1
        self._ws = None

    def __len__(self):
        return self.n

    def collate_device(self, sel):
        """device-side batch: dict of device tensors (X, obs_idx i32, time_ptr i32, time_idx i32, start_X,
        n_obs_ot i32) sized for the worst case plus ``counts`` = [K, N] (device)."""
        d = _bind(_ext.cuda_lib())
        sel = torch.as_tensor(sel, dtype=torch.int64, device=self.device).contiguous()
        B = int(sel.numel())
        nw = (B + 31) // 32
        need = (self.nb_steps * nw + 3 * self.nb_steps + 16) * 4
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        nmax = B * self.nb_steps
        i32 = dict(dtype=torch.int32, device=self.device)
        out = {"X": torch.empty(nmax, self.dim, dtype=torch.float32, device=self.device),
               "obs_idx": torch.empty(nmax, **i32), "time_ptr": torch.empty(self.nb_steps + 1, **i32),
               "time_idx": torch.empty(self.nb_steps, **i32),
               "start_X": torch.empty(B, self.dim, dtype=torch.float32, device=self.device),
               "n_obs_ot": torch.empty(B, **i32), "counts": torch.empty(2, **i32)}
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            rc = d.njode_collate(p(self.paths), p(self.observed), self.n, self.dim, self.nb_steps, p(sel), B,
                                 p(out["X"]), p(out["obs_idx"]), p(out["time_ptr"]), p(out["time_idx"]),
                                 p(out["start_X"]), p(out["n_obs_ot"]), p(out["counts"]), p(self._ws),
                                 self._ws.numel(), _stream(self.device))
        _ext.cuda_lib().check(rc, "njode_collate")
        return out

    @staticmethod
    def _apply_functions(X, func_names):
        """CustomCollateFnGen (NJODE/data_utils.py:352-416): append f(X) for every supported function name ('exp',
        'power-x') along the feature axis -- on the device, the values never visit the host"""
        if not func_names:
            return X
        cols = [X]
        for name in func_names:
            if name in ("exp", "exponential"):
                cols.append(torch.exp(X.double()).float())
            elif "power-" in name:
                cols.append(torch.pow(X.double(), float(name.split("-")[1])).float())
        return torch.cat(cols, dim=1)

    def collate(self, sel, func_names=None, on_device=False):
        """the reference's collate dict (NJODE/data_utils.py:311-315); ``func_names`` = the 'func_appl_X' option of
        train.py (e.g. ["power-2"]: the model then also learns the conditional second moment).

        ``on_device=True``: ``time_ptr`` / ``obs_idx`` / ``n_obs_ot`` stay device tensors (int32) that
        ``NJODE.forward`` / ``prepare_batch`` take as they are -- the only device->host traffic of the batch is the
        8-byte (K, N) pair, plus the K grid indices when some grid time has no observation in the batch (``times`` is
        host data by contract: the Euler schedule is built from it in float64 on the host)."""
        o = self.collate_device(sel)
        o["X"] = self._apply_functions(o["X"], func_names)
        o["start_X"] = self._apply_functions(o["start_X"], func_names)
        K, N = (int(v) for v in o["counts"].cpu())
        # current_time += dt once per grid step in float64 (data_utils.py:293-296)
        grid_t = np.cumsum(np.full(self.nb_steps, self.dt, dtype=np.float64))
        if on_device:
            times = grid_t if K == self.nb_steps else grid_t[o["time_idx"][:K].cpu().numpy() - 1]
            return {"times": times, "time_ptr": o["time_ptr"][:K + 1], "obs_idx": o["obs_idx"][:N], "start_X": o["start_X"],
                    "n_obs_ot": o["n_obs_ot"], "X": o["X"][:N], "true_paths": None, "observed_dates": None}
        tidx = o["time_idx"][:K].cpu().numpy()
        return {"times": grid_t[tidx - 1], "time_ptr": o["time_ptr"][:K + 1].cpu().numpy().astype(np.int64),
                "obs_idx": o["obs_idx"][:N].cpu().to(torch.int64), "start_X": o["start_X"],
                "n_obs_ot": o["n_obs_ot"].cpu().to(torch.int64), "X": o["X"][:N],
                "true_paths": None, "observed_dates": None}
