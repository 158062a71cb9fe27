"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

CPU (NumPy, float64) restatement of the reference's Euler-Maruyama path generators and of its
collate, used to check njode_b200/csrc/njode_sde.cu.  Only tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline leg may import it.

Parity status.  The update rules below restate /root/reference/NJODE/stock_model.py line by line
(cited per function) and are PINNED against the real reference by tests/test_sde_oracle.py: fed with
the *same* normal numbers, ``euler_paths`` reproduces ``generate_paths`` of the reference bit for bit
(checked live in the build container, and through the committed fixture tests/golden/sde_ref.npz
elsewhere).  The random stream itself cannot be shared with the reference (NumPy's Mersenne
Twister, consumed in (path, step) order): the CUDA kernel and this file both use Philox-4x32-10
keyed by (seed; path id, step, coordinate), so the kernel is compared bit-for-bit (to fp64 libm
differences) with this file, and in distribution with the reference.
"""
import numpy as np

from .njode_oracle import philox4x32_10

STREAM_MASK = 0xFFFFFFFF


def philox_normals(seed, path_ids, nb_steps, dim):
    """(n1, n2) float64 [n_paths, dim, nb_steps] for steps k = 1..nb_steps: Box-Muller (cos branch) of
    Philox words (x0, x1) and (x2, x3); counter = (path lo, path hi, k, coordinate)."""
    path_ids = np.asarray(path_ids, dtype=np.uint64)
    lo = (path_ids & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None, None]
    hi = (path_ids >> np.uint64(32)).astype(np.uint32)[:, None, None]
    k = np.arange(1, nb_steps + 1, dtype=np.uint32)[None, None, :]
    j = np.arange(dim, dtype=np.uint32)[None, :, None]
    x0, x1, x2, x3 = philox4x32_10(lo, hi, k, j, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)

    def bm(a, b):
        u1 = (a.astype(np.float64) + 1.0) * (1.0 / 4294967296.0)
        u2 = b.astype(np.float64) * (1.0 / 4294967296.0)
        return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    return bm(x0, x1), bm(x2, x3)


def philox_mask(seed, path_ids, nb_steps, obs_perc):
    """observed int32 [n_paths, nb_steps+1]: column k uses word (k & 3) of Philox(counter = (path lo,
    path hi, k >> 2, 0xFFFFFFFF)); column 0 is drawn like every other column and never used as an
    observation (NJODE/data_utils.py:79-81, 292-307)."""
    path_ids = np.asarray(path_ids, dtype=np.uint64)
    lo = (path_ids & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None]
    hi = (path_ids >> np.uint64(32)).astype(np.uint32)[:, None]
    n1 = nb_steps + 1
    k4 = np.arange((n1 + 3) // 4, dtype=np.uint32)[None, :]
    w = philox4x32_10(lo, hi, k4, np.uint32(STREAM_MASK), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = np.stack(w, axis=-1).reshape(len(path_ids), -1)[:, :n1].astype(np.float64) * (1.0 / 4294967296.0)
    obs = (u < obs_perc).astype(np.int32)
    return obs


def euler_paths(model, hp, n1, n2, start_X=None, return_var=False):
    """the reference's update rules on given standard normals n1, n2 [n_paths, dim, nb_steps]
    (n2 only for the Heston models).  -> float64 [n_paths, out_dim, nb_steps + 1]"""
    n_paths, dim, steps = n1.shape
    dt = hp["maturity"] / steps
    sine = hp.get("sine_coeff")
    coeff = (lambda t: 1.0) if sine is None else (lambda t: 1.0 + np.sin(sine * t))     # stock_model.py:29-32
    S = np.empty((n_paths, dim, steps + 1))
    S[:, :, 0] = np.asarray(hp["S0"], dtype=np.float64) if start_X is None else start_X
    V = None
    sq = np.sqrt(dt)
    if model == "BlackScholes":                       # stock_model.py:356-375
        for k in range(1, steps + 1):
            dW = n1[:, :, k - 1] * sq
            x = S[:, :, k - 1]
            S[:, :, k] = x + hp["drift"] * coeff((k - 1) * dt) * x * dt + hp["volatility"] * x * dW
    elif model == "OrnsteinUhlenbeck":                # stock_model.py:397-418
        for k in range(1, steps + 1):
            dW = n1[:, :, k - 1] * sq
            x = S[:, :, k - 1]
            S[:, :, k] = x + (-hp["speed"] * coeff((k - 1) * dt) * (x - hp["mean"])) * dt + hp["volatility"] * dW
    elif model == "Heston":                           # stock_model.py:181-221 (spot uses the NEW variance)
        V = np.empty_like(S)
        V[:, :, 0] = hp["mean"]
        for k in range(1, steps + 1):
            dW = n1[:, :, k - 1] * sq
            dZ = (hp["correlation"] * n1[:, :, k - 1] + np.sqrt(1 - hp["correlation"] ** 2) * n2[:, :, k - 1]) * sq
            v = V[:, :, k - 1]
            with np.errstate(invalid="ignore"):
                V[:, :, k] = v + (-hp["speed"] * (v - hp["mean"])) * dt + hp["volatility"] * np.sqrt(v) * dZ
                x = S[:, :, k - 1]
                S[:, :, k] = x + hp["drift"] * coeff((k - 1) * dt) * x * dt + np.sqrt(V[:, :, k]) * x * dW
    elif model == "HestonWOFeller":                   # stock_model.py:288-335 (log-Euler, v+ = max(v, 0))
        V = np.empty_like(S)
        v0 = hp.get("v0")
        V[:, :, 0] = hp["mean"] if v0 is None else v0
        for k in range(1, steps + 1):
            dW = n1[:, :, k - 1] * sq
            dZ = (hp["correlation"] * n1[:, :, k - 1] + np.sqrt(1 - hp["correlation"] ** 2) * n2[:, :, k - 1]) * sq
            vp = np.maximum(V[:, :, k - 1], 0)
            S[:, :, k] = np.exp(np.log(S[:, :, k - 1]) + (hp["drift"] * coeff((k - 1) * dt) - 0.5 * vp) * dt + np.sqrt(vp) * dW)
            V[:, :, k] = V[:, :, k - 1] + (-hp["speed"] * (vp - hp["mean"])) * dt + hp["volatility"] * np.sqrt(vp) * dZ
        if hp.get("return_vol"):
            S = np.concatenate([S, V], axis=1)        # stock_model.py:329-330
    else:
        raise ValueError(model)
    if return_var:
        return S, V
    return S


def generate(model, hp, seed, first_path, n_paths, obs_perc=None):
    """what njode_sde_generate produces: (paths, observed or None, nb_obs or None)"""
    dim = int(np.size(hp["S0"]))
    ids = np.arange(first_path, first_path + n_paths, dtype=np.uint64)
    n1, n2 = philox_normals(seed, ids, hp["nb_steps"], dim)
    paths = euler_paths(model, hp, n1, n2)
    if obs_perc is None:
        return paths, None, None
    obs = philox_mask(seed, ids, hp["nb_steps"], obs_perc)
    return paths, obs, obs[:, 1:].sum(axis=1).astype(np.int32)


def collate(paths, observed, nb_obs, dt):
    """NJODE/data_utils.py:278-316 as plain loops (small cases only): rows ordered (time, batch position)."""
    B, d, n1 = paths.shape
    times, time_ptr, obs_idx, X = [], [0], [], []
    current_time = 0.0
    counter = 0
    for t in range(n1):
        if t > 0:
            current_time += dt
        if t == 0:
            continue
        any_obs = False
        for i in range(B):
            if observed[i, t] == 1:
                counter += 1
                X.append(paths[i, :, t])
                obs_idx.append(i)
                any_obs = True
        if any_obs:
            times.append(current_time)
            time_ptr.append(counter)
    return {"times": np.array(times), "time_ptr": np.array(time_ptr), "obs_idx": np.array(obs_idx, dtype=np.int64),
            "start_X": paths[:, :, 0].astype(np.float32), "n_obs_ot": np.asarray(nb_obs),
            "X": np.array(X, dtype=np.float32).reshape(len(obs_idx), d)}
