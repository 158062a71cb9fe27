"""Host-side schedule and work-unit builder.

The reference decides *on the host, in float64* (float32 when ``times`` is float32, NumPy >= 2
weak-scalar promotion) how many Euler steps every batch takes and how long each one is
(NJODE/models.py:430-439 and 497-505).  ``build_schedule`` executes exactly those expressions, so
the step list -- and therefore ``path_t`` -- is equal to the reference's by construction; the
device only ever sees the fp32-rounded scalars the reference's tensor ops would see.

``build_units`` turns the collate contract's (time_ptr, obs_idx) (NJODE/data_utils.py:292-315)
into a per-path CSR of observation rows and the list of work units the kernels march:
whole paths (masked model / return_path), or (path, inter-observation segment) pairs for the
non-masked model, where the state after a jump does not depend on the state before it
(NJODE/models.py:463-470) so every segment is independent.
"""
import collections

import numpy as np
import torch

UNIT_WRITES_HT = 1 << 30

_Schedule = collections.namedtuple(
    "_Schedule", "S K E step_dt step_t jump_step jump_tau step_event jump_event path_t")
_cache = collections.OrderedDict()
_CACHE_MAX = 16


def build_schedule(times, delta_t, T, until_T, return_path):
    """event list of one forward call.  step k: dt = step_dt[k] (fl32 of delta_t_), time at the
    start = step_t[k]; jump i happens after jump_step[i] steps; *_event = index into path_t."""
    times = np.asarray(times)
    key = (times.tobytes(), times.dtype.str, float(delta_t), float(T), bool(until_T), bool(return_path))
    hit = _cache.get(key)
    if hit is not None:
        _cache.move_to_end(key)
        return hit
    step_dt, step_t, step_event, jump_step, jump_event = [], [], [], [], []
    path_t = [0]
    current_time = 0.0
    for obs_time in times:
        # NJODE/models.py:432-439 (expressions kept verbatim, incl. operand types)
        while current_time < (obs_time - 1e-10 * delta_t):
            if current_time < obs_time - delta_t:
                delta_t_ = delta_t
            else:
                delta_t_ = obs_time - current_time
            step_t.append(current_time)
            step_dt.append(delta_t_)
            current_time = current_time + delta_t_          # ode_step, NJODE/models.py:376
            step_event.append(len(path_t))
            path_t.append(current_time)
        jump_step.append(len(step_dt))
        jump_event.append(len(path_t))
        path_t.append(obs_time)
    if until_T:                                              # NJODE/models.py:497-505
        while current_time < T - 1e-10 * delta_t:
            if current_time < T - delta_t:
                delta_t_ = delta_t
            else:
                delta_t_ = T - current_time
            step_t.append(current_time)
            step_dt.append(delta_t_)
            current_time = current_time + delta_t_
            step_event.append(len(path_t))
            path_t.append(current_time)
    sched = _Schedule(
        S=len(step_dt), K=len(times), E=len(path_t) if return_path else 0,
        step_dt=np.array(step_dt, dtype=np.float64).astype(np.float32),
        step_t=np.array(step_t, dtype=np.float64).astype(np.float32),
        jump_step=np.array(jump_step, dtype=np.int32),
        jump_tau=times.astype(np.float64).astype(np.float32),     # NJODE/models.py:487
        step_event=np.array(step_event, dtype=np.int32),
        jump_event=np.array(jump_event, dtype=np.int32),
        path_t=np.array(path_t) if return_path else None)
    _cache[key] = sched
    if len(_cache) > _CACHE_MAX:
        _cache.popitem(last=False)
    return sched


_CondExpSchedule = collections.namedtuple(
    "_CondExpSchedule", "step_dt step_t jump_index jump_step rec_t rec_step rec_jumps")


def cond_exp_schedule(times, delta_t, T, start_time=None):
    """event list of ``StockModel.compute_cond_exp`` (NJODE/stock_model.py:75-146).  Same Euler marching rule as
    ``build_schedule`` but its own observation filter: an observation time beyond ``T + 1e-10`` ends the list, one at or
    before the current time is skipped (the regime-switching ``Combined`` model restarts the loop at ``start_time``),
    and the march always continues to ``T``.  Returns, all float64 / int64 NumPy arrays:
      step_dt, step_t [S]     length of step k and the time at its start
      jump_index, jump_step   [J] index into ``times`` of every processed observation time, steps done before it
      rec_t, rec_step, rec_jumps [E]  the records of ``path_t``: time stamp, steps done, jumps applied at that record
    """
    current_time = start_time if start_time else 0.0
    step_dt, step_t, jump_index, jump_step = [], [], [], []
    rec_t, rec_step, rec_jumps = [], [], []
    if not start_time:
        rec_t.append(0.); rec_step.append(0); rec_jumps.append(0)

    def march(target, current_time):
        while current_time < (target - 1e-10 * delta_t):
            delta_t_ = delta_t if current_time < target - delta_t else target - current_time
            step_t.append(current_time)
            step_dt.append(delta_t_)
            current_time = current_time + delta_t_
            rec_t.append(current_time); rec_step.append(len(step_dt)); rec_jumps.append(len(jump_step))
        return current_time

    for i, obs_time in enumerate(times):
        if obs_time > T + 1e-10:
            break
        if obs_time <= current_time:
            continue
        current_time = march(obs_time, current_time)
        jump_index.append(i)
        jump_step.append(len(step_dt))
        rec_t.append(obs_time); rec_step.append(len(step_dt)); rec_jumps.append(len(jump_step))
    march(T, current_time)
    f64, i64 = np.float64, np.int64
    return _CondExpSchedule(np.array(step_dt, dtype=f64), np.array(step_t, dtype=f64), np.array(jump_index, dtype=i64),
                            np.array(jump_step, dtype=i64), np.array(rec_t, dtype=f64), np.array(rec_step, dtype=i64),
                            np.array(rec_jumps, dtype=i64))


def build_csr(time_ptr, obs_idx, B):
    """rows of every path in time order.  obs_idx: int64 numpy [N]; rows are time-major already
    (NJODE/data_utils.py:298-307), so a stable sort by path keeps each path's rows time-ordered."""
    time_ptr = np.asarray(time_ptr, dtype=np.int64)
    obs_idx = np.asarray(obs_idx, dtype=np.int64)
    K = len(time_ptr) - 1
    N = int(time_ptr[-1]) if K >= 0 and len(time_ptr) else 0
    if len(obs_idx) != N:
        raise AssertionError("len(obs_idx) != time_ptr[-1]")
    if N and (obs_idx.min() < 0 or obs_idx.max() >= B):
        raise IndexError("obs_idx out of range")
    row_jump = np.repeat(np.arange(K, dtype=np.int32), np.diff(time_ptr))
    path_rows = _stable_argsort_small(obs_idx, B).astype(np.int32)
    path_ptr = np.zeros(B + 1, dtype=np.int32)
    np.cumsum(np.bincount(obs_idx, minlength=B), out=path_ptr[1:])
    if N > 1:
        # the contract has at most one row per (time, path) (NJODE/data_utils.py:302-306): inside a
        # path the observation-time index must strictly increase
        rj = row_jump[path_rows]
        same_path = np.ones(N - 1, dtype=bool)
        same_path[path_ptr[1:-1][(path_ptr[1:-1] > 0) & (path_ptr[1:-1] < N)] - 1] = False
        if np.any((rj[1:] == rj[:-1]) & same_path):
            raise ValueError("a path has two observation rows at the same observation time")
    return path_ptr, path_rows, row_jump


def _stable_argsort_small(keys, bound):
    """stable argsort of non-negative integer keys < bound; numpy's stable sort is an O(N) radix sort
    for 16-bit keys, so use it whenever the keys fit"""
    if bound <= 65536:
        return np.argsort(keys.astype(np.uint16), kind="stable")
    return np.argsort(keys, kind="stable")


def build_units(sched, path_ptr, path_rows, row_jump, B, segments):
    """returns (units [n,6] int32, n_loss_units).  Layout of a unit: see include/njode_b200.h.
    The first n_loss_units units are the ones that contribute to the loss (sorted longest first);
    the remaining ones only produce hT (the tail after a path's last observation)."""
    S = sched.S
    N = len(path_rows)
    if not segments:
        units = np.empty((B, 6), dtype=np.int32)
        units[:, 0] = np.arange(B)
        units[:, 1] = 0
        units[:, 2] = S
        units[:, 3] = path_ptr[:-1]
        units[:, 4] = path_ptr[1:]
        units[:, 5] = UNIT_WRITES_HT
        return units, B
    q = np.arange(N, dtype=np.int64)
    sorted_path = np.repeat(np.arange(B, dtype=np.int32), np.diff(path_ptr))
    js = sched.jump_step[row_jump[path_rows]] if N else np.zeros(0, dtype=np.int32)
    first = np.ones(N, dtype=bool)
    if N:
        first[1:] = sorted_path[1:] != sorted_path[:-1]
    prev_js = np.where(first, 0, np.concatenate(([0], js[:-1]))) if N else js
    prev_row = np.where(first, -1, np.concatenate(([-1], path_rows[:-1]))) if N else path_rows
    loss_units = np.empty((N, 6), dtype=np.int32)
    loss_units[:, 0] = sorted_path
    loss_units[:, 1] = prev_js
    loss_units[:, 2] = js
    loss_units[:, 3] = q
    loss_units[:, 4] = q + 1
    loss_units[:, 5] = prev_row + 1
    tails = np.empty((B, 6), dtype=np.int32)
    has = path_ptr[1:] > path_ptr[:-1]
    last_q = np.maximum(path_ptr[1:] - 1, 0)
    tails[:, 0] = np.arange(B)
    tails[:, 1] = np.where(has, js[last_q] if N else 0, 0)
    tails[:, 2] = S
    tails[:, 3] = path_ptr[1:]
    tails[:, 4] = path_ptr[1:]
    tails[:, 5] = (np.where(has, path_rows[last_q] + 1 if N else 0, 0)) | UNIT_WRITES_HT
    # longest first (descending length, stable): ascending sort of (S - length)
    o1 = _stable_argsort_small(S - (loss_units[:, 2] - loss_units[:, 1]), S + 1)
    o2 = _stable_argsort_small(S - (tails[:, 2] - tails[:, 1]), S + 1)
    return np.concatenate((loss_units[o1], tails[o2]), axis=0), N


def build_index_torch(obs, time_ptr, jump_step, B, S, segments, T1, T2):
    """device-side (any torch device) version of build_csr + build_units: the per-path CSR of observation
    rows and the work units, built with a handful of tensor ops (stable radix sorts, prefix sums) right
    where the batch lives, so the host ships only the raw collate arrays.

    obs [N] path index of every row (time-major rows, NJODE/data_utils.py:298-307), time_ptr [K+1],
    jump_step [K] (host schedule).  Returns (path_ptr, path_rows, row_jump, unit_desc) as int32 tensors,
    n_loss_units, and ``stats`` (int64 [6] on the device): units with length >= T1 / >= T2 in the loss
    run and in the tail run, a duplicate-(time, path) flag and an out-of-range-index flag."""
    dev = obs.device
    N = int(obs.numel())
    i64 = dict(dtype=torch.int64, device=dev)
    obs = obs.to(torch.int64)
    zero = torch.zeros((), **i64)
    bad = ((obs < 0) | (obs >= B)).any().to(torch.int64) if N else zero
    obs_c = obs.clamp(0, max(B - 1, 0))
    counts = torch.bincount(obs_c, minlength=B)[:B] if N else torch.zeros(B, **i64)
    path_ptr = torch.zeros(B + 1, **i64)
    torch.cumsum(counts, 0, out=path_ptr[1:])
    sorted_path, path_rows = torch.sort(obs_c, stable=True)
    r = torch.arange(N, **i64)
    row_jump = torch.searchsorted(time_ptr.to(torch.int64), r, right=True) - 1
    rj_sorted = row_jump[path_rows]
    first = torch.ones(N, dtype=torch.bool, device=dev)
    if N > 1:
        first[1:] = sorted_path[1:] != sorted_path[:-1]
        dup = ((rj_sorted[1:] == rj_sorted[:-1]) & ~first[1:]).any().to(torch.int64)
    else:
        dup = zero
    ar_b = torch.arange(B, **i64)
    if not segments:
        units = torch.stack([ar_b, torch.zeros(B, **i64), torch.full((B,), S, **i64), path_ptr[:-1], path_ptr[1:],
                             torch.full((B,), UNIT_WRITES_HT, **i64)], 1)
        stats = torch.stack([zero, zero, zero, zero, dup, bad])
        n_loss = B
    else:
        js = jump_step.to(torch.int64)[rj_sorted]
        prev_js = torch.where(first, torch.zeros_like(js), torch.roll(js, 1))
        prev_row = torch.where(first, torch.full_like(path_rows, -1), torch.roll(path_rows, 1))
        loss_units = torch.stack([sorted_path, prev_js, js, r, r + 1, prev_row + 1], 1)
        has = counts > 0
        if N:
            last_q = (path_ptr[1:] - 1).clamp(min=0)
            tail_s0 = torch.where(has, js[last_q], torch.zeros(B, **i64))
            tail_start = torch.where(has, path_rows[last_q] + 1, torch.zeros(B, **i64)) | UNIT_WRITES_HT
        else:
            tail_s0 = torch.zeros(B, **i64)
            tail_start = torch.full((B,), UNIT_WRITES_HT, **i64)
        tails = torch.stack([ar_b, tail_s0, torch.full((B,), S, **i64), path_ptr[1:], path_ptr[1:], tail_start], 1)
        len_l, len_t = js - prev_js, S - tail_s0
        o1 = torch.sort(S - len_l, stable=True).indices          # longest first, ties in path-major order
        o2 = torch.sort(S - len_t, stable=True).indices
        units = torch.cat((loss_units[o1], tails[o2]), 0)
        stats = torch.stack([(len_l >= T1).sum(), (len_l >= T2).sum(), (len_t >= T1).sum(), (len_t >= T2).sum(), dup, bad])
        n_loss = N
    i32 = torch.int32
    return (path_ptr.to(i32), path_rows.to(i32), row_jump.to(i32), units.to(i32).contiguous().view(-1), n_loss, stats)
