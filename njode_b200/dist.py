"""Data-parallel training over paths (no counterpart in the reference, whose
"parallel" training is a joblib farm over hyper-parameter configs, NJODE/parallel_train.py:214-223).

One process per GPU.  Paths are independent given the batch-global Euler schedule, so every rank
runs the unchanged forward/backward kernels on its contiguous shard of the batch; the only
exchange is ONE all-reduce(sum) per step over the flat fp32 gradient buffer (the kernels already
write all parameter gradients into one contiguous buffer), issued on the stream the backward
kernel ran on.  The loss normalisation uses the GLOBAL batch size (NJODE/models.py:105-106 divides
by batch_size) and dropout keys use GLOBAL path ids, so results do not depend on the rank count.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_batch(batch, rank, world, M_key="M"):
    """contiguous path shard of a collated batch (NJODE/data_utils.py:311-315 contract).  ``times``
    stays the global list: every rank steps through every observation time of the *global* batch
    (the reference's Euler grid is batch-global, NJODE/models.py:430-439); a rank without rows at
    some time just gets an empty time_ptr slot.  Returns (local batch, first global path id)."""
    B = int(batch["start_X"].shape[0])
    lo, hi = (B * rank) // world, (B * (rank + 1)) // world
    obs_idx = batch["obs_idx"].numpy() if torch.is_tensor(batch["obs_idx"]) else np.asarray(batch["obs_idx"])
    time_ptr = np.asarray(batch["time_ptr"])
    keep = (obs_idx >= lo) & (obs_idx < hi)
    row_time = np.repeat(np.arange(len(time_ptr) - 1), np.diff(time_ptr))
    counts = np.bincount(row_time[keep], minlength=len(time_ptr) - 1)
    out = dict(batch)
    out["time_ptr"] = np.concatenate(([0], np.cumsum(counts)))
    out["obs_idx"] = torch.from_numpy((obs_idx[keep] - lo).astype(np.int64))
    keep_t = torch.from_numpy(keep)
    out["X"] = batch["X"][keep_t]
    if batch.get(M_key) is not None:
        out[M_key] = batch[M_key][keep_t]
    out["start_X"] = batch["start_X"][lo:hi]
    out["n_obs_ot"] = batch["n_obs_ot"][lo:hi]
    for k in ("true_paths", "observed_dates"):
        if k in out:
            out[k] = out[k][lo:hi]
    return out, lo


class DataParallel:
    """wraps an ``njode_b200.models.NJODE``: broadcasts rank 0's parameters, and makes every
    backward end with one all-reduce(sum) of the flat gradient buffer."""

    def __init__(self, model, global_batch_size=None, group=None):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        model._ensure_flat()
        dist.broadcast(model._flat, src=0, group=group)
        model._grad_sync = self._sync
        self.set_batch(global_batch_size, 0)

    def set_batch(self, global_batch_size, first_path_id):
        self.model.batch_size_norm = global_batch_size
        self.model.path_id_offset = int(first_path_id)

    def _sync(self, flat_grads):
        # same stream as the backward kernel (torch's current stream): ordered after it
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=self.group)

    def reduce_loss(self, loss):
        """sum of the per-rank partial losses (each already divided by the global batch size)"""
        t = loss.detach().clone().to(self.model._flat.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def __call__(self, *a, **k):
        return self.model(*a, **k)
