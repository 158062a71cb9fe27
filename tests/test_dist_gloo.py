"""CPU, world_size 2, gloo: the data-parallel path of njode_b200.dist (path shards, global schedule,
global loss normalisation, one all-reduce of the flat gradient buffer) reproduces the single-process
result.  The kernels are the host simulation of the CUDA source (test infrastructure)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dropout, q):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cases
    import hostsim_util
    from njode_b200 import models
    from njode_b200 import dist as njdist
    hostsim_util.install()
    cfg = cases.demo_cfg(dropout_rate=dropout, input_size=2, output_size=2)
    batch = cases.grid_batch(37, 2, 20, 0.25, seed=3)
    B = 37

    def run(model, b, first, sync_seed):
        if dropout:
            model.train()
            torch.manual_seed(sync_seed)          # same dropout seed on every rank / in the reference run
        else:
            model.eval()
        hT, loss = model(b["times"], b["time_ptr"], b["X"], b["obs_idx"], 0.05, 1.0, b["start_X"], b["n_obs_ot"])
        loss.backward()
        return hT.detach(), loss.detach()

    # single-process reference on the full batch (identical on both ranks)
    torch.manual_seed(100)
    ref = models.NJODE(**cfg)
    r_hT, r_loss = run(ref, batch, 0, 7)
    r_grads = [p.grad.clone() for p in ref.parameters()]
    # data parallel: rank-dependent init is overwritten by the broadcast from rank 0
    torch.manual_seed(100 if rank == 0 else 555)
    model = models.NJODE(**cfg)
    dp = njdist.DataParallel(model, global_batch_size=B)
    local, first = njdist.shard_batch(batch, rank, world)
    dp.set_batch(B, first)
    hT, loss = run(model, local, first, 7)
    total = dp.reduce_loss(loss)
    ok = True
    for g, p in zip(r_grads, model.parameters()):
        # fp32 sums in a different order (two shards): tolerance relative to the tensor's scale
        ok &= bool(torch.allclose(p.grad, g, rtol=2e-5, atol=2e-6 * float(g.abs().max()) + 1e-7))
    lo, hi = (B * rank) // world, (B * (rank + 1)) // world
    ok &= bool(torch.allclose(hT, r_hT[lo:hi], rtol=1e-5, atol=1e-6))
    ok &= bool(abs(float(total) - float(r_loss)) <= 1e-5 * abs(float(r_loss)))
    ok &= len(local["times"]) == len(batch["times"])            # every rank keeps the global schedule
    q.put((rank, ok, float(total), float(r_loss)))
    dist.destroy_process_group()


@pytest.mark.parametrize("dropout", [0.0, 0.2])
def test_two_rank_data_parallel_matches_single_process(dropout):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dropout, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, total, ref in res:
        assert ok, (rank, total, ref)


def test_shard_batch_partitions_rows():
    sys.path.insert(0, HERE)
    import cases
    from njode_b200 import dist as njdist
    batch = cases.grid_batch(23, 1, 12, 0.3, seed=9)
    parts = [njdist.shard_batch(batch, r, 3) for r in range(3)]
    assert sum(len(p[0]["obs_idx"]) for p in parts) == len(batch["obs_idx"])
    assert [p[1] for p in parts] == [0, 7, 15]
    for p, first in parts:
        assert np.array_equal(p["times"], batch["times"])
        assert int(p["time_ptr"][-1]) == len(p["obs_idx"]) == len(p["X"])
