"""world_size-2 gloo test of the one exchange step of the multi-GPU path: the
all_gather + merge of per-rank (argmin, top-k) records equals np.argmin / stable argsort on the
whole candidate set.  Runs on CPU."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, k, m, q):
    import torch.distributed as dist

    from bayesianoptimization_b200.sharding import allgather_selection, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ys = np.random.RandomState(123).randn(m)
    ys[5] = ys[900]  # a cross-shard tie
    s, e = shard_range(m, rank, world)
    loc = ys[s:e]
    vals = np.full(k + 1, np.nan)
    idxs = np.full(k + 1, -1, dtype=np.int64)
    vals[0], idxs[0] = loc[np.argmin(loc)], s + int(np.argmin(loc))
    order = np.argsort(loc, kind="stable")[:k]
    vals[1:1 + len(order)], idxs[1:1 + len(order)] = loc[order], s + order
    bi, bv, top = allgather_selection(vals, idxs, k)
    if rank == 0:
        q.put((bi, bv, list(top), int(np.argmin(ys)), list(np.argsort(ys, kind="stable")[:k])))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_selection_exchange():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world, k, m = 2, 10, 1001
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, m, q)) for r in range(world)]
    for p in procs:
        p.start()
    bi, bv, top, ref_i, ref_top = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert bi == ref_i and top == ref_top
