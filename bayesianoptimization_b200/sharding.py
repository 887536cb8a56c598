"""Multi-GPU layer: candidates shard across ranks, the O(N^2) model state
is replicated (every rank factorises the same data deterministically), and ONE small exchange
merges the per-rank (argmin, top-k) records - an all_gather of (k+1) 16-byte records per rank
over NCCL/NVLink (gloo on CPU for the tests).  No data-path collective: the fused kernel never
waits on a peer."""
from __future__ import annotations

import numpy as np


def shard_range(m: int, rank: int, world: int):
    """Contiguous shard [start, stop) of m candidates for ``rank`` (remainder to the low ranks)."""
    base, rem = divmod(m, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _order_nan_first(v, i):
    # np.argmin: NaN is the minimum, first occurrence wins
    return (0 if np.isnan(v) else 1, 0.0 if np.isnan(v) else v, i)


def _order_nan_last(v, i):
    # np.argsort: NaN sorts last; ties by index
    return (1 if np.isnan(v) else 0, 0.0 if np.isnan(v) else v, i)


def merge_selection(values: np.ndarray, indices: np.ndarray, k: int):
    """values/indices: (world, k+1) per-rank records with GLOBAL indices (index < 0 = empty slot);
    record 0 = local argmin (NaN first), records 1..k = local top-k (NaN last).
    Returns (best_idx, best_val, topk_idx[<=k]) identical to np.argmin / stable argsort[:k] on the
    concatenated array."""
    values = np.asarray(values, dtype=np.float64)
    indices = np.asarray(indices, dtype=np.int64)
    cands = [(_order_nan_first(values[r, 0], indices[r, 0]), r) for r in range(values.shape[0]) if indices[r, 0] >= 0]
    key, r = min(cands)
    best_idx, best_val = int(indices[r, 0]), float(values[r, 0])
    pool = [
        (_order_nan_last(values[r, j], indices[r, j]), int(indices[r, j]))
        for r in range(values.shape[0]) for j in range(1, values.shape[1]) if indices[r, j] >= 0
    ]
    pool.sort()
    return best_idx, best_val, np.array([i for _, i in pool[:k]], dtype=np.int64)


def allgather_selection(local_values, local_indices, k: int, device=None):
    """One collective: all_gather of the (k+1) local records.  Works on any initialised
    torch.distributed backend (nccl: tensors on ``device``; gloo: CPU)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    rec = torch.empty((k + 1, 2), dtype=torch.float64)
    rec[:, 0] = torch.as_tensor(np.asarray(local_values, dtype=np.float64))
    # indices travel as float64 bit patterns (exact for |idx| < 2^53)
    rec[:, 1] = torch.as_tensor(np.asarray(local_indices, dtype=np.int64)).to(torch.float64)
    if device is not None:
        rec = rec.to(device)
    out = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(out, rec)
    allr = torch.stack(out).cpu().numpy()
    return merge_selection(allr[:, :, 0], allr[:, :, 1].astype(np.int64), k)


def sharded_argmin_topk(acq, x_shard, k: int, start_index: int, device=None):
    """Multi-GPU version of AcquisitionFunction._random_sample_minimize's selection
    (R/bayes_opt/acquisition.py:312-317): every rank evaluates its contiguous shard of the candidate
    matrix with the fused kernel (``acq`` = FusedAcquisition on this rank's replica of the model),
    then ONE all_gather merges the (argmin, top-k) records.  Returns global (best_idx, best_val,
    topk_idx) - identical on every rank and identical to the single-GPU result."""
    import numpy as np

    idx, val, top = acq.argmin_topk(x_shard, k)
    vals = np.full(k + 1, np.nan)
    idxs = np.full(k + 1, -1, dtype=np.int64)
    ys_top = acq(x_shard[top]) if len(top) else np.empty(0)
    vals[0], idxs[0] = val, start_index + idx
    vals[1:1 + len(top)], idxs[1:1 + len(top)] = ys_top, start_index + top
    return allgather_selection(vals, idxs, k, device=device)
