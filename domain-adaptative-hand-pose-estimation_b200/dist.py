"""Batch sharding across the GPUs of one box + the single collective of the hot path.

Every stage is independent per sample, so ranks own contiguous slices of the batch
and exchange nothing until the end: ONE all-reduce(sum) of the float64 partial vector
``[mse_sum, kl_sum, n_maps, n_elems, hits[K], valid[K]]`` (4+2K doubles = 368 B at K=21; integer
counts < 2^53 are exact in float64), after which every rank finalises exactly what
``accuracy()`` / ``loss.mean()`` would give on the concatenated batch.  NCCL over NVLink on the
GPUs; the same code runs on ``gloo`` for the CPU tests of the host logic.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of ``total`` samples owned by ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def is_distributed(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def allreduce_partial(partial: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum of the partial vector over the ranks of ``group`` (no-op when single-process)."""
    if is_distributed(group):
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def finalize_partial_host(partial, K: int):
    """Host mirror of ``hp_pipeline_finalize`` for a partial vector already on the host:
    -> dict(mse, kl, avg_acc, cnt, acc[K]).  Same order of float64 operations as
    utils/keypoint_detection.py:53-60, 80-90."""
    p = np.asarray(partial, dtype=np.float64)
    hits = p[4:4 + K].astype(np.int64)
    valid = p[4 + K:4 + 2 * K].astype(np.int64)
    acc = np.full(K, -1.0)
    total, cnt = 0.0, 0
    for k in range(K):
        if valid[k] > 0:
            acc[k] = hits[k] * 1.0 / valid[k]
        if acc[k] >= 0:
            total = total + acc[k]
            cnt += 1
    return dict(mse=p[0] / p[2], kl=p[1] / p[2], avg_acc=(total / cnt if cnt != 0 else 0.0), cnt=cnt, acc=acc,
                hits=hits, valid=valid)
