"""Multi-GPU plumbing: one process per GPU, scenes sharded across ranks.

Scenes are independent in every stage (social pooling couples agents only inside one scene), so
the forward path needs NO data-path collective: rank r simply owns scenes b = r (mod world).  The
only cross-rank value is the masked-mean cost, which must be normalised by the GLOBAL number of
existing agents to equal the single-GPU cost/counter of model/model.py:376 — one tiny all-reduce
of (sum, count).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_scenes(n_scenes: int, rank: int, world: int):
    """Indices of the scenes owned by `rank` (round-robin, so ragged tails spread evenly)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return list(range(rank, n_scenes, world))


def global_masked_cost(local_sum: torch.Tensor, local_count: torch.Tensor) -> torch.Tensor:
    """cost = sum_ranks(sum) / sum_ranks(count).  Works on any backend (nccl on GPUs, gloo in tests)."""
    t = torch.stack([local_sum.reshape(()).double(), local_count.reshape(()).double()])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return (t[0] / t[1]).float()


def _active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def global_count_(count: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks of the number of existing agents — the shared normaliser of `cost`
    (model/model.py:376) that every rank's backward divides by, so that the SUM of the per-rank gradients
    equals the gradient of the single-GPU cost over the union of the scenes."""
    if _active():
        dist.all_reduce(count, op=dist.ReduceOp.SUM)
    return count


def all_reduce_gradients_(grad_flat: torch.Tensor) -> torch.Tensor:
    """The train step's one collective: in-place SUM of the flat fp32 gradient buffer over ranks
    (NCCL over NVLink on GPUs; gloo in the CPU tests).  Every rank then applies the identical clip + Adam."""
    if _active():
        dist.all_reduce(grad_flat, op=dist.ReduceOp.SUM)
    return grad_flat
