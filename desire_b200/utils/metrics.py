"""ADE / FDE evaluator for the K ranked samples (DESIRE paper, SDD table: errors of the oracle-best
sample and of the top-10 % ranked samples).  Pure NumPy on the host arrays `DESIREModel.sample_and_rank` returns —
evaluation glue, not part of the hot path.

Shapes: Y_pred [B,N,K,T,2], Y_true [B,N,T,2], scores [B,N,K] (higher = better, the last IOC iteration's),
mask [B,N] (existing agents, D8)."""
from __future__ import annotations

import numpy as np


def displacement_errors(Y_pred, Y_true):
    """-> (ade [B,N,K], fde [B,N,K]): mean / final L2 displacement of every sample."""
    d = np.linalg.norm(np.asarray(Y_pred, np.float64) - np.asarray(Y_true, np.float64)[:, :, None], axis=-1)   # [B,N,K,T]
    return d.mean(-1), d[..., -1]


def evaluate(Y_pred, Y_true, scores, mask, top_frac=0.1):
    """Masked means over existing agents of
       ade/fde_best   — the sample closest to the ground truth (oracle choice, the paper's "best of K"),
       ade/fde_top1   — the sample the IOC module ranks first,
       ade/fde_topk   — the best among the ceil(top_frac*K) highest-ranked samples (paper: top 10 %),
       ade/fde_mean   — average over all K samples."""
    ade, fde = displacement_errors(Y_pred, Y_true)
    mask = np.asarray(mask, bool)
    if not mask.any():
        raise ValueError("evaluate(): no existing agent in the batch")
    K = ade.shape[-1]
    order = np.argsort(-np.asarray(scores, np.float64), axis=-1)                       # best-ranked first
    kk = max(1, int(np.ceil(top_frac * K)))
    top = order[..., :kk]
    take = lambda a, idx: np.take_along_axis(a, idx, axis=-1)
    out = {
        "ade_best": ade.min(-1), "fde_best": fde.min(-1),
        "ade_top1": take(ade, order[..., :1])[..., 0], "fde_top1": take(fde, order[..., :1])[..., 0],
        "ade_topk": take(ade, top).min(-1), "fde_topk": take(fde, top).min(-1),
        "ade_mean": ade.mean(-1), "fde_mean": fde.mean(-1),
    }
    res = {k: float(v[mask].mean()) for k, v in out.items()}
    res["n_agents"] = int(mask.sum())
    res["top_k"] = kk
    return res
