"""Synthetic workloads of SURVEY.md §8d: random-walk agents on the unit square, a random scene
image, explicit eps — everything seeded so ranks, tests and the oracle see identical bits."""
from __future__ import annotations

import torch

from .config import DesireConfig


def make_batch(cfg: DesireConfig, B: int, seed: int = 0, n_missing: int = 0):
    """-> input_data [B,N,Tp,3], target_data [B,N,Tf,3] (id,x,y), eps [B*N,K,Z], scene [B,Hi,Wi,3].
    Start U(0,1)^2, velocity N(0,0.02^2) random walk; ids 1..N; the last `n_missing` agents of every
    odd scene get id 0 (non-existent, D8).  eps uses seed+2, the scene seed+3."""
    g = torch.Generator().manual_seed(seed)
    N, Tp, Tf = cfg.max_num_obj, cfg.seq_length, cfg.pred_length
    start = torch.rand(B, N, 1, 2, generator=g)
    vel = torch.randn(B, N, Tp + Tf, 2, generator=g) * 0.02
    traj = start + torch.cumsum(vel, 2)
    ids = torch.arange(1, N + 1, dtype=torch.float32).view(1, N, 1, 1).expand(B, N, Tp + Tf, 1).clone()
    if n_missing:
        ids[1::2, N - n_missing:] = 0
    data = torch.cat([ids, traj], -1).float()
    inp, tgt = data[:, :, :Tp].contiguous(), data[:, :, Tp:].contiguous()
    eps = torch.randn(B * N, cfg.K, cfg.Z, generator=torch.Generator().manual_seed(seed + 2))
    scene = torch.rand(B, cfg.scene_size, cfg.scene_size, 3, generator=torch.Generator().manual_seed(seed + 3))
    return inp, tgt, eps, scene
