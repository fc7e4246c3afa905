"""Microbenchmark of the fused social pooling + fc kernel at the bench workload's per-chain and whole-batch sizes.
    python tools/bench_social.py [B N K H]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desire_b200 import _lib
from desire_b200.config import DesireConfig, logpolar_tables

B, N, K, H = [int(x) for x in sys.argv[1:5]] if len(sys.argv) >= 5 else (32, 60, 20, 128)
lib = _lib.load()
cfg = DesireConfig(d_dim=H, max_num_obj=N, num_samples=K)
r2, dirs = [t.cuda() for t in logpolar_tables(cfg)]
g = torch.Generator().manual_seed(0)
pos = (torch.rand(B, N, K, 2, generator=g) * 0.6).cuda()
h = torch.tanh(torch.randn(B, N, K, H, generator=g)).cuda()
obs = torch.ones(B * N, 8, 3).cuda()
W = (torch.randn(cfg.G * H, H, generator=g) / H ** 0.5).cuda()
b = torch.zeros(H).cuda()
out = torch.empty(B * N * K, H, device="cuda")
nbytes = lib.desire_social_fc_workspace_bytes(H, cfg.G)
ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())


def call():
    _lib.check(lib.desire_social_fc_fwd(p(pos), 2, p(h), H, p(obs), 8, B, N, K, H, cfg.n_rad, cfg.n_ang, p(r2), p(dirs), p(W), p(b),
                                        p(out), p(ws), nbytes, None), "social_fc")


for _ in range(3):
    call()
torch.cuda.synchronize()
lib.desire_prof_enable(1)
for _ in range(20):
    call()
torch.cuda.synchronize()
n, ms = C.c_long(), C.c_double()
lib.desire_prof_read(4, C.byref(n), C.byref(ms))
lib.desire_prof_enable(0)
R = B * N * K
flop = 2.0 * R * cfg.G * H * H
us = ms.value / max(n.value, 1) * 1e3
print("social fc B%d N%d K%d H%d (R=%d): %d launches, %.1f us each, %.1f TFLOP/s algorithmic (%s)" % (
    B, N, K, H, R, n.value, us, flop / (us * 1e-6) / 1e12, "v1 smem" if os.environ.get("DESIRE_SOCIAL_V1") == "1" else "tmem"))
