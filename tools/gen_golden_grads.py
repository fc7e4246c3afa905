"""Generate tests/golden/grads_*.npz: gradients of `cost` (stage 1) and `ioc_cost` (stage 2, D13) from float64 autograd of the
oracle's differentiable twin, fixed seeds (data 0, weights 1, eps 2, scene 3), single social bin (smooth — no bin edges).

    python tools/gen_golden_grads.py

Like tools/gen_golden.py these are vectors of OUR restatement (the reference never runs its optimiser); they freeze the
twin and travel to the GPU box, where the backward kernels are checked against them without running autograd."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import np_batch, np_params, np_tables, oracle_forward, small_cfg  # noqa: E402

ONE_BIN = dict(n_rad=1, n_ang=1, r_min=1e-6, r_max=1e3)
CASES = {"grads_N6_K3_H32_onebin": (dict(d_dim=32, max_num_obj=6, num_samples=3, scene_size=24, ioc_iters=2), 2, 2)}
STRIDE = 7      # tensors above 4096 elements keep every 7th element of the flattened gradient


def reference_grads(cfg, B, miss):
    from oracle import desire_oracle_torch as OT
    P = np_params(cfg, dtype=np.float64)
    batch = np_batch(cfg, B, 0, miss, dtype=np.float64)
    tables = np_tables(cfg, np.float64)
    gen = oracle_forward(cfg, P, batch, tables)
    ocfg = dict(K=cfg.K, Z=cfg.Z, ioc_iters=cfg.ioc_iters)
    Pt = OT.to_torch(P)
    c1 = OT.generate_forward(Pt, ocfg, batch[0], batch[1], batch[2])["cost"]
    c2 = OT.ioc_train_forward(Pt, ocfg, gen, batch[0], batch[1], batch[3], *tables)["ioc_cost"]
    (c1 + c2).backward()
    g = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in Pt.items()}
    return g, float(c1.detach()), float(c2.detach())


def sample(a):
    a = np.asarray(a).reshape(-1)
    return a[::STRIDE] if a.size > 4096 else a


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, (kw, B, miss) in CASES.items():
        cfg = small_cfg(**dict(kw, **ONE_BIN))
        g, c1, c2 = reference_grads(cfg, B, miss)
        arrs = {"g_" + k: sample(v).astype(np.float32) for k, v in g.items()}
        arrs["cost"] = np.array([c1, c2])
        arrs["meta"] = np.array([B, miss] + [kw[k] for k in ("d_dim", "max_num_obj", "num_samples", "scene_size", "ioc_iters")])
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrs)
        print(name, c1, c2, sum(v.size for v in arrs.values()), "floats")


if __name__ == "__main__":
    main()
