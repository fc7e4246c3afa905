"""Shared test helpers: small configs, numpy views of the product's params, rel-L2."""
import numpy as np

from desire_b200.config import DesireConfig, init_params, logpolar_tables
from desire_b200.synthetic import make_batch

TOL = 1e-4   # north_star: decoded trajectories and IOC scores within 1e-4 rel-L2 of the fp32 oracle


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def small_cfg(**kw):
    d = dict(d_dim=48, max_num_obj=8, num_samples=3, scene_size=32, ioc_iters=2)
    d.update(kw)
    cfg = DesireConfig(**d)
    cfg.validate()
    return cfg


def np_params(cfg, seed=1, dtype=np.float32):
    return {k: v.numpy().astype(dtype) for k, v in init_params(cfg, seed).items()}


def np_batch(cfg, B, seed=0, n_missing=0, dtype=np.float32):
    return tuple(t.numpy().astype(dtype) for t in make_batch(cfg, B, seed, n_missing))


def np_tables(cfg, dtype=np.float32):
    r2, dirs = logpolar_tables(cfg)
    return r2.numpy().astype(dtype), dirs.numpy().astype(dtype)


def oracle_forward(cfg, P, batch, tables):
    from oracle import desire_oracle as O
    inp, tgt, eps, scene = batch
    return O.forward(P, dict(K=cfg.K, Z=cfg.Z, ioc_iters=cfg.ioc_iters), inp, tgt, eps, scene, *tables)
