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


def oracle_forward(cfg, P, batch, tables, margins=False):
    from oracle import desire_oracle as O
    inp, tgt, eps, scene = batch
    return O.forward(P, dict(K=cfg.K, Z=cfg.Z, ioc_iters=cfg.ioc_iters, margins=margins, exist_mode=cfg.exist_mode), inp, tgt, eps, scene, *tables)


# A (scene, sample) group's IOC outputs may differ from the oracle's by more than TOL only if some step of some
# iteration binned a pair that sits within BIN_MARGIN (position units; positions are O(1), so ~50 fp32 ulps) of a
# log-polar edge in the ORACLE's own trajectory: two fp32 implementations whose trajectories agree to ~1e-6 can
# then put that neighbour into different bins, which changes the group's pooled features by O(1/neighbours).
# (Typical smallest margin of a group of 10-40 agents over 12 steps x 2 iterations: 1e-6 .. 5e-5.)
BIN_MARGIN = 3e-6


def check_ioc_groups(got, ref, cfg, B, verbose=True):
    """Per (scene b, sample k) comparison of ioc_scores / Y_refined.  Every group must meet TOL unless the oracle's
    bin margin of that group is below BIN_MARGIN.  Returns (#groups, #groups excused by the margin)."""
    N, K, T = cfg.max_num_obj, cfg.K, cfg.pred_length
    y_g = np.asarray(got["Y_refined"]).reshape(B, N, K, T * 2).transpose(0, 2, 1, 3).reshape(B * K, -1)
    y_r = np.asarray(ref["Y_refined"]).reshape(B, N, K, T * 2).transpose(0, 2, 1, 3).reshape(B * K, -1)
    s_g = np.asarray(got["ioc_scores"]).reshape(-1, B, N, K).transpose(1, 3, 0, 2).reshape(B * K, -1)
    s_r = np.asarray(ref["ioc_scores"]).reshape(-1, B, N, K).transpose(1, 3, 0, 2).reshape(B * K, -1)
    margin = np.asarray(ref["bin_margin"]).reshape(B * K)
    excused = np.zeros(B * K, bool)
    for name, g, r in (("Y_refined", y_g, y_r), ("ioc_scores", s_g, s_r)):
        errs = np.array([rel_l2(g[i], r[i]) for i in range(B * K)])
        over = errs > TOL
        if verbose:
            print("%-12s groups: median %.2e, over tol %d/%d, worst %.2e; margins of the groups over tol: %s" % (
                name, np.median(errs), over.sum(), len(errs), errs.max(), np.sort(margin[over])[:8]))
        unexplained = over & ~(margin < BIN_MARGIN)
        assert not unexplained.any(), (name, "groups over tolerance with no pair near a bin edge",
                                       np.nonzero(unexplained)[0][:10], errs[unexplained][:10], margin[unexplained][:10])
        excused |= over
    return B * K, int(excused.sum())
