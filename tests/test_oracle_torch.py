"""The differentiable twin (oracle/desire_oracle_torch.py) must reproduce the NumPy oracle's forward values
(float64, round-off only) — that is what ties the autograd gradients used by the -m gpu backward tests to the
oracle the forward path is held to.  Plus a finite-difference spot check of autograd itself and the Adam
reference against torch.optim.Adam."""
import numpy as np
import torch

from helpers import np_batch, np_params, np_tables, oracle_forward, small_cfg

KEYS = ["rho_i", "H_x", "H_y", "vae_inputs", "z_mean", "z_log_sigma_sq", "zval", "x_reconstr_mean", "x_z",
        "output_states", "Yhat", "kld_rows", "recon_rows", "cost"]


def _both(cfg, B, missing):
    from oracle import desire_oracle_torch as OT
    P = np_params(cfg, dtype=np.float64)
    batch = np_batch(cfg, B, n_missing=missing, dtype=np.float64)
    ref = oracle_forward(cfg, P, batch, np_tables(cfg, np.float64))
    Pt = OT.to_torch(P)
    out = OT.generate_forward(Pt, dict(K=cfg.K, Z=cfg.Z), batch[0], batch[1], batch[2])
    return P, Pt, batch, ref, out


def test_forward_matches_numpy_oracle():
    for (H, N, K, B, missing) in [(48, 8, 3, 2, 3), (16, 5, 2, 1, 0)]:
        cfg = small_cfg(d_dim=H, max_num_obj=N, num_samples=K)
        _, _, _, ref, out = _both(cfg, B, missing)
        for k in KEYS:
            a, b = out[k].detach().numpy(), np.asarray(ref[k])
            assert a.shape == b.shape, k
            assert np.allclose(a, b, rtol=1e-10, atol=1e-12), (k, np.abs(a - b).max())


def test_autograd_matches_finite_differences():
    from oracle import desire_oracle_torch as OT
    cfg = small_cfg(d_dim=16, max_num_obj=4, num_samples=2)
    P, Pt, batch, _, out = _both(cfg, 1, 1)
    out["cost"].backward()
    rng = np.random.default_rng(0)
    for name in ["encx_wg", "venc_c2_w", "vdec_d3_w", "vdec_d2_g", "w_post_vae", "dec1_wc", "output_w", "venc_fc_b"]:
        idx = tuple(rng.integers(0, s) for s in P[name].shape)
        h = 1e-6
        vals = []
        for sgn in (+1, -1):
            P2 = {k: v.copy() for k, v in P.items()}
            P2[name][idx] += sgn * h
            o = OT.generate_forward(OT.to_torch(P2, requires_grad=False), dict(K=cfg.K, Z=cfg.Z), *batch[:3])
            vals.append(float(o["cost"]))
        fd = (vals[0] - vals[1]) / (2 * h)
        ag = float(Pt[name].grad[idx])
        assert abs(fd - ag) <= 1e-5 * max(1.0, abs(ag)), (name, idx, fd, ag)


def test_unused_parameters_have_no_gradient():
    """rho_i / feature_pooling never reach `cost` (model/model.py:291-311 is computed, then unused), and a bias
    in front of a batch-norm cannot change its output."""
    cfg = small_cfg(d_dim=16, max_num_obj=4, num_samples=2)
    _, Pt, _, _, out = _both(cfg, 1, 0)
    out["cost"].backward()
    assert Pt["temporal_w"].grad is None or float(Pt["temporal_w"].grad.abs().max()) == 0.0
    for n in ("venc_c1_b", "venc_c2_b", "venc_c3_b", "vdec_d1_b", "vdec_d2_b", "vdec_d3_b", "vdec_d4_b"):
        assert float(Pt[n].grad.abs().max()) < 1e-9, n


def test_adam_reference_matches_torch_adam():
    from oracle import desire_oracle_torch as OT
    rng = np.random.default_rng(1)
    p = {"a": rng.standard_normal((5, 3)), "b": rng.standard_normal(7)}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(x) for k, x in p.items()}
    tp = {k: torch.tensor(x.copy(), requires_grad=True) for k, x in p.items()}
    opt = torch.optim.Adam(list(tp.values()), lr=3e-3, eps=1e-8)
    for step in range(1, 4):
        g = {k: rng.standard_normal(x.shape) for k, x in p.items()}
        p, m, v = OT.adam_reference(p, g, m, v, step, 3e-3)
        for k in tp:
            tp[k].grad = torch.tensor(g[k])
        opt.step()
    for k in p:
        # torch puts eps outside the bias-corrected sqrt (eps_hat difference ~1e-8 relative)
        assert np.allclose(p[k], tp[k].detach().numpy(), rtol=1e-6, atol=1e-7)


def test_ioc_twin_matches_numpy_oracle_forward_and_loss():
    from oracle import desire_oracle_torch as OT
    for (H, N, K, B, missing) in [(16, 5, 2, 2, 1), (32, 6, 3, 1, 0)]:
        cfg = small_cfg(d_dim=H, max_num_obj=N, num_samples=K)
        P, Pt, batch, ref, _ = _both(cfg, B, missing)
        r2, dirs = np_tables(cfg, np.float64)
        out = OT.ioc_train_forward(Pt, dict(K=cfg.K, Z=cfg.Z, ioc_iters=cfg.ioc_iters), ref, batch[0], batch[1], batch[3],
                                   r2, dirs)
        for k in ("ioc_scores", "Y_refined", "ioc_rows", "ioc_cost", "scene_features"):
            a, b = out[k].detach().numpy(), np.asarray(ref[k])
            assert np.allclose(a, b, rtol=1e-9, atol=1e-11), (k, np.abs(a - b).max())


def test_ioc_twin_autograd_matches_finite_differences():
    from oracle import desire_oracle_torch as OT
    cfg = small_cfg(d_dim=16, max_num_obj=4, num_samples=2, n_rad=1, n_ang=1, r_min=1e-6, r_max=1e3)
    P, Pt, batch, ref, _ = _both(cfg, 1, 1)
    r2, dirs = np_tables(cfg, np.float64)
    ocfg = dict(K=cfg.K, Z=cfg.Z, ioc_iters=cfg.ioc_iters)
    out = OT.ioc_train_forward(Pt, ocfg, ref, batch[0], batch[1], batch[3], r2, dirs)
    out["ioc_cost"].backward()
    rng = np.random.default_rng(0)
    # only parameters whose effect does not pass through a stop_gradient can be checked by finite differences:
    # the heads and the last iteration's path are; use iters=1 so every IOC weight qualifies
    cfg1 = small_cfg(d_dim=16, max_num_obj=4, num_samples=2, n_rad=1, n_ang=1, r_min=1e-6, r_max=1e3, ioc_iters=1)
    ocfg1 = dict(K=cfg1.K, Z=cfg1.Z, ioc_iters=1)
    Pt1 = OT.to_torch(P)
    OT.ioc_train_forward(Pt1, ocfg1, ref, batch[0], batch[1], batch[3], r2, dirs)["ioc_cost"].backward()
    for name in ["dec2_wg", "dec2_wc", "ioc_sp_w", "ioc_vel_w", "ioc_score_w", "ioc_reg_w", "scene_c3_w", "scene_c1_w"]:
        idx = tuple(rng.integers(0, s) for s in P[name].shape)
        h = 1e-6
        vals = []
        for sgn in (+1, -1):
            P2 = {k: v.copy() for k, v in P.items()}
            P2[name][idx] += sgn * h
            o = OT.ioc_train_forward(OT.to_torch(P2, requires_grad=False), ocfg1, ref, batch[0], batch[1], batch[3], r2, dirs)
            vals.append(float(o["ioc_cost"]))
        fd = (vals[0] - vals[1]) / (2 * h)
        ag = float(Pt1[name].grad[idx])
        assert abs(fd - ag) <= 1e-5 * max(1.0, abs(ag)), (name, idx, fd, ag)
