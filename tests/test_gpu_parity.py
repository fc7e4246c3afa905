"""-m gpu: the CUDA path through the C-ABI vs the CPU oracle on identical seeded inputs.
Tolerance: rel-L2 <= 1e-4 (north_star) on every named tensor of the path."""
import numpy as np
import pytest
import torch

from helpers import TOL, np_batch, np_params, np_tables, oracle_forward, rel_l2, small_cfg

pytestmark = pytest.mark.gpu

CHECK = ["rho_i", "H_x", "H_y", "vae_inputs", "z_mean", "z_log_sigma_sq", "zval", "x_reconstr_mean", "x_z",
         "output_states", "Yhat", "feature_pooling", "kld_rows", "recon_rows", "cost", "scene_features",
         "ioc_scores", "Y_refined"]


def run_gpu(cfg, B, seed=0, n_missing=0):
    from desire_b200.config import init_params
    from desire_b200.engine import HotPath
    from desire_b200.synthetic import make_batch
    hp = HotPath(cfg, init_params(cfg, 1), B)
    inp, tgt, eps, scene = [t.cuda() for t in make_batch(cfg, B, seed, n_missing)]
    out = hp.run(inp, tgt, eps, scene)
    torch.cuda.synchronize()
    return {k: v.detach().cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("H,N,K,B,missing", [(48, 8, 1, 2, 0), (48, 8, 3, 2, 3), (128, 12, 4, 3, 2), (16, 5, 2, 1, 0),
                                                (64, 10, 5, 2, 1), (256, 6, 6, 2, 0), (32, 40, 2, 2, 0)])
def test_full_path_matches_oracle(H, N, K, B, missing):
    cfg = small_cfg(d_dim=H, max_num_obj=N, num_samples=K)
    got = run_gpu(cfg, B, n_missing=missing)
    ref = oracle_forward(cfg, np_params(cfg), np_batch(cfg, B, n_missing=missing), np_tables(cfg))
    bad = {}
    for k in CHECK:
        e = rel_l2(got[k].reshape(-1), np.asarray(ref[k]).reshape(-1))
        print("%-18s rel-L2 %.3e" % (k, e))
        if not e <= TOL:
            bad[k] = e
    assert not bad, bad
