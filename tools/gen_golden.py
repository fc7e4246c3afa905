"""Generate tests/golden/*.npz from the CPU oracle with fixed seeds (data 0, weights 1, eps 2, scene 3).

    python tools/gen_golden.py

The reference has no golden vectors and cannot run, so these are vectors of OUR
restatement; they freeze it (any later edit of the oracle or of the initialisers that changes a
number fails tests/test_golden.py) and travel to the GPU box, where the CUDA path is checked
against them without importing the oracle's inputs again.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import np_batch, np_params, np_tables, oracle_forward, small_cfg  # noqa: E402

CASES = {
    # name: (cfg kwargs, B, n_missing)
    "cfg1_like_N8_K1_H48": (dict(d_dim=48, max_num_obj=8, num_samples=1, scene_size=32, ioc_iters=2), 1, 0),
    "small_N6_K3_H32_missing": (dict(d_dim=32, max_num_obj=6, num_samples=3, scene_size=24, ioc_iters=2), 2, 2),
}
KEEP = ["rho_i", "H_x", "H_y", "z_mean", "z_log_sigma_sq", "x_z", "Yhat", "kld_rows", "recon_rows", "cost",
        "ioc_scores", "Y_refined"]


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (kw, B, miss) in CASES.items():
        cfg = small_cfg(**kw)
        out = oracle_forward(cfg, np_params(cfg), np_batch(cfg, B, 0, miss), np_tables(cfg))
        arrs = {k: np.asarray(out[k], np.float32) for k in KEEP}
        # a strided sample of the big tensors keeps the fixture small but position-sensitive
        arrs["output_states_s"] = np.asarray(out["output_states"], np.float32)[:, ::3, ::5]
        arrs["feature_pooling_s"] = np.asarray(out["feature_pooling"], np.float32)[:, ::4, ::17]
        arrs["x_reconstr_mean_s"] = np.asarray(out["x_reconstr_mean"], np.float32)[:, ::37]
        arrs["scene_features_s"] = np.asarray(out["scene_features"], np.float32)[:, ::3, ::3, ::5]
        arrs["meta"] = np.array([B, miss] + [kw[k] for k in ("d_dim", "max_num_obj", "num_samples", "scene_size", "ioc_iters")])
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrs)
        print(name, {k: v.shape for k, v in arrs.items()})


if __name__ == "__main__":
    main()
