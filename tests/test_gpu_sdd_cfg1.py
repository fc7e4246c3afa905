"""-m gpu: BASELINE configs[0] — real SDD bookstore/video0 window through the DataLoader, N=8 agents (first 8 of
the crowded frames), T_p=8, T_f=12, K=1, hidden=48 — CUDA path vs oracle; plus the train.py loop on the same data."""
import os

import numpy as np
import pytest
import torch

from helpers import TOL, np_params, np_tables, rel_l2, small_cfg

pytestmark = pytest.mark.gpu
SDD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdd") + "/"


def test_cfg1_real_sdd_window_matches_oracle():
    from desire_b200.model.model import DESIREModel
    from desire_b200.utils.data_loader import DataLoader
    from oracle import desire_oracle as O
    # pixel coordinates normalised by the bookstore frame size so positions live on the unit square of the scene map
    dl = DataLoader(2, 8, 8, 1, preprocess=True, data_dir=SDD, cache=False, pred_length=12, clip=True,
                    normalize=(1424.0, 1088.0))
    xb, yb, _ = dl.next_batch(random_update=False)
    x, y = DataLoader.to_model_layout(xb), DataLoader.to_model_layout(yb)
    assert (x[:, :, 0, 0] != 0).sum() >= 10                       # real agents present, id-0 slot empty
    cfg = small_cfg(d_dim=48, max_num_obj=8, num_samples=1, n_rad=1, n_ang=1, r_min=1e-6, r_max=1e3)
    model = DESIREModel(cfg, seed=1, use_graph=False)
    g = torch.Generator().manual_seed(2)
    eps = torch.randn(2 * 8, 1, cfg.Z, generator=g)
    scene = torch.rand(2, cfg.scene_size, cfg.scene_size, 3, generator=g)
    out = model.forward(x, y, eps.numpy(), scene.numpy())
    torch.cuda.synchronize()
    ref = O.forward(np_params(cfg), dict(K=1, Z=cfg.Z, ioc_iters=cfg.ioc_iters), x, y, eps.numpy(), scene.numpy(), *np_tables(cfg))
    for k in ("H_x", "Yhat", "cost", "ioc_scores", "Y_refined"):
        e = rel_l2(out[k].cpu().numpy().reshape(-1), np.asarray(ref[k]).reshape(-1))
        print("%-12s rel-L2 %.3e" % (k, e))
        assert e <= TOL, (k, e)


def test_train_loop_runs_and_logs(tmp_path, capsys):
    from desire_b200 import train as T
    argv = ["--data_dir", SDD, "--save_dir", str(tmp_path / "save"), "--batch_size", "2", "--max_num_obj", "8",
            "--clip_objects", "--d_dim", "32", "--num_samples", "2", "--leave_dataset", "1", "--num_epochs", "2",
            "--scene_size", "32", "--norm_w", "1424", "--norm_h", "1088", "--save_every", "1"]
    args = T.build_parser().parse_args(argv)
    losses = T.train(args)
    out = capsys.readouterr().out
    assert len(losses) == 2 * 4 and all(np.isfinite(losses))     # num_batches = 2*int(int(40/10)/2) = 4 per epoch
    assert "train_loss = " in out and "time/batch = " in out and "(epoch 1)" in out
    assert os.path.exists(tmp_path / "save" / "config.pkl")
    assert any(f.startswith("social_model.ckpt-") for f in os.listdir(tmp_path / "save"))
