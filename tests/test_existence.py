"""D8 existence (model/model.py:206,351-366): agents that leave the scene, are missing at the last observed frame, or
enter after the first frame must not contribute to the cost, the IOC loss or their neighbours' social pooling.
CPU: the oracle's mask; GPU: the CUDA path (desire_existence_fwd feeding every kernel) vs the oracle on a batch that
holds such agents — the windows DataLoader._window builds for them are zero-filled rows."""
import numpy as np
import pytest
import torch

from helpers import np_params, np_tables, oracle_forward, rel_l2, small_cfg, TOL

ONE_BIN = dict(n_rad=1, n_ang=1, r_min=1e-6, r_max=1e3)


def leaving_entering_batch(cfg, B=2):
    from desire_b200.synthetic import make_batch
    inp, tgt, eps, scene = [t.clone() for t in make_batch(cfg, B, 0, 0)]
    Tp = cfg.seq_length
    tgt[0, 1, 5:] = 0            # leaves at target frame 5 (DataLoader leaves the rows zero)
    inp[0, 2, Tp - 1] = 0        # missing at the last observed frame
    inp[1, 3, :3] = 0            # enters at observed frame 3
    tgt[1, 4, 0] = 0             # flickers out for one target frame
    gone = {(0, 1), (0, 2), (1, 3), (1, 4)}
    return (inp, tgt, eps, scene), gone


def test_oracle_mask_and_cost_ignore_absent_agents():
    from oracle import desire_oracle as O
    cfg = small_cfg(d_dim=16, max_num_obj=6, num_samples=2, scene_size=16, ioc_iters=1, **ONE_BIN)
    batch, gone = leaving_entering_batch(cfg)
    inp, tgt = batch[0].numpy(), batch[1].numpy()
    m1 = O.existence_mask(inp, tgt, 1)
    m0 = O.existence_mask(inp, tgt, 0)
    assert {(b, n) for b, n in zip(*np.nonzero(~m1))} == gone
    assert {(b, n) for b, n in zip(*np.nonzero(~m0))} == {(1, 3)}          # frame-0 id only
    out = oracle_forward(cfg, np_params(cfg), [t.numpy() for t in batch], np_tables(cfg))
    rows = (out["recon_rows"] + out["kld_rows"]).reshape(2, 6)
    assert np.isclose(out["cost"], rows[m1].mean(), rtol=1e-6)
    cfg0 = small_cfg(d_dim=16, max_num_obj=6, num_samples=2, scene_size=16, ioc_iters=1, exist_mode=0, **ONE_BIN)
    out0 = oracle_forward(cfg0, np_params(cfg0), [t.numpy() for t in batch], np_tables(cfg0))
    assert np.isclose(out0["cost"], rows[m0].mean(), rtol=1e-6) and not np.isclose(out0["cost"], out["cost"], rtol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 0])
def test_cuda_path_matches_oracle_with_leaving_and_entering_agents(mode):
    from desire_b200.config import init_params
    from desire_b200.engine import HotPath, existing_agents
    cfg = small_cfg(d_dim=64, max_num_obj=10, num_samples=3, exist_mode=mode, **ONE_BIN)
    batch, gone = leaving_entering_batch(cfg)
    hp = HotPath(cfg, init_params(cfg, 1), 2)
    dev = [t.cuda() for t in batch]
    out = hp.run(*dev)
    torch.cuda.synchronize()
    ref = oracle_forward(cfg, np_params(cfg), [t.numpy() for t in batch], np_tables(cfg))
    n_exist = int(existing_agents(dev[0], dev[1], mode).sum())
    assert n_exist == (2 * 10 - len(gone) if mode == 1 else 2 * 10 - 1)
    assert float(hp.buf["cost"][1]) == n_exist
    for k in ("cost", "Yhat", "ioc_scores", "Y_refined"):
        e = rel_l2(out[k].cpu().numpy().reshape(-1), np.asarray(ref[k]).reshape(-1))
        print("mode %d %-12s rel-L2 %.3e" % (mode, k, e))
        assert e <= TOL, (k, e)
