"""-m gpu: the reference-facing Python API (DESIREModel) on top of the C-ABI — host-buffer entry point, optimiser step,
checkpoint / resume."""
import numpy as np
import pytest
import torch

from helpers import np_batch, rel_l2, small_cfg

pytestmark = pytest.mark.gpu


def test_sample_and_rank_host_buffers_match_device_path():
    """sample_and_rank (numpy in, pinned staging, two CUDA-graph replays with the scene copy overlapped, numpy out) ==
    forward() on device tensors, for two different batches through the same captured graphs."""
    from desire_b200.model.model import DESIREModel
    cfg = small_cfg(d_dim=64, max_num_obj=10, num_samples=4)
    B, N, K, T = 3, cfg.max_num_obj, cfg.K, cfg.pred_length
    m = DESIREModel(cfg, seed=1)
    for seed in (0, 7):
        inp, tgt, eps, scene = np_batch(cfg, B, seed, 2)
        y, s, cost = m.sample_and_rank(inp, tgt, eps, scene)
        ref = m.forward(torch.from_numpy(inp).cuda(), torch.from_numpy(tgt).cuda(), torch.from_numpy(eps).cuda(),
                        torch.from_numpy(scene).cuda())
        torch.cuda.synchronize()
        assert y.shape == (B, N, K, T, 2) and s.shape == (cfg.ioc_iters, B, N, K)
        assert rel_l2(y.reshape(-1), ref["Y_refined"].cpu().numpy().reshape(-1)) <= 1e-6
        assert rel_l2(s.reshape(-1), ref["ioc_scores"].cpu().numpy().reshape(-1)) <= 1e-6
        assert abs(cost - float(ref["cost"])) <= 1e-6 * abs(cost)


def test_train_step_updates_shared_weights_and_resume_continues_the_trajectory(tmp_path):
    from desire_b200 import train as T
    from desire_b200.model.model import DESIREModel
    cfg = small_cfg(d_dim=32, max_num_obj=8, num_samples=3)
    B = 2
    inp, tgt, eps, scene = np_batch(cfg, B, 0, 1)

    def steps(model, n):
        return [float(model.train_step(inp, tgt, eps, scene)[0]) for _ in range(n)]

    a = DESIREModel(cfg, seed=1)
    a.batch_size, a.learning_rate = B, 1e-3
    w0 = a.weights["dec1_wc"].clone()
    ca = steps(a, 2)
    assert not torch.equal(w0, a.weights["dec1_wc"])              # the forward path's weight views are the updated ones
    T.save_checkpoint(a, str(tmp_path / "ck"), next_step=2)
    ca += steps(a, 2)
    b = DESIREModel(cfg, seed=5)                                  # different init: everything must come from the file
    b.batch_size, b.learning_rate = B, 1e-3
    assert T.load_checkpoint(b, str(tmp_path / "ck")) == 2
    cb = steps(b, 2)
    assert ca[3] < ca[0]
    assert np.allclose(cb, ca[2:], rtol=2e-4), (ca, cb)           # atomics: last-bit differences only
    out = b.forward(inp, tgt, eps, scene)                         # inference path sees the trained weights
    assert np.isfinite(float(out["cost"]))


def test_sample_keeps_the_reference_shape_contract():
    """model/model.py:613-688: traj [obs, N, 3] -> [obs + num, N, 3]."""
    from desire_b200.model.model import DESIREModel
    cfg = small_cfg(d_dim=32, max_num_obj=6, num_samples=3)
    m = DESIREModel(cfg, seed=1)
    inp, tgt, _, _ = np_batch(cfg, 1, 0, 0)
    traj = inp[0].transpose(1, 0, 2)                               # [obs, N, 3]
    true = np.concatenate([inp[0], tgt[0]], 1).transpose(1, 0, 2)
    out = m.sample(None, traj, None, None, true, num=cfg.pred_length)
    assert out.shape == (cfg.seq_length + cfg.pred_length, cfg.max_num_obj, 3)
    assert np.array_equal(out[:cfg.seq_length], traj)
