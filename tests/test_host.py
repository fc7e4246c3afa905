"""CPU: host-side logic — config rules, parameter inventory, synthetic generator, loud failure without CUDA."""
import numpy as np
import pytest
import torch

from desire_b200.config import DesireConfig, init_params, logpolar_tables, param_shapes
from desire_b200.synthetic import make_batch
from helpers import small_cfg


def test_config_rejects_rnn_size_that_breaks_the_cvae():
    # model/model.py:57-59 + :440-441: the decoder always emits 32x32, so S must be 32 (D5)
    with pytest.raises(ValueError):
        DesireConfig(rnn_size=128).validate()
    DesireConfig(rnn_size=512).validate()


def test_param_shapes_follow_reference():
    cfg = DesireConfig(d_dim=16)
    sh = param_shapes(cfg)
    assert sh["temporal_w"][0] == (8, 2, 100)            # [1,T,2,C] of model.py:427-429, squeezed
    assert sh["w_hidden_enc1"][0] == (32, 1024)          # [2*d_dim, S*S]  model.py:434-435
    assert sh["w_post_vae"][0] == (1024, 16)             # model.py:440-441
    assert sh["vdec_d1_w"][0] == (4, 4, 128, 128)        # [kh,kw,out,in]  conv_util.py:83
    assert sh["encx_wg"][0] == (2 + 16, 32)
    assert sh["dec2_wg"][0] == (cfg.dec2_in + 16, 32)
    n = sum(int(np.prod(s)) for s, _ in sh.values())
    assert n > 1_000_000


def test_init_is_deterministic_and_follows_rules():
    cfg = small_cfg()
    a, b = init_params(cfg, 1), init_params(cfg, 1)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a["w_post_vae"], init_params(cfg, 2)["w_post_vae"])
    assert torch.all(a["encx_bg"] == 1) and torch.all(a["encx_bc"] == 0)      # TF GRUCell bias_start
    assert a["temporal_w"].abs().max() <= 0.2 + 1e-6                           # truncated normal, sigma 0.1
    assert 0.9 < a["w_hidden_enc1"].std() < 1.1                                # random_normal sigma 1


def test_synthetic_batch_layout():
    cfg = small_cfg()
    inp, tgt, eps, scene = make_batch(cfg, 3, seed=0, n_missing=2)
    assert inp.shape == (3, 8, 8, 3) and tgt.shape == (3, 8, 12, 3)
    assert eps.shape == (24, 3, 128) and scene.shape == (3, 32, 32, 3)
    assert torch.all(inp[0, :, 0, 0] == torch.arange(1, 9))                    # ids 1..N
    assert torch.all(inp[1, 6:, :, 0] == 0) and torch.all(inp[2, :, :, 0] != 0)
    assert torch.equal(make_batch(cfg, 3, seed=0, n_missing=2)[0], inp)


def test_logpolar_tables():
    cfg = small_cfg()
    r2, dirs = logpolar_tables(cfg)
    assert r2.shape == (7,) and dirs.shape == (6, 2)
    assert abs(float(r2[0]) - cfg.r_min ** 2) < 1e-9 and abs(float(r2[-1]) - cfg.r_max ** 2) < 1e-6
    assert torch.allclose((dirs ** 2).sum(1), torch.ones(6), atol=1e-6)


def test_model_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from desire_b200._lib import DesireError
    from desire_b200.model.model import DESIREModel
    with pytest.raises((DesireError, RuntimeError, AssertionError)):
        DESIREModel(small_cfg(), device="cpu")


def test_flatten_params_views_alias_one_aligned_flat_buffer():
    """Train-step host logic: every parameter is a view into ONE flat buffer (256-byte aligned starts, zero padding), so
    clip / Adam / the gradient all-reduce can work on the flat tensor."""
    import torch
    from desire_b200.config import init_params
    from desire_b200.engine import FLAT_ALIGN, flatten_params
    from helpers import small_cfg
    cfg = small_cfg(d_dim=16, max_num_obj=4, num_samples=2)
    P = init_params(cfg, 1)
    flat, views, offs = flatten_params(P, "cpu")
    assert set(views) == set(P) and flat.dtype == torch.float32
    used = torch.zeros_like(flat, dtype=torch.bool)
    for k, (o, cnt, shp) in offs.items():
        assert o % FLAT_ALIGN == 0 and cnt == P[k].numel() and tuple(views[k].shape) == tuple(P[k].shape)
        assert torch.equal(views[k], P[k])
        assert views[k].data_ptr() == flat.data_ptr() + 4 * o          # a view, not a copy
        assert not used[o:o + cnt].any()
        used[o:o + cnt] = True
    assert float(flat[~used].abs().sum()) == 0.0                        # padding stays zero
    views["output_b"].add_(1.0)                                         # writes go through to the flat buffer
    o, cnt, _ = offs["output_b"]
    assert torch.equal(flat[o:o + cnt], views["output_b"].reshape(-1))


def test_config_raises_on_flags_the_build_does_not_implement():
    # reference flags that change semantics (train.py:32,34,86; model/model.py:48,130,137-141) must not be dropped silently
    for kw in (dict(stride=2), dict(num_layers=2), dict(model="lstm"), dict(model="rnn")):
        with pytest.raises(ValueError):
            DesireConfig(**kw).validate()
    DesireConfig(stride=1, num_layers=1, model="gru").validate()


def test_vae_layers_reject_inference_phase_and_other_activations():
    # model/model.py:457-462,476-481 pass phase=train + ELU; anything else is not what the kernels compute
    from desire_b200.model.model import DESIREModel
    chk = DESIREModel._check_activ_phase
    chk("vae_encoder", None, None)
    chk("vae_encoder", type("elu", (), {"__name__": "elu"}), "train")
    with pytest.raises(ValueError):
        chk("vae_decoder", None, "infer")
    with pytest.raises(ValueError):
        chk("vae_decoder", "relu", None)
