"""utils/convolutional_vae_util.py surface: size rules and spec expansion (CPU), deconv2d / vae_encoder / vae_decoder
through the kernel library vs the oracle (GPU)."""
import numpy as np
import pytest

from helpers import TOL, np_params, rel_l2, small_cfg


def test_deconv_output_size_rules_and_errors():
    from desire_b200.utils.convolutional_vae_util import _kernel, _stride, get2d_deconv_output_size
    assert get2d_deconv_output_size(1, 1, 4, 4, 1, 1, "VALID") == (4, 4)          # model.py:465 chain
    assert get2d_deconv_output_size(4, 4, 5, 5, 1, 1, "VALID") == (8, 8)
    assert get2d_deconv_output_size(8, 8, 5, 5, 2, 2, "SAME") == (16, 16)
    assert get2d_deconv_output_size(16, 16, 5, 5, 2, 2, "SAME") == (32, 32)
    assert get2d_deconv_output_size(None, 7, 3, 3, 2, 3, "SAME") == (None, 21)
    with pytest.raises(ValueError):
        get2d_deconv_output_size(4, 4, 5, 5, 1, 1, "FULL")
    assert _kernel(5) == [5, 5] and _kernel([3]) == [3, 3] and _kernel((2, 4)) == [2, 4]
    assert _stride(None) == [1, 1, 1, 1] and _stride(2) == [1, 2, 2, 1] and _stride([3]) == [1, 3, 3, 1]
    assert _stride((2, 3)) == [1, 2, 3, 1] and _stride([1, 2, 2, 1]) == [1, 2, 2, 1]


@pytest.mark.gpu
@pytest.mark.parametrize("R,Hin,Cin,k,s,edges,Cout,bn,act,bias", [
    (5, 1, 128, 4, 1, "VALID", 128, True, "elu", True), (70, 4, 128, 5, 1, "VALID", 64, True, "elu", True),
    (70, 8, 64, 5, 2, "SAME", 32, True, "elu", True), (9, 16, 32, 5, 2, "SAME", 1, True, "sigmoid", True),
    (6, 6, 10, 3, 2, "VALID", 7, False, "relu", False), (4, 5, 12, 3, 1, "SAME", 16, False, None, True)])
def test_deconv2d_matches_oracle(R, Hin, Cin, k, s, edges, Cout, bn, act, bias):
    import torch
    from oracle import desire_oracle as O
    from desire_b200.utils.convolutional_vae_util import deconv2d
    rng = np.random.default_rng(R + Hin)
    x = rng.standard_normal((R, Hin, Hin, Cin)).astype(np.float32)
    P = {"weights": (rng.standard_normal((k, k, Cout, Cin)) * 0.2).astype(np.float32)}
    if bias:
        P["bias"] = rng.standard_normal(Cout).astype(np.float32)
    if bn:
        P["gamma"], P["beta"] = (rng.random(Cout) + 0.5).astype(np.float32), rng.standard_normal(Cout).astype(np.float32)
    ref = O.deconv2d_tf(x, P["weights"], P.get("bias", np.zeros(Cout, np.float32)), s, edges)
    if bn:
        ref = O.bn_rowwise(ref, P["gamma"], P["beta"])
    ref = {"elu": O.elu, "relu": O.relu, "sigmoid": O.sigmoid, None: lambda v: v}[act](ref)
    y, _ = deconv2d(torch.from_numpy(x).cuda(), k, Cout, stride=s, activation_fn=act, bias=bias, edges=edges,
                    batch_normalize=bn, params={n: torch.from_numpy(v) for n, v in P.items()})
    torch.cuda.synchronize()
    assert tuple(y.shape) == ref.shape
    assert rel_l2(y.cpu().numpy().reshape(-1), ref.reshape(-1)) <= TOL


@pytest.mark.gpu
def test_deconv2d_creates_reference_style_params_and_rejects_bad_input():
    import torch
    from desire_b200.utils.convolutional_vae_util import deconv2d
    x = torch.randn(3, 4, 4, 8, device="cuda")
    y, p = deconv2d(x, 5, 16, stride=2, batch_normalize=True, activation_fn="elu")
    assert tuple(y.shape) == (3, 8, 8, 16) and tuple(p["weights"].shape) == (5, 5, 16, 8)
    assert float(p["weights"].abs().max()) <= (6.0 / (25 * 24)) ** 0.5 + 1e-6 and float(p["bias"].abs().max()) == 0
    with pytest.raises(ValueError):
        deconv2d(torch.randn(3, 4, 8, device="cuda"), 5, 6)
    with pytest.raises(ValueError):
        deconv2d(x, 5, 6, init=lambda s: np.zeros(s), stddev=0.1)


@pytest.mark.gpu
def test_model_vae_encoder_decoder_methods_match_oracle():
    import torch
    from oracle import desire_oracle as O
    from desire_b200.model.model import DESIREModel
    cfg = small_cfg(d_dim=32, max_num_obj=6, num_samples=2)
    m = DESIREModel(cfg, seed=1)
    P = np_params(cfg)
    rng = np.random.default_rng(0)
    v = rng.random((70, 1024)).astype(np.float32)
    mu, lv = m.vae_encoder(torch.from_numpy(v), cfg.Z)
    rm, rl = O.vae_encoder(v, P, cfg.Z)
    z = rng.standard_normal((90, cfg.Z)).astype(np.float32)
    xr = m.vae_decoder(torch.from_numpy(z), 1024)
    torch.cuda.synchronize()
    assert rel_l2(mu.cpu().numpy(), rm) <= TOL and rel_l2(lv.cpu().numpy(), rl) <= TOL
    assert rel_l2(xr.cpu().numpy(), O.vae_decoder(z, P)) <= TOL
    with pytest.raises(ValueError):
        m.vae_decoder(torch.from_numpy(z), 512)
