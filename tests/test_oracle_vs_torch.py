"""CPU: pin the numpy oracle op by op against an INDEPENDENT implementation built from
torch.nn.functional (different code path: cuDNN-style conv/conv_transpose with explicit TF pad/crop,
instance_norm, grid_sample, a hand-written TF-GRU in torch) and against closed-form known answers.
The reference ships no golden vectors, so this is what anchors the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import np_batch, np_params, np_tables, rel_l2, small_cfg
from oracle import desire_oracle as O

rng = np.random.default_rng(7)


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).double()


# ------------------------------------------------------------------ convs
def torch_conv_tf(x, w, b, stride, padding):
    xt = t(x).permute(0, 3, 1, 2)
    wt = t(w).permute(3, 2, 0, 1)                      # [kh,kw,in,out] -> [out,in,kh,kw]
    if padding == "SAME":
        Hi, k = x.shape[1], w.shape[0]
        out = -(-Hi // stride)
        tot = max((out - 1) * stride + k - Hi, 0)
        xt = F.pad(xt, (tot // 2, tot - tot // 2, tot // 2, tot - tot // 2))
    y = F.conv2d(xt, wt, t(b), stride=stride)
    return y.permute(0, 2, 3, 1).numpy()


def torch_deconv_tf(x, w, b, stride, padding):
    xt = t(x).permute(0, 3, 1, 2)
    wt = t(w).permute(3, 2, 0, 1)                      # [kh,kw,out,in] -> [in,out,kh,kw]
    y = F.conv_transpose2d(xt, wt, None, stride=stride)
    if padding == "SAME":
        Hi, k = x.shape[1], w.shape[0]
        Ho = Hi * stride
        pb = max((Hi - 1) * stride + k - Ho, 0) // 2
        y = y[:, :, pb:pb + Ho, pb:pb + Ho]
    return (y + t(b).view(1, -1, 1, 1)).permute(0, 2, 3, 1).numpy()


@pytest.mark.parametrize("Hi,Ci,Co,k,s,pad", [(32, 1, 32, 5, 2, "SAME"), (16, 32, 64, 5, 2, "SAME"), (8, 64, 128, 5, 1, "VALID"),
                                              (9, 3, 4, 5, 2, "SAME"), (12, 3, 16, 5, 1, "SAME")])
def test_conv2d_tf(Hi, Ci, Co, k, s, pad):
    x = rng.normal(size=(3, Hi, Hi, Ci))
    w = rng.normal(size=(k, k, Ci, Co))
    b = rng.normal(size=Co)
    assert rel_l2(O.conv2d_tf(x, w, b, s, pad), torch_conv_tf(x, w, b, s, pad)) < 1e-12


@pytest.mark.parametrize("Hi,Ci,Co,k,s,pad", [(1, 128, 128, 4, 1, "VALID"), (4, 128, 64, 5, 1, "VALID"), (8, 64, 32, 5, 2, "SAME"),
                                              (16, 32, 1, 5, 2, "SAME")])
def test_deconv2d_tf(Hi, Ci, Co, k, s, pad):
    x = rng.normal(size=(2, Hi, Hi, Ci))
    w = rng.normal(size=(k, k, Co, Ci))
    b = rng.normal(size=Co)
    got = O.deconv2d_tf(x, w, b, s, pad)
    assert got.shape[1] == O.deconv_output_size(Hi, k, s, pad)
    assert rel_l2(got, torch_deconv_tf(x, w, b, s, pad)) < 1e-12


def test_deconv_is_transpose_of_tf_same_conv():
    """conv2d_transpose(SAME) must be the adjoint of conv2d(SAME): <conv(x), y> == <x, deconv(y)>."""
    x = rng.normal(size=(2, 16, 16, 3))
    y = rng.normal(size=(2, 8, 8, 5))
    w = rng.normal(size=(5, 5, 3, 5))                  # conv layout [kh,kw,in,out] == deconv layout [kh,kw,out',in'] with out'=3,in'=5
    lhs = (O.conv2d_tf(x, w, np.zeros(5), 2, "SAME") * y).sum()
    rhs = (x * O.deconv2d_tf(y, w, np.zeros(3), 2, "SAME")).sum()
    assert abs(lhs - rhs) < 1e-9 * abs(lhs)


def test_deconv_output_sizes_match_reference_rules():
    # utils/convolutional_vae_util.py:154-157,164-167 and the decoder chain model/model.py:465-468
    assert O.deconv_output_size(1, 4, 1, "VALID") == 4
    assert O.deconv_output_size(4, 5, 1, "VALID") == 8
    assert O.deconv_output_size(8, 5, 2, "SAME") == 16
    assert O.deconv_output_size(16, 5, 2, "SAME") == 32
    with pytest.raises(ValueError):
        O.deconv_output_size(4, 5, 1, "FULL")


def test_bn_rowwise_is_instance_norm():
    x = rng.normal(size=(4, 8, 8, 16))
    g, b = rng.normal(size=16), rng.normal(size=16)
    ref = F.instance_norm(t(x).permute(0, 3, 1, 2), weight=t(g), bias=t(b), eps=1e-3).permute(0, 2, 3, 1).numpy()
    assert rel_l2(O.bn_rowwise(x, g, b), ref) < 1e-12


# ------------------------------------------------------------------ CVAE stacks end to end vs torch
def test_vae_encoder_decoder_vs_torch():
    cfg = small_cfg()
    P = np_params(cfg, dtype=np.float64)
    # non-trivial BN affine so gamma/beta placement is tested
    for k in P:
        if k.endswith("_g"):
            P[k] = 1.0 + 0.1 * rng.normal(size=P[k].shape)
        if k.endswith("_be") or (k.startswith("v") and k.endswith("_b")):
            P[k] = 0.1 * rng.normal(size=P[k].shape)
    v = np.maximum(rng.normal(size=(3, 1024)), 0)
    mu, lv = O.vae_encoder(v, P, cfg.Z)
    x = v.reshape(3, 32, 32, 1)
    for n, s, pad in (("venc_c1", 2, "SAME"), ("venc_c2", 2, "SAME"), ("venc_c3", 1, "VALID")):
        y = torch_conv_tf(x, P[n + "_w"], P[n + "_b"], s, pad)
        y = F.instance_norm(t(y).permute(0, 3, 1, 2), weight=t(P[n + "_g"]), bias=t(P[n + "_be"]), eps=1e-3)
        x = F.elu(y).permute(0, 2, 3, 1).numpy()
    p = x.reshape(3, -1) @ P["venc_fc_w"] + P["venc_fc_b"]
    assert rel_l2(mu, p[:, :cfg.Z]) < 1e-10 and rel_l2(lv, p[:, cfg.Z:]) < 1e-10

    z = rng.normal(size=(4, cfg.Z))
    xr = O.vae_decoder(z, P)
    x = z.reshape(4, 1, 1, -1)
    for n, s, pad, act in (("vdec_d1", 1, "VALID", F.elu), ("vdec_d2", 1, "VALID", F.elu), ("vdec_d3", 2, "SAME", F.elu),
                           ("vdec_d4", 2, "SAME", torch.sigmoid)):
        y = torch_deconv_tf(x, P[n + "_w"], P[n + "_b"], s, pad)
        y = F.instance_norm(t(y).permute(0, 3, 1, 2), weight=t(P[n + "_g"]), bias=t(P[n + "_be"]), eps=1e-3)
        x = act(y).permute(0, 2, 3, 1).numpy()
    assert xr.shape == (4, 1024)
    assert rel_l2(xr, x.reshape(4, -1)) < 1e-10


# ------------------------------------------------------------------ GRU (TF-1.x semantics)
def test_gru_cell_semantics():
    H, I = 6, 3
    x, h = rng.normal(size=(4, I)), rng.normal(size=(4, H))
    wg, bg = rng.normal(size=(I + H, 2 * H)), rng.normal(size=2 * H)
    wc, bc = rng.normal(size=(I + H, H)), rng.normal(size=H)
    got = O.gru_cell(x, h, wg, bg, wc, bc)
    xt, ht = t(x), t(h)
    g = torch.sigmoid(torch.cat([xt, ht], 1) @ t(wg) + t(bg))
    r, u = g[:, :H], g[:, H:]                                    # order (r, u)
    c = torch.tanh(torch.cat([xt, r * ht], 1) @ t(wc) + t(bc))   # reset before the matmul
    assert rel_l2(got, (u * ht + (1 - u) * c).numpy()) < 1e-12
    # known answer: zero weights, TF bias init (gate 1, candidate 0): h' = sigmoid(1) * h
    z = O.gru_cell(x, h, np.zeros_like(wg), np.ones_like(bg), np.zeros_like(wc), np.zeros_like(bc))
    assert np.allclose(z, h / (1 + np.exp(-1.0)))


def test_gru_differs_from_cudnn_style():
    """Guard: applying the reset AFTER the matmul (cuDNN/PyTorch GRU) is a different function."""
    H = 5
    x, h = rng.normal(size=(3, 2)), rng.normal(size=(3, H))
    wg, bg = rng.normal(size=(2 + H, 2 * H)), rng.normal(size=2 * H)
    wc, bc = rng.normal(size=(2 + H, H)), rng.normal(size=H)
    g = O.sigmoid(np.concatenate([x, h], 1) @ wg + bg)
    r, u = g[:, :H], g[:, H:]
    c_after = np.tanh(x @ wc[:2] + r * (h @ wc[2:]) + bc)
    assert rel_l2(O.gru_cell(x, h, wg, bg, wc, bc), u * h + (1 - u) * c_after) > 1e-3


def test_decoder_hoisting_identity():
    """Decoder-1 feeds the same input every step: hoisting x@W_x out of the loop is exact algebra."""
    H = 8
    xz, h0 = rng.normal(size=(5, H)), rng.normal(size=(5, H))
    wg, bg = rng.normal(size=(2 * H, 2 * H)), rng.normal(size=2 * H)
    wc, bc = rng.normal(size=(2 * H, H)), rng.normal(size=H)
    hs = O.gru_decode(xz, h0, wg, bg, wc, bc, 4)
    xg, xc = xz @ wg[:H] + bg, xz @ wc[:H] + bc
    h = h0
    for s in range(4):
        g = O.sigmoid(xg + h @ wg[H:])
        r, u = g[:, :H], g[:, H:]
        h = u * h + (1 - u) * np.tanh(xc + (r * h) @ wc[H:])
        assert rel_l2(hs[:, s], h) < 1e-12


# ------------------------------------------------------------------ small ops / known answers
def test_tconv_is_depthwise_valid_conv():
    M, T, Cm = 5, 8, 100
    X = rng.normal(size=(M, T, 2))
    w, b = rng.normal(size=(T, 2, Cm)), rng.normal(size=2 * Cm)
    # torch depthwise conv1d over time: input [M, 2, T], groups=2, out channel = c*Cm + j
    wt = t(w).permute(1, 2, 0).reshape(2 * Cm, 1, T)
    ref = F.relu(F.conv1d(t(X).permute(0, 2, 1), wt, t(b), groups=2))[:, :, 0].numpy()
    assert rel_l2(O.tconv(X, w, b), ref) < 1e-12


def test_kld_known_answers():
    # N(0,1) posterior: zero divergence; closed form otherwise (model/model.py:587-589)
    assert O.kld_loss(np.zeros((3, 4)), np.zeros((3, 4))) == 0.0
    mu, lv = np.array([[1.0, -2.0]]), np.array([[0.5, -1.0]])
    exp = -0.5 * ((1 + 0.5 - 1 - np.exp(0.5)) + (1 - 1.0 - 4 - np.exp(-1.0)))
    assert abs(O.kld_rows(mu, lv)[0] - exp) < 1e-12


def test_reparam_layout_and_literal_form():
    M, K, Z = 3, 4, 8
    mu, lv, eps = rng.normal(size=(M, Z)), rng.normal(size=(M, Z)), rng.normal(size=(M, K, Z))
    z = O.reparam(mu, lv, eps)
    assert z.shape == (M * K, Z)
    assert np.allclose(z[1 * K + 2], mu[1] + np.sqrt(np.exp(lv[1])) * eps[1, 2])


def test_mask_gate_softmax_rows():
    R, K, H = 6, 3, 16
    xr, w, b = rng.random((R, 1024)), rng.normal(size=(1024, H)) * 0.05, rng.normal(size=H)
    Hx = rng.normal(size=(R // K, H))
    xz = O.mask_gate(xr, w, b, Hx, K)
    ref = F.softmax(F.relu(t(xr) @ t(w) + t(b)), 1).numpy() * np.repeat(Hx, K, 0)
    assert rel_l2(xz, ref) < 1e-12


def test_masked_cost_skips_id0():
    rows = np.array([1.0, 2.0, 3.0, 4.0])
    assert O.masked_cost(rows, np.array([1, 0, 1, 0])) == 2.0


def test_feature_pool_and_split_readout_follow_reference_shapes():
    # model/model.py:286-311: 7 decoder states split into T chunks, chunk[0],[1] = (x,y); pool = [x*rho[:C], y*rho[C:]]
    R, Td, T, H, Cm = 2, 7, 8, 16, 100
    hs = rng.normal(size=(R, Td, H))
    y = O.readout_split(hs, T)
    assert y.shape == (R, Td, T, 2)
    assert y[1, 3, 5, 0] == hs[1, 3, 5 * 2] and y[1, 3, 5, 1] == hs[1, 3, 5 * 2 + 1]
    rho = rng.random((R, 2 * Cm))
    fp = O.feature_pool(y.reshape(R, Td * T, 2), rho, 1).reshape(R, Td, T, 2 * Cm)
    assert fp.shape == (R, 7, 8, 200)
    assert np.allclose(fp[0, 2, 4, :Cm], y[0, 2, 4, 0] * rho[0, :Cm])
    assert np.allclose(fp[0, 2, 4, Cm:], y[0, 2, 4, 1] * rho[0, Cm:])


# ------------------------------------------------------------------ stage 2
def test_bilinear_gather_vs_grid_sample():
    B, Hm, Wm, Cs, R = 2, 9, 13, 5, 40
    fmap = rng.normal(size=(B, Hm, Wm, Cs))
    pos = rng.random((B, R, 2)) * 1.2 - 0.1                   # some points outside -> clamped to the border
    got = O.bilinear_gather(fmap, pos)
    grid = torch.from_numpy(np.clip(pos, 0, 1) * 2 - 1).double().view(B, R, 1, 2)
    ref = F.grid_sample(t(fmap).permute(0, 3, 1, 2), grid, mode="bilinear", padding_mode="border", align_corners=True)
    assert rel_l2(got, ref[:, :, :, 0].permute(0, 2, 1).numpy()) < 1e-10


def test_logpolar_bins_match_atan2_log_definition():
    cfg = small_cfg()
    r2e, dirs = np_tables(cfg, np.float64)
    d = rng.normal(size=(4000, 2)) * 0.2
    got = O.logpolar_bin(d[:, 0], d[:, 1], r2e, dirs)
    r = np.hypot(d[:, 0], d[:, 1])
    rb = np.floor(cfg.n_rad * np.log(r / cfg.r_min) / np.log(cfg.r_max / cfg.r_min)).astype(int)
    ab = np.floor((np.arctan2(d[:, 1], d[:, 0]) + np.pi) / (2 * np.pi) * cfg.n_ang).astype(int) % cfg.n_ang
    ref = np.where((rb >= 0) & (rb < cfg.n_rad), rb * cfg.n_ang + ab, -1)
    assert (got != ref).mean() < 2e-3                          # only razor-edge points may differ
    assert (got >= 0).mean() > 0.5


def test_social_pool_vectorised_equals_loops_and_excludes_self_and_missing():
    cfg = small_cfg()
    r2e, dirs = np_tables(cfg)
    B, N, K, H = 2, 7, 3, 5
    pos = (rng.random((B, N, K, 2)) * 0.4).astype(np.float32)
    h = rng.normal(size=(B, N, K, H)).astype(np.float32)
    mask = np.ones((B, N), bool)
    mask[1, 4:] = False
    a = O.social_pool(pos, h, mask, r2e, dirs)
    b = O.social_pool_loops(pos, h, mask, r2e, dirs)
    assert np.array_equal(a != 0, b != 0) and rel_l2(a, b) < 1e-6
    # a lone agent pools nothing
    solo = O.social_pool(pos[:1, :1], h[:1, :1], np.ones((1, 1), bool), r2e, dirs)
    assert not solo.any()
    # masked neighbours contribute nothing: changing their h changes nothing
    h2 = h.copy()
    h2[1, 4:] += 100
    assert np.array_equal(O.social_pool(pos, h2, mask, r2e, dirs)[1, :4], a[1, :4])


def test_forward_runs_and_fp32_floor_vs_fp64():
    cfg = small_cfg()
    out32 = O.forward(np_params(cfg), dict(K=cfg.K, Z=cfg.Z, ioc_iters=cfg.ioc_iters), *np_batch(cfg, 2, n_missing=2), *np_tables(cfg))
    out64 = O.forward(np_params(cfg, dtype=np.float64), dict(K=cfg.K, Z=cfg.Z, ioc_iters=cfg.ioc_iters),
                      *np_batch(cfg, 2, n_missing=2, dtype=np.float64), *np_tables(cfg, np.float64))
    for k in ("Yhat", "Y_refined", "ioc_scores", "cost"):
        assert out32[k].dtype == np.float32 or np.isscalar(out32[k]) or out32[k].dtype == np.float32
        assert rel_l2(out32[k], out64[k]) < 5e-5, k            # the fp32 rounding floor of the oracle itself
