"""Differentiable twin of oracle/desire_oracle.py (TEST INFRASTRUCTURE — never imported by the product).

PARITY UNPINNED BY THE REFERENCE (see desire_oracle.py): the reference never runs its optimiser
(`self.gradients = tf.gradients(self.cost, tvars)` / Adam at model/model.py:388-394 are built but
train.py only evaluates `model.cost`), so there is nothing to pin gradients against.  This module restates
the SAME forward arithmetic as desire_oracle.py with torch tensor ops on the CPU (float64 by default) so
that `torch.autograd` yields the gradients the backward kernels must reproduce; tests/test_oracle_torch.py
checks that its forward values agree with the NumPy oracle to round-off, which is what ties the gradient
reference to the oracle the forward path is held to.

Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
S_IMG = 32


def elu(x):
    return torch.where(x > 0, x, torch.exp(torch.clamp(x, max=0)) - 1.0)


def tconv(X, temporal_w, temporal_b):
    """model/model.py:116-133.  X [M,T,2], w [T,2,C] -> [M,2C]"""
    out = torch.einsum("mtc,tcj->mcj", X, temporal_w).reshape(X.shape[0], -1) + temporal_b
    return torch.relu(out)


def gru_cell(x, h, wg, bg, wc, bc):
    """TF-1.x GRUCell: gates (r|u), reset before the matmul."""
    H = h.shape[1]
    g = torch.sigmoid(torch.cat([x, h], 1) @ wg + bg)
    r, u = g[:, :H], g[:, H:]
    c = torch.tanh(torch.cat([x, r * h], 1) @ wc + bc)
    return u * h + (1.0 - u) * c


def gru_encode(X, wg, bg, wc, bc):
    h = torch.zeros(X.shape[0], wc.shape[1], dtype=X.dtype)
    for t in range(X.shape[1]):
        h = gru_cell(X[:, t], h, wg, bg, wc, bc)
    return h


def _same_pad(inp, k, s):
    out = -(-inp // s)
    total = max((out - 1) * s + k - inp, 0)
    return out, total // 2, total - total // 2


def conv2d_tf(x, w, b, stride, padding):
    """x NHWC, w [kh,kw,in,out]; TF SAME alignment (total//2 before)."""
    kh, kw = w.shape[0], w.shape[1]
    xn = x.permute(0, 3, 1, 2)
    if padding == "SAME":
        _, pt, pb = _same_pad(x.shape[1], kh, stride)
        _, pl, pr = _same_pad(x.shape[2], kw, stride)
        xn = F.pad(xn, (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), None, stride)
    return y.permute(0, 2, 3, 1) + b


def deconv2d_tf(x, w, b, stride, padding):
    """x NHWC, w [kh,kw,out,in] (utils/convolutional_vae_util.py:83); SAME keeps full[pad : pad + in*s]."""
    kh, kw = w.shape[0], w.shape[1]
    Hi, Wi = x.shape[1], x.shape[2]
    # conv_transpose2d weight layout [in, out, kh, kw]
    full = F.conv_transpose2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), None, stride)
    if padding == "SAME":
        Ho, Wo = Hi * stride, Wi * stride
        pt = max((Hi - 1) * stride + kh - Ho, 0) // 2
        pl = max((Wi - 1) * stride + kw - Wo, 0) // 2
        full = full[:, :, pt:pt + Ho, pl:pl + Wo]
    return full.permute(0, 2, 3, 1) + b


def bn_rowwise(x, gamma, beta, eps=BN_EPS):
    mean = x.mean(dim=(1, 2), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(1, 2), keepdim=True)
    return gamma * ((x - mean) / torch.sqrt(var + eps)) + beta


def vae_encoder(v, P, Z):
    M = v.shape[0]
    x = v.reshape(M, S_IMG, S_IMG, 1)
    x = elu(bn_rowwise(conv2d_tf(x, P["venc_c1_w"], P["venc_c1_b"], 2, "SAME"), P["venc_c1_g"], P["venc_c1_be"]))
    x = elu(bn_rowwise(conv2d_tf(x, P["venc_c2_w"], P["venc_c2_b"], 2, "SAME"), P["venc_c2_g"], P["venc_c2_be"]))
    x = elu(bn_rowwise(conv2d_tf(x, P["venc_c3_w"], P["venc_c3_b"], 1, "VALID"), P["venc_c3_g"], P["venc_c3_be"]))
    p = x.reshape(M, -1) @ P["venc_fc_w"] + P["venc_fc_b"]
    return p[:, :Z], p[:, Z:]


def reparam(mu, logvar, eps):
    z = mu[:, None, :] + torch.sqrt(torch.exp(logvar))[:, None, :] * eps
    return z.reshape(-1, mu.shape[1])


def vae_decoder(z, P):
    R = z.shape[0]
    x = z.reshape(R, 1, 1, -1)
    x = elu(bn_rowwise(deconv2d_tf(x, P["vdec_d1_w"], P["vdec_d1_b"], 1, "VALID"), P["vdec_d1_g"], P["vdec_d1_be"]))
    x = elu(bn_rowwise(deconv2d_tf(x, P["vdec_d2_w"], P["vdec_d2_b"], 1, "VALID"), P["vdec_d2_g"], P["vdec_d2_be"]))
    x = elu(bn_rowwise(deconv2d_tf(x, P["vdec_d3_w"], P["vdec_d3_b"], 2, "SAME"), P["vdec_d3_g"], P["vdec_d3_be"]))
    x = torch.sigmoid(bn_rowwise(deconv2d_tf(x, P["vdec_d4_w"], P["vdec_d4_b"], 2, "SAME"), P["vdec_d4_g"], P["vdec_d4_be"]))
    return x.reshape(R, -1)


def mask_gate(xr, w, b, Hx, K):
    beta = torch.softmax(torch.relu(xr @ w + b), dim=1)
    return beta * Hx.repeat_interleave(K, dim=0)


def gru_decode(x_z, h0, wg, bg, wc, bc, T):
    h = h0
    outs = []
    for _ in range(T):
        h = gru_cell(x_z, h, wg, bg, wc, bc)
        outs.append(h)
    return torch.stack(outs, 1)


def kld_rows(mu, logvar):
    return -0.5 * torch.sum(1.0 + logvar - mu * mu - torch.exp(logvar), dim=1)


def recon_rows(Yhat, Y, K):
    M = Y.shape[0]
    d = Yhat.reshape(M, K, *Y.shape[1:]) - Y[:, None]
    return (d * d).sum(dim=(2, 3)).mean(dim=1)


def masked_cost(rows, mask):
    mask = mask.to(rows.dtype)
    return (rows * mask).sum() / mask.sum()


def to_torch(P, dtype=torch.float64, requires_grad=True):
    out = {}
    for k, v in P.items():
        t = torch.as_tensor(np.asarray(v)).to(dtype).clone()
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def generate_forward(P, cfg, input_data, target_data, eps):
    """Sample generation a2-a13, same statement as desire_oracle.forward (stage 1).  Tensors of P carry
    requires_grad; returns a dict of every named intermediate with `cost` a scalar tensor."""
    dt = next(iter(P.values())).dtype
    input_data = torch.as_tensor(np.asarray(input_data)).to(dt)
    target_data = torch.as_tensor(np.asarray(target_data)).to(dt)
    eps = torch.as_tensor(np.asarray(eps)).to(dt)
    B, N, Tp, _ = input_data.shape
    Tf = target_data.shape[2]
    K, Z = cfg["K"], cfg["Z"]
    M = B * N
    X = input_data.reshape(M, Tp, 3)[..., 1:3]
    Y = target_data.reshape(M, Tf, 3)[..., 1:3]
    mask = (input_data[:, :, 0, 0] != 0).reshape(M)
    out = {}
    out["rho_i"] = tconv(X, P["temporal_w"], P["temporal_b"])
    Hx = gru_encode(X, P["encx_wg"], P["encx_bg"], P["encx_wc"], P["encx_bc"])
    Hy = gru_encode(Y, P["ency_wg"], P["ency_bg"], P["ency_wc"], P["ency_bc"])
    out["H_x"], out["H_y"] = Hx, Hy
    v = torch.relu(torch.cat([Hx, Hy], 1) @ P["w_hidden_enc1"] + P["b_hidden_enc1"])
    out["vae_inputs"] = v
    mu, logvar = vae_encoder(v, P, Z)
    out["z_mean"], out["z_log_sigma_sq"] = mu, logvar
    z = reparam(mu, logvar, eps)
    out["zval"] = z
    xr = vae_decoder(z, P)
    out["x_reconstr_mean"] = xr
    x_z = mask_gate(xr, P["w_post_vae"], P["b_post_vae"], Hx, K)
    out["x_z"] = x_z
    hs = gru_decode(x_z, Hx.repeat_interleave(K, dim=0), P["dec1_wg"], P["dec1_bg"], P["dec1_wc"], P["dec1_bc"], Tf)
    out["output_states"] = hs
    x_last = X[:, -1].repeat_interleave(K, dim=0)
    Yhat = hs @ P["output_w"] + P["output_b"] + x_last[:, None, :]
    out["Yhat"] = Yhat
    out["kld_rows"] = kld_rows(mu, logvar)
    out["recon_rows"] = recon_rows(Yhat, Y, K)
    out["cost"] = masked_cost(out["recon_rows"] + out["kld_rows"], mask)
    return out


# --------------------------------------------------------------------------- stage 2 (D11 forward, D13 loss)
def scene_cnn(img, P):
    x = torch.relu(conv2d_tf(img, P["scene_c1_w"], P["scene_c1_b"], 2, "SAME"))
    x = torch.relu(conv2d_tf(x, P["scene_c2_w"], P["scene_c2_b"], 1, "SAME"))
    return torch.relu(conv2d_tf(x, P["scene_c3_w"], P["scene_c3_b"], 1, "SAME"))


def _bilinear_const(fmap, pos_np):
    """Bilinear gather with CONSTANT positions (numpy): linear in fmap.  pos [B,R,2] -> [B,R,Cs]."""
    B, Hm, Wm, Cs = fmap.shape
    px = np.clip(pos_np[..., 0] * (Wm - 1), 0, Wm - 1)
    py = np.clip(pos_np[..., 1] * (Hm - 1), 0, Hm - 1)
    x0, y0 = np.floor(px).astype(np.int64), np.floor(py).astype(np.int64)
    x1, y1 = np.minimum(x0 + 1, Wm - 1), np.minimum(y0 + 1, Hm - 1)
    fx = torch.as_tensor(px - x0).to(fmap.dtype)[..., None]
    fy = torch.as_tensor(py - y0).to(fmap.dtype)[..., None]
    bi = np.arange(B)[:, None]
    v00, v01, v10, v11 = fmap[bi, y0, x0], fmap[bi, y0, x1], fmap[bi, y1, x0], fmap[bi, y1, x1]
    top = v00 + fx * (v01 - v00)
    bot = v10 + fx * (v11 - v10)
    return top + fy * (bot - top)


def _pool_const(pos_np, h, mask_np, r2_edges, dirs, bin_dtype=None):
    """Log-polar social pooling with CONSTANT bins (numpy, same binning code as the NumPy oracle): linear in h.
    pos [B,N,K,2], h [B,N,K,H] torch -> [B,N,K,G,H].  bin_dtype=np.float32 bins in the kernels' arithmetic."""
    from oracle import desire_oracle as O
    if bin_dtype is not None:
        pos_np, r2_edges, dirs = pos_np.astype(bin_dtype), r2_edges.astype(bin_dtype), dirs.astype(bin_dtype)
    B, N, K, H = h.shape
    G = (r2_edges.shape[0] - 1) * dirs.shape[0]
    outs = []
    eye = np.eye(N, dtype=bool)[:, :, None]
    for b in range(B):
        dx = pos_np[b, None, :, :, 0] - pos_np[b, :, None, :, 0]
        dy = pos_np[b, None, :, :, 1] - pos_np[b, :, None, :, 1]
        bins = O.logpolar_bin(dx, dy, r2_edges, dirs)
        ok = (bins >= 0) & (mask_np[b][None, :, None] > 0) & ~eye
        onehot = ((bins[..., None] == np.arange(G)) & ok[..., None]).astype(np.float64)     # [i,j,K,G]
        cnt = np.maximum(onehot.sum(1), 1)                                                   # [i,K,G]
        A = torch.as_tensor(onehot / cnt[:, None]).to(h.dtype)
        outs.append(torch.einsum("ijkg,jkh->ikgh", A, h[b]))
    return torch.stack(outs, 0)


def ioc_train_forward(P, cfg, gen, input_data, target_data, scene_img, r2_edges, dirs, bin_dtype=None):
    """Stage 2 with the D13 training loss.  `gen` = the (detached) outputs of stage 1 as numpy: Yhat, H_x,
    feature_pooling — the IOC module treats them as constants (stage-wise training), and inside an iteration
    every feature is computed from stop_gradient(Y_it); only Y_{it+1} = Y_it + dY_it carries gradient.
    Returns dict(ioc_scores [iters,MK], Y_refined, ioc_rows [M], ioc_cost)."""
    dt = next(iter(P.values())).dtype
    inp = np.asarray(input_data, np.float64)
    B, N, Tp, _ = inp.shape
    K, iters = cfg["K"], cfg["ioc_iters"]
    M, MK = B * N, B * N * K
    mask_np = (inp[:, :, 0, 0] != 0)
    Y_true = torch.as_tensor(np.asarray(target_data, np.float64).reshape(M, -1, 3)[..., 1:3]).to(dt)
    T = Y_true.shape[1]
    x_last = np.repeat(inp.reshape(M, Tp, 3)[:, -1, 1:3], K, 0)
    Hx = torch.as_tensor(np.asarray(gen["H_x"], np.float64)).to(dt)
    H = Hx.shape[1]
    fpool = torch.as_tensor(np.asarray(gen["feature_pooling"], np.float64)).to(dt)
    fmap = scene_cnn(torch.as_tensor(np.asarray(scene_img, np.float64)).to(dt), P)
    Y = torch.as_tensor(np.asarray(gen["Yhat"], np.float64)).to(dt)
    r2_edges, dirs = np.asarray(r2_edges), np.asarray(dirs)
    rows = torch.zeros(M, dtype=dt)
    scores = []
    for _ in range(iters):
        Yd = Y.detach().numpy()
        h2 = Hx.repeat_interleave(K, dim=0)
        s = torch.zeros(MK, dtype=dt)
        prev = x_last
        for t in range(T):
            v = torch.as_tensor(Yd[:, t] - prev).to(dt)
            fv = torch.relu(v @ P["ioc_vel_w"] + P["ioc_vel_b"])
            fs = _bilinear_const(fmap, Yd[:, t].reshape(B, N * K, 2)).reshape(MK, -1)
            pooled = _pool_const(Yd[:, t].reshape(B, N, K, 2), h2.reshape(B, N, K, H), mask_np, r2_edges, dirs, bin_dtype)
            fsp = torch.relu(pooled.reshape(MK, -1) @ P["ioc_sp_w"] + P["ioc_sp_b"])
            x_t = torch.cat([fv, fs, fpool[:, t], fsp], 1)
            h2 = gru_cell(x_t, h2, P["dec2_wg"], P["dec2_bg"], P["dec2_wc"], P["dec2_bc"])
            s = s + h2 @ P["ioc_score_w"] + P["ioc_score_b"]
            prev = Yd[:, t]
        dY = (h2 @ P["ioc_reg_w"] + P["ioc_reg_b"]).reshape(MK, T, 2)
        d = torch.as_tensor(np.sqrt(((Yd.reshape(M, K, T, 2) - Y_true.numpy()[:, None]) ** 2).sum(-1)).max(-1)).to(dt)
        q = torch.softmax(-d, dim=1)
        ce = -(q * torch.log_softmax(s.reshape(M, K), dim=1)).sum(dim=1)
        Y = Y + dY
        e = Y.reshape(M, K, T, 2) - Y_true[:, None]
        rows = rows + ce + (e * e).sum(dim=(2, 3)).mean(dim=1)
        scores.append(s)
    out = {"ioc_scores": torch.stack(scores, 0) if scores else torch.zeros(0, MK, dtype=dt), "Y_refined": Y,
           "ioc_rows": rows, "scene_features": fmap}
    out["ioc_cost"] = masked_cost(rows, torch.as_tensor(mask_np.reshape(M)))
    return out


def adam_reference(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, clip=0.0):
    """clip_by_global_norm over the whole list + TF-1.x AdamOptimizer update, on dicts of numpy arrays
    (float64).  Returns (p', m', v')."""
    if clip and clip > 0:
        norm = np.sqrt(sum(float((x.astype(np.float64) ** 2).sum()) for x in g.values()))
        s = clip / max(norm, clip)
    else:
        s = 1.0
    lr_t = lr * np.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p2, m2, v2 = {}, {}, {}
    for k in p:
        gi = g[k].astype(np.float64) * s
        m2[k] = beta1 * m[k] + (1 - beta1) * gi
        v2[k] = beta2 * v[k] + (1 - beta2) * gi * gi
        p2[k] = p[k] - lr_t * m2[k] / (np.sqrt(v2[k]) + eps)
    return p2, m2, v2
