"""CPU oracle for the DESIRE hot path (TEST INFRASTRUCTURE — never imported by the product).

PARITY UNPINNED BY THE REFERENCE.  tdavchev/DESIRE ships no tests, golden vectors or seeds
and cannot run (TensorFlow 1.3 + prettytensor are absent here and the graph build raises even
with them).  This file is therefore a *restatement*: every function follows
the reference line range cited in its docstring where code exists, the TF-1.x / prettytensor
library semantics where the reference only names a class, and the written spec in DESIGN.md
("Resolved spec") where the reference is silent (stage 2 / IOC).  It is
pinned two ways instead (tests/test_oracle_*.py):
  * op by op against an independent implementation built from torch.nn.functional on CPU
    (conv2d / conv_transpose2d with explicit TF-style pad+crop, instance_norm, F.grid_sample);
  * against closed-form known answers (kld, deconv output sizes, GRU fixed points).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  All arithmetic is numpy in the dtype of the inputs (float32 for parity,
float64 to report the fp32 floor); every random tensor (weights, eps) is an explicit argument.

Layouts (row-major, NHWC for images):
  M = B*N agent rows (scene-major: row = b*N + n), MK = M*K rows (row = m*K + k).
"""
from __future__ import annotations

import numpy as np

C_MULT = 100          # channel_multiplier, model/model.py:46
BN_EPS = 1e-3         # variance_epsilon, model/model.py:460,479
S_IMG = 32            # int(sqrt(2*rnn_size)) at rnn_size=512, model/model.py:57-59


# --------------------------------------------------------------------------- activations
def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def relu(x):
    return np.maximum(x, 0)


def elu(x):
    # tf.nn.elu: x if x > 0 else exp(x) - 1
    return np.where(x > 0, x, np.exp(np.minimum(x, 0)) - 1.0).astype(x.dtype)


# --------------------------------------------------------------------------- a2 temporal conv
def tconv(X, temporal_w, temporal_b):
    """rho_i, model/model.py:116-133 (weights :427-431).

    tf.nn.depthwise_conv2d(VALID) with filter [1, T, 2, C] over an input whose width is T gives
    one output column; output channel = c*C + j (TF depthwise channel order).
    X [M,T,2] (x,y only), temporal_w [T,2,C], temporal_b [2C]  ->  [M, 2C]
    """
    M = X.shape[0]
    out = np.einsum("mtc,tcj->mcj", X, temporal_w).reshape(M, -1) + temporal_b
    return relu(out)


# --------------------------------------------------------------------------- TF-1.x GRUCell
def gru_cell(x, h, wg, bg, wc, bc):
    """tf.contrib.rnn.GRUCell (TF 1.x; the reference only names the class, model/model.py:137,144).

    gates  = sigmoid([x,h] @ wg + bg);  r, u = split(gates, 2)      (order r then u)
    cand   = tanh([x, r*h] @ wc + bc)                               (reset BEFORE the matmul)
    h'     = u*h + (1-u)*cand
    wg [I+H, 2H], wc [I+H, H]
    """
    H = h.shape[1]
    g = sigmoid(np.concatenate([x, h], 1) @ wg + bg)
    r, u = g[:, :H], g[:, H:]
    c = np.tanh(np.concatenate([x, r * h], 1) @ wc + bc)
    return u * h + (1.0 - u) * c


def gru_encode(X, wg, bg, wc, bc):
    """static_rnn over the T frames from a zero state, model/model.py:152-167,233-241.
    X [M,T,I] -> final state [M,H]."""
    H = wc.shape[1]
    h = np.zeros((X.shape[0], H), X.dtype)
    for t in range(X.shape[1]):
        h = gru_cell(X[:, t], h, wg, bg, wc, bc)
    return h


# --------------------------------------------------------------------------- a5 fc_c
def fc(x, w, b, act=None):
    y = x @ w + b
    if act == "relu":
        y = relu(y)
    elif act == "elu":
        y = elu(y)
    elif act == "sigmoid":
        y = sigmoid(y)
    return y


def fc_c(Hx, Hy, w, b):
    """relu(xw_plus_b(concat(H_x,H_y))), model/model.py:243-251."""
    return fc(np.concatenate([Hx, Hy], 1), w, b, "relu")


# --------------------------------------------------------------------------- conv / deconv, TF rules
def _same_pad(inp, k, s):
    out = -(-inp // s)
    total = max((out - 1) * s + k - inp, 0)
    return out, total // 2, total - total // 2


def conv2d_tf(x, w, b, stride, padding):
    """tf.nn.conv2d NHWC, filter [kh,kw,in,out] (prettytensor conv2d, used at model/model.py:484-486).
    SAME pads total//2 before, the rest after (TF library rule)."""
    M, Hi, Wi, Ci = x.shape
    kh, kw, _, Co = w.shape
    if padding == "SAME":
        Ho, pt, pb = _same_pad(Hi, kh, stride)
        Wo, pl, pr = _same_pad(Wi, kw, stride)
        x = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    else:
        Ho = (Hi - kh) // stride + 1
        Wo = (Wi - kw) // stride + 1
    out = np.zeros((M, Ho, Wo, Co), x.dtype)
    for ky in range(kh):
        for kx in range(kw):
            patch = x[:, ky:ky + stride * (Ho - 1) + 1:stride, kx:kx + stride * (Wo - 1) + 1:stride, :]
            out += patch @ w[ky, kx]
    return out + b


def deconv_output_size(inp, k, s, padding):
    """get2d_deconv_output_size, utils/convolutional_vae_util.py:141-169."""
    if padding == "VALID":
        return (inp - 1) * s + k
    if padding == "SAME":
        return inp * s
    raise ValueError("Invalid value for padding: %r" % padding)


def deconv2d_tf(x, w, b, stride, padding):
    """tf.nn.conv2d_transpose NHWC, filter [kh,kw,out,in]; utils/convolutional_vae_util.py:83,113-121.

    conv2d_transpose is the input-gradient of the forward conv whose SAME padding is
    total//2 before: full[(i*s+k)] += x[i]*w[k]; SAME keeps full[pad_before : pad_before+in*s]."""
    M, Hi, Wi, Ci = x.shape
    kh, kw, Co, _ = w.shape
    Hf, Wf = (Hi - 1) * stride + kh, (Wi - 1) * stride + kw
    full = np.zeros((M, Hf, Wf, Co), x.dtype)
    for ky in range(kh):
        for kx in range(kw):
            full[:, ky:ky + stride * (Hi - 1) + 1:stride, kx:kx + stride * (Wi - 1) + 1:stride, :] += x @ w[ky, kx].T
    Ho, Wo = deconv_output_size(Hi, kh, stride, padding), deconv_output_size(Wi, kw, stride, padding)
    if padding == "SAME":
        pt = max((Hi - 1) * stride + kh - Ho, 0) // 2
        pl = max((Wi - 1) * stride + kw - Wo, 0) // 2
        full = full[:, pt:pt + Ho, pl:pl + Wo, :]
    return full + b


def bn_rowwise(x, gamma, beta, eps=BN_EPS):
    """prettytensor batch_normalize in Phase.train on the reference's per-object batch of ONE
    (model/model.py:457-462,476-481): moments over (N,H,W) == over H*W of that row; biased
    variance; gamma*(x-mean)/sqrt(var+eps)+beta (scale_after_normalization=True)."""
    mean = x.mean(axis=(1, 2), keepdims=True)
    var = ((x - mean) ** 2).mean(axis=(1, 2), keepdims=True)
    return gamma * ((x - mean) / np.sqrt(var + eps)) + beta


# --------------------------------------------------------------------------- a6 / a8 conv-CVAE
def vae_encoder(v, P, latent_size):
    """model/model.py:471-492.  v [M, S*S] -> (mean [M,Z], logvar [M,Z]).
    conv -> +bias -> BN -> ELU per layer (prettytensor conv2d op order); no BN on the final fc
    (DESIGN.md D5)."""
    M = v.shape[0]
    x = v.reshape(M, S_IMG, S_IMG, 1)
    x = elu(bn_rowwise(conv2d_tf(x, P["venc_c1_w"], P["venc_c1_b"], 2, "SAME"), P["venc_c1_g"], P["venc_c1_be"]))
    x = elu(bn_rowwise(conv2d_tf(x, P["venc_c2_w"], P["venc_c2_b"], 2, "SAME"), P["venc_c2_g"], P["venc_c2_be"]))
    x = elu(bn_rowwise(conv2d_tf(x, P["venc_c3_w"], P["venc_c3_b"], 1, "VALID"), P["venc_c3_g"], P["venc_c3_be"]))
    p = x.reshape(M, -1) @ P["venc_fc_w"] + P["venc_fc_b"]
    return p[:, :latent_size], p[:, latent_size:]


def reparam(mu, logvar, eps):
    """zval = mean + sqrt(exp(logvar))*eps, model/model.py:260-264 (D6: keep sqrt(exp()) literally).
    mu, logvar [M,Z], eps [M,K,Z] -> [M*K, Z]."""
    z = mu[:, None, :] + np.sqrt(np.exp(logvar))[:, None, :] * eps
    return z.reshape(-1, mu.shape[1])


def vae_decoder(z, P):
    """model/model.py:453-469 with deconv2d of utils/convolutional_vae_util.py:31-135
    (conv_transpose -> +bias -> BN -> activation).  z [R,Z] -> [R, 1024]."""
    R = z.shape[0]
    x = z.reshape(R, 1, 1, -1)
    x = elu(bn_rowwise(deconv2d_tf(x, P["vdec_d1_w"], P["vdec_d1_b"], 1, "VALID"), P["vdec_d1_g"], P["vdec_d1_be"]))
    x = elu(bn_rowwise(deconv2d_tf(x, P["vdec_d2_w"], P["vdec_d2_b"], 1, "VALID"), P["vdec_d2_g"], P["vdec_d2_be"]))
    x = elu(bn_rowwise(deconv2d_tf(x, P["vdec_d3_w"], P["vdec_d3_b"], 2, "SAME"), P["vdec_d3_g"], P["vdec_d3_be"]))
    x = sigmoid(bn_rowwise(deconv2d_tf(x, P["vdec_d4_w"], P["vdec_d4_b"], 2, "SAME"), P["vdec_d4_g"], P["vdec_d4_be"]))
    return x.reshape(R, -1)


# --------------------------------------------------------------------------- a9 mask
def softmax(x):
    e = np.exp(x - x.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


def mask_gate(xr, w_post_vae, b_post_vae, Hx, K):
    """beta = softmax(relu(xr@W2+b2)); x_z = beta * H_x, model/model.py:271-280.
    xr [M*K,1024], Hx [M,H] -> [M*K,H]."""
    beta = softmax(relu(xr @ w_post_vae + b_post_vae))
    return beta * np.repeat(Hx, K, axis=0)


# --------------------------------------------------------------------------- a10 / a11 decoder 1
def gru_decode(x_z, h0, wg, bg, wc, bc, T):
    """seq2seq.rnn_decoder with the same input every step, model/model.py:279-285 (D1, D4).
    x_z [R,H] constant input, h0 [R,H] -> all states [R,T,H]."""
    h = h0
    outs = []
    for _ in range(T):
        h = gru_cell(x_z, h, wg, bg, wc, bc)
        outs.append(h)
    return np.stack(outs, 1)


def readout_linear(hs, output_w, output_b, x_last):
    """D3: explicit linear H->2 per step (the author's commented output_w/output_b,
    model/model.py:315-321,445-449), anchored at the last observed position.
    hs [R,T,H], x_last [R,2] -> Yhat [R,T,2]."""
    return hs @ output_w + output_b + x_last[:, None, :]


def readout_split(hs, n_chunks):
    """Reference read-out, model/model.py:286-289,301,306: each state is split into n_chunks
    chunks; elements 0 and 1 of chunk t are read as (x, y).  hs [R,Td,H] -> [R,Td,n_chunks,2]."""
    R, Td, H = hs.shape
    cs = H // n_chunks
    return hs.reshape(R, Td, n_chunks, cs)[..., :2]


def feature_pool(Yhat, rho, K):
    """[y_x * rho[:C], y_y * rho[C:]], model/model.py:291-311 (D12).
    Yhat [M*K, T, 2], rho [M, 2C] -> [M*K, T, 2C]."""
    C = rho.shape[1] // 2
    r = np.repeat(rho, K, axis=0)[:, None, :]
    return np.concatenate([Yhat[..., 0:1] * r[..., :C], Yhat[..., 1:2] * r[..., C:]], -1)


# --------------------------------------------------------------------------- a12 / a13 losses
def kld_rows(mu, logvar):
    """latent_loss per row, model/model.py:587-589."""
    return -0.5 * np.sum(1.0 + logvar - mu * mu - np.exp(logvar), axis=1)


def kld_loss(mu, logvar):
    """reduce_mean(latent_loss), model/model.py:591."""
    return kld_rows(mu, logvar).mean()


def recon_rows(Yhat, Y, K):
    """D7: mean over K of sum_t ||Y - Yhat_k||^2 per agent.  Yhat [M*K,T,2], Y [M,T,2] -> [M]."""
    M = Y.shape[0]
    d = Yhat.reshape(M, K, *Y.shape[1:]) - Y[:, None]
    return (d * d).sum(axis=(2, 3)).mean(axis=1)


def masked_cost(loss_rows, mask):
    """cost = sum_{existing} loss / #existing, model/model.py:193-196,351-366,374-376 (D8)."""
    mask = mask.astype(loss_rows.dtype)
    return (loss_rows * mask).sum() / mask.sum()


# --------------------------------------------------------------------------- a14 stage 2 (D11)
def scene_cnn(img, P):
    """D11 scene CNN rho(I): conv5 s2 ->16, conv5 s1 ->32, conv5 s1 ->32, SAME, ReLU each.
    img [B,Hi,Wi,3] -> [B,Hi/2,Wi/2,32]."""
    x = relu(conv2d_tf(img, P["scene_c1_w"], P["scene_c1_b"], 2, "SAME"))
    x = relu(conv2d_tf(x, P["scene_c2_w"], P["scene_c2_b"], 1, "SAME"))
    x = relu(conv2d_tf(x, P["scene_c3_w"], P["scene_c3_b"], 1, "SAME"))
    return x


def bilinear_gather(fmap, pos):
    """Bilinear sample of the scene feature map at normalised positions (D11).
    fmap [B,Hm,Wm,Cs]; pos [B,R,2] (x,y) in [0,1] -> map px = x*(Wm-1), py = y*(Hm-1), clamped
    to the border.  -> [B,R,Cs]."""
    B, Hm, Wm, Cs = fmap.shape
    dt = fmap.dtype
    px = np.clip(pos[..., 0] * dt.type(Wm - 1), 0, Wm - 1).astype(dt)
    py = np.clip(pos[..., 1] * dt.type(Hm - 1), 0, Hm - 1).astype(dt)
    x0 = np.floor(px).astype(np.int64)
    y0 = np.floor(py).astype(np.int64)
    x1 = np.minimum(x0 + 1, Wm - 1)
    y1 = np.minimum(y0 + 1, Hm - 1)
    fx = (px - x0.astype(dt))[..., None]
    fy = (py - y0.astype(dt))[..., None]
    bi = np.arange(B)[:, None]
    v00, v01 = fmap[bi, y0, x0], fmap[bi, y0, x1]
    v10, v11 = fmap[bi, y1, x0], fmap[bi, y1, x1]
    top = v00 + fx * (v01 - v00)
    bot = v10 + fx * (v11 - v10)
    return top + fy * (bot - top)


def logpolar_tables(n_rad, n_ang, r_min, r_max, dtype=np.float32):
    """Bin tables shared verbatim by oracle and kernel so that binning needs only +,*,compare:
    squared radial edges r_e^2 (n_rad+1) and the n_ang sector boundary directions (cos, sin)."""
    e = r_min * (r_max / r_min) ** (np.arange(n_rad + 1, dtype=np.float64) / n_rad)
    th = -np.pi + 2 * np.pi * np.arange(n_ang, dtype=np.float64) / n_ang
    return (e * e).astype(dtype), np.stack([np.cos(th), np.sin(th)], 1).astype(dtype)


def logpolar_bin(dx, dy, r2_edges, dirs):
    """-> bin index in [0, n_rad*n_ang) or -1.  Radial bin = #edges passed - 1; angular sector s is
    the first s with cross(e_s,d) >= 0 and cross(e_{s+1},d) < 0 (fallback n_ang-1).  Each product
    and difference is rounded separately (the kernel uses __fmul_rn/__fsub_rn to match)."""
    n_rad = r2_edges.shape[0] - 1
    n_ang = dirs.shape[0]
    r2 = dx * dx + dy * dy
    rb = (r2[..., None] >= r2_edges).sum(-1) - 1
    valid = (rb >= 0) & (rb < n_rad)
    cr = dirs[:, 0] * dy[..., None] - dirs[:, 1] * dx[..., None]        # [..., n_ang]
    ge = cr >= 0
    cond = ge & ~np.roll(ge, -1, axis=-1)
    ab = np.where(cond.any(-1), cond.argmax(-1), n_ang - 1)
    return np.where(valid, rb * n_ang + ab, -1)


def logpolar_margin(dx, dy, r2_edges, dirs):
    """Distance, in position units, from d = (dx, dy) to the nearest edge of the log-polar grid: the radial
    circles |d| = r_e and, for d inside the radial range, the sector boundary rays.  A perturbation of the
    positions smaller than this cannot change logpolar_bin(d).  Test infrastructure for the parity tests
    (which (scene, sample) groups may legitimately differ between two fp32 implementations)."""
    r = np.sqrt(dx * dx + dy * dy)
    edges = np.sqrt(np.asarray(r2_edges, np.float64))
    m_r = np.abs(r[..., None].astype(np.float64) - edges).min(-1)
    cr = np.abs(dirs[:, 0].astype(np.float64) * dy[..., None] - dirs[:, 1].astype(np.float64) * dx[..., None])
    inside = (r >= edges[0]) & (r < edges[-1])
    m_a = np.where(inside, cr.min(-1), np.inf)
    return np.minimum(m_r, m_a)


def group_margin(pos, mask, r2_edges, dirs):
    """min over the pairs (i, j != i, j existing) of logpolar_margin, per (scene, sample): pos [B,N,K,2] -> [B,K]."""
    B, N, K, _ = pos.shape
    out = np.full((B, K), np.inf)
    eye = np.eye(N, dtype=bool)[:, :, None]
    for b in range(B):
        dx = pos[b, None, :, :, 0] - pos[b, :, None, :, 0]
        dy = pos[b, None, :, :, 1] - pos[b, :, None, :, 1]
        m = logpolar_margin(dx, dy, r2_edges, dirs)
        m = np.where((mask[b][None, :, None] > 0) & ~eye, m, np.inf)
        out[b] = m.reshape(N * N, K).min(0)
    return out


def social_pool(pos, h, mask, r2_edges, dirs):
    """D11 log-polar social pooling: for row (b,i,k) average the hidden vectors h[b,j,k] of the
    other existing agents j != i of the same scene and sample index into the bin of
    (pos_j - pos_i).  pos [B,N,K,2], h [B,N,K,H], mask [B,N] -> [B,N,K,G,H].  Vectorised; the
    literal triple loop is social_pool_loops (tests check they agree)."""
    B, N, K, H = h.shape
    G = (r2_edges.shape[0] - 1) * dirs.shape[0]
    out = np.zeros((B, N, K, G, H), h.dtype)
    eye = np.eye(N, dtype=bool)[:, :, None]
    for b in range(B):
        dx = pos[b, None, :, :, 0] - pos[b, :, None, :, 0]       # [i, j, K] = pos_j - pos_i
        dy = pos[b, None, :, :, 1] - pos[b, :, None, :, 1]
        bins = logpolar_bin(dx, dy, r2_edges, dirs)               # [i, j, K]
        ok = (bins >= 0) & (mask[b][None, :, None] > 0) & ~eye
        if N > 128:
            # crowds (BASELINE configs[4], N=1024): the same sums as one BLAS product per sample,
            # [(i,g), j] @ [j, H] — the einsum below takes minutes there
            for k in range(K):
                oh = ((bins[:, :, k, None] == np.arange(G)) & ok[:, :, k, None]).astype(h.dtype)     # [i,j,G]
                s = (oh.transpose(0, 2, 1).reshape(N * G, N) @ h[b, :, k]).reshape(N, G, H)
                out[b, :, k] = s / np.maximum(oh.sum(1), 1)[..., None]
            continue
        onehot = ((bins[..., None] == np.arange(G)) & ok[..., None]).astype(h.dtype)   # [i,j,K,G]
        s = np.einsum("ijkg,jkh->ikgh", onehot, h[b])
        cnt = onehot.sum(1)                                        # [i,K,G]
        out[b] = s / np.maximum(cnt, 1)[..., None]
    return out


def social_pool_loops(pos, h, mask, r2_edges, dirs):
    """Literal per-pair statement of social_pool (small cases only)."""
    B, N, K, H = h.shape
    G = (r2_edges.shape[0] - 1) * dirs.shape[0]
    out = np.zeros((B, N, K, G, H), h.dtype)
    for b in range(B):
        for i in range(N):
            for k in range(K):
                cnt = np.zeros(G, h.dtype)
                for j in range(N):
                    if j == i or not mask[b, j]:
                        continue
                    d = pos[b, j, k] - pos[b, i, k]
                    g = int(logpolar_bin(d[0:1], d[1:2], r2_edges, dirs)[0])
                    if g < 0:
                        continue
                    out[b, i, k, g] += h[b, j, k]
                    cnt[g] += 1
                nz = cnt > 0
                out[b, i, k, nz] /= cnt[nz, None]
    return out


def ioc_refine(Yhat, x_last, Hx, fpool, fmap, mask, P, dims, iters, r2_edges, dirs, snapshots=None, margins=None):
    """D11 ranking & refinement (absent in the reference, marker model/model.py:312-313).

    Per iteration, per step t and row (b,n,k):
      fv  = relu((Y_t - Y_{t-1}) @ vel_w + vel_b)                    velocity, Y_{-1} = x_last
      fs  = bilinear(rho(I)[b], Y_t)                                 scene context
      fsp = relu(social_pool(Y_t, h2_{t-1}) @ sp_w + sp_b)           interaction
      x_t = [fv, fs, feature_pooling_t, fsp]                         (D12 feeds feature_pooling)
      h2_t = GRU2(x_t, h2_{t-1}),  h2_{-1} = H_x;   s += h2_t @ score_w + score_b
    dY = h2_T @ reg_w + reg_b;  Y <- Y + dY.   Returns (scores [iters, MK], Y refined [MK,T,2]).
    dims = (B, N, K)."""
    B, N, K = dims
    MK, T, _ = Yhat.shape
    H = Hx.shape[1]
    Y = Yhat.copy()
    scores = []
    xl = x_last
    h_init = np.repeat(Hx, K, axis=0)
    for _ in range(iters):
        h2 = h_init
        s = np.zeros(MK, Y.dtype)
        prev = xl
        for t in range(T):
            fv = relu((Y[:, t] - prev) @ P["ioc_vel_w"] + P["ioc_vel_b"])
            fs = bilinear_gather(fmap, Y[:, t].reshape(B, N * K, 2)).reshape(MK, -1)
            pooled = social_pool(Y[:, t].reshape(B, N, K, 2), h2.reshape(B, N, K, H), mask, r2_edges, dirs)
            if margins is not None:      # [B,K] per (iteration, step): how close this step's binning is to an edge
                margins.append(group_margin(Y[:, t].reshape(B, N, K, 2), mask, r2_edges, dirs))
            fsp = relu(pooled.reshape(MK, -1) @ P["ioc_sp_w"] + P["ioc_sp_b"])
            x_t = np.concatenate([fv, fs, fpool[:, t], fsp], 1)
            h2 = gru_cell(x_t, h2, P["dec2_wg"], P["dec2_bg"], P["dec2_wc"], P["dec2_bc"])
            s = s + h2 @ P["ioc_score_w"] + P["ioc_score_b"]
            prev = Y[:, t]
        dY = (h2 @ P["ioc_reg_w"] + P["ioc_reg_b"]).reshape(MK, T, 2)
        if snapshots is not None:
            snapshots.append(Y)          # the trajectories this iteration scored (Y_it)
        Y = Y + dY
        scores.append(s)
    if snapshots is not None:
        snapshots.append(Y)
    return np.stack(scores, 0), Y


def ioc_loss_rows(scores, snaps, Y_true, K):
    """D13 (paper sec. 3.3-3.4; absent in the reference): per agent m, summed over the IOC iterations,
      CE_it  = - sum_k q log p,   p = softmax_k(score_it),  q = softmax_k(-max_t ||Y - Y_it(k)||_2)
      REG_it = mean_k sum_t ||Y - Y_{it+1}(k)||^2
    scores [iters, M*K]; snaps = [Y_0 .. Y_iters] each [M*K,T,2]; Y_true [M,T,2] -> [M]."""
    M = Y_true.shape[0]
    rows = np.zeros(M, Y_true.dtype)
    for it in range(scores.shape[0]):
        d = np.sqrt(((snaps[it].reshape(M, K, -1, 2) - Y_true[:, None]) ** 2).sum(-1)).max(-1)     # [M,K]
        q = softmax(-d)
        s = scores[it].reshape(M, K)
        logp = s - s.max(axis=1, keepdims=True)
        logp = logp - np.log(np.exp(logp).sum(axis=1, keepdims=True))
        ce = -(q * logp).sum(axis=1)
        e = snaps[it + 1].reshape(M, K, -1, 2) - Y_true[:, None]
        reg = (e * e).sum(axis=(2, 3)).mean(axis=1)
        rows = rows + ce + reg
    return rows


def philox_randn(seed, offset, n):
    """The device noise source restated (desire_randn_fwd; replaces tf.random_normal of model/model.py:262):
    Philox4x32-10 (Salmon et al. 2011) on counter (i, offset) under key `seed`, four words -> four 24-bit uniforms
    -> two Box-Muller pairs.  -> float32 [n]."""
    quads = (n + 3) // 4
    i = np.arange(quads, dtype=np.uint64)
    M = np.uint64(0xFFFFFFFF)
    c = [i & M, i >> np.uint64(32), np.full(quads, offset & 0xFFFFFFFF, np.uint64), np.full(quads, (offset >> 32) & 0xFFFFFFFF, np.uint64)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(0xD2511F53) * c[0], np.uint64(0xCD9E8D57) * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & M, p1 & M, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & M, p0 & M]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & M, (k1 + np.uint64(0xBB67AE85)) & M
    z = np.empty((quads, 4), np.float32)
    for h in range(2):
        u1 = ((c[2 * h] >> np.uint64(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
        u2 = ((c[2 * h + 1] >> np.uint64(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
        rad = np.sqrt(np.float32(-2.0) * np.log(u1))
        th = np.float32(6.283185307179586) * u2
        z[:, 2 * h], z[:, 2 * h + 1] = rad * np.cos(th), rad * np.sin(th)
    return z.reshape(-1)[:n]


def existence_mask(input_data, target_data, mode=1):
    """D8.  model/model.py:351-366 leaves an object out of the cost when `obj_id` (its id at the first frame of the
    observed window, :214) or `target_obj_id` (undefined in the reference; its id in the target data) equals the
    non-existent id 0 (:206).  mode 0: observed frame 0 only.  mode 1: the object must also be present at the last
    observed frame (the read-out is anchored there, D3) and at every target frame.  -> [B,N] bool."""
    m = input_data[:, :, 0, 0] != 0
    if mode == 1:
        m = m & (input_data[:, :, -1, 0] != 0) & (target_data[:, :, :, 0] != 0).all(-1)
    return m


# --------------------------------------------------------------------------- whole path
def forward(P, cfg, input_data, target_data, eps, scene_img, r2_edges, dirs):
    """Sample generation (a2-a13) + ranking/refinement (a14).

    input_data [B,N,Tp,3] (id,x,y), target_data [B,N,Tf,3], eps [B*N,K,Z], scene_img [B,Hi,Wi,3].
    cfg: dict(K, H, Z, ioc_iters).  Returns a dict of every named intermediate."""
    B, N, Tp, _ = input_data.shape
    Tf = target_data.shape[2]
    K, Z = cfg["K"], cfg["Z"]
    M = B * N
    X = input_data.reshape(M, Tp, 3)[..., 1:3]
    Y = target_data.reshape(M, Tf, 3)[..., 1:3]
    mask = existence_mask(input_data, target_data, cfg.get("exist_mode", 1))     # D8: id 0 == non-existent
    out = {}
    out["rho_i"] = tconv(X, P["temporal_w"], P["temporal_b"])
    Hx = gru_encode(X, P["encx_wg"], P["encx_bg"], P["encx_wc"], P["encx_bc"])
    Hy = gru_encode(Y, P["ency_wg"], P["ency_bg"], P["ency_wc"], P["ency_bc"])
    out["H_x"], out["H_y"] = Hx, Hy
    v = fc_c(Hx, Hy, P["w_hidden_enc1"], P["b_hidden_enc1"])
    out["vae_inputs"] = v
    mu, logvar = vae_encoder(v, P, Z)
    out["z_mean"], out["z_log_sigma_sq"] = mu, logvar
    z = reparam(mu, logvar, eps)
    out["zval"] = z
    xr = vae_decoder(z, P)
    out["x_reconstr_mean"] = xr
    x_z = mask_gate(xr, P["w_post_vae"], P["b_post_vae"], Hx, K)
    out["x_z"] = x_z
    hs = gru_decode(x_z, np.repeat(Hx, K, 0), P["dec1_wg"], P["dec1_bg"], P["dec1_wc"], P["dec1_bc"], Tf)
    out["output_states"] = hs
    x_last = np.repeat(X[:, -1], K, 0)
    Yhat = readout_linear(hs, P["output_w"], P["output_b"], x_last)
    out["Yhat"] = Yhat
    fpool = feature_pool(Yhat, out["rho_i"], K)
    out["feature_pooling"] = fpool
    out["kld_rows"] = kld_rows(mu, logvar)
    out["recon_rows"] = recon_rows(Yhat, Y, K)
    out["cost"] = masked_cost(out["recon_rows"] + out["kld_rows"], mask.reshape(M))
    fmap = scene_cnn(scene_img, P)
    out["scene_features"] = fmap
    snaps = []
    margins = [] if cfg.get("margins") else None
    scores, Yref = ioc_refine(Yhat, x_last, Hx, fpool, fmap, mask, P, (B, N, K),
                              cfg["ioc_iters"], r2_edges, dirs, snaps, margins)
    out["ioc_scores"], out["Y_refined"] = scores, Yref
    if margins:
        out["bin_margin"] = np.min(np.stack(margins, 0), 0)      # [B,K], min over iterations and steps
    out["ioc_rows"] = ioc_loss_rows(scores, snaps, Y, K)
    out["ioc_cost"] = masked_cost(out["ioc_rows"], mask.reshape(M)) if cfg["ioc_iters"] > 0 else np.zeros((), Y.dtype)
    return out
