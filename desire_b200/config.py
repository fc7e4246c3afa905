"""Sizes of the DESIRE hot path and the parameter inventory (shapes + initialisers).

The first block mirrors the reference's argparse flags (train.py:28-88) and the derived sizes in
DESIREModel.__init__ (model/model.py:43-60); the second block holds the knobs the build adds
(DESIGN.md "Resolved spec": D1 K samples, D2 pred_length, D11 stage-2 sizes).
"""
from __future__ import annotations

import dataclasses
import math
from collections import OrderedDict

import torch


@dataclasses.dataclass
class DesireConfig:
    # reference flags (train.py:30-88)
    rnn_size: int = 512            # only sets the CVAE image side S = int(sqrt(2*rnn_size)) (model.py:57-59)
    d_dim: int = 16                # H, GRU hidden size (BASELINE's "hidden")
    latent_size: int = 128         # Z
    seq_length: int = 8            # T_p
    max_num_obj: int = 60          # N
    stride: int = 1
    num_layers: int = 1            # train.py:32; MultiRNNCell depth (model.py:137-141) — only 1 is built
    model: str = "gru"             # train.py:34 "rnn, gru, or lstm" — the reference graph only ever builds GRUCells
    # added knobs
    pred_length: int = 12          # T_f   (D2)
    num_samples: int = 20          # K     (D1)
    ioc_iters: int = 2             # D11
    scene_size: int = 256          # scene image side (pixels); feature map is scene_size/2
    scene_channels: int = 32       # C_s
    vel_dim: int = 16              # F_v
    n_rad: int = 6                 # log-polar radial bins
    n_ang: int = 6                 # log-polar angular bins
    r_min: float = 0.01
    r_max: float = 0.5
    channel_multiplier: int = 100  # model.py:46
    exist_mode: int = 1            # D8: 1 = an agent exists if present at observed frame 0, at the last observed frame and
                                   # at every target frame (obj_id / target_obj_id of model.py:351-366); 0 = frame 0 only

    @property
    def H(self):
        return self.d_dim

    @property
    def Z(self):
        return self.latent_size

    @property
    def K(self):
        return self.num_samples

    @property
    def S(self):
        return int(math.sqrt(2 * self.rnn_size))

    @property
    def G(self):
        return self.n_rad * self.n_ang

    @property
    def dec2_in(self):
        """Decoder-2 input width: velocity fc + scene gather + feature_pooling + social fc."""
        return self.vel_dim + self.scene_channels + 2 * self.channel_multiplier + self.d_dim

    def validate(self):
        if self.S != 32:
            # the CVAE decoder always emits 32x32 (model.py:465-468 + convolutional_vae_util.py:154-157),
            # so w_post_vae [S*S, H] (model.py:440-441) only type-checks at S == 32 (D5)
            raise ValueError("rnn_size must give S = int(sqrt(2*rnn_size)) == 32 (got %d)" % self.S)
        # flags the reference accepts and this build does not implement must not be dropped silently
        if self.stride != 1:
            raise ValueError("stride=%d: only the reference default stride=1 of the temporal convolution is built "
                             "(model.py:48,130)" % self.stride)
        if self.num_layers != 1:
            raise ValueError("num_layers=%d: only single-layer GRUs are built (MultiRNNCell of model.py:137-141 with "
                             "the default num_layers=1)" % self.num_layers)
        if str(self.model).lower() != "gru":
            raise ValueError("model=%r: the path is built for GRU cells only (the reference graph constructs "
                             "rnn.GRUCell regardless of this flag, model.py:137,144)" % (self.model,))
        if self.d_dim % 4:
            raise ValueError("d_dim must be a multiple of 4")
        if self.latent_size % 4:
            raise ValueError("latent_size must be a multiple of 4")


def param_shapes(cfg: DesireConfig) -> "OrderedDict[str, tuple]":
    """Every trainable tensor of the path with its init rule (D10)."""
    H, Z, Tp, Tf, C, S = cfg.H, cfg.Z, cfg.seq_length, cfg.pred_length, cfg.channel_multiplier, cfg.S
    Cs, Fv, G = cfg.scene_channels, cfg.vel_dim, cfg.G
    sh = OrderedDict()

    def gru(prefix, I):
        sh[prefix + "_wg"] = ((I + H, 2 * H), "glorot")
        sh[prefix + "_bg"] = ((2 * H,), "one")          # TF GRUCell gate bias_start=1.0
        sh[prefix + "_wc"] = ((I + H, H), "glorot")
        sh[prefix + "_bc"] = ((H,), "zero")

    def bn(prefix, ch):
        sh[prefix + "_g"] = ((ch,), "one")
        sh[prefix + "_be"] = ((ch,), "zero")

    sh["temporal_w"] = ((Tp, 2, C), "trunc0.1")          # model.py:427-429 ([1,T,2,C] squeezed)
    sh["temporal_b"] = ((2 * C,), "normal1")             # model.py:430-431
    gru("encx", 2)
    gru("ency", 2)
    sh["w_hidden_enc1"] = ((2 * H, S * S), "normal1")    # model.py:434-437
    sh["b_hidden_enc1"] = ((S * S,), "normal1")
    for name, shp in (("venc_c1", (5, 5, 1, 32)), ("venc_c2", (5, 5, 32, 64)), ("venc_c3", (5, 5, 64, 128))):
        sh[name + "_w"] = (shp, "xavier_conv")
        sh[name + "_b"] = ((shp[3],), "zero")
        bn(name, shp[3])
    sh["venc_fc_w"] = ((4 * 4 * 128, 2 * Z), "glorot")
    sh["venc_fc_b"] = ((2 * Z,), "zero")
    for name, shp in (("vdec_d1", (4, 4, 128, Z)), ("vdec_d2", (5, 5, 64, 128)),
                      ("vdec_d3", (5, 5, 32, 64)), ("vdec_d4", (5, 5, 1, 32))):
        sh[name + "_w"] = (shp, "xavier_deconv")        # [kh,kw,out,in], conv_util.py:83
        sh[name + "_b"] = ((shp[2],), "zero")
        bn(name, shp[2])
    sh["w_post_vae"] = ((S * S, H), "normal1")           # model.py:440-443
    sh["b_post_vae"] = ((H,), "normal1")
    gru("dec1", H)                                       # D4: separate decoder-1 parameters
    sh["output_w"] = ((H, 2), "glorot")                  # D3
    sh["output_b"] = ((2,), "zero")
    # stage 2 (D11)
    sh["scene_c1_w"] = ((5, 5, 3, 16), "xavier_conv")
    sh["scene_c1_b"] = ((16,), "zero")
    sh["scene_c2_w"] = ((5, 5, 16, 32), "xavier_conv")
    sh["scene_c2_b"] = ((32,), "zero")
    sh["scene_c3_w"] = ((5, 5, 32, Cs), "xavier_conv")
    sh["scene_c3_b"] = ((Cs,), "zero")
    sh["ioc_vel_w"] = ((2, Fv), "glorot")
    sh["ioc_vel_b"] = ((Fv,), "zero")
    sh["ioc_sp_w"] = ((G * H, H), "glorot")
    sh["ioc_sp_b"] = ((H,), "zero")
    gru("dec2", cfg.dec2_in)
    sh["ioc_score_w"] = ((H,), "glorot_vec")
    sh["ioc_score_b"] = ((1,), "zero")
    sh["ioc_reg_w"] = ((H, 2 * Tf), "glorot")
    sh["ioc_reg_b"] = ((2 * Tf,), "zero")
    return sh


def init_params(cfg: DesireConfig, seed: int = 1, device="cpu") -> "OrderedDict[str, torch.Tensor]":
    """D10: temporal_w truncated-normal(0.1), the reference's explicit weights N(0,1)
    (model.py:427-443), library defaults elsewhere (Glorot-uniform GRU kernels, gate bias 1,
    xavier conv kernels, BN gamma 1 / beta 0).  Generated on CPU with a seeded torch.Generator
    so every rank and the oracle see the same bits."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for name, (shape, rule) in param_shapes(cfg).items():
        if rule == "zero":
            t = torch.zeros(shape)
        elif rule == "one":
            t = torch.ones(shape)
        elif rule == "normal1":
            t = torch.randn(shape, generator=g)
        elif rule == "trunc0.1":
            t = torch.randn(shape, generator=g).clamp_(-2, 2) * 0.1
        elif rule in ("glorot", "glorot_vec"):
            fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (shape[0], 1)
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shape, generator=g) * 2 - 1) * lim
        elif rule in ("xavier_conv", "xavier_deconv"):
            patch = shape[0] * shape[1]
            lim = math.sqrt(6.0 / (patch * (shape[2] + shape[3])))   # layers.xavier_init(n_in*patch, n_out*patch)
            t = (torch.rand(shape, generator=g) * 2 - 1) * lim
        else:
            raise ValueError(rule)
        out[name] = t.to(torch.float32).contiguous().to(device)
    return out


def logpolar_tables(cfg: DesireConfig):
    """Squared radial edges [n_rad+1] and sector boundary directions [n_ang,2] in fp32 — computed
    once on the host and handed verbatim to the kernel (and, in tests, to the oracle) so that
    binning is exact arithmetic on shared constants."""
    e = [cfg.r_min * (cfg.r_max / cfg.r_min) ** (i / cfg.n_rad) for i in range(cfg.n_rad + 1)]
    r2 = torch.tensor([x * x for x in e], dtype=torch.float64).to(torch.float32)
    th = [-math.pi + 2 * math.pi * i / cfg.n_ang for i in range(cfg.n_ang)]
    dirs = torch.tensor([[math.cos(t), math.sin(t)] for t in th], dtype=torch.float64).to(torch.float32)
    return r2, dirs
