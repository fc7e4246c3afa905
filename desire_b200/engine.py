"""Host-side driver of the hot path: owns the device buffers and issues the C-ABI calls in order.

PyTorch is plumbing here (device memory, streams, CUDA graphs); every arithmetic op of the path
runs inside libdesire_b200.so.  All buffers are allocated once per (config, batch) so a step can
be captured into a CUDA graph and replayed.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .config import DesireConfig, logpolar_tables


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


FLAT_ALIGN = 64   # floats (256 B): the kernels read weights with 128-bit loads


def flatten_params(params: dict, device):
    """One flat fp32 device buffer holding every tensor (each start 256-byte aligned, padding zero) and the
    dict of views into it.  The train step all-reduces / clips / updates the flat buffer in one go."""
    offs, n = {}, 0
    for k, v in params.items():
        offs[k] = (n, v.numel(), tuple(v.shape))
        n += (v.numel() + FLAT_ALIGN - 1) // FLAT_ALIGN * FLAT_ALIGN
    flat = torch.zeros(n, dtype=torch.float32, device=device)
    views = {}
    for k, (o, cnt, shp) in offs.items():
        views[k] = flat[o:o + cnt].view(shp)
        views[k].copy_(params[k])
    return flat, views, offs


def existing_agents(obs, tgt, mode=1):
    """[B,N] bool: the agents that desire_existence_fwd keeps (host-side mirror, used for the count normaliser)."""
    m = obs[:, :, 0, 0] != 0
    if mode == 1:
        m = m & (obs[:, :, -1, 0] != 0)
        if tgt is not None:
            m = m & (tgt[:, :, :, 0] != 0).all(-1)
    return m


class HotPath:
    """Sample generation (a2-a13) + IOC ranking/refinement (a14) for a fixed batch of B scenes."""

    def __init__(self, cfg: DesireConfig, params: dict, B: int, device="cuda:0"):
        cfg.validate()
        self.cfg, self.B, self.device = cfg, B, torch.device(device)
        if self.device.type != "cuda":
            raise _lib.DesireError("the DESIRE hot path only runs on a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        # tensors already on the device (e.g. views of DESIREModel's flat buffer) are used in place
        self.P = {k: v.to(self.device, torch.float32).contiguous() for k, v in params.items()}
        r2, dirs = logpolar_tables(cfg)
        self.r2_edges, self.dirs = r2.to(self.device), dirs.to(self.device)
        N, K, H, Zl, Tp, Tf, Cm = cfg.max_num_obj, cfg.K, cfg.H, cfg.Z, cfg.seq_length, cfg.pred_length, cfg.channel_multiplier
        M, R = B * N, B * N * K
        self.M, self.R = M, R
        Hm = (cfg.scene_size + 1) // 2
        self.Hm = Hm
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)
        self.buf = dict(
            rho_i=f(M, 2 * Cm), HxHy=f(M, 2 * H), vae_inputs=f(M, cfg.S * cfg.S), mu_logvar=f(M, 2 * Zl),
            zval=f(R, Zl), x_reconstr_mean=f(R, cfg.S * cfg.S), x_z=f(R, H), output_states=f(R, Tf, H),
            Yhat=f(R, Tf, 2), feature_pooling=f(R, Tf, 2 * Cm), kld_rows=f(M), recon_rows=f(M), cost=f(2),
            scene_features=f(B, Hm, Hm, cfg.scene_channels), Y_refined=f(R, Tf, 2), obs_eff=f(B, N, Tp, 3),
            eps=f(M, K, Zl),
            ioc_scores=f(max(cfg.ioc_iters, 1), R),
        )
        self._structs()
        lib = self.lib
        self.ioc_dims = _lib.IocDims(B, N, K, H, Tf, Cm, cfg.vel_dim, cfg.scene_channels, cfg.n_rad, cfg.n_ang,
                                     Hm, Hm, cfg.ioc_iters)
        ws = max(
            lib.desire_cvae_encode_workspace_bytes(M, Zl), lib.desire_cvae_decode_workspace_bytes(R, Zl),
            lib.desire_mask_softmax_workspace_bytes(R, H), lib.desire_gru_decode_workspace_bytes(R, H),
            lib.desire_scene_cnn_workspace_bytes(B, cfg.scene_size, cfg.scene_size),
            lib.desire_ioc_workspace_bytes(C.byref(self.ioc_dims)), 256)
        self.ws_bytes = ws
        self.ws = torch.empty(ws, dtype=torch.uint8, device=self.device)
        # the scene CNN only depends on the images: it gets its own scratch and runs on a side stream, concurrently
        # with the (small-M, SM-underfilling) encoder / CVAE-encoder kernels of the generation stage
        self.ws_scene_bytes = max(lib.desire_scene_cnn_workspace_bytes(B, cfg.scene_size, cfg.scene_size), 256)
        self.ws_scene = torch.empty(self.ws_scene_bytes, dtype=torch.uint8, device=self.device)
        self.side = torch.cuda.Stream(self.device)
        # IOC ranking/refinement is a chain of ~50 dependent launches per iteration whose tile counts do not divide the
        # SM count (320 / 300 tiles on 148 SMs: every launch ends in a mostly idle round).  Scenes are independent,
        # so the batch is cut into `ioc_chains` groups of scenes, each an independent chain on its own stream / graph
        # branch with its own scratch: the idle tail of one chain's launch is filled by the other chains' CTAs.
        import os
        self.ioc_chains = 1
        want = int(os.environ.get("DESIRE_IOC_CHAINS", "0"))        # 0 = automatic
        if want > 1:
            if B % want == 0 and B >= 2 * want:
                self.ioc_chains = want
        elif want == 0 and B >= 8 and N <= 128:
            # automatic: the fewest chains whose launches fit one round of the machine (measured at B=32: 1 chain
            # 19.3 ms/step, 2 chains 22.5 — two 160-tile launches fight over 148 SMs —, 4 chains 18.7, 8 chains 19.6)
            npad = 8
            while npad < N:
                npad *= 2
            n_sm = torch.cuda.get_device_properties(self.device).multi_processor_count
            for c in range(2, B // 2 + 1):
                if B % c == 0 and -(-(B // c) * K // (128 // npad)) <= n_sm:
                    self.ioc_chains = c
                    break
        if self.ioc_chains > 1:
            Bc = B // self.ioc_chains
            self.ioc_dims_c = _lib.IocDims(Bc, N, K, H, Tf, Cm, cfg.vel_dim, cfg.scene_channels, cfg.n_rad, cfg.n_ang,
                                           Hm, Hm, cfg.ioc_iters)
            self.ws_ioc_bytes = max(lib.desire_ioc_workspace_bytes(C.byref(self.ioc_dims_c)), 256)
            self.ws_ioc = [torch.empty(self.ws_ioc_bytes, dtype=torch.uint8, device=self.device)
                           for _ in range(self.ioc_chains)]
            self.ioc_streams = [torch.cuda.Stream(self.device) for _ in range(self.ioc_chains - 1)]
            self.scores_c = [torch.empty(max(cfg.ioc_iters, 1), Bc * N * K, dtype=torch.float32, device=self.device)
                             for _ in range(self.ioc_chains)]
        # the two encoders run side by side on the tensor-core recurrence, each with its own scratch
        self.side2 = torch.cuda.Stream(self.device)
        self.ws_enc_bytes = max(lib.desire_gru_encode_workspace_bytes(M, max(Tp, Tf), H), 256)
        self.ws_encx = torch.empty(self.ws_enc_bytes, dtype=torch.uint8, device=self.device)
        self.ws_ency = torch.empty(self.ws_enc_bytes, dtype=torch.uint8, device=self.device)
        # {seed, offset} of the device noise source (desire_randn_fwd), read by the kernel at execution time: a captured
        # graph draws fresh eps when set_noise() bumps the offset between replays
        self.rng_state = torch.tensor([2, 0], dtype=torch.int64, device=self.device)
        self._rng_host = torch.empty(2, dtype=torch.int64).pin_memory()
        self.serial = False      # True: no parallel branches (bench.py's per-kernel timing pass needs launches that do not overlap)
        self.graph = None
        self.graph_gen = self.graph_rank = None
        self.static_in = None

    def _structs(self):
        P = self.P
        gru = lambda n: _lib.GruW(*[P[n + s].data_ptr() for s in ("_wg", "_bg", "_wc", "_bc")])
        cbn = lambda n: _lib.ConvBnW(*[P[n + s].data_ptr() for s in ("_w", "_b", "_g", "_be")])
        self.w_encx, self.w_ency, self.w_dec1 = gru("encx"), gru("ency"), gru("dec1")
        self.w_venc = _lib.CvaeEncW(cbn("venc_c1"), cbn("venc_c2"), cbn("venc_c3"),
                                    P["venc_fc_w"].data_ptr(), P["venc_fc_b"].data_ptr())
        self.w_vdec = _lib.CvaeDecW(cbn("vdec_d1"), cbn("vdec_d2"), cbn("vdec_d3"), cbn("vdec_d4"))
        self.w_scene = _lib.SceneCnnW(*[P["scene_" + n].data_ptr() for n in ("c1_w", "c1_b", "c2_w", "c2_b", "c3_w", "c3_b")])
        self.w_ioc = _lib.IocW(P["ioc_vel_w"].data_ptr(), P["ioc_vel_b"].data_ptr(), P["ioc_sp_w"].data_ptr(),
                               P["ioc_sp_b"].data_ptr(), gru("dec2"), P["ioc_score_w"].data_ptr(),
                               P["ioc_score_b"].data_ptr(), P["ioc_reg_w"].data_ptr(), P["ioc_reg_b"].data_ptr(),
                               self.r2_edges.data_ptr(), self.dirs.data_ptr())

    def set_noise(self, seed, offset):
        """Select the eps draw of the next pass that runs with eps=None (async 16-byte copy on the current stream)."""
        self._rng_host[0], self._rng_host[1] = int(seed), int(offset)
        self.rng_state.copy_(self._rng_host, non_blocking=True)

    # ------------------------------------------------------------------ one pass of the hot path
    def run(self, obs, tgt, eps, scene, stages=("generate", "rank")):
        """obs [B,N,Tp,3], tgt [B,N,Tf,3], eps [M,K,Z] or None, scene [B,Hi,Wi,3] — contiguous fp32 CUDA tensors.
        eps=None draws the noise on the device (desire_randn_fwd with the state of set_noise(); the reference draws it
        inside the graph too, model/model.py:262) into buf["eps"].
        Enqueues everything on the current stream; returns the dict of output buffers (views)."""
        cfg, lib, b, P = self.cfg, self.lib, self.buf, self.P
        N, K, H, Zl, Tp, Tf, Cm = cfg.max_num_obj, cfg.K, cfg.H, cfg.Z, cfg.seq_length, cfg.pred_length, cfg.channel_multiplier
        M, R, S2 = self.M, self.R, cfg.S * cfg.S
        draw = eps is None
        if draw:
            eps = b["eps"]
        for t, shp in ((obs, (self.B, N, Tp, 3)), (tgt, (self.B, N, Tf, 3)), (eps, (M, K, Zl))):
            if tuple(t.shape) != shp or t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
                raise ValueError("expected contiguous float32 %s on %s, got %s %s" % (shp, self.device, tuple(t.shape), t.dtype))
        cur = torch.cuda.current_stream(self.device)
        st = C.c_void_p(cur.cuda_stream)
        ws, wsb = _p(self.ws), self.ws_bytes
        ck = _lib.check
        if "rank" in stages:
            stages = tuple(x for x in stages if x != "rank") + ("scene", "ioc")
        if "generate" in stages:
            stages = stages + ("encode", "decode")      # the two halves of the generation stage (eps enters the second)
        if "scene" in stages:
            if tuple(scene.shape) != (self.B, cfg.scene_size, cfg.scene_size, 3) or not scene.is_contiguous():
                raise ValueError("scene must be contiguous [B,%d,%d,3]" % (cfg.scene_size, cfg.scene_size))
            fork = "generate" in stages and not self.serial   # overlap with the generation stage when both run in this call
            s_str = self.side if fork else cur
            if fork:
                self.side.wait_stream(cur)
            ck(lib.desire_scene_cnn_fwd(_p(scene), self.B, cfg.scene_size, cfg.scene_size, cfg.scene_channels,
                                        C.byref(self.w_scene), _p(b["scene_features"]), _p(self.ws_scene),
                                        self.ws_scene_bytes, C.c_void_p(s_str.cuda_stream)), "scene_cnn")
        if "encode" in stages:
            # D8: from here on every kernel sees obs_eff, whose frame-0 id is zero for agents that do not exist where
            # the losses / decoders need them (absent at the last observed frame or in the target, cfg.exist_mode)
            ck(lib.desire_existence_fwd(_p(obs), _p(tgt), M, Tp, Tf, cfg.exist_mode, _p(b["obs_eff"]), st), "existence")
        obs = b["obs_eff"]          # stages run in order on one HotPath: later stages reuse the encode stage's copy
        if "encode" in stages:
            ck(lib.desire_tconv_fwd(_p(obs), M, Tp, Cm, _p(P["temporal_w"]), _p(P["temporal_b"]), _p(b["rho_i"]), st), "tconv")
            s_y = cur if self.serial else self.side2
            if not self.serial:
                self.side2.wait_stream(cur)
            ck(lib.desire_gru_encode_ws_fwd(_p(tgt), M, Tf, H, C.byref(self.w_ency), C.c_void_p(b["HxHy"].data_ptr() + 4 * H),
                                            2 * H, _p(self.ws_ency), self.ws_enc_bytes,
                                            C.c_void_p(s_y.cuda_stream)), "gru_encode_y")
            ck(lib.desire_gru_encode_ws_fwd(_p(obs), M, Tp, H, C.byref(self.w_encx), _p(b["HxHy"]), 2 * H,
                                            _p(self.ws_encx), self.ws_enc_bytes, st), "gru_encode_x")
            if not self.serial:
                cur.wait_stream(self.side2)
            ck(lib.desire_fc_fwd(_p(b["HxHy"]), 2 * H, _p(P["w_hidden_enc1"]), S2, _p(P["b_hidden_enc1"]),
                                 _p(b["vae_inputs"]), S2, M, S2, 2 * H, 1, 0, st), "fc_c")
            ck(lib.desire_cvae_encode_fwd(_p(b["vae_inputs"]), M, Zl, C.byref(self.w_venc), _p(b["mu_logvar"]), ws, wsb, st), "cvae_encode")
        if "decode" in stages:
            if draw:
                ck(lib.desire_randn_fwd(_p(self.rng_state), _p(eps), M * K * Zl, st), "randn")
            ck(lib.desire_reparam_fwd(_p(b["mu_logvar"]), _p(eps), M, K, Zl, _p(b["zval"]), st), "reparam")
            ck(lib.desire_cvae_decode_fwd(_p(b["zval"]), R, Zl, C.byref(self.w_vdec), _p(b["x_reconstr_mean"]), ws, wsb, st), "cvae_decode")
            ck(lib.desire_mask_softmax_fwd(_p(b["x_reconstr_mean"]), R, S2, H, K, _p(P["w_post_vae"]), _p(P["b_post_vae"]),
                                           _p(b["HxHy"]), 2 * H, _p(b["x_z"]), ws, wsb, st), "mask_softmax")
            ck(lib.desire_gru_decode_fwd(_p(b["x_z"]), _p(b["HxHy"]), 2 * H, R, K, H, Tf, C.byref(self.w_dec1),
                                         _p(b["output_states"]), ws, wsb, st), "gru_decode")
            ck(lib.desire_readout_pool_fwd(_p(b["output_states"]), R, K, Tf, H, 0, 1, _p(P["output_w"]), _p(P["output_b"]),
                                           _p(obs), Tp, _p(b["rho_i"]), Cm, _p(b["Yhat"]), _p(b["feature_pooling"]), st), "readout_pool")
            ck(lib.desire_kld_rows_fwd(_p(b["mu_logvar"]), M, Zl, _p(b["kld_rows"]), st), "kld_rows")
            ck(lib.desire_recon_rows_fwd(_p(b["Yhat"]), _p(tgt), M, K, Tf, _p(b["recon_rows"]), st), "recon_rows")
            ck(lib.desire_masked_cost_fwd(_p(b["recon_rows"]), _p(b["kld_rows"]), _p(obs), M, Tp, _p(b["cost"]), st), "masked_cost")
        if "ioc" in stages:
            if "scene" in stages and "generate" in stages and not self.serial:
                cur.wait_stream(self.side)         # join the scene-CNN branch
            b["Y_refined"].copy_(b["Yhat"])
            if self.ioc_chains == 1 or self.serial:
                ck(lib.desire_ioc_factored_fwd(C.byref(self.ioc_dims), C.byref(self.w_ioc), _p(b["scene_features"]), _p(obs),
                                               Tp, _p(b["HxHy"]), 2 * H, _p(b["feature_pooling"]), _p(b["rho_i"]),
                                               _p(b["Yhat"]), _p(b["Y_refined"]), _p(b["ioc_scores"]), ws, wsb, st), "ioc")
            else:
                nc, Bc = self.ioc_chains, self.B // self.ioc_chains
                Mc, Rc = Bc * N, Bc * N * K
                off = lambda t, rows_per_scene_block, c: C.c_void_p(t.data_ptr() + 4 * c * rows_per_scene_block)
                # fork BEFORE anything of chain 0 is enqueued: every side stream waits for the same point of `cur`
                # (wait_stream after chain 0's launches would make chains 1.. start only when chain 0 has finished)
                fork = torch.cuda.Event()
                fork.record(cur)
                for s_side in self.ioc_streams:
                    s_side.wait_event(fork)
                for c in range(nc):
                    s_c = cur if c == 0 else self.ioc_streams[c - 1]
                    ck(lib.desire_ioc_factored_fwd(C.byref(self.ioc_dims_c), C.byref(self.w_ioc),
                                                   off(b["scene_features"], Bc * self.Hm * self.Hm * cfg.scene_channels, c),
                                                   off(obs, Mc * Tp * 3, c), Tp, off(b["HxHy"], Mc * 2 * H, c), 2 * H,
                                                   off(b["feature_pooling"], Rc * Tf * 2 * Cm, c),
                                                   off(b["rho_i"], Mc * 2 * Cm, c), off(b["Yhat"], Rc * Tf * 2, c),
                                                   off(b["Y_refined"], Rc * Tf * 2, c),
                                                   _p(self.scores_c[c]), _p(self.ws_ioc[c]), self.ws_ioc_bytes,
                                                   C.c_void_p(s_c.cuda_stream)), "ioc chain %d" % c)
                    with torch.cuda.stream(s_c):
                        b["ioc_scores"][:, c * Rc:(c + 1) * Rc].copy_(self.scores_c[c])
                for c in range(1, nc):
                    cur.wait_stream(self.ioc_streams[c - 1])
        return self.outputs()

    # ------------------------------------------------------------------ CUDA graph: capture once, replay per step
    def capture(self, obs, tgt, eps, scene):
        """Capture one whole pass (every kernel of run()) into a CUDA graph over static input buffers.
        ~10^3 launches per step otherwise leave the GPU waiting on the host between small kernels."""
        self.static_in = [t.clone() if t is not None else None for t in (obs, tgt, eps, scene)]
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):                      # warm-up outside capture (lazy attribute setting, cuBLAS-free)
                self.run(*self.static_in)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run(*self.static_in)
        self.graph = g
        return g

    def capture_split(self, obs, tgt, eps, scene):
        """Two graphs over the same static buffers: sample generation, then ranking/refinement.  The host-buffer
        entry point (replay_split) uses the gap to stage the scene images — the largest input, which only the
        second graph reads — while the first graph is already running."""
        self.static_in = [t.clone() if t is not None else None for t in (obs, tgt, eps, scene)]
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self.run(*self.static_in)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph_enc, self.graph_gen, self.graph_scene, self.graph_rank = (torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(),
                                                                              torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph())
        with torch.cuda.graph(self.graph_enc):
            self.run(*self.static_in, stages=("encode",))
        with torch.cuda.graph(self.graph_gen):
            self.run(*self.static_in, stages=("decode",))
        with torch.cuda.graph(self.graph_scene):
            self.run(*self.static_in, stages=("scene",))
        with torch.cuda.graph(self.graph_rank):
            self.run(*self.static_in, stages=("ioc",))
        self.copy_stream = torch.cuda.Stream(self.device)
        self.copy_done = torch.cuda.Event()
        self.eps_done = torch.cuda.Event()
        self.gen_done = torch.cuda.Event()
        self.rank_done = torch.cuda.Event()
        self.gen_done.record(torch.cuda.current_stream(self.device))
        self.rank_done.record(torch.cuda.current_stream(self.device))
        self.split_draws = eps is None               # the decoder graph holds the device noise kernel
        return self.graph_gen, self.graph_rank

    def replay_split(self, obs, tgt, stage_eps, stage_scene):
        """Host-buffer replay.  obs/tgt: pinned host tensors (small); stage_eps() / stage_scene() return the pinned eps
        and scene tensors and are called only AFTER GPU work that does not need them has been launched, so their
        host-side memcpy, their H2D copies (second stream) and the scene CNN overlap the encoder graph and the
        decoder graph respectively (stage_eps=None: the graphs were captured with eps=None and draw it on the device):
            main:  H2D obs,tgt | encode graph ............ | wait eps | decode graph ............... | wait scene | IOC graph
            host:              | memcpy eps -> pinned      |          | memcpy scene -> pinned       |
            copy:                          | H2D eps       |                       | H2D scene, scene CNN |
        Calls may be issued back to back without a host synchronisation in between (DESIREModel.submit): the events
        below keep pass t+1's copies and scene CNN off the buffers pass t's graphs are still reading."""
        cur = torch.cuda.current_stream(self.device)
        self.static_in[0].copy_(obs, non_blocking=True)
        self.static_in[1].copy_(tgt, non_blocking=True)
        self.graph_enc.replay()
        if stage_eps is not None:
            eps = stage_eps()
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.gen_done)          # the previous pass's decoder graph read static eps
                self.static_in[2].copy_(eps, non_blocking=True)
                self.eps_done.record(self.copy_stream)
            cur.wait_event(self.eps_done)
        self.graph_gen.replay()
        self.gen_done.record(cur)
        scene = stage_scene()
        with torch.cuda.stream(self.copy_stream):
            self.static_in[3].copy_(scene, non_blocking=True)
            self.copy_stream.wait_event(self.rank_done)             # the previous pass's IOC graph read scene_features
            self.graph_scene.replay()              # the scene CNN runs next to the rest of the decoder graph
            self.copy_done.record(self.copy_stream)
        cur.wait_event(self.copy_done)
        self.graph_rank.replay()
        self.rank_done.record(cur)
        return self.outputs()

    def replay(self, obs=None, tgt=None, eps=None, scene=None, non_blocking=True):
        """Copy new inputs (device or pinned-host tensors) into the static buffers and replay the graph."""
        if self.graph is None:
            raise _lib.DesireError("replay() before capture()")
        for dst, src in zip(self.static_in, (obs, tgt, eps, scene)):
            if src is not None and dst is not None:
                dst.copy_(src, non_blocking=non_blocking)
        self.graph.replay()
        return self.outputs()

    def outputs(self):
        b, cfg = self.buf, self.cfg
        H, Zl = cfg.H, cfg.Z
        out = dict(b)
        out["H_x"], out["H_y"] = b["HxHy"][:, :H], b["HxHy"][:, H:]
        out["z_mean"], out["z_log_sigma_sq"] = b["mu_logvar"][:, :Zl], b["mu_logvar"][:, Zl:]
        out["cost"] = b["cost"][0]
        out["ioc_scores"] = b["ioc_scores"][:cfg.ioc_iters]
        return out


class TrainPath(HotPath):
    """Forward (sample generation) + backward of `cost` + clip_by_global_norm + Adam for a fixed batch of B
    scenes (model/model.py:388-394).  `flat` is the flat parameter buffer `params` are views of
    (flatten_params); gradients and the Adam moments live in flat buffers of the same layout, so the
    multi-GPU step is ONE all-reduce of `grad_flat`."""

    def __init__(self, cfg: DesireConfig, flat: torch.Tensor, params: dict, offsets: dict, B: int, device="cuda:0",
                 train_ioc=True, opt_state=None):
        """opt_state: (adam_m, adam_v, step) with step a one-element list — the optimiser state belongs to the
        parameters, not to a batch size: DESIREModel passes the same triple to every TrainPath it builds."""
        super().__init__(cfg, params, B, device)
        self.train_ioc = bool(train_ioc) and cfg.ioc_iters > 0
        for k, (o, cnt, shp) in offsets.items():
            if self.P[k].data_ptr() != flat.data_ptr() + 4 * o:
                raise ValueError("parameter %s is not a view of the flat buffer" % k)
        self.flat, self.offsets = flat, offsets
        self.grad_flat = torch.zeros_like(flat)
        self.G = {k: self.grad_flat[o:o + cnt].view(shp) for k, (o, cnt, shp) in offsets.items()}
        if opt_state is None:
            opt_state = (torch.zeros_like(flat), torch.zeros_like(flat), [0])
        self.adam_m, self.adam_v, self._step = opt_state
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.count = torch.ones(1, dtype=torch.float32, device=self.device)
        N, K, H, Zl, Tf = cfg.max_num_obj, cfg.K, cfg.H, cfg.Z, cfg.pred_length
        M, R, S2 = self.M, self.R, cfg.S * cfg.S
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)
        self.dbuf = dict(dYhat=f(R, Tf, 2), dhs=f(R, Tf, H), dx_z=f(R, H), dxr=f(R, S2), dz=f(R, Zl),
                         d_mu_logvar=f(M, 2 * Zl), dv=f(M, S2), dHxHy=f(M, 2 * H))
        G = self.G
        gru = lambda n: _lib.GruG(*[G[n + s].data_ptr() for s in ("_wg", "_bg", "_wc", "_bc")])
        cbn = lambda n: _lib.ConvBnG(*[G[n + s].data_ptr() for s in ("_w", "_b", "_g", "_be")])
        self.g_encx, self.g_ency, self.g_dec1 = gru("encx"), gru("ency"), gru("dec1")
        self.g_venc = _lib.CvaeEncG(cbn("venc_c1"), cbn("venc_c2"), cbn("venc_c3"),
                                    G["venc_fc_w"].data_ptr(), G["venc_fc_b"].data_ptr())
        self.g_vdec = _lib.CvaeDecG(cbn("vdec_d1"), cbn("vdec_d2"), cbn("vdec_d3"), cbn("vdec_d4"))
        self.g_scene = _lib.SceneCnnG(*[G["scene_" + n].data_ptr() for n in ("c1_w", "c1_b", "c2_w", "c2_b", "c3_w", "c3_b")])
        self.g_ioc = _lib.IocG(G["ioc_vel_w"].data_ptr(), G["ioc_vel_b"].data_ptr(), G["ioc_sp_w"].data_ptr(),
                               G["ioc_sp_b"].data_ptr(), gru("dec2"), G["ioc_score_w"].data_ptr(),
                               G["ioc_score_b"].data_ptr(), G["ioc_reg_w"].data_ptr(), G["ioc_reg_b"].data_ptr())
        self.dbuf["dfmap"] = torch.empty_like(self.buf["scene_features"])
        self.buf["ioc_cost"] = f(2)
        lib = self.lib
        bws = max(lib.desire_gru_decode_bwd_workspace_bytes(R, H, Tf), lib.desire_mask_softmax_bwd_workspace_bytes(R, H),
                  lib.desire_cvae_decode_bwd_workspace_bytes(R, Zl), lib.desire_cvae_encode_bwd_workspace_bytes(M, Zl),
                  lib.desire_gru_encode_bwd_workspace_bytes(M, max(cfg.seq_length, Tf), H))
        if self.train_ioc:
            bws = max(bws, lib.desire_ioc_train_workspace_bytes(C.byref(self.ioc_dims)),
                      lib.desire_scene_cnn_bwd_workspace_bytes(B, cfg.scene_size, cfg.scene_size))
        if bws > self.ws_bytes:
            self.ws_bytes = bws
            self.ws = torch.empty(bws, dtype=torch.uint8, device=self.device)
        self.train_graph = None

    @property
    def step_no(self):
        return self._step[0]

    @step_no.setter
    def step_no(self, v):
        self._step[0] = int(v)

    def set_count(self, obs, tgt=None):
        """Number of existing agents (D8, cfg.exist_mode) of this rank's scenes, summed over ranks: the normaliser of
        `cost` (model/model.py:376) every rank must share."""
        from .dist import global_count_
        self.count.copy_(existing_agents(obs, tgt, self.cfg.exist_mode).sum().to(torch.float32).reshape(1))
        return global_count_(self.count)

    def backward(self, obs, tgt, eps, scene=None):
        """Gradients of `cost` (+ `ioc_cost` when train_ioc, D13) w.r.t. every parameter into grad_flat (zeroed
        here).  Call after run(stages=("generate",)); uses self.count as the normaliser.  With train_ioc the IOC
        forward runs inside (Y_refined, ioc_scores, ioc_cost are outputs of this call)."""
        cfg, lib, b, d, P, G = self.cfg, self.lib, self.buf, self.dbuf, self.P, self.G
        N, K, H, Zl, Tp, Tf = cfg.max_num_obj, cfg.K, cfg.H, cfg.Z, cfg.seq_length, cfg.pred_length
        obs = b["obs_eff"]          # the existence-masked copy written by the forward's encode stage (same inputs)
        if eps is None:
            eps = b["eps"]          # the noise the forward drew on the device
        M, R, S2 = self.M, self.R, cfg.S * cfg.S
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        ws, wsb = _p(self.ws), self.ws_bytes
        ck = _lib.check
        self.grad_flat.zero_()
        d["dHxHy"].zero_()
        ck(lib.desire_cost_bwd(_p(b["Yhat"]), _p(tgt), _p(b["mu_logvar"]), _p(obs), _p(self.count), M, K, Tf, Tp, Zl,
                               _p(d["dYhat"]), _p(d["d_mu_logvar"]), st), "cost_bwd")
        ck(lib.desire_readout_bwd(_p(b["output_states"]), _p(d["dYhat"]), R, Tf, H, _p(P["output_w"]), _p(d["dhs"]),
                                  _p(G["output_w"]), _p(G["output_b"]), st), "readout_bwd")
        ck(lib.desire_gru_decode_bwd(_p(b["x_z"]), _p(b["HxHy"]), 2 * H, R, K, H, Tf, C.byref(self.w_dec1),
                                     _p(b["output_states"]), _p(d["dhs"]), _p(d["dx_z"]), _p(d["dHxHy"]), 2 * H,
                                     C.byref(self.g_dec1), ws, wsb, st), "gru_decode_bwd")
        ck(lib.desire_mask_softmax_bwd(_p(b["x_reconstr_mean"]), R, S2, H, K, _p(P["w_post_vae"]), _p(P["b_post_vae"]),
                                       _p(b["HxHy"]), 2 * H, _p(d["dx_z"]), _p(d["dxr"]), _p(d["dHxHy"]), 2 * H,
                                       _p(G["w_post_vae"]), _p(G["b_post_vae"]), ws, wsb, st), "mask_softmax_bwd")
        ck(lib.desire_cvae_decode_bwd(_p(b["zval"]), R, Zl, C.byref(self.w_vdec), _p(d["dxr"]), _p(d["dz"]),
                                      C.byref(self.g_vdec), ws, wsb, st), "cvae_decode_bwd")
        ck(lib.desire_reparam_bwd(_p(b["mu_logvar"]), _p(eps), _p(d["dz"]), M, K, Zl, _p(d["d_mu_logvar"]), st), "reparam_bwd")
        ck(lib.desire_cvae_encode_bwd(_p(b["vae_inputs"]), M, Zl, C.byref(self.w_venc), _p(d["d_mu_logvar"]), _p(d["dv"]),
                                      C.byref(self.g_venc), ws, wsb, st), "cvae_encode_bwd")
        ck(lib.desire_fc_bwd(_p(b["HxHy"]), 2 * H, _p(P["w_hidden_enc1"]), S2, _p(b["vae_inputs"]), S2, _p(d["dv"]), S2,
                             M, S2, 2 * H, 1, _p(d["dHxHy"]), 2 * H, 1, _p(G["w_hidden_enc1"]), S2,
                             _p(G["b_hidden_enc1"]), st), "fc_c_bwd")
        ck(lib.desire_gru_encode_bwd(_p(obs), M, Tp, H, C.byref(self.w_encx), _p(d["dHxHy"]), 2 * H,
                                     C.byref(self.g_encx), ws, wsb, st), "gru_encode_x_bwd")
        ck(lib.desire_gru_encode_bwd(_p(tgt), M, Tf, H, C.byref(self.w_ency),
                                     C.c_void_p(d["dHxHy"].data_ptr() + 4 * H), 2 * H,
                                     C.byref(self.g_ency), ws, wsb, st), "gru_encode_y_bwd")
        if self.train_ioc:
            if scene is None:
                raise ValueError("train_ioc needs the scene images")
            ck(lib.desire_scene_cnn_fwd(_p(scene), self.B, cfg.scene_size, cfg.scene_size, cfg.scene_channels,
                                        C.byref(self.w_scene), _p(b["scene_features"]), ws, wsb, st), "scene_cnn")
            d["dfmap"].zero_()
            ck(lib.desire_ioc_train(C.byref(self.ioc_dims), C.byref(self.w_ioc), _p(b["scene_features"]), _p(obs), Tp,
                                    _p(tgt), _p(b["HxHy"]), 2 * H, _p(b["feature_pooling"]), _p(b["Yhat"]),
                                    _p(self.count), _p(b["Y_refined"]), _p(b["ioc_scores"]), _p(b["ioc_cost"]),
                                    C.byref(self.g_ioc), _p(d["dfmap"]), ws, wsb, st), "ioc_train")
            ck(lib.desire_scene_cnn_bwd(_p(scene), self.B, cfg.scene_size, cfg.scene_size, cfg.scene_channels,
                                        C.byref(self.w_scene), _p(d["dfmap"]), C.byref(self.g_scene), ws, wsb, st),
               "scene_cnn_bwd")
        return self.G

    def apply(self, lr, clip=10.0, beta1=0.9, beta2=0.999, eps=1e-8):
        """All-reduce (sum) of the flat gradient over ranks, clip_by_global_norm, Adam."""
        from .dist import all_reduce_gradients_
        lib = self.lib
        all_reduce_gradients_(self.grad_flat)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        n = self.flat.numel()
        self.step_no += 1
        _lib.check(lib.desire_sumsq_fwd(_p(self.grad_flat), n, _p(self.sumsq), 0, st), "sumsq")
        _lib.check(lib.desire_adam_step(_p(self.flat), _p(self.grad_flat), _p(self.adam_m), _p(self.adam_v), n,
                                        _p(self.sumsq), lr, beta1, beta2, eps, self.step_no, clip, 1.0, st), "adam")

    def capture_train(self, obs, tgt, eps, scene):
        """Capture forward (generate stage) + backward into one CUDA graph over static inputs."""
        self.static_in = [t.clone() if t is not None else None for t in (obs, tgt, eps, scene)]
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self.run(*self.static_in, stages=("generate",))
                self.backward(*self.static_in)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run(*self.static_in, stages=("generate",))
            self.backward(*self.static_in)
        self.train_graph = g
        return g

    def train_step(self, obs, tgt, eps, scene, lr, clip=10.0, use_graph=True):
        """One optimiser step on this rank's scenes.  Returns the (local) cost tensor [cost, count]."""
        if use_graph:
            if self.train_graph is None:
                self.set_count(obs, tgt)
                self.capture_train(obs, tgt, eps, scene)
            for dst, src in zip(self.static_in, (obs, tgt, eps, scene)):
                if src is not None and dst is not None and src.data_ptr() != dst.data_ptr():
                    dst.copy_(src, non_blocking=True)
            self.set_count(self.static_in[0], self.static_in[1])
            self.train_graph.replay()
        else:
            self.set_count(obs, tgt)
            self.run(obs, tgt, eps, scene, stages=("generate",))
            self.backward(obs, tgt, eps, scene)
        self.apply(lr, clip)
        return self.buf["cost"]
