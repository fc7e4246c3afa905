"""DESIREModel — host-side mirror of the reference's model/model.py:29-688 on the B200 kernel library.

Same constructor (`DESIREModel(args)` with the train.py flags), same attribute names
(`input_data`, `target_data`, `temporal_data`, `learning_rate`, `cost`, `final_states`,
`final_output`, `rho_i`, `output_states`, `feature_pooling`) and the same `[agents, time, (id,x,y)]`
layouts (model/model.py:86-105).  Where the reference builds a TF graph and evaluates it one
sequence per `sess.run` (train.py:146-181), this class owns device buffers for a whole minibatch
of scenes and runs the fused path through the C-ABI (engine.HotPath): attributes are filled by
`forward()` instead of being graph nodes.  There is no CPU path: constructing the model without a
CUDA device or without libdesire_b200.so raises.
"""
from __future__ import annotations

import argparse
from collections import OrderedDict

import numpy as np
import torch

from ..config import DesireConfig, init_params
from ..engine import HotPath, TrainPath, flatten_params


def config_from_args(args) -> DesireConfig:
    """Map the reference's argparse namespace (train.py:28-88) + the added flags onto DesireConfig."""
    g = lambda n, d: getattr(args, n, d)
    return DesireConfig(
        rnn_size=g("rnn_size", 512), d_dim=g("d_dim", 16), latent_size=g("latent_size", 128),
        seq_length=g("seq_length", 8), max_num_obj=g("max_num_obj", 60), stride=g("stride", 1),
        num_layers=g("num_layers", 1), model=g("model", "gru"),
        pred_length=g("pred_length", 12), num_samples=g("num_samples", 20), ioc_iters=g("ioc_iters", 2),
        scene_size=g("scene_size", 256), scene_channels=g("scene_channels", 32), vel_dim=g("vel_dim", 16),
        n_rad=g("n_rad", 6), n_ang=g("n_ang", 6), r_min=g("r_min", 0.01), r_max=g("r_max", 0.5))


class DESIREModel(object):
    """Stochastic IOC RNN encoder-decoder (sample generation + ranking/refinement) on one GPU."""

    def __init__(self, args, device="cuda:0", seed=1, params=None, use_graph=True):
        self.args = args
        self.use_graph = use_graph
        cfg = args if isinstance(args, DesireConfig) else config_from_args(args)
        cfg.validate()
        self.cfg = cfg
        self.device = torch.device(device)
        # names kept from model/model.py:43-60
        self.filter_height, self.filter_width = 1, cfg.seq_length
        self.in_channels, self.channel_multiplier = 2, cfg.channel_multiplier
        self.input_size = 3
        self.decoder_output = cfg.d_dim
        self.rnn_size = cfg.rnn_size
        self.seq_length = cfg.seq_length
        self.batch_size = getattr(args, "batch_size", 1)
        self.latent_size = cfg.latent_size
        self.input_shape = [cfg.S, cfg.S]
        self.vae_input_size = cfg.S * cfg.S
        self.max_num_obj = cfg.max_num_obj
        self.learning_rate = float(getattr(args, "learning_rate", 0.005))
        # every trainable tensor is a view into ONE flat device buffer (the train step clips / all-reduces /
        # updates that buffer in one go); `weights` keeps the reference's name -> tensor dict
        init = params if params is not None else self.define_weights(seed)
        self.flat_weights, self.weights, self._offsets = flatten_params(init, self.device)
        self.grad_clip = float(getattr(args, "grad_clip", 10.0))
        # Adam moments + step live next to the flat weight buffer and are shared by the train paths of every batch size
        self.adam_m, self.adam_v = torch.zeros_like(self.flat_weights), torch.zeros_like(self.flat_weights)
        self.adam_step = [0]
        self._paths = {}
        self._train_paths = {}
        self._pinned = {}
        self._submit_no, self._noise_offset, self._slot_done = 0, 0, {}
        # filled by forward(); same names as the reference's graph nodes
        self.input_data = self.target_data = self.target_data_enc = self.temporal_data = None
        self.rho_i = self.output_states = self.feature_pooling = None
        self.cost = self.final_states = self.final_output = self.gradients = None
        self.build_model()

    # ------------------------------------------------------------------ reference API
    def define_weights(self, seed=1):
        """model/model.py:420-451 (+ library-default initialisers, D10)."""
        return init_params(self.cfg, seed, self.device)

    def build_model(self):
        """model/model.py:79-403 built a graph for ONE sequence; here it sizes the buffers for
        `batch_size` scenes and loads the kernel library (raises if it is missing)."""
        self._path(max(int(self.batch_size), 1))

    @staticmethod
    def _check_activ_phase(who, activ, phase):
        """The reference calls these layers with activ=tf.nn.elu and phase=pt.Phase.train only
        (model/model.py:258,266,457-462,476-481): per-row train-phase batch statistics are what the kernels implement.
        Anything else would silently compute something different, so it raises."""
        if activ is not None and getattr(activ, "__name__", str(activ)).lower() not in ("elu",):
            raise ValueError("%s: only the ELU activation of the reference (model.py:258,266) is built, got %r" % (who, activ))
        if phase is not None and str(getattr(phase, "name", phase)).lower() not in ("train", "phase.train"):
            raise ValueError("%s: only phase=train (batch statistics, model.py:457-462,476-481) is built; the "
                             "inference phase with learned moments (learned_moments_update_rate=0.0003) is not, "
                             "got %r" % (who, phase))

    def vae_encoder(self, inputs, latent_size=None, activ=None, phase=None):
        """model/model.py:471-492: inputs [M, 1024] (the fc_c features viewed as 32x32x1) -> (mean [M,Z], logvar [M,Z]).
        `activ` / `phase`: None or the reference's values (ELU, train phase); anything else raises."""
        import ctypes as C
        from .. import _lib
        self._check_activ_phase("vae_encoder", activ, phase)
        x = inputs.to(self.device, torch.float32).contiguous().reshape(-1, self.vae_input_size)
        Zl = int(latent_size or self.cfg.Z)
        if Zl != self.cfg.Z:
            raise ValueError("vae_encoder: latent_size %d does not match the model's %d" % (Zl, self.cfg.Z))
        M = x.shape[0]
        hp = self._path(max(int(self.batch_size), 1))
        lib = hp.lib
        out = torch.empty(M, 2 * Zl, dtype=torch.float32, device=self.device)
        wsb = lib.desire_cvae_encode_workspace_bytes(M, Zl)
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=self.device)
        _lib.check(lib.desire_cvae_encode_fwd(C.c_void_p(x.data_ptr()), M, Zl, C.byref(hp.w_venc), C.c_void_p(out.data_ptr()),
                                              C.c_void_p(ws.data_ptr()), wsb,
                                              C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "vae_encoder")
        return out[:, :Zl], out[:, Zl:]

    def vae_decoder(self, zval, projection_size=None, activ=None, phase=None):
        """model/model.py:453-469: zval [R, Z] -> x_reconstr_mean [R, 1024] (32x32x1 flattened, sigmoid)."""
        import ctypes as C
        from .. import _lib
        self._check_activ_phase("vae_decoder", activ, phase)
        z = zval.to(self.device, torch.float32).contiguous().reshape(-1, self.cfg.Z)
        if projection_size is not None and int(projection_size) != self.vae_input_size:
            raise ValueError("vae_decoder: the decoder always emits %d values (model.py:465-468)" % self.vae_input_size)
        R = z.shape[0]
        hp = self._path(max(int(self.batch_size), 1))
        lib = hp.lib
        out = torch.empty(R, self.vae_input_size, dtype=torch.float32, device=self.device)
        wsb = lib.desire_cvae_decode_workspace_bytes(R, self.cfg.Z)
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=self.device)
        _lib.check(lib.desire_cvae_decode_fwd(C.c_void_p(z.data_ptr()), R, self.cfg.Z, C.byref(hp.w_vdec),
                                              C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), wsb,
                                              C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "vae_decoder")
        return out

    def get_name(self):
        """model/model.py:405-412."""
        return "cvae_input_%dx%d_latent%d_edim%d_ddim%d" % (
            self.input_shape[0], self.input_shape[1], self.latent_size,
            getattr(self.args, "e_dim", 256), self.cfg.d_dim)

    # ------------------------------------------------------------------ the hot path
    def _path(self, B) -> HotPath:
        if B not in self._paths:
            self._paths[B] = HotPath(self.cfg, self.weights, B, self.device)
        return self._paths[B]

    def _train_path(self, B) -> TrainPath:
        if B not in self._train_paths:
            self._train_paths[B] = TrainPath(self.cfg, self.flat_weights, self.weights, self._offsets, B, self.device,
                                             opt_state=(self.adam_m, self.adam_v, self.adam_step))
        return self._train_paths[B]

    def train_step(self, input_data, target_data, eps=None, scene=None, seed=None):
        """One optimiser step on a minibatch (D9): forward of the sample-generation stage, gradients of `cost`
        (what `tf.gradients(self.cost, tvars)` builds at model/model.py:388-391), clip_by_global_norm(grad_clip)
        and Adam(self.learning_rate) (:391-394) — the op the reference creates and never runs.  Under
        torch.distributed the flat gradient is all-reduced (sum) and `cost` is normalised by the global number
        of existing agents.  Returns the device tensor [cost, count] of this rank's scenes."""
        cfg = self.cfg
        to_dev = lambda n, x: x if (torch.is_tensor(x) and x.is_cuda) else self._stage(n, x.numpy() if torch.is_tensor(x) else x)
        obs, tgt = to_dev("obs", input_data), to_dev("tgt", target_data)
        B = obs.shape[0]
        tp = self._train_path(B)
        if eps is None:
            tp.set_noise(2 if seed is None else seed, 0)        # drawn on the device inside the step (desire_randn_fwd)
        else:
            eps = to_dev("eps", eps)
        if scene is None:
            scene = torch.zeros(B, cfg.scene_size, cfg.scene_size, 3, device=self.device)
        else:
            scene = to_dev("scene", scene)
        cost = tp.train_step(obs, tgt, eps, scene, lr=self.learning_rate, clip=self.grad_clip, use_graph=self.use_graph)
        self.input_data, self.target_data, self.target_data_enc, self.temporal_data = obs, tgt, tgt, obs
        self.cost, self.final_states, self.gradients = cost[0], tp.buf["HxHy"][:, :cfg.H], tp.G
        return cost

    def _stage(self, name, arr):
        """numpy -> pinned host staging buffer (reused) -> device, async on the current stream."""
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
        key = (name, tuple(t.shape))
        if key not in self._pinned:
            self._pinned[key] = (torch.empty(t.shape, dtype=torch.float32).pin_memory(),
                                 torch.empty(t.shape, dtype=torch.float32, device=self.device))
        pin, dev = self._pinned[key]
        pin.copy_(t)
        dev.copy_(pin, non_blocking=True)
        return dev

    def _pin(self, name, arr, slot=0):
        """numpy -> reused pinned host buffer (one per `slot`: back-to-back submissions alternate slots so the memcpy
        for pass t+1 never lands in a buffer whose H2D copy for pass t is still in flight)."""
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
        key = ("pin", name, tuple(t.shape), slot)
        if key not in self._pinned:
            self._pinned[key] = torch.empty(t.shape, dtype=torch.float32).pin_memory()
        self._pinned[key].copy_(t)
        return self._pinned[key]

    def forward(self, input_data, target_data, eps=None, scene=None, stages=("generate", "rank"), seed=None):
        """input_data [B,N,Tp,3], target_data [B,N,Tf,3] (id,x,y), agent-major as the reference's
        placeholders (model/model.py:91-105) with a leading scene axis; eps [B*N,K,Z] (None: drawn on the device from
        `seed`, D1 — desire_randn_fwd, element e of the draw is a function of (seed, e) only); scene [B,Hi,Wi,3].
        CUDA tensors are used in place, numpy/CPU inputs are staged through pinned memory.  Returns the dict of
        device outputs."""
        cfg = self.cfg
        to_dev = lambda n, x: x if (torch.is_tensor(x) and x.is_cuda) else self._stage(n, x.numpy() if torch.is_tensor(x) else x)
        obs = to_dev("obs", input_data)
        tgt = to_dev("tgt", target_data)
        B = obs.shape[0]
        hp = self._path(B)
        if eps is None:
            hp.set_noise(2 if seed is None else seed, 0)
        else:
            eps = to_dev("eps", eps)
        if scene is None:
            scene = torch.zeros(B, cfg.scene_size, cfg.scene_size, 3, device=self.device)
        else:
            scene = to_dev("scene", scene)
        out = hp.run(obs, tgt, eps, scene, stages)
        self.input_data, self.target_data, self.target_data_enc = obs, tgt, tgt
        self.temporal_data = obs
        self.rho_i, self.output_states, self.feature_pooling = out["rho_i"], out["output_states"], out["feature_pooling"]
        self.cost, self.final_states, self.final_output = out["cost"], out["H_x"], out["Y_refined"]
        return out

    # ------------------------------------------------------------------ host-buffer entry points
    def submit(self, input_data, target_data, eps=None, scene=None, seed=None):
        """Enqueue one pass of sample generation + IOC ranking/refinement for HOST buffers (numpy) and return a handle
        for result() WITHOUT waiting for the GPU: pinned staging -> H2D -> CUDA graphs -> D2H into pinned result
        buffers.  Two staging / result slots alternate, so a caller that submits pass t+1 before collecting pass t
        overlaps its host-side staging with the GPU work (the serving loop; sample_and_rank is submit + result).
        eps=None draws the noise on the device (reference: tf.random_normal inside the graph, model/model.py:262);
        successive submissions use successive offsets of the stream selected by `seed`."""
        cfg = self.cfg
        if torch.is_tensor(input_data) or not self.use_graph:
            out = self.forward(input_data, target_data, eps, scene, seed=seed)
            B = out["Y_refined"].shape[0] // (cfg.max_num_obj * cfg.K)
            hp = self._path(B)
            slot = 0
        else:
            B = int(np.shape(input_data)[0])
            hp = self._path(B)
            slot = self._submit_no & 1
            self._submit_no += 1
            ev = self._slot_done.get((B, slot))
            if ev is not None:
                ev.synchronize()           # the pass that last used this slot's pinned buffers has finished
            if scene is None:
                scene = np.zeros((B, cfg.scene_size, cfg.scene_size, 3), np.float32)
            if hp.graph_gen is None or hp.split_draws != (eps is None):
                pins = [self._pin(n, a, slot) for n, a in (("obs", input_data), ("tgt", target_data), ("scene", scene))]
                e0 = None if eps is None else self._pin("eps", eps, slot).to(self.device)
                hp.capture_split(pins[0].to(self.device), pins[1].to(self.device), e0, pins[2].to(self.device))
            pins = [self._pin(n, a, slot) for n, a in (("obs", input_data), ("tgt", target_data))]
            if eps is None:
                hp.set_noise(2 if seed is None else seed, self._noise_offset)
                self._noise_offset += 1
            # eps and the scene images (the large inputs) are staged while graphs that do not need them run
            out = hp.replay_split(*pins, stage_eps=None if eps is None else (lambda: self._pin("eps", eps, slot)),
                                  stage_scene=lambda: self._pin("scene", scene, slot))
        key = ("res", B, slot)
        if key not in self._pinned:
            self._pinned[key] = (torch.empty_like(out["Y_refined"], device="cpu").pin_memory(),
                                 torch.empty_like(out["ioc_scores"], device="cpu").pin_memory(),
                                 torch.empty(2, dtype=torch.float32).pin_memory())
        y, s, c = self._pinned[key]
        y.copy_(out["Y_refined"], non_blocking=True)
        s.copy_(out["ioc_scores"], non_blocking=True)
        c.copy_(hp.buf["cost"], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._slot_done[(B, slot)] = ev
        return (B, slot, ev)

    def result(self, handle):
        """Wait for a submit() and return (Y_refined [B,N,K,Tf,2], scores [iters,B,N,K], cost) as numpy views of the
        slot's pinned result buffers (valid until the slot is reused two submissions later)."""
        B, slot, ev = handle
        ev.synchronize()
        y, s, c = self._pinned[("res", B, slot)]
        N, K, Tf = self.cfg.max_num_obj, self.cfg.K, self.cfg.pred_length
        return (y.numpy().reshape(B, N, K, Tf, 2), s.numpy().reshape(-1, B, N, K), float(c[0]))

    def sample_and_rank(self, input_data, target_data, eps=None, scene=None, seed=None):
        """End-to-end public call with HOST buffers: stage inputs, run sample generation + IOC
        ranking/refinement, copy the ranked result back.  Returns (Y_refined [B,N,K,Tf,2],
        scores [iters,B,N,K], cost) as numpy."""
        return self.result(self.submit(input_data, target_data, eps, scene, seed))

    def rank_stream(self, batches, seed=None):
        """Throughput loop over an iterable of (input_data, target_data, scene) host batches: pass t+1 is staged and
        enqueued while the GPU runs pass t.  Yields result() tuples in order."""
        pending = None
        for inp, tgt, scene in batches:
            h = self.submit(inp, tgt, None, scene, seed)
            if pending is not None:
                yield self.result(pending)
            pending = h
        if pending is not None:
            yield self.result(pending)

    def sample(self, sess, traj, grid, dimensions, true_traj, num=10):
        """Signature of model/model.py:613 kept as an entry point (the reference body is a broken
        Social-LSTM leftover).  traj [obs_len, N, 3] -> [obs_len+num, N, 3]: the
        top-ranked of the K refined samples per agent is appended.  `sess`, `grid`, `dimensions`
        are accepted and ignored; `true_traj` ([obs_len+num, N, 3]) feeds the future encoder."""
        cfg = self.cfg
        traj = np.asarray(traj, np.float32)
        true_traj = np.asarray(true_traj, np.float32)
        obs_len = traj.shape[0]
        if obs_len != cfg.seq_length or num != cfg.pred_length:
            raise ValueError("sample(): traj must hold seq_length=%d frames and num must equal pred_length=%d"
                             % (cfg.seq_length, cfg.pred_length))
        inp = np.ascontiguousarray(traj.transpose(1, 0, 2))[None]                       # [1,N,Tp,3]
        tgt = np.ascontiguousarray(true_traj[obs_len:obs_len + num].transpose(1, 0, 2))[None]
        y, s, _ = self.sample_and_rank(inp, tgt)
        best = s[-1, 0].argmax(-1)                                                      # [N]
        pred = y[0, np.arange(cfg.max_num_obj), best]                                   # [N,Tf,2]
        ids = traj[-1, :, 0]
        new = np.concatenate([np.broadcast_to(ids[None, :, None], (num, cfg.max_num_obj, 1)),
                              pred.transpose(1, 0, 2)], -1)
        return np.vstack((traj, new.astype(np.float32)))

    # closed-form helpers kept for API parity (model/model.py:494-611); torch, any device
    def kld_loss(self, inputs, x_reconstr_mean, z_log_sigma_sq, z_mean):
        latent = -0.5 * torch.sum(1.0 + z_log_sigma_sq - z_mean ** 2 - torch.exp(z_log_sigma_sq), 1)
        return latent.mean()

    def get_coef(self, output):
        z_mux, z_muy, z_sx, z_sy, z_corr = torch.split(output, output.shape[1] // 5, 1)
        return [z_mux, z_muy, torch.exp(z_sx), torch.exp(z_sy), torch.tanh(z_corr)]

    def tf_2d_normal(self, x_val, y_val, mux, muy, sx_val, sy_val, rho):
        normx, normy = x_val - mux, y_val - muy
        sxsy = sx_val * sy_val
        z_val = (normx / sx_val) ** 2 + (normy / sy_val) ** 2 - 2 * (rho * normx * normy) / sxsy
        neg_rho = 1 - rho ** 2
        return torch.exp(-z_val / (2 * neg_rho)) / (2 * np.pi * sxsy * torch.sqrt(neg_rho))

    def get_reconstr_loss(self, z_mux, z_muy, z_sx, z_sy, z_corr, x_data, y_data):
        result0 = self.tf_2d_normal(x_data, y_data, z_mux, z_muy, z_sx, z_sy, z_corr)
        return torch.sum(-torch.log(torch.clamp(result0, min=1e-20)))

    def sample_gaussian_2d(self, mux, muy, sx_val, sy_val, rho):
        cov = [[sx_val * sx_val, rho * sx_val * sy_val], [rho * sx_val * sy_val, sy_val * sy_val]]
        x_val = np.random.multivariate_normal([mux, muy], cov, 1)
        return x_val[0][0], x_val[0][1]
