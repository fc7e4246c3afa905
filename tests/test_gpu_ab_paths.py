"""-m gpu: the A/B switches keep the superseded kernels reachable (DESIRE_SOCIAL_V1: first-design fused social kernel;
DESIRE_NO_FUSE4: the decoder's last layer as its own kernel; DESIRE_GEMM_NO_PERSIST: one tile per CTA for the tall GEMMs;
DESIRE_GRU_V2: second-design recurrence).  Each switch is read once per process, so every variant runs in its own
interpreter on the same seeded inputs; all of them must agree with the default path to fp32-class accuracy."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, json, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
from helpers import small_cfg
from desire_b200.config import init_params
from desire_b200.engine import HotPath
from desire_b200.synthetic import make_batch
cfg = small_cfg(d_dim=128, max_num_obj=20, num_samples=5, scene_size=64, n_rad=1, n_ang=1, r_min=1e-6, r_max=1e3)
B = 4
hp = HotPath(cfg, init_params(cfg, 1), B)
out = hp.run(*[t.cuda() for t in make_batch(cfg, B, 0, 2)])
torch.cuda.synchronize()
np.savez(sys.argv[1], **{k: out[k].cpu().numpy() for k in ("x_reconstr_mean", "output_states", "Yhat", "Y_refined", "ioc_scores")})
""" % (ROOT, ROOT)


def _run(tmp_path, name, env):
    f = str(tmp_path / (name + ".npz"))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", SCRIPT, f], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return dict(np.load(f))


def test_superseded_kernels_agree_with_the_default_path(tmp_path):
    ref = _run(tmp_path, "default", {})
    for name, env in (("social_v1", {"DESIRE_SOCIAL_V1": "1"}), ("no_fuse4", {"DESIRE_NO_FUSE4": "1"}),
                      ("gemm_tiled", {"DESIRE_GEMM_NO_PERSIST": "1"}), ("gru_v2", {"DESIRE_GRU_V2": "1"})):
        got = _run(tmp_path, name, env)
        for k, v in ref.items():
            err = float(np.linalg.norm(got[k].astype(np.float64) - v) / max(np.linalg.norm(v), 1e-30))
            print("%-12s %-16s rel-L2 vs default %.2e" % (name, k, err))
            assert err < 5e-5, (name, k, err)
