"""-m gpu: the scene CNN (desire_scene_cnn_fwd) against the oracle's float64 convolutions at map sizes that exercise the
tile-resident implicit-GEMM kernel (conv5_tc.cu: layers 2 and 3, and layer 1 in its space-to-depth form): several row tiles with a partial last one, a map narrower
than a tile, two column tiles with a partial second one, odd sizes; and that kernel against the im2col GEMM it replaces."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import rel_l2
from oracle import desire_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def make(B, Hi, Wi, Cs, seed):
    rng = np.random.default_rng(seed)
    P = {}
    for name, ci, co in (("c1", 3, 16), ("c2", 16, 32), ("c3", 32, Cs)):
        P["scene_%s_w" % name] = (rng.normal(size=(5, 5, ci, co)) / np.sqrt(25 * ci)).astype(np.float32)
        P["scene_%s_b" % name] = (0.1 * rng.normal(size=(co,))).astype(np.float32)
    img = rng.uniform(-1, 1, size=(B, Hi, Wi, 3)).astype(np.float32)
    return img, P


def scene_cnn_gpu(lib, img, P, Cs):
    from desire_b200 import _lib
    B, Hi, Wi, _ = img.shape
    Ho, Wo = (Hi + 1) // 2, (Wi + 1) // 2
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    keep = [d(P["scene_%s_%s" % (n, k)]) for n in ("c1", "c2", "c3") for k in ("w", "b")]
    w = _lib.SceneCnnW(*[t.data_ptr() for t in keep])
    img_d = d(img)
    out = torch.full((B, Ho, Wo, Cs), -7.0, device="cuda")
    nbytes = lib.desire_scene_cnn_workspace_bytes(B, Hi, Wi)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    _lib.check(lib.desire_scene_cnn_fwd(_ptr(img_d), B, Hi, Wi, Cs, C.byref(w), _ptr(out), _ptr(ws), nbytes, None), "scene_cnn")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def oracle64(img, P):
    P64 = {k: v.astype(np.float64) for k, v in P.items()}
    return O.scene_cnn(img.astype(np.float64), P64)


# (odd image sides: layer 1 keeps the im2col GEMM; Cs = 64: layer 3 does)
CASES = [(2, 64, 64, 32), (2, 256, 256, 32), (1, 300, 300, 32), (3, 30, 50, 32), (1, 14, 530, 16), (1, 2, 2, 64), (2, 31, 45, 32)]


@pytest.mark.parametrize("B,Hi,Wi,Cs", CASES)
def test_scene_cnn_matches_oracle(lib, B, Hi, Wi, Cs):
    img, P = make(B, Hi, Wi, Cs, 11 + Hi)
    got = scene_cnn_gpu(lib, img, P, Cs)
    ref = oracle64(img, P)
    assert got.shape == ref.shape
    assert np.isfinite(got).all()
    # three layers of 3xBF16 products (the lo*lo term, 2^-16 relative per product, is dropped) with FP32 accumulation over
    # K = 75 / 400 / 800: measured 8e-6
    assert rel_l2(got, ref) < 2e-5, rel_l2(got, ref)
    assert np.abs(got - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())


def test_implicit_gemm_agrees_with_im2col_path():
    """DESIRE_NO_CONV5=1 routes all three layers through the im2col GEMM (gemm_tc.cu); both paths multiply the same BF16 hi/lo
    splits and accumulate in FP32, in a different order."""
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from desire_b200 import _lib\n"
        "import test_gpu_scene_cnn as T\n"
        "lib = _lib.load()\n"
        "img, P = T.make(2, 200, 140, 32, 5)\n"
        "np.save(sys.argv[1], T.scene_cnn_gpu(lib, img, P, 32))\n" % (ROOT, os.path.join(ROOT, "tests")))
    outs = []
    for flag in ("0", "1"):
        path = "/tmp/scene_cnn_ab_%s.npy" % flag
        env = dict(os.environ, DESIRE_NO_CONV5=flag)
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
        outs.append(np.load(path))
    assert rel_l2(outs[0], outs[1]) < 5e-6, rel_l2(outs[0], outs[1])
