"""Scene CNN alone at the bench shape (32 maps of 256 x 256 -> 128 x 128 x 32): ms per call, and per launch under ncu
(tools/r2z.sh)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desire_b200 import _lib

B, S, Cs = (int(sys.argv[1]), int(sys.argv[2]), 32) if len(sys.argv) >= 3 else (32, 256, 32)
lib = _lib.load()
g = torch.Generator().manual_seed(0)
ws_ = [(torch.randn(5, 5, ci, co, generator=g) / (25 * ci) ** 0.5).cuda() for ci, co in ((3, 16), (16, 32), (32, Cs))]
bs_ = [torch.zeros(co).cuda() for co in (16, 32, Cs)]
w = _lib.SceneCnnW(ws_[0].data_ptr(), bs_[0].data_ptr(), ws_[1].data_ptr(), bs_[1].data_ptr(), ws_[2].data_ptr(), bs_[2].data_ptr())
img = torch.rand(B, S, S, 3, generator=g).cuda()
out = torch.empty(B, S // 2, S // 2, Cs, device="cuda")
nbytes = lib.desire_scene_cnn_workspace_bytes(B, S, S)
ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
call = lambda: _lib.check(lib.desire_scene_cnn_fwd(p(img), B, S, S, Cs, C.byref(w), p(out), p(ws), nbytes, None), "scene_cnn")
for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    call()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
px = B * (S // 2) ** 2
flop = 2.0 * px * (75 * 16 + 400 * 32 + 800 * Cs)
print("scene CNN B%d %dx%d: %.3f ms per call, %.1f TFLOP/s algorithmic (%s)" % (
    B, S, S, ms, flop / (ms * 1e-3) / 1e12, "im2col GEMMs" if os.environ.get("DESIRE_NO_CONV5") == "1" else "tile-resident implicit GEMM"))
