"""Micro-benchmark of the Decoder-1 GRU recurrence alone (desire_gru_decode_fwd) at a given shape.
    python tools/bench_gru.py --rows 38400 --hidden 128 --steps 12 [--mode 3]"""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from desire_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=38400)
ap.add_argument("--hidden", type=int, default=128)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--k", type=int, default=20)
ap.add_argument("--mode", type=int, default=3)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
lib = _lib.load()
lib.desire_set_gemm_mode(a.mode)
R, H, T, K = a.rows, a.hidden, a.steps, a.k
g = torch.Generator().manual_seed(0)
xz = torch.randn(R, H, generator=g).cuda()
Hx = torch.randn(R // K, H, generator=g).cuda()
lim = (6.0 / (3 * H)) ** 0.5
w = dict(wg=((torch.rand(2 * H, 2 * H, generator=g) * 2 - 1) * lim).cuda(), bg=torch.ones(2 * H).cuda(),
         wc=((torch.rand(2 * H, H, generator=g) * 2 - 1) * lim).cuda(), bc=torch.zeros(H).cuda())
gw = _lib.GruW(*[w[k].data_ptr() for k in ("wg", "bg", "wc", "bc")])
hs = torch.empty(R, T, H, device="cuda")
wsb = lib.desire_gru_decode_workspace_bytes(R, H)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
lib.desire_prof_enable(1)
for _ in range(a.iters + 2):
    _lib.check(lib.desire_gru_decode_fwd(xz.data_ptr(), Hx.data_ptr(), H, R, K, H, T, C.byref(gw), hs.data_ptr(), ws.data_ptr(), wsb, None), "gru")
torch.cuda.synchronize()
n, ms = C.c_long(0), C.c_double(0)
lib.desire_prof_read(0, C.byref(n), C.byref(ms))
per = ms.value / n.value
fl = R * 6.0 * H * H * T
print("gru_decode R=%d H=%d T=%d mode=%d: %.3f ms/launch (%d launches)  %.1f TFLOP/s algorithmic, x%d MMA passes" % (R, H, T, a.mode, per, n.value, fl / per / 1e9, 3 if a.mode == 3 else 1))
