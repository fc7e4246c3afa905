"""Time desire_gemm_tc_fwd at a short-K / huge-M shape (the Decoder-2 input projection): python tools/bench_gemm.py [M N K]
DESIRE_GEMM_NO_PERSIST=1 selects the one-tile-per-CTA kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desire_b200 import _lib

M, N, K = [int(x) for x in sys.argv[1:4]] if len(sys.argv) >= 4 else (460800, 384, 48)
lib = _lib.load()
A = torch.randn(M, K, device="cuda")
W = torch.randn(K, N, device="cuda")
b = torch.randn(N, device="cuda")
Cd = torch.empty(M, N, device="cuda")
wsb = lib.desire_gemm_tc_workspace_bytes(N, K)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")


def call():
    _lib.check(lib.desire_gemm_tc_fwd(A.data_ptr(), K, W.data_ptr(), N, 0, b.data_ptr(), Cd.data_ptr(), N, M, N, K, 0, 0,
                                      ws.data_ptr(), wsb, None), "gemm")


for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    call()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
gb = (M * K + M * N) * 4 / 1e9
print("gemm M%d N%d K%d: %.3f ms per call (incl. the weight pack), %.2f TB/s of A + C traffic (%s)" % (
    M, N, K, ms, gb / ms, "one tile per CTA" if os.environ.get("DESIRE_GEMM_NO_PERSIST") == "1" else "persistent"))
