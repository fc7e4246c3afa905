"""tcgen05.mma issue-to-retire rate per shape / operand source (tools/r2n.sh)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desire_b200 import _lib

lib = _lib.load()
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for grid in (148,):
    for mode in (5, 10, 11, 13):
        for N in (128,):
            for it in range(2):
                _lib.check(lib.desire_selftest_mma_rate(mode, N, 4092, grid, C.c_void_p(out.data_ptr()), None), "rate")
                torch.cuda.synchronize()
            cyc = out.item() / 4092
            print("grid %3d  %s  N=%3d: %.1f cycles per MMA (floor %d) -> %.0f%% of the tensor peak" % (
                grid, ("SS", "TS", "SS two accumulators", "TS two accumulators", "TS two issuing threads (cycles per MMA of ONE thread)", "TS elected lane, uniform operands", "SS elected lane, uniform operands", "TS elected + commit per 6 MMAs", "TS elected + wait, fence, commit per 6 MMAs", "TS elected + wait, fence per 6 MMAs", "TS elected, accumulators alternate every MMA", "TS elected, accumulators alternate every 6 MMAs", "", "TS two issuing warps, elected lanes (cycles per MMA of ONE warp; 128 = the pipe is shared without loss)")[mode], N, cyc, N // 2, 100 * (N / 2) / cyc))

out = torch.zeros(32, dtype=torch.int64, device="cuda")
for it in range(2):
    _lib.check(lib.desire_selftest_mma_rate(12, 128, 64, 148, C.c_void_p(out.data_ptr()), None), "rate")
    torch.cuda.synchronize()
print("cycles until the n-th MMA (N=128) was issued:", ", ".join("%d: %d" % (4 * k + 4, out[1 + k].item()) for k in range(16)))

for mode, what in ((14, "nothing else running"), (15, "16 more warps polling an mbarrier"),
                   (16, "16 more warps streaming 16-byte shared-memory stores and loads"),
                   (17, "16 more warps issuing tcgen05.ld on idle columns")):
    for it in range(2):
        _lib.check(lib.desire_selftest_mma_rate(mode, 128, 4000, 148, C.c_void_p(out.data_ptr()), None), "rate")
    torch.cuda.synchronize()
    print("stage mix of the fused social kernel (24 TS + 16 SS MMAs, N=128, two issuing warps, %s): %.0f cycles per stage (floor 2560)" % (what, out[0].item() / 100))

for mode, what in ((18, "all of it"), (19, "handshakes only (no tcgen05.ld / st)"), (20, "fc MMAs in four blocks with a commit each"),
                   (21, "P released after the conversion"), (22, "all of it + fence.proxy.async per thread and stage"),
                   (23, "all of it + two 16-byte shared-memory stores and fence.proxy.async per thread and stage")):
    for it in range(2):
        _lib.check(lib.desire_selftest_mma_rate(mode, 128, 4000, 148, C.c_void_p(out.data_ptr()), None), "rate")
    torch.cuda.synchronize()
    print("stage protocol of the fused social kernel in miniature, %s: %.0f cycles per stage (floor 2560)" % (what, out[0].item() / 100))
