import ctypes as C, os, sys, torch
sys.path.insert(0, "/root/repo")
from desire_b200 import _lib
lib = _lib.load()
out = torch.zeros(32, dtype=torch.int64, device="cuda")
for mode in (6, 5):
    for N in (16, 32, 48, 64, 96, 128):
        for it in range(2):
            _lib.check(lib.desire_selftest_mma_rate(mode, N, 4092, 148, C.c_void_p(out.data_ptr()), None), "rate")
            torch.cuda.synchronize()
        print("mode %d (%s) N=%3d: %.1f cycles per MMA (N/2 = %d)" % (mode, "SS" if mode == 6 else "TS", N, out[0].item() / 4092, N // 2))

for shift in (0, 1, 2, 4, 7):
    for N in (32, 128):
        for it in range(2):
            _lib.check(lib.desire_selftest_mma_rate(100 * shift + 6, N, 4092, 148, C.c_void_p(out.data_ptr()), None), "rate")
            torch.cuda.synchronize()
        print("SS, A operand starting %3d bytes into its core matrix, N=%3d: %.1f cycles per MMA" % (16 * shift, N, out[0].item() / 4092))
