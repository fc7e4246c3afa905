"""-m gpu: the tcgen05 GEMM (desire_gemm_tc_fwd) against a float64 matmul.
3xBF16 mode must sit at FP32-class accuracy (<= 2e-5 rel-L2, leaving headroom under the 1e-4 bar of the
whole path); the single-pass BF16 mode is checked against its own expected ~3e-3."""
import ctypes as C

import numpy as np
import pytest
import torch

from desire_b200 import _lib
from helpers import rel_l2

pytestmark = pytest.mark.gpu

SHAPES = [  # M, N, K, trans_w, act, accumulate, bias
    (128, 128, 32, 0, 0, 0, 1),
    (128, 16, 8, 0, 0, 0, 0),
    (300, 48, 100, 0, 1, 0, 1),          # ragged M, N, K
    (1000, 1600, 128, 1, 0, 0, 0),       # CVAE deconv2 shape (W stored [N,K])
    (2048, 800, 64, 1, 0, 0, 0),         # deconv3
    (777, 25, 32, 1, 0, 0, 0),           # deconv4: N=25 padded to 32
    (4096, 128, 4608, 0, 1, 0, 1),       # social fc
    (1536, 384, 248, 0, 0, 0, 1),        # decoder-2 input projection, K tail
    (640, 24, 128, 0, 0, 1, 1),          # regression refine: accumulate into C, N=24
    (513, 2048, 128, 1, 2, 0, 1),        # deconv1 shape + ELU
    # short K, huge M: the persistent weight-stationary kernel (ragged last tile; one and two K stages; N = 3H of both sizes)
    (100000 + 77, 384, 48, 0, 0, 0, 1),
    (90000, 512, 64, 0, 1, 0, 1),
    (76000, 96, 24, 1, 0, 0, 0),
    (80000 + 5, 768, 48, 0, 0, 0, 1),    # N = 3H at H = 256: the packed weights do not fit shared memory -> the tiled kernel
]


def run(M, N, K, trans, act, accumulate, bias, mode, seed=0):
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g)
    W = torch.randn((N, K) if trans else (K, N), generator=g)
    b = torch.randn(N, generator=g) if bias else None
    C0 = torch.randn(M, N, generator=g)
    ref = A.double() @ (W.double().t() if trans else W.double())
    if bias:
        ref = ref + b.double()
    if accumulate:
        ref = ref + C0.double()
    if act == 1:
        ref = torch.relu(ref)
    elif act == 2:
        ref = torch.nn.functional.elu(ref)
    Ad, Wd, Cd = A.cuda(), W.cuda(), C0.cuda().clone()
    bd = b.cuda() if bias else None
    wsb = lib.desire_gemm_tc_workspace_bytes(N, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    lib.desire_set_gemm_mode(mode)
    try:
        rc = lib.desire_gemm_tc_fwd(Ad.data_ptr(), K, Wd.data_ptr(), K if trans else N, trans,
                                    bd.data_ptr() if bias else None, Cd.data_ptr(), N, M, N, K, act, accumulate,
                                    ws.data_ptr(), wsb, None)
        _lib.check(rc, "gemm_tc")
        torch.cuda.synchronize()
    finally:
        lib.desire_set_gemm_mode(3)
    return rel_l2(Cd.cpu().numpy(), ref.numpy())


@pytest.mark.parametrize("shape", SHAPES, ids=[str(s[:3]) for s in SHAPES])
def test_gemm_tc_3xbf16(shape):
    e = run(*shape, mode=3)
    print("3xBF16", shape[:3], "rel-L2 %.3e" % e)
    assert e < 2e-5


@pytest.mark.parametrize("shape", SHAPES[:4], ids=[str(s[:3]) for s in SHAPES[:4]])
def test_gemm_tc_bf16_single_pass(shape):
    e = run(*shape, mode=1)
    print("BF16  ", shape[:3], "rel-L2 %.3e" % e)
    assert e < 1e-2


def test_strided_views():
    """lda/ldc larger than K/N (H_x|H_y style column slices)."""
    lib = _lib.load()
    M, N, K = 256, 64, 96
    A = torch.randn(M, K + 8).cuda()
    W = torch.randn(K, N).cuda()
    Cbuf = torch.zeros(M, N + 32).cuda()
    wsb = lib.desire_gemm_tc_workspace_bytes(N, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    _lib.check(lib.desire_gemm_tc_fwd(A.data_ptr(), K + 8, W.data_ptr(), N, 0, None, Cbuf.data_ptr() + 16 * 4, N + 32,
                                      M, N, K, 0, 0, ws.data_ptr(), wsb, None), "gemm_tc")
    torch.cuda.synchronize()
    ref = (A[:, :K].double() @ W.double()).cpu().numpy()
    assert rel_l2(Cbuf[:, 16:16 + N].cpu().numpy(), ref) < 2e-5
    assert float(Cbuf[:, :16].abs().max()) == 0 and float(Cbuf[:, 16 + N:].abs().max()) == 0
