"""-m gpu: hardware self-tests of tcgen05 forms the kernels rely on."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tcgen05_mma_with_a_operand_in_tensor_memory(lib):
    """tcgen05.mma [d], [a_tmem], b_desc: a BF16 A operand written with tcgen05.st as packed pairs (lower half = the smaller k)
    gives the same product as the shared-memory operand, and both match the BF16-rounded reference."""
    from desire_b200 import _lib
    g = torch.Generator().manual_seed(0)
    A = torch.randn(128, 32, generator=g)
    B = torch.randn(64, 32, generator=g)
    ref = (A.bfloat16().double() @ B.bfloat16().double().T).numpy()
    Ad, Bd = A.cuda(), B.cuda()
    res = {}
    for order in (0, 1):
        ss, ts = torch.zeros(128, 64, device="cuda"), torch.zeros(128, 64, device="cuda")
        _lib.check(lib.desire_selftest_tsmma(C.c_void_p(Ad.data_ptr()), C.c_void_p(Bd.data_ptr()), C.c_void_p(ss.data_ptr()),
                                             C.c_void_p(ts.data_ptr()), order, None), "selftest")
        torch.cuda.synchronize()
        res[order] = (ss.cpu().numpy(), ts.cpu().numpy())
        print("order %d: |ss-ref| %.3e  |ts-ref| %.3e" % (order, np.abs(res[order][0] - ref).max(), np.abs(res[order][1] - ref).max()))
    assert np.abs(res[0][0] - ref).max() < 1e-4
    assert np.abs(res[0][1] - ref).max() < 1e-4, "TMEM A operand: packed pairs with the smaller k in the lower half"
    assert np.abs(res[1][1] - ref).max() > 1e-2          # the swapped order is a different product

