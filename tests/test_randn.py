"""The device noise source (desire_randn_fwd: Philox4x32-10 + Box-Muller; replaces tf.random_normal of
model/model.py:262).  CPU: the oracle's restatement against the published Philox4x32-10 known-answer vectors
(Random123 kat_vectors: counter/key all zeros, all ones, and the digits of pi) and basic moments.
GPU: the kernel reproduces the oracle element by element for any (seed, offset) and ragged n."""
import ctypes as C

import numpy as np
import pytest

from oracle import desire_oracle as O


def _philox_words(counter, key):
    """Raw Philox4x32-10 through the same round function the oracle's generator uses."""
    M = 0xFFFFFFFF
    c, (k0, k1) = list(counter), key
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M, p1 & M, ((p0 >> 32) ^ c[3] ^ k1) & M, p0 & M]
        k0, k1 = (k0 + 0x9E3779B9) & M, (k1 + 0xBB67AE85) & M
    return c


KAT = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


@pytest.mark.parametrize("ctr,key,out", KAT)
def test_philox_known_answers(ctr, key, out):
    assert tuple(_philox_words(ctr, key)) == out
    # the vectorised generator runs the same rounds: element 0..3 of a draw come from counter (0, 0, offset) under key seed
    seed, offset = key[0] | (key[1] << 32), ctr[2] | (ctr[3] << 32)
    if ctr[0] == 0 and ctr[1] == 0:
        z = O.philox_randn(seed, offset, 4)
        u = [((w >> 8) + 0.5) * 2.0 ** -24 for w in out]
        ref = [np.sqrt(-2 * np.log(u[0])) * np.cos(2 * np.pi * u[1]), np.sqrt(-2 * np.log(u[0])) * np.sin(2 * np.pi * u[1]),
               np.sqrt(-2 * np.log(u[2])) * np.cos(2 * np.pi * u[3]), np.sqrt(-2 * np.log(u[2])) * np.sin(2 * np.pi * u[3])]
        assert np.allclose(z, ref, atol=2e-6)


def test_oracle_noise_is_standard_normal_and_offsets_are_independent():
    z = O.philox_randn(42, 0, 400001)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3 and np.isfinite(z).all()
    z2 = O.philox_randn(42, 1, 400001)
    assert abs(np.corrcoef(z, z2)[0, 1]) < 5e-3
    assert np.array_equal(z[:1001], O.philox_randn(42, 0, 1001))          # a prefix of the same stream


@pytest.mark.gpu
@pytest.mark.parametrize("seed,offset,n", [(1, 0, 4), (2 ** 40 + 17, 2 ** 33 + 5, 1000003), (7, 123, 38400 * 128)])
def test_kernel_reproduces_oracle_noise(lib, seed, offset, n):
    import torch
    from desire_b200 import _lib
    st = torch.tensor([seed, offset], dtype=torch.int64, device="cuda")
    out = torch.full((n + 8,), 7.0, device="cuda")
    _lib.check(lib.desire_randn_fwd(C.c_void_p(st.data_ptr()), C.c_void_p(out.data_ptr()), n, None), "randn")
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.all(got[n:] == 7.0)
    ref = O.philox_randn(seed, offset, n)
    assert np.abs(got[:n] - ref).max() <= 2e-5      # logf / sincosf differ from numpy's by a few ulp
