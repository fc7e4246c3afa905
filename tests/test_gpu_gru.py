"""-m gpu: the GRU recurrences through the C-ABI vs the oracle's TF-1.x GRUCell (oracle/desire_oracle.py:gru_cell,
reference model/model.py:137-148,233-241,279-285), stand-alone, over the shapes that select each kernel:
H in {128, 256} -> gru_tc3_kernel (TMA-staged xp, register state), other multiples of 32 -> gru_tc_kernel,
the rest -> the FP32 recurrence.  Ragged row counts exercise the zero-filled TMA boxes and the store guards."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu


def _weights(H, I, seed):
    rng = np.random.default_rng(seed)
    lim = (6.0 / (I + 3 * H)) ** 0.5
    return dict(wg=rng.uniform(-lim, lim, (I + H, 2 * H)).astype(np.float32), bg=np.ones(2 * H, np.float32),
                wc=rng.uniform(-lim, lim, (I + H, H)).astype(np.float32), bc=(0.1 * rng.normal(size=H)).astype(np.float32))


def _gruw(w):
    from desire_b200 import _lib
    d = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
    return d, _lib.GruW(*[d[k].data_ptr() for k in ("wg", "bg", "wc", "bc")])


@pytest.mark.parametrize("H,M,K,T", [(128, 7, 10, 12), (128, 64, 2, 1), (128, 33, 9, 3), (256, 5, 13, 12), (256, 40, 4, 2),
                                     (128, 300, 20, 12), (256, 103, 5, 5), (64, 20, 5, 12), (48, 9, 3, 7)])
def test_decoder1_recurrence(lib, H, M, K, T):
    from desire_b200 import _lib
    from oracle import desire_oracle as O
    R = M * K
    rng = np.random.default_rng(H + R)
    x_z = rng.normal(size=(R, H)).astype(np.float32)
    Hx = rng.normal(size=(M, H)).astype(np.float32)
    w = _weights(H, H, 3)
    ref = O.gru_decode(x_z.astype(np.float64), np.repeat(Hx, K, 0).astype(np.float64),
                       *[w[k].astype(np.float64) for k in ("wg", "bg", "wc", "bc")], T)
    wd, gw = _gruw(w)
    xz_d, hx_d = torch.from_numpy(x_z).cuda(), torch.from_numpy(Hx).cuda()
    hs = torch.full((R, T, H), 7.0, device="cuda")
    wsb = lib.desire_gru_decode_workspace_bytes(R, H)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    _lib.check(lib.desire_gru_decode_fwd(xz_d.data_ptr(), hx_d.data_ptr(), H, R, K, H, T, C.byref(gw), hs.data_ptr(),
                                         ws.data_ptr(), wsb, None), "gru_decode")
    torch.cuda.synchronize()
    got = hs.cpu().numpy()
    for t in (0, T - 1):
        e = rel_l2(got[:, t], ref[:, t])
        print("H=%d R=%d step %d rel-L2 %.3e" % (H, R, t, e))
        assert e <= 2e-5, (t, e)
    assert rel_l2(got, ref) <= 2e-5


@pytest.mark.parametrize("H,M,T", [(128, 200, 8), (256, 70, 12), (128, 1920, 12), (64, 50, 8)])
def test_encoder_recurrence(lib, H, M, T):
    """Per-step input projection (xp varies with t), zero initial state, final state only."""
    from desire_b200 import _lib
    from oracle import desire_oracle as O
    rng = np.random.default_rng(H + M + T)
    traj = rng.normal(size=(M, T, 3)).astype(np.float32)
    w = _weights(H, 2, 5)
    ref = O.gru_encode(traj[:, :, 1:3].astype(np.float64), *[w[k].astype(np.float64) for k in ("wg", "bg", "wc", "bc")])
    wd, gw = _gruw(w)
    tr_d = torch.from_numpy(traj).cuda()
    out = torch.full((M, 2 * H), 7.0, device="cuda")
    wsb = lib.desire_gru_encode_workspace_bytes(M, T, H)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    _lib.check(lib.desire_gru_encode_ws_fwd(tr_d.data_ptr(), M, T, H, C.byref(gw), out.data_ptr(), 2 * H, ws.data_ptr(), wsb,
                                            None), "gru_encode")
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    e = rel_l2(got[:, :H], ref)
    print("encoder H=%d M=%d T=%d rel-L2 %.3e" % (H, M, T, e))
    assert e <= 2e-5
    assert float(np.abs(got[:, H:] - 7.0).max()) == 0     # the other half of the [H_x | H_y] buffer is untouched
