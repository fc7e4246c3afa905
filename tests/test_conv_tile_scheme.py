"""CPU: the two index identities the tile-resident convolution kernel (desire_b200/csrc/conv5_tc.cu) rests on, restated in
NumPy and checked against the oracle's conv2d_tf (the GPU tests in test_gpu_scene_cnn.py check the kernel itself).

1. A filter tap is a row shift: with the zero-padded tile stored as rows p = r * WP + c, the 5x5 SAME convolution of the
   tile is  D[p] = sum_{ky,kx} A[p + ky*WP + kx] @ W[ky, kx]  for the positions p = r * WP + c with c < TW; the positions in
   the four halo columns are garbage and dropped.
2. Space-to-depth: a 5x5 / stride-2 / SAME convolution of an image with even sides equals a 3x3 / stride-1 / SAME
   convolution of S[Y, X, (py, px, c)] = img[2Y + py, 2X + px, c] with w3[a+1, b+1, (py, px, c)] = w[2a + py + 1, 2b + px + 1, c]
   (zero where that tap does not exist) — c5_pack_s2d_kernel builds exactly this w3."""
import numpy as np
import pytest

from oracle import desire_oracle as O


def conv_by_row_shifts(x, w, TH, TW):
    """x [H, W, Ci] (one map), w [5, 5, Ci, Co]: the kernel's tile scheme with TH x TW output tiles, float64."""
    H, W, Ci = x.shape
    Co = w.shape[3]
    WP = TW + 4
    out = np.zeros((H, W, Co))
    for y0 in range(0, H, TH):
        for x0 in range(0, W, TW):
            rows = min(TH, H - y0)
            tile = np.zeros(((rows + 4) * WP + 4 * WP + 8, Ci))          # + slack: shifted reads of the last positions
            for r in range(rows + 4):
                for c in range(WP):
                    iy, ix = y0 - 2 + r, x0 - 2 + c
                    if 0 <= iy < H and 0 <= ix < W:
                        tile[r * WP + c] = x[iy, ix]
            npos = rows * WP
            D = np.zeros((npos, Co))
            for ky in range(5):
                for kx in range(5):
                    s = ky * WP + kx
                    D += tile[s:s + npos] @ w[ky, kx]
            for r in range(rows):
                for c in range(min(TW, W - x0)):
                    out[y0 + r, x0 + c] = D[r * WP + c]
    return out


@pytest.mark.parametrize("H,W,TH,TW", [(9, 11, 7, 128), (16, 20, 7, 8), (5, 5, 3, 4)])
def test_tap_is_a_row_shift(H, W, TH, TW):
    rng = np.random.default_rng(H * 100 + W)
    x = rng.normal(size=(H, W, 6))
    w = rng.normal(size=(5, 5, 6, 4))
    ref = O.conv2d_tf(x[None], w, np.zeros(4), 1, "SAME")[0]
    got = conv_by_row_shifts(x, w, TH, TW)
    assert np.abs(got - ref).max() < 1e-12


def s2d_weights(w):
    """w [5, 5, 3, Co] -> w3 [3, 3, 12, Co], the index rule of c5_pack_s2d_kernel."""
    Co = w.shape[3]
    w3 = np.zeros((3, 3, 12, Co))
    for a in (-1, 0, 1):
        for b in (-1, 0, 1):
            for ci in range(12):
                py, px, c = ci // 6, (ci // 3) % 2, ci % 3
                ky, kx = 2 * a + py + 1, 2 * b + px + 1
                if 0 <= ky < 5 and 0 <= kx < 5:
                    w3[a + 1, b + 1, ci] = w[ky, kx, c]
    return w3


@pytest.mark.parametrize("Hi,Wi", [(8, 8), (6, 14), (2, 2)])
def test_stride2_conv_is_a_3x3_conv_on_the_space_to_depth_image(Hi, Wi):
    rng = np.random.default_rng(Hi + Wi)
    img = rng.normal(size=(2, Hi, Wi, 3))
    w = rng.normal(size=(5, 5, 3, 5))
    b = rng.normal(size=(5,))
    ref = O.conv2d_tf(img, w, b, 2, "SAME")
    # S[Y, X, (py, px, c)] = img[2Y + py, 2X + px, c]
    S = img.reshape(2, Hi // 2, 2, Wi // 2, 2, 3).transpose(0, 1, 3, 2, 4, 5).reshape(2, Hi // 2, Wi // 2, 12)
    got = O.conv2d_tf(S, s2d_weights(w), b, 1, "SAME")
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 1e-12
