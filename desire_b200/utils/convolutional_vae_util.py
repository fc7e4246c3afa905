"""The reference's utils/convolutional_vae_util.py surface on top of the kernel library: `deconv2d`,
`get2d_deconv_output_size`, `_kernel`, `_stride` with the reference's argument names and defaults (:31-44,:141-142,
:172,:190) and its ValueErrors (:76-80,:96,:159).

The reference registers `deconv2d` as a prettytensor method that creates its variables in a graph-wide store; there is
no variable store here, so the parameters travel explicitly: `params` is a dict with 'weights' [kh,kw,depth,in]
(the reference's filter layout, :83), and optionally 'bias' [depth] (:117-121), 'gamma'/'beta' [depth] for
batch_normalize.  When `params` is None they are created the way the reference would (xavier over
depth*patch / in*patch, :88-91; truncated normal with `stddev`, zeros with stddev == 0; zero bias; gamma 1, beta 0) and
returned next to the output.  Operation order is the reference's: conv2d_transpose -> +bias -> batch-normalise ->
activation (:113-134); batch-norm uses per-row statistics."""
from __future__ import annotations

import ctypes as C
import math

import torch

from .. import _lib

PAD_SAME = "SAME"
PAD_VALID = "VALID"
_ACTS = {None: 0, "relu": 1, "elu": 2, "sigmoid": 3}


def get2d_deconv_output_size(input_height, input_width, filter_height, filter_width, row_stride, col_stride, padding_type):
    """Rows and columns of a transposed convolution's output: VALID (in-1)*stride + filter, SAME in*stride; an unknown
    (None) extent stays None."""
    def one(n, f, s):
        if n is None or f is None:
            return None
        if padding_type == PAD_VALID:
            return (int(n) - 1) * int(s) + int(f)
        if padding_type == PAD_SAME:
            return int(n) * int(s)
        raise ValueError("Invalid value for padding: %r" % padding_type)
    return one(input_height, filter_height, row_stride), one(input_width, filter_width, col_stride)


def _kernel(kernel_spec):
    """int or length-1/2 sequence -> [kh, kw]."""
    if isinstance(kernel_spec, int):
        return [kernel_spec, kernel_spec]
    if len(kernel_spec) == 1:
        return [kernel_spec[0], kernel_spec[0]]
    assert len(kernel_spec) == 2
    return list(kernel_spec)


def _stride(stride_spec):
    """None, int or length-1/2/4 sequence -> [1, sh, sw, 1]."""
    if stride_spec is None:
        return [1, 1, 1, 1]
    if isinstance(stride_spec, int):
        return [1, stride_spec, stride_spec, 1]
    if len(stride_spec) == 1:
        return [1, stride_spec[0], stride_spec[0], 1]
    if len(stride_spec) == 2:
        return [1, stride_spec[0], stride_spec[1], 1]
    assert len(stride_spec) == 4
    return list(stride_spec)


def _act_code(activation_fn):
    fn = activation_fn[0] if isinstance(activation_fn, (tuple, list)) else activation_fn
    name = getattr(fn, "__name__", fn)
    if name not in _ACTS:
        raise ValueError("deconv2d: unsupported activation %r (None, 'relu', 'elu', 'sigmoid')" % (name,))
    return _ACTS[name]


def deconv2d(input_layer, kernel, depth, name=None, stride=None, activation_fn=None, l2loss=None, init=None, stddev=None,
             bias=True, edges=PAD_SAME, batch_normalize=False, phase=None, params=None, seed=0):
    """input_layer: CUDA float32 tensor [batch, H, W, in] (NHWC, H == W).  Returns (output [batch, H', W', depth], params)."""
    if input_layer.dim() != 4:
        raise ValueError("Cannot perform conv2d on tensor with shape %s" % (tuple(input_layer.shape),))
    if init is not None and stddev is not None:
        raise ValueError("Do not set both init and stddev.")
    if not input_layer.is_cuda:
        raise _lib.DesireError("deconv2d runs on a CUDA device only (no CPU fallback)")
    kh, kw = _kernel(kernel)
    st = _stride(stride)
    R, Hin, Win, Cin = (int(v) for v in input_layer.shape)
    if Hin != Win or kh != kw or st[1] != st[2]:
        raise ValueError("deconv2d: the kernel library handles square inputs, kernels and strides")
    out_rows, out_cols = get2d_deconv_output_size(Hin, Win, kh, kw, st[1], st[2], edges)
    dev = input_layer.device
    if params is None:
        g = torch.Generator().manual_seed(seed)
        size = (kh, kw, depth, Cin)
        if init is not None:
            wgt = torch.as_tensor(init(size) if callable(init) else init, dtype=torch.float32).reshape(size)
        elif stddev is None:
            lim = math.sqrt(6.0 / (kh * kw * (depth + Cin)))            # layers.xavier_init(depth*patch, in*patch)
            wgt = (torch.rand(size, generator=g) * 2 - 1) * lim
        elif stddev:
            wgt = torch.randn(size, generator=g).clamp_(-2, 2) * stddev
        else:
            wgt = torch.zeros(size)
        params = {"weights": wgt}
        if bias:
            params["bias"] = torch.zeros(depth)
        if batch_normalize:
            params["gamma"], params["beta"] = torch.ones(depth), torch.zeros(depth)
    params = {k: v.to(dev, torch.float32).contiguous() for k, v in params.items()}
    if tuple(params["weights"].shape) != (kh, kw, depth, Cin):
        raise ValueError("deconv2d: weights must be [kh,kw,depth,in] = %s" % ((kh, kw, depth, Cin),))
    x = input_layer.to(torch.float32).contiguous()
    lib = _lib.load()
    same = 1 if edges == PAD_SAME else 0
    y = torch.empty(R, out_rows, out_cols, depth, dtype=torch.float32, device=dev)
    wsb = lib.desire_deconv2d_workspace_bytes(R, Hin, Cin, kh, st[1], same, depth)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
    bn = batch_normalize and "gamma" in params
    _lib.check(lib.desire_deconv2d_fwd(p(x), R, Hin, Cin, p(params["weights"]), kh, st[1], same, depth,
                                       p(params.get("bias") if bias else None), p(params["gamma"] if bn else None),
                                       p(params["beta"] if bn else None), _act_code(activation_fn), p(y), p(ws), wsb,
                                       C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "deconv2d")
    return y, params
