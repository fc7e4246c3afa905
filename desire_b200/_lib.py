"""ctypes binding of libdesire_b200.so — the stub a reference maintainer would add.

The library is the ONLY compute path: if it is missing this module raises (there is no CPU or
PyTorch fallback).  Signatures mirror include/desire_abi.h one to one.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libdesire_b200.so")

c_float_p = C.c_void_p  # device pointers travel as integers


class GruW(C.Structure):
    _fields_ = [("wg", C.c_void_p), ("bg", C.c_void_p), ("wc", C.c_void_p), ("bc", C.c_void_p)]


class ConvBnW(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p)]


class CvaeEncW(C.Structure):
    _fields_ = [("c1", ConvBnW), ("c2", ConvBnW), ("c3", ConvBnW), ("fc_w", C.c_void_p), ("fc_b", C.c_void_p)]


class CvaeDecW(C.Structure):
    _fields_ = [("d1", ConvBnW), ("d2", ConvBnW), ("d3", ConvBnW), ("d4", ConvBnW)]


class SceneCnnW(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("c1_w", "c1_b", "c2_w", "c2_b", "c3_w", "c3_b")]


class IocW(C.Structure):
    _fields_ = [("vel_w", C.c_void_p), ("vel_b", C.c_void_p), ("sp_w", C.c_void_p), ("sp_b", C.c_void_p),
                ("dec2", GruW), ("score_w", C.c_void_p), ("score_b", C.c_void_p),
                ("reg_w", C.c_void_p), ("reg_b", C.c_void_p), ("r2_edges", C.c_void_p), ("dirs", C.c_void_p)]


class GruG(C.Structure):
    _fields_ = [("wg", C.c_void_p), ("bg", C.c_void_p), ("wc", C.c_void_p), ("bc", C.c_void_p)]


class ConvBnG(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p)]


class CvaeEncG(C.Structure):
    _fields_ = [("c1", ConvBnG), ("c2", ConvBnG), ("c3", ConvBnG), ("fc_w", C.c_void_p), ("fc_b", C.c_void_p)]


class CvaeDecG(C.Structure):
    _fields_ = [("d1", ConvBnG), ("d2", ConvBnG), ("d3", ConvBnG), ("d4", ConvBnG)]


class IocG(C.Structure):
    _fields_ = [("vel_w", C.c_void_p), ("vel_b", C.c_void_p), ("sp_w", C.c_void_p), ("sp_b", C.c_void_p),
                ("dec2", GruG), ("score_w", C.c_void_p), ("score_b", C.c_void_p),
                ("reg_w", C.c_void_p), ("reg_b", C.c_void_p)]


class SceneCnnG(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("c1_w", "c1_b", "c2_w", "c2_b", "c3_w", "c3_b")]


class IocDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("B", "N", "K", "H", "Tf", "C", "Fv", "Cs", "n_rad", "n_ang", "Hm", "Wm", "iters")]


P, I, L, Z, F = C.c_void_p, C.c_int, C.c_long, C.c_size_t, C.c_float

# name -> (restype, argtypes); the single source of truth the ABI test checks against the header
SIGNATURES = {
    "desire_version": (I, []),
    "desire_last_error": (C.c_char_p, []),
    "desire_launch_count": (L, []),
    "desire_selftest_tsmma": (I, [P, P, P, P, I, P]),
    "desire_selftest_mma_rate": (I, [I, I, I, I, P, P]),
    "desire_fallback_count": (L, [I]),
    "desire_set_gemm_mode": (I, [I]),
    "desire_get_gemm_mode": (I, []),
    "desire_prof_enable": (I, [I]),
    "desire_prof_read": (I, [I, C.POINTER(C.c_long), C.POINTER(C.c_double)]),
    "desire_fc_fwd": (I, [P, I, P, I, P, P, I, I, I, I, I, I, P]),
    "desire_gemm_tc_workspace_bytes": (Z, [I, I]),
    "desire_gemm_tc_fwd": (I, [P, I, P, I, I, P, P, I, I, I, I, I, I, P, Z, P]),
    "desire_tconv_fwd": (I, [P, I, I, I, P, P, P, P]),
    "desire_gru_encode_fwd": (I, [P, I, I, I, C.POINTER(GruW), P, I, P]),
    "desire_gru_encode_workspace_bytes": (Z, [I, I, I]),
    "desire_gru_encode_ws_fwd": (I, [P, I, I, I, C.POINTER(GruW), P, I, P, Z, P]),
    "desire_cvae_encode_workspace_bytes": (Z, [I, I]),
    "desire_cvae_encode_fwd": (I, [P, I, I, C.POINTER(CvaeEncW), P, P, Z, P]),
    "desire_reparam_fwd": (I, [P, P, I, I, I, P, P]),
    "desire_cvae_decode_workspace_bytes": (Z, [I, I]),
    "desire_cvae_decode_fwd": (I, [P, I, I, C.POINTER(CvaeDecW), P, P, Z, P]),
    "desire_mask_softmax_workspace_bytes": (Z, [I, I]),
    "desire_mask_softmax_fwd": (I, [P, I, I, I, I, P, P, P, I, P, P, Z, P]),
    "desire_gru_decode_workspace_bytes": (Z, [I, I]),
    "desire_gru_decode_fwd": (I, [P, P, I, I, I, I, I, C.POINTER(GruW), P, P, Z, P]),
    "desire_readout_pool_fwd": (I, [P, I, I, I, I, I, I, P, P, P, I, P, I, P, P, P]),
    "desire_kld_rows_fwd": (I, [P, I, I, P, P]),
    "desire_recon_rows_fwd": (I, [P, P, I, I, I, P, P]),
    "desire_masked_cost_fwd": (I, [P, P, P, I, I, P, P]),
    "desire_existence_fwd": (I, [P, P, I, I, I, I, P, P]),
    "desire_randn_fwd": (I, [P, P, Z, P]),
    "desire_scene_cnn_workspace_bytes": (Z, [I, I, I]),
    "desire_scene_cnn_fwd": (I, [P, I, I, I, I, C.POINTER(SceneCnnW), P, P, Z, P]),
    "desire_scene_gather_fwd": (I, [P, I, I, I, I, P, L, I, P, I, P]),
    "desire_social_pool_fwd": (I, [P, L, P, I, P, I, I, I, I, I, I, I, P, P, P, P]),
    "desire_social_fc_workspace_bytes": (C.c_size_t, [I, I]),
    "desire_social_fc_fwd": (I, [P, L, P, I, P, I, I, I, I, I, I, I, P, P, P, P, P, P, C.c_size_t, P]),
    "desire_ioc_workspace_bytes": (Z, [C.POINTER(IocDims)]),
    "desire_ioc_fwd": (I, [C.POINTER(IocDims), C.POINTER(IocW), P, P, I, P, I, P, P, P, P, Z, P]),
    "desire_ioc_factored_fwd": (I, [C.POINTER(IocDims), C.POINTER(IocW), P, P, I, P, I, P, P, P, P, P, P, Z, P]),
    # ---- train step
    "desire_cost_bwd": (I, [P, P, P, P, P, I, I, I, I, I, P, P, P]),
    "desire_readout_bwd": (I, [P, P, I, I, I, P, P, P, P, P]),
    "desire_gru_decode_bwd_workspace_bytes": (Z, [I, I, I]),
    "desire_gru_decode_bwd": (I, [P, P, I, I, I, I, I, C.POINTER(GruW), P, P, P, P, I, C.POINTER(GruG), P, Z, P]),
    "desire_mask_softmax_bwd_workspace_bytes": (Z, [I, I]),
    "desire_mask_softmax_bwd": (I, [P, I, I, I, I, P, P, P, I, P, P, P, I, P, P, P, Z, P]),
    "desire_cvae_decode_bwd_workspace_bytes": (Z, [I, I]),
    "desire_cvae_decode_bwd": (I, [P, I, I, C.POINTER(CvaeDecW), P, P, C.POINTER(CvaeDecG), P, Z, P]),
    "desire_reparam_bwd": (I, [P, P, P, I, I, I, P, P]),
    "desire_cvae_encode_bwd_workspace_bytes": (Z, [I, I]),
    "desire_cvae_encode_bwd": (I, [P, I, I, C.POINTER(CvaeEncW), P, P, C.POINTER(CvaeEncG), P, Z, P]),
    "desire_fc_bwd": (I, [P, I, P, I, P, I, P, I, I, I, I, I, P, I, I, P, I, P, P]),
    "desire_gru_encode_bwd_workspace_bytes": (Z, [I, I, I]),
    "desire_gru_encode_bwd": (I, [P, I, I, I, C.POINTER(GruW), P, I, C.POINTER(GruG), P, Z, P]),
    "desire_ioc_train_workspace_bytes": (Z, [C.POINTER(IocDims)]),
    "desire_ioc_train": (I, [C.POINTER(IocDims), C.POINTER(IocW), P, P, I, P, P, I, P, P, P, P, P, P, C.POINTER(IocG), P,
                             P, Z, P]),
    "desire_scene_cnn_bwd_workspace_bytes": (Z, [I, I, I]),
    "desire_scene_cnn_bwd": (I, [P, I, I, I, I, C.POINTER(SceneCnnW), P, C.POINTER(SceneCnnG), P, Z, P]),
    "desire_wgrad_workspace_bytes": (Z, [I, I]),
    "desire_wgrad_tn": (I, [P, I, P, I, P, I, I, I, I, P, Z, P]),
    "desire_deconv2d_workspace_bytes": (Z, [I, I, I, I, I, I, I]),
    "desire_deconv2d_fwd": (I, [P, I, I, I, P, I, I, I, I, P, P, P, I, P, P, Z, P]),
    "desire_sumsq_fwd": (I, [P, L, P, I, P]),
    "desire_adam_step": (I, [P, P, P, P, L, P, F, F, F, F, I, F, F, P]),
}

_lib = None


class DesireError(RuntimeError):
    pass


def load():
    """dlopen the kernel library (once).  Raises if it has not been built — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DesireError(
            "libdesire_b200.so is missing (%s). Build it with `python -m desire_b200.csrc.build`; "
            "there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.desire_version() != 1:
        raise DesireError("ABI version mismatch: library %d, binding 1" % lib.desire_version())
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().desire_last_error()
        raise DesireError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else ""))
