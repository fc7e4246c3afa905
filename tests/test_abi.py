"""CPU: the C-ABI library loads and exports exactly the entry points include/desire_abi.h declares,
and the ctypes binding covers all of them (no compute calls — there is no GPU here)."""
import os
import re
import subprocess

from desire_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "desire_abi.h")


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(desire_[a-z0-9_]+)\s*\(", src))


def test_header_symbols_are_exported_and_bound():
    names = declared()
    assert {"desire_version", "desire_gru_decode_fwd", "desire_ioc_fwd", "desire_social_pool_fwd",
            "desire_scene_gather_fwd", "desire_cvae_decode_fwd"} <= names
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (desire_[a-z0-9_]+)", out))
    assert names <= exported, names - exported
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)


def test_library_loads_and_reports_version():
    lib = _lib.load()
    assert lib.desire_version() == 1
    assert lib.desire_launch_count() >= 0


def test_workspace_queries_need_no_gpu():
    lib = _lib.load()
    assert lib.desire_cvae_decode_workspace_bytes(100, 128) >= 100 * 51200 * 4
    assert lib.desire_gru_decode_workspace_bytes(64, 128) >= 64 * 3 * 128 * 4
    d = _lib.IocDims(2, 8, 3, 48, 12, 100, 16, 32, 6, 6, 16, 16, 2)
    import ctypes as C
    assert lib.desire_ioc_workspace_bytes(C.byref(d)) > 0


def test_bad_arguments_are_rejected_with_a_message():
    lib = _lib.load()
    rc = lib.desire_fc_fwd(None, 4, None, 4, None, None, 4, 1, 4, 4, 0, 0, None)
    assert rc == -1
    assert b"desire_fc_fwd" in lib.desire_last_error()
    rc = lib.desire_reparam_fwd(None, None, 1, 1, 8, None, None)
    assert rc == -1


def test_only_sm100a_code_in_the_library():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
