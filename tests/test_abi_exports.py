"""CPU: the C-ABI library loads and exports every symbol include/recom_b200.h declares."""

import ctypes
import os
import re

from gerrychain_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    from gerrychain_b200.build import build

    build()
    header = open(os.path.join(ROOT, "include", "recom_b200.h")).read()
    declared = set(re.findall(r"\b(rc_[a-z_]+)\s*\(", header))
    assert declared == set(_abi.EXPORTS)
    lib = ctypes.CDLL(_abi.lib_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert _abi.load().rc_abi_version() == 1


def test_struct_sizes_match_header():
    assert ctypes.sizeof(_abi.RcReplayStep) == 24
    assert ctypes.sizeof(_abi.RcGraph) == 8 + 5 * 8 + 8 + 2 * 8
    assert ctypes.sizeof(_abi.RcState) == 9 * 8 + 16 + 8
    assert ctypes.sizeof(_abi.RcConfig) == 112  # == sizeof(RcConfig) in include/recom_b200.h


def test_create_rejects_bad_arguments_without_gpu():
    lib = _abi.load()
    h = ctypes.c_void_p()
    rc = lib.rc_engine_create(None, None, 0, ctypes.byref(h))
    assert rc == _abi.RC_ERR_INVALID_ARG
    assert b"null" in lib.rc_last_error()
