"""CPU: the C-ABI library builds, loads and exports every symbol declared in include/medsam2_b200.h
(no compute calls without a GPU), and argument validation returns errors instead of launching."""
import ctypes
import os

import pytest


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from medsam2_b200 import native
    protos = native.parse_header()
    assert len(protos) >= 29
    lib = ctypes.CDLL(native.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert native.lib().ms2_version() >= 100


def test_argument_validation_without_gpu():
    from medsam2_b200 import native
    l = native.lib()
    # odd height is rejected before any launch (mirrors the reference's AT_ASSERTM, connected_components.cu:225-227)
    rc = l.ms2_cc_label(ctypes.c_void_p(8), ctypes.c_void_p(8), ctypes.c_void_p(8), None, 1, 5, 4, None)
    assert rc != 0 and b"even" in l.ms2_last_error()
    rc = l.ms2_gemm(None, 0, 1, None, 0, None, None, None, 0, None, 0, 1, 1, 1, 1, 0, 0, None)
    assert rc != 0


def test_header_cites_reference_boundary():
    from medsam2_b200 import native
    src = open(native.header_path()).read()
    assert "connected_components.cu" in src and "get_connected_componnets" in src
