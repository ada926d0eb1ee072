"""The C-ABI library loads (no GPU needed) and exports every symbol include/univtg_b200.h declares."""
import ctypes
import os
import re

from univtg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "univtg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(univtg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/univtg_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in univtg_b200/_lib.py"


def test_abi_version_and_sizes():
    lib = _lib.load_library()
    assert lib.univtg_abi_version() == 1
    cfg = _lib.Config(1024, 8, 1024, 4, 2, 2818, 512, 0)
    assert lib.univtg_num_params(ctypes.byref(cfg)) == 8 * 2 + 1 + 12 * 4 + 12 + 1
    pb = lib.univtg_packed_bytes(ctypes.byref(cfg))
    # 16-bit copies of every GEMM weight (K padded to 64) + fp32 vectors: between 2 and 2.3 bytes per used parameter
    assert 2.0 * 43.3e6 < pb < 2.3 * 43.3e6
    shp = _lib.Shape(32, 75, 32, 0)
    assert lib.univtg_workspace_bytes(ctypes.byref(cfg), ctypes.byref(shp)) > 50e6


def test_bad_config_is_rejected_with_message():
    lib = _lib.load_library()
    cfg = _lib.Config(1000, 8, 1024, 4, 2, 2818, 512, 0)
    assert lib.univtg_packed_bytes(ctypes.byref(cfg)) == 0
    assert "hidden_dim" in _lib.last_error()
