"""The C-ABI library loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

from unmicst_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "unmicst_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(umx_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in _declared_symbols():
        assert hasattr(L, name), name
    assert b"sm_100a" in L.umx_version()


def test_struct_sizes_match_header_layout():
    assert ctypes.sizeof(_lib.umx_model_desc) == 16 * 4
    assert ctypes.sizeof(_lib.umx_tensor) == 8 + 8 + 8 + 32
    assert ctypes.sizeof(_lib.umx_premap) == 48
    assert ctypes.sizeof(_lib.umx_opts) == 16 + 8 + 8 + 4 + 20
    assert ctypes.sizeof(_lib.umx_prof_entry) == 48 + 8 + 24


def test_device_count_is_never_an_error():
    assert _lib.lib().umx_device_count() >= 0
