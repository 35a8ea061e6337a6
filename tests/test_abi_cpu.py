"""`-m "not gpu"`: the C-ABI library loads without a GPU and exports every symbol include/troy_b200.h declares;
host-side logic that needs no device (argument checks, NAF, key indices)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "troy_b200.h")).read()
    return sorted(set(re.findall(r"\b(troyb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from troy_b200 import lib as tlib
    L = tlib.Library(tlib.LIB_PATH)  # the CUDA build; dlopen only, no compute calls
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L.cdll, n), n
    assert set(names) == set(tlib.EXPORTED_SYMBOLS)


def test_calls_before_initialize_fail_loudly():
    from troy_b200 import lib as tlib
    import ctypes as C
    L = tlib.Library(tlib.LIB_PATH)
    if L.cdll.troyb_initialized():
        pytest.skip("already initialised in this process")
    p = C.c_void_p()
    with pytest.raises(tlib.TroyInvalidArgument, match="KernelProvider not initialized"):
        L.call("troyb_malloc", C.byref(p), 64, None)


def test_missing_library_is_an_import_error(tmp_path):
    from troy_b200 import lib as tlib
    with pytest.raises(ImportError, match="no CPU fallback"):
        tlib.Library(str(tmp_path / "libtroy_b200.so"))


def test_host_helpers():
    from troy_b200.host import GaloisKeysCuda, RelinKeysCuda, _naf
    assert _naf(7) == [-1, 8] and _naf(-7) == [1, -8] and _naf(12) == [-4, 16] and _naf(1) == [1]
    assert GaloisKeysCuda.get_index(3) == 1 and GaloisKeysCuda.get_index(15) == 7
    assert RelinKeysCuda.get_index(2) == 0
    with pytest.raises(ValueError):
        RelinKeysCuda.get_index(1)
