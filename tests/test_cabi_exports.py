"""The C-ABI library must load (no GPU needed) and export every symbol include/*.h declares."""
import os
import re

from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "namaste_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nmb_[a-z0-9_]+)\s*\(", txt)))


def test_library_loads_and_exports_all_declared_symbols():
    from namaste_b200 import build, _lib
    build.build_library()
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 10
    for nm in names:
        assert hasattr(lib, nm), "symbol %s declared in include/namaste_b200.h is not exported" % nm
        assert nm in _lib.SIGNATURES, "symbol %s has no ctypes signature" % nm
    assert b"sm_100a" in lib.nmb_version()


def test_argument_errors_do_not_need_a_gpu():
    from namaste_b200 import _lib
    lib = _lib.load()
    assert lib.nmb_lnprob(None, None, 1, None, None, None, 0, 64, None) == _lib.NMB_ERR_ARG
    assert b"null" in lib.nmb_last_error()
