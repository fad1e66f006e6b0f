"""The C-ABI library builds, loads on a GPU-less host and exports every symbol the header declares.
No compute call is made here."""
import ctypes
import os
import subprocess

import pytest

from movedepth_b200 import _lib
from movedepth_b200.build import build, LIB


@pytest.fixture(scope="module")
def libpath():
    return build()


def test_library_builds_for_sm100a(libpath):
    assert os.path.exists(libpath) and libpath == LIB
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", libpath], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_every_declared_symbol_is_exported(libpath):
    declared = _lib.declared_symbols()
    assert len(declared) >= 12
    handle = ctypes.CDLL(libpath)
    missing = [s for s in declared if not hasattr(handle, s)]
    assert not missing, missing
    assert set(declared) <= set(_lib.SIGNATURES), set(declared) - set(_lib.SIGNATURES)


def test_loader_binds_and_reports_version(libpath):
    h = _lib.lib()
    assert h.mvd_version() == 1
    assert h.mvd_last_error_string() is not None


def test_tma_path_is_in_the_binary(libpath):
    """The cost-volume kernel stages tiles with TMA: the SASS must contain UTMALDG."""
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", libpath], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass


def test_ctypes_signatures_match_the_header_prototypes():
    """Arity and pointer-ness of every ctypes binding in _lib.SIGNATURES against the prototype in include/movedepth_b200.h
    (an ABI drift -- e.g. a parameter added in the .cu and the header but not in the loader -- would otherwise only show up
    as garbage arguments on a GPU box)."""
    import re
    text = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER).read(), flags=re.S)
    protos = dict(re.findall(r"\b(mvd_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S))
    assert set(protos) == set(_lib.declared_symbols())
    for name, params in protos.items():
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        argtypes = _lib.SIGNATURES[name][0]
        assert len(plist) == len(argtypes), (name, len(plist), len(argtypes), params)
        for p, t in zip(plist, argtypes):
            is_ptr = "*" in p
            bound_as_ptr = t in (ctypes.c_void_p, ctypes.c_char_p) or (isinstance(t, type) and issubclass(t, ctypes._Pointer))
            assert is_ptr == bound_as_ptr, (name, p, t)
