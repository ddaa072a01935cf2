"""The C-ABI shared library loads and exports every symbol include/arcface_b200.h declares (no compute
without a GPU), and the ctypes table in multimodalsimilar_b200/_lib.py covers exactly that set."""
import ctypes
import os
import re

import pytest
import torch

from multimodalsimilar_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "arcface_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(arcface_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built():
    assert os.path.exists(_lib.LIB_PATH), "run `make -C multimodalsimilar_b200/csrc` (or __graft_entry__.build())"


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "header declares %s but the library does not export it" % n
    assert sorted(_lib.SIGNATURES) == names


def test_ctypes_arity_matches_header():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, params in re.findall(r"\b(arcface_b200_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert len(_lib.SIGNATURES[name][1]) == n, "%s: header has %d parameters, ctypes table %d" % (
            name, n, len(_lib.SIGNATURES[name][1]))


def test_version_and_error_string():
    lib = _lib.load()
    major, minor = ctypes.c_int32(-1), ctypes.c_int32(-1)
    assert lib.arcface_b200_version(ctypes.byref(major), ctypes.byref(minor)) == 0
    assert (major.value, minor.value) == (0, 1)
    assert isinstance(_lib.last_error(), str)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_no_device_is_an_error_not_a_fallback():
    lib = _lib.load()
    rc = lib.arcface_b200_device_ok()
    assert rc < 0
    assert _lib.last_error() != ""
    with pytest.raises(_lib.ArcfaceB200Error):
        _lib.call("arcface_b200_device_ok")
    n = ctypes.c_int32(0)
    assert lib.arcface_b200_forward_parts(512, 512, 1000000, ctypes.byref(n)) < 0  # needs the device's SM count


def test_missing_library_raises(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libarcface_b200.so")
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.load()
