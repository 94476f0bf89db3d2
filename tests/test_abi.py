"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares.
No compute call is made here (there is no GPU in this container) -- and creating a context
without a GPU must fail loudly, never fall back."""
import ctypes
import os
import re

import pytest

from helpers import ROOT


def _lib():
    from orb_slam3_comments_ghr_b200 import matcher
    if not os.path.exists(matcher.lib_path()):
        import __graft_entry__ as g
        g.build()
    return matcher.load_library()


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "orbmatch_b200.h")).read()
    names = sorted(set(re.findall(r"\b(orbgpu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    L = _lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared but not exported: {missing}"


def test_version_and_no_silent_fallback():
    L = _lib()
    assert b"sm_100a" in L.orbgpu_version()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        h = ctypes.c_void_p()
        rc = L.orbgpu_create(0, ctypes.byref(h))
        assert rc != 0, "orbgpu_create must fail without a GPU (no CPU fallback)"
        assert b"fallback" in L.orbgpu_last_error() or b"CUDA" in L.orbgpu_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "orb_slam3_comments_ghr_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cc")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "pyoracle" not in src and "orb_oracle" not in src and "libref_orbmatcher" not in src, fn
