"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/camus_b200.h
declares, and fails loudly (no fallback) when there is no GPU. No compute calls here."""
import ctypes as C
import os
import re

import pytest

from camus_b200 import build, cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_cuda()
    return cabi.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "camus_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(camus_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in camus_b200.h but not exported"
    assert sorted(cabi.EXPORTED) == names


def test_create_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.camus_b200_create(C.byref(h), 1, None)
    assert rc != 0 and not h.value
    msg = lib.camus_b200_last_error(None).decode()
    assert "no CPU fallback" in msg
    with pytest.raises(cabi.CamusB200Error):
        cabi.CamusB200(0)


def test_create_argument_checks(lib):
    """n_gpus >= 1, no device listed twice (checked before any CUDA call); a multi-GPU handle on a
    box without GPUs fails like the single one."""
    h = C.c_void_p()
    assert lib.camus_b200_create(C.byref(h), 0, None) == 1 and not h.value
    assert lib.camus_b200_create(None, 1, None) == 1
    twice = (C.c_int * 2)(0, 0)
    assert lib.camus_b200_create(C.byref(h), 2, twice) == 1 and not h.value
    assert "twice" in lib.camus_b200_last_error(None).decode()
    import torch
    if not torch.cuda.is_available():
        assert lib.camus_b200_create(C.byref(h), 2, None) != 0 and not h.value


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under camus_b200/ or include/ may reference it."""
    bad = []
    for base in ("camus_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"^\s*(from|import)\s+oracle\b|#include\s+\".*oracle|liboracle", txt, flags=re.M):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_header_is_plain_c():
    """The boundary is a C ABI (cgo binds it): include/camus_b200.h must compile as C99 with no torch / C++ types."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "camus_b200.h")
    subprocess.run(["gcc", "-x", "c", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", hdr], check=True)
    text = open(hdr).read()
    assert "torch" not in text and "std::" not in text and "at::" not in text


def test_dp_and_ingest_entry_points_fail_loudly_without_a_gpu(lib):
    """camus_b200_dp_create has no CPU fallback either; the null-argument checks come before any CUDA call."""
    import numpy as np
    import torch
    lib.camus_b200_dp_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.camus_b200_dp_last_error.argtypes = [C.c_void_p]
    lib.camus_b200_dp_last_error.restype = C.c_char_p
    h = C.c_void_p()
    assert lib.camus_b200_dp_create(C.byref(h), 0, 0, 0, None, None, None) == 1 and not h.value
    assert b"dp_create" in lib.camus_b200_dp_last_error(None)
    if not torch.cuda.is_available():
        end, dep, edge = np.array([3, 2, 3], np.int32), np.array([0, 1, 1], np.int32), np.zeros(9, np.uint64)
        rc = lib.camus_b200_dp_create(C.byref(h), 0, 0, 3, end.ctypes.data, dep.ctypes.data, edge.ctypes.data)
        assert rc != 0 and not h.value and b"no CPU fallback" in lib.camus_b200_dp_last_error(None)
    assert lib.camus_b200_add_gene_trees_newick(None, b"x", 1, 0.0, None, None, None) == 1
    assert lib.camus_b200_set_tip_labels(None, 0, None, None) == 1
