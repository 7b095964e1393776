"""The C-ABI library: it builds, loads, exports every symbol include/b200_ldu.h declares, and fails
loudly without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from multiregionfoam_b200 import build, ldu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build_library()


def test_header_symbols_are_exported(lib_path):
    hdr = open(os.path.join(ROOT, "include", "b200_ldu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(b200_[a-z_0-9A-Z]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    L = C.CDLL(lib_path)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(ldu.ABI_SYMBOLS) == declared


def test_block_header_symbols_are_exported(lib_path):
    """include/b200_blk.h: the block-coupled (vector4) entry points."""
    from multiregionfoam_b200 import blockldu
    hdr = open(os.path.join(ROOT, "include", "b200_blk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(b200_blk_[a-z_0-9A-Z]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    L = C.CDLL(lib_path)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(blockldu.ABI_SYMBOLS) == declared


def test_smoother_header_symbols_are_exported(lib_path):
    """include/b200_smooth.h: the Gauss-Seidel smoother entry points."""
    from multiregionfoam_b200 import smoother
    hdr = open(os.path.join(ROOT, "include", "b200_smooth.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(b200_gs_[a-z_0-9A-Z]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    L = C.CDLL(lib_path)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(smoother.ABI_SYMBOLS) == declared


def test_version_and_error_text(lib_path):
    L = ldu.load()
    assert L.b200_version() >= 100


def test_no_cpu_fallback(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(ldu.B200Error) as e:
        ldu.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multiregionfoam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                bad = [ln for ln in txt.splitlines() if re.search(r"^\s*(#include|import|from)\b.*oracle", ln) or "dlopen" in ln and "oracle" in ln]
                assert not bad, (f, bad)
                assert "pyoracle" not in txt and "ldu_oracle" not in txt and "pyblk" not in txt, f
