"""CPU: the C-ABI library loads and exports every symbol include/mkssd_b200.h declares; without a
GPU it fails loudly (no CPU fallback, nothing routed through oracle/)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "mkssd_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(mk_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_exported(lib_built):
    L = lib_built.load()
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(lib_built.EXPORTS) == names, "metakssd_b200.EXPORTS out of sync with the header"


def test_error_strings(lib_built):
    L = lib_built.load()
    assert L.mk_strerror(0) == b"ok"
    assert b"crowd" in L.mk_strerror(-5)
    assert b"no CPU fallback" in L.mk_strerror(-3)


def test_struct_layouts_match_header(lib_built):
    assert C.sizeof(lib_built.MkInfo) == 13 * 4
    assert C.sizeof(lib_built.api.MkSketch) == 48          # (+ `borrowed` since round 2)
    assert C.sizeof(lib_built.api.MkProfile) == 10 * 8
    assert C.sizeof(lib_built.api.MkSpeciesStat) == 24
    assert C.sizeof(lib_built.MksParams) == 40


def test_no_silent_fallback_without_gpu(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU box")
    assert lib_built.device_count() == 0
    sid, perm = lib_built.make_shuf(1, 4)
    with pytest.raises(lib_built.MkError) as ei:
        lib_built.Sketcher(perm, 9, 4, 1)
    assert ei.value.code == -3          # MK_ERR_CUDA


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "metakssd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "liboracle" not in src and "kssd_oracle" not in src, f


def test_parameter_validation_matches_reference_ranges(oracle, lib_built):
    """get_hashsz() (command_dist.c:286-315) accepts 4(k-L)-15 in 0..24; both sides agree."""
    for k, subk, L in [(11, 6, 3), (11, 5, 2), (10, 6, 3), (9, 4, 1), (7, 6, 3), (13, 6, 4)]:
        p = oracle.params(k, subk, L)
        assert p.hashsize > 0 and p.hashlimit == int(p.hashsize * 0.6)
    for k, subk, L in [(6, 6, 3), (16, 7, 4)]:
        with pytest.raises(ValueError):
            oracle.params(k, subk, L)
