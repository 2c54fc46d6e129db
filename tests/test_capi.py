"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/kpreg_b200.h declares; the product package never touches the oracle."""
import ctypes
import os
import re
import subprocess

import pytest

import kpreg_b200
from kpreg_b200 import _lib
from conftest import ROOT

PKG = os.path.dirname(kpreg_b200.__file__)


def _declared():
    hdr = open(os.path.join(ROOT, "include", "kpreg_b200.h")).read()
    return sorted(set(re.findall(r"KPREG_API[^;(]*?\b(kpreg_\w+)\s*\(", hdr)))


def _ensure_built():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for must in ("kpreg_subsample_batch", "kpreg_grid_build", "kpreg_grid_query", "kpreg_kpconv_forward",
                 "kpreg_kpconv_backward", "kpreg_max_pool_forward", "kpreg_max_pool_backward", "kpreg_kabsch"):
        assert must in names
    assert sorted(_lib.EXPORTED_SYMBOLS) == names  # the ctypes binding covers exactly the header


def test_library_loads_and_exports_every_declared_symbol():
    _ensure_built()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert _lib.load().kpreg_version() >= 100
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (kpreg_\w+)", out))
    assert exported == set(_declared())  # nothing else leaks out of the shared object


def test_library_is_sm100a_only():
    _ensure_built()
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_size_queries_work_without_gpu():
    _ensure_built()
    assert _lib.size_query("kpreg_subsample_workspace_bytes", 100000, 4) > 100000 * 50
    assert _lib.size_query("kpreg_grid_workspace_bytes", 100000, 4) > 100000 * 16
    assert _lib.size_query("kpreg_kpconv_workspace_bytes", 1000, 1000, 15, 32, 32, 0) >= 1000 * 15 * 32 * 4


def test_tile_width_and_row_predicate_queries_without_gpu():
    """kpreg_linear_tile_cols / kpreg_segment_norm_rowpos_supported are pure host functions.  The in-place chain of the wide
    res2net units (res2net.py) relies on ONE output tile covering a whole group: w = 112 -> 128 columns, w = 224 -> 256 columns
    (fp16 split) for both the first layer (K = w) and the pair layers (K = pad32(w) + w)."""
    _ensure_built()
    lib = _lib.load()
    assert lib.kpreg_linear_tile_cols(64, 32) == 32 and lib.kpreg_linear_tile_cols(128, 64) == 64
    assert lib.kpreg_linear_tile_cols(112, 112) == 128 and lib.kpreg_linear_tile_cols(128 + 112, 112) == 128
    if os.environ.get("KPREG_GEMM_TF32", "")[:1] != "1" and os.environ.get("KPREG_GEMM_NO_N256", "")[:1] != "1":
        assert lib.kpreg_linear_tile_cols(224, 224) == 256 and lib.kpreg_linear_tile_cols(224 + 224, 224) == 256
        assert lib.kpreg_linear_tile_cols(512, 256) == 256
    assert lib.kpreg_linear_tile_cols(2048, 1024) == 128     # long reductions: 128-column tiles
    assert lib.kpreg_linear_tile_cols(0, 8) == 0
    assert [bool(lib.kpreg_segment_norm_rowpos_supported(c)) for c in (16, 32, 64, 128, 256)] == [False, True, True, False, False]


def test_chain_kernel_support_matrix_without_gpu():
    """kpreg_chain_supported / kpreg_chain_pack_bytes are pure host functions: the res2net widths of the shipped configs
    (w = floor(C * 14 / 64) for C = 64 .. 1024, 7 chained layers) map to the register-resident kernel up to w = 56."""
    _ensure_built()
    lib = _lib.load()
    want = {14: True, 28: True, 56: True, 112: False, 224: False}
    for w, ok in want.items():
        assert bool(lib.kpreg_chain_supported(w, 7)) is ok, w
    assert not lib.kpreg_chain_supported(27, 7) and not lib.kpreg_chain_supported(28, 0)
    # packed size = fragments (hi + lo, zero-padded to 8-channel tiles) + shifts, 256-byte granules
    for w in (14, 28, 56):
        nt = (w + 7) // 8
        floats = 7 * (nt * nt * 128 + 8 * nt)
        assert _lib.size_query("kpreg_chain_pack_bytes", w, 7) == (floats * 4 + 255) // 256 * 256
    assert _lib.size_query("kpreg_chain_pack_bytes", 56, 7) <= 200 * 1024  # resident in one CTA's shared memory
    with pytest.raises(RuntimeError):
        _lib.size_query("kpreg_chain_pack_bytes", 112, 7)


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: no product module may import, load or execute it."""
    offenders = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"kp_oracle|libkporacle|libkpref|oracle/", text):
                    offenders.append(f)
    assert offenders == []


def test_cpu_tensors_are_rejected_not_silently_computed():
    import torch
    from kpreg_b200 import ops
    with pytest.raises(RuntimeError):
        ops.subsample(torch.zeros(4, 3), torch.tensor([4], dtype=torch.int32), 0.1)
    with pytest.raises(RuntimeError):
        from kpreg_b200.se3_torch import compute_rigid_transform
        compute_rigid_transform(torch.zeros(4, 3), torch.zeros(4, 3))
