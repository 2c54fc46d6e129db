"""ctypes binding of ``libkpreg_b200.so`` (C ABI declared in ``include/kpreg_b200.h``).

The product path has no CPU fallback: if the CUDA library is missing or a call is made without a
CUDA device, this module raises.  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KPREG_B200_LIB") or os.path.join(_HERE, "libkpreg_b200.so")  # override: development A/B builds

_c_i64 = ctypes.c_int64
_c_int = ctypes.c_int
_c_f32 = ctypes.c_float
_c_ptr = ctypes.c_void_p
_c_size = ctypes.c_size_t

_SIGNATURES = {
    "kpreg_version": (_c_int, []),
    "kpreg_last_error": (ctypes.c_char_p, []),
    "kpreg_launch_count": (ctypes.c_ulonglong, []),
    "kpreg_profile": (_c_int, [_c_int]),
    "kpreg_profile_reserve": (_c_int, [_c_int]),
    "kpreg_profile_read": (_c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_ulonglong)]),
    "kpreg_subsample_workspace_bytes": (_c_int, [_c_i64, _c_int, ctypes.POINTER(_c_size)]),
    "kpreg_subsample_batch": (_c_int, [_c_ptr, _c_ptr, _c_i64, _c_int, _c_f32, _c_int, _c_ptr, _c_ptr,
                                       _c_ptr, _c_size, _c_ptr]),
    "kpreg_grid_workspace_bytes": (_c_int, [_c_i64, _c_int, ctypes.POINTER(_c_size)]),
    "kpreg_grid_build": (_c_int, [_c_ptr, _c_ptr, _c_i64, _c_int, _c_f32, _c_ptr, _c_size, _c_ptr, _c_ptr]),
    "kpreg_grid_query": (_c_int, [_c_ptr, _c_i64, _c_int, _c_ptr, _c_ptr, _c_i64, _c_f32, _c_int, _c_int,
                                  _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr]),
    "kpreg_pack_rows": (_c_int, [_c_ptr, _c_i64, _c_int, _c_int, _c_int, _c_ptr, _c_ptr]),
    "kpreg_kpconv_workspace_bytes": (_c_int, [_c_i64, _c_i64, _c_int, _c_int, _c_int, _c_int,
                                              ctypes.POINTER(_c_size)]),
    "kpreg_kpconv_forward": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i64,
                                      _c_int, _c_int, _c_int, _c_int, _c_f32, _c_int, _c_int, _c_int, _c_ptr, _c_ptr,
                                      _c_ptr, _c_size, _c_ptr]),
    "kpreg_kpconv_forward_rowpos": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i64,
                                             _c_int, _c_int, _c_int, _c_int, _c_f32, _c_int, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr,
                                             _c_ptr, _c_size, _c_ptr]),
    "kpreg_kpconv_backward": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64,
                                       _c_i64, _c_int, _c_int, _c_int, _c_int, _c_f32, _c_int, _c_int, _c_ptr, _c_ptr,
                                       _c_ptr, _c_ptr, _c_size, _c_ptr]),
    "kpreg_max_pool_forward": (_c_int, [_c_ptr, _c_ptr, _c_int, _c_i64, _c_i64, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr,
                                        _c_ptr]),
    "kpreg_max_pool_backward": (_c_int, [_c_ptr, _c_ptr, _c_i64, _c_i64, _c_int, _c_ptr, _c_ptr]),
    "kpreg_split_weights": (_c_int, [_c_ptr, _c_int, _c_int, _c_int, _c_ptr, _c_size, _c_ptr]),
    "kpreg_gemm_supported": (_c_int, [_c_i64, _c_int, _c_int, _c_int, _c_ptr]),
    "kpreg_linear_workspace_bytes": (_c_int, [_c_int, _c_int, ctypes.POINTER(_c_size)]),
    "kpreg_linear_forward": (_c_int, [_c_ptr, _c_int, _c_ptr, _c_i64, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_int,
                                      _c_f32, _c_ptr, _c_int, _c_ptr, _c_int, _c_ptr, _c_int, _c_ptr, _c_int, _c_int, _c_int, _c_ptr,
                                      _c_size, _c_ptr]),
    "kpreg_linear_backward_workspace_bytes": (_c_int, [_c_i64, _c_int, _c_int, ctypes.POINTER(_c_size)]),
    "kpreg_linear_backward": (_c_int, [_c_ptr, _c_int, _c_ptr, _c_int, _c_ptr, _c_i64, _c_int, _c_int, _c_ptr, _c_int, _c_ptr, _c_ptr,
                                       _c_size, _c_ptr]),
    "kpreg_linear_tile_cols": (_c_int, [_c_int, _c_int]),
    "kpreg_linear_pair_forward": (_c_int, [_c_ptr, _c_int, _c_int, _c_ptr, _c_int, _c_int, _c_ptr, _c_i64, _c_int, _c_ptr, _c_ptr, _c_int,
                                           _c_f32, _c_ptr, _c_int, _c_int, _c_ptr, _c_int, _c_ptr]),
    "kpreg_segment_norm_workspace_bytes": (_c_int, [_c_int, _c_int, ctypes.POINTER(_c_size)]),
    "kpreg_segment_norm_forward": (_c_int, [_c_ptr, _c_int, _c_ptr, _c_int, _c_i64, _c_int, _c_f32, _c_ptr, _c_int, _c_int,
                                            _c_f32, _c_ptr, _c_int, _c_ptr, _c_size, _c_ptr]),
    "kpreg_segment_norm_rowpos_supported": (_c_int, [_c_int]),
    "kpreg_segment_norm_forward_rowpos": (_c_int, [_c_ptr, _c_int, _c_ptr, _c_int, _c_i64, _c_int, _c_f32, _c_ptr, _c_int, _c_int,
                                                   _c_f32, _c_ptr, _c_int, _c_ptr, _c_ptr, _c_size, _c_ptr]),
    "kpreg_segment_norm_backward": (_c_int, [_c_ptr, _c_int, _c_ptr, _c_int, _c_ptr, _c_int, _c_i64, _c_int, _c_f32, _c_ptr, _c_int,
                                             _c_ptr, _c_size, _c_ptr]),
    "kpreg_chain_supported": (_c_int, [_c_int, _c_int]),
    "kpreg_chain_pack_bytes": (_c_int, [_c_int, _c_int, ctypes.POINTER(_c_size)]),
    "kpreg_chain_pack": (_c_int, [_c_ptr, _c_ptr, _c_int, _c_int, _c_ptr, _c_size, _c_ptr]),
    "kpreg_chain_forward": (_c_int, [_c_ptr, _c_int, _c_ptr, _c_int, _c_int, _c_i64, _c_ptr, _c_int, _c_ptr, _c_int, _c_int,
                                     _c_ptr]),
    "kpreg_front_supported": (_c_int, [_c_int, _c_int, _c_int]),
    "kpreg_front_pack_bytes": (_c_int, [_c_int, _c_int, _c_int, ctypes.POINTER(_c_size)]),
    "kpreg_front_pack": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_ptr, _c_size, _c_ptr]),
    "kpreg_front_forward": (_c_int, [_c_ptr, _c_int, _c_int, _c_ptr, _c_int, _c_int, _c_i64, _c_ptr, _c_int, _c_int, _c_ptr]),
    "kpreg_kabsch": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_f32, _c_int, _c_ptr, _c_ptr]),
    "kpreg_overlap_pool": (_c_int, [_c_ptr, _c_ptr, _c_int, _c_i64, _c_i64, _c_int, _c_ptr, _c_ptr]),
    "kpreg_sine_embed": (_c_int, [_c_ptr, _c_i64, _c_int, _c_int, _c_int, _c_f32, _c_ptr, _c_ptr, _c_ptr]),
    "kpreg_pack_coarse_workspace_bytes": (_c_int, [_c_int, ctypes.POINTER(_c_size)]),
    "kpreg_pack_coarse": (_c_int, [_c_ptr, _c_int, _c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_f32, _c_ptr, _c_int, _c_int,
                                   _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_size, _c_ptr]),
    "kpreg_shuffle_gather": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr]),
    "kpreg_remap_pairs": (_c_int, [_c_ptr, _c_i64, _c_ptr, _c_i64, _c_ptr, _c_i64, _c_ptr, _c_ptr, _c_ptr]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_ERRORS = {1: "invalid argument", 2: "workspace too small", 3: "CUDA error", 4: "grid too large to index"}

_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """Load the CUDA library (once).  Raises ImportError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C <package>/csrc`.  kpreg_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    """Raise RuntimeError for a non-zero return code (the reference wrappers' error convention)."""
    if rc != 0:
        detail = _ERRORS.get(rc, f"error {rc}")
        if rc == 3:
            detail += ": " + load().kpreg_last_error().decode(errors="replace")
        raise RuntimeError(f"{what}: {detail}")


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: kpreg_b200 has no CPU path")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr(device: torch.device) -> int:
    """cudaStream_t of torch's current stream on ``device`` (the raw getter avoids building a Stream object)."""
    if _raw_stream is not None:
        return _raw_stream(device.index if device.index is not None else torch.cuda.current_device())
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(load().kpreg_launch_count())


FAMILIES = ("subsample", "grid_build", "grid_query", "kpconv_gather", "kpconv_contract", "max_pool", "kabsch", "other",
            "linear", "segment_norm")


def profile(enable: bool) -> None:
    """Start (and clear) / stop per-kernel-family device timing."""
    check(load().kpreg_profile(1 if enable else 0), "kpreg_profile")


def profile_reserve(n_records: int) -> None:
    """Pre-create the CUDA events of ``n_records`` timed scopes (so that a measured region creates none)."""
    check(load().kpreg_profile_reserve(int(n_records)), "kpreg_profile_reserve")


def profile_read() -> Dict[str, Tuple[float, int]]:
    """{family: (summed device ms, timed launches)} since the last profile(True)."""
    ms = (ctypes.c_double * len(FAMILIES))()
    cnt = (ctypes.c_ulonglong * len(FAMILIES))()
    check(load().kpreg_profile_read(ms, cnt), "kpreg_profile_read")
    return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(FAMILIES)}


class _Workspaces:
    """One growing scratch buffer per (device, stream): kernels on a stream are ordered, so the
    buffer can be reused call after call without extra synchronisation."""

    def __init__(self) -> None:
        self._bufs: Dict[Tuple[int, int], torch.Tensor] = {}

    def get(self, nbytes: int, device: torch.device) -> torch.Tensor:
        key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr(device))
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            # keep the old buffer alive until queued kernels are done: record it on the stream
            if buf is not None:
                buf.record_stream(torch.cuda.current_stream(device))
            buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf


workspaces = _Workspaces()


_SIZE_CACHE: Dict[tuple, int] = {}


def size_query(fn_name: str, *args) -> int:
    """Workspace size from a kpreg_*_workspace_bytes entry point (pure functions of their arguments: memoised)."""
    key = (fn_name,) + args
    hit = _SIZE_CACHE.get(key)
    if hit is None:
        out = _c_size(0)
        check(getattr(load(), fn_name)(*args, ctypes.byref(out)), fn_name)
        hit = int(out.value)
        if len(_SIZE_CACHE) < 65536:
            _SIZE_CACHE[key] = hit
    return hit
