"""ctypes binding of libvodb.so (the C ABI declared in include/vodb.h).

The CUDA library is the product: there is no CPU fallback. If the shared library is missing or
cannot be loaded, every entry point raises `VodbUnavailableError` (loudly, at first use).
"""
from __future__ import annotations

import ctypes
import pathlib

_PKG = pathlib.Path(__file__).resolve().parent
LIB_PATH = _PKG / "libvodb.so"

F32, BF16, F16 = 0, 1, 2
MODE_EXACT, MODE_TENSOR, MODE_TENSOR_X2, MODE_TENSOR_X3 = 0, 1, 2, 3
QUIRK_INVERTED_SUPPORT = 1
MAX_K = 2048
DTYPE_NAMES = {"float32": F32, "f32": F32, "fp32": F32, "bfloat16": BF16, "bf16": BF16, "float16": F16, "f16": F16,
               "fp16": F16}


class VodbError(RuntimeError):
    """An error reported by libvodb.so (negative return code + vodb_last_error())."""


class VodbUnavailableError(VodbError):
    """libvodb.so is not built / not loadable, or no CUDA device is visible."""


_c = ctypes
_vp = _c.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "vodb_last_error": (_c.c_char_p, []),
    "vodb_abi_version": (_c.c_int, []),
    "vodb_device_count": (_c.c_int, []),
    "vodb_store_create": (_c.c_int, [_c.POINTER(_vp), _c.c_int, _c.c_int64, _c.c_int, _c.c_int, _c.c_int64]),
    "vodb_store_destroy": (None, [_vp]),
    "vodb_store_add": (_c.c_int, [_vp, _vp, _c.c_int, _c.c_int, _c.c_int64, _c.c_int64, _vp]),
    "vodb_store_fill_synthetic": (_c.c_int, [_vp, _c.c_uint64, _c.c_int64, _c.c_int64, _c.c_int, _vp]),
    "vodb_store_read": (_c.c_int, [_vp, _c.c_int64, _c.c_int64, _vp, _c.c_int, _vp]),
    "vodb_store_ntotal": (_c.c_int64, [_vp]),
    "vodb_store_dim": (_c.c_int, [_vp]),
    "vodb_store_dtype": (_c.c_int, [_vp]),
    "vodb_store_device": (_c.c_int, [_vp]),
    "vodb_store_bytes": (_c.c_int64, [_vp]),
    "vodb_search": (_c.c_int, [_vp, _vp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _vp, _vp, _c.c_int, _vp]),
    "vodb_store_prepare_tensor": (_c.c_int, [_vp, _vp]),
    "vodb_search_check": (_c.c_int, [_vp, _vp]),
    "vodb_search_stats": (_c.c_int, [_vp, _c.POINTER(_c.c_int64)]),
    "vodb_store_set_profiling": (_c.c_int, [_vp, _c.c_int]),
    "vodb_store_profile": (_c.c_int, [_vp, _c.POINTER(_c.c_double)]),
    "vodb_xchg_create": (_c.c_int, [_c.POINTER(_vp), _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _vp]),
    "vodb_xchg_connect": (_c.c_int, [_vp, _vp]),
    "vodb_xchg_destroy": (None, [_vp]),
    "vodb_search_sharded": (_c.c_int, [_vp, _vp, _vp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _vp,
                                       _vp, _c.c_int, _vp]),
    "vodb_merge_topk": (_c.c_int, [_c.c_int, _vp, _vp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _vp, _vp, _c.c_int, _vp]),
    "vodb_merge_results": (_c.c_int, [_c.c_int, _c.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _c.c_int, _c.c_int, _c.c_int,
                                      _c.c_double, _c.c_int, _c.c_int, _vp, _vp, _vp, _vp, _vp, _c.c_int, _vp]),
    "vodb_sample": (_c.c_int, [_c.c_int, _vp, _vp, _vp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_float,
                               _c.c_int, _c.c_int, _c.c_uint64, _c.c_uint64, _vp, _vp, _vp, _vp, _c.c_int, _vp]),
    "vodb_sample_results": (_c.c_int, [_c.c_int, _vp, _vp, _vp, _vp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_float,
                                       _c.c_int, _c.c_int, _c.c_uint64, _c.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                       _vp]),
    "vodb_plan_scan": (_c.c_int, [_c.c_int64, _c.c_int, _c.c_int, _c.c_int, _vp, _vp, _c.c_int]),
    "vodb_retrieve_sample": (_c.c_int, [_vp, _vp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _vp, _c.c_int,
                                        _c.c_int, _c.c_int, _c.c_float, _c.c_int, _c.c_int, _c.c_uint64, _c.c_uint64,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib: ctypes.CDLL | None = None


def exported_symbols() -> list[str]:
    """Every symbol include/vodb.h declares (kept in sync by tests/test_abi.py)."""
    return sorted(_SIGNATURES)


def load() -> ctypes.CDLL:
    """Load libvodb.so and set the prototypes. Raises VodbUnavailableError when it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise VodbUnavailableError(
            f"{LIB_PATH} is missing: build it with `python -m vod_b200.build` (nvcc, sm_100a). "
            "vod_b200 has no CPU fallback."
        )
    try:
        lib = ctypes.CDLL(str(LIB_PATH))
    except OSError as exc:  # e.g. libcudart not found
        raise VodbUnavailableError(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.vodb_abi_version() != 1:
        raise VodbUnavailableError(f"libvodb.so ABI version {lib.vodb_abi_version()} != 1")
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().vodb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> None:
    if rc < 0:
        raise VodbError(f"{what} failed (code {rc}): {last_error()}")


def require_gpu() -> int:
    """Number of visible CUDA devices; raises VodbUnavailableError if there is none."""
    lib = load()
    n = lib.vodb_device_count()
    if n <= 0:
        raise VodbUnavailableError(f"no CUDA device visible to libvodb.so ({last_error() or 'device count 0'})")
    return n
