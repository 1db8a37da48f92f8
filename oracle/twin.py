"""oracle/twin.py — ctypes binding of oracle/libvodb_twin.so. TEST INFRASTRUCTURE ONLY.

`sample()` is the bit-exact CPU twin of `vodb_sample` (see oracle/sample_twin.c for the
reference lines it restates); `synth_rows()` is the CPU side of `vodb_store_fill_synthetic`.
"""
from __future__ import annotations

import ctypes
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_LIB: ctypes.CDLL | None = None


def build(force: bool = False) -> pathlib.Path:
    so = _HERE / "libvodb_twin.so"
    src = _HERE / "sample_twin.c"
    hdr = _HERE.parent / "vod_b200" / "csrc" / "vodb_math.h"
    stale = (not so.exists()) or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime)
    if force or stale:
        subprocess.run(["make", "-C", str(_HERE), "-B", "libvodb_twin.so"], check=True, capture_output=True)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(str(build()))
        f32p = ctypes.POINTER(ctypes.c_float)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.twin_sample.restype = ctypes.c_int
        L.twin_sample.argtypes = [f32p, u8p, f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                                  ctypes.c_uint64, i64p, f32p, u8p, f32p]
        for name in ("twin_logf", "twin_expf", "twin_log1pf"):
            getattr(L, name).restype = ctypes.c_float
            getattr(L, name).argtypes = [ctypes.c_float]
        L.twin_exp1_noise.restype = ctypes.c_float
        L.twin_exp1_noise.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32]
        L.twin_synth_rows.restype = None
        L.twin_synth_rows.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, f32p]
        L.twin_philox.restype = None
        L.twin_philox.argtypes = [ctypes.c_uint32] * 6 + [ctypes.POINTER(ctypes.c_uint32)]
        _LIB = L
    return _LIB


def _p(a: np.ndarray | None, ty):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ty))


def sample(scores, labels, *, k_positive, k_total, normalized=True, temperature=1.0, max_support=-1,
           quirks=1, noise=None, seed=0, offset=0):
    """CPU twin of vodb_sample. Returns (ids i64[B,k], logw f32[B,k], labels bool[B,k], lse f32[B,2])."""
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    assert scores.ndim == 2
    B, K = scores.shape
    lab = None if labels is None else np.ascontiguousarray(np.asarray(labels) > 0, dtype=np.uint8)
    nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float32)
    ids = np.empty((B, k_total), np.int64)
    logw = np.empty((B, k_total), np.float32)
    olab = np.empty((B, k_total), np.uint8)
    lse = np.zeros((B, 2), np.float32)
    rc = lib().twin_sample(_p(scores, ctypes.c_float), _p(lab, ctypes.c_uint8), _p(nz, ctypes.c_float), B, K,
                           int(k_positive), int(k_total), int(bool(normalized)), float(temperature),
                           int(max_support if max_support else -1), int(quirks), int(seed), int(offset),
                           _p(ids, ctypes.c_int64), _p(logw, ctypes.c_float), _p(olab, ctypes.c_uint8),
                           _p(lse, ctypes.c_float))
    if rc != 0:
        raise ValueError("twin_sample: bad arguments")
    return ids, logw, olab.astype(np.bool_), lse


def exp1_noise(seed: int, offset: int, B: int, K: int) -> np.ndarray:
    L = lib()
    out = np.empty((B, K), np.float32)
    for b in range(B):
        for j in range(K):
            out[b, j] = L.twin_exp1_noise(seed, offset, b, j)
    return out


def synth_rows(seed: int, row0: int, n: int, dim: int, dtype: int = 0, unit_norm: bool = False) -> np.ndarray:
    """float32 [n, dim] synthetic embeddings, already rounded to dtype (0=f32, 1=bf16, 2=f16)."""
    out = np.empty((n, dim), np.float32)
    lib().twin_synth_rows(int(seed), int(row0), int(n), int(dim), int(dtype), int(bool(unit_norm)),
                          _p(out, ctypes.c_float))
    return out


def unary(name: str, x: np.ndarray) -> np.ndarray:
    f = getattr(lib(), f"twin_{name}")
    x = np.asarray(x, np.float32)
    return np.array([f(float(v)) for v in x.ravel()], np.float32).reshape(x.shape)
