"""Labeled priority sampling of k passages from the retrieved top-K — the dataloader step after retrieval.

Mirrors the reference's sampling interface, same names / argument meaning / return types:

    sample_search_results(*, search_results, raw_scores, total, max_pos_sections, temperature, max_support_size)
        -> PrioritySampledSections                     src/vod_dataloaders/core/sample.py:22-84
    labeled_priority_sampling(scores, labels, k_positive, k_total, normalized, temperature, max_support_size)
        -> (samples i64, log_weights, labels bool, lse)  src/vod_dataloaders/core/sample.py:87-157
    priority_sampling_1d(scores, k, temperature, max_support_size) -> (ids, log_weights)   sample.py:222-242

The per-row work (log-softmax, Exp(1) priority keys, top-(k+1), importance weights, per-label
self-normalisation) runs in one fused CUDA kernel (`vodb_sample`, vod_b200/csrc/sample.cu). Differences from
the reference, all opt-in or documented:
  * noise is counter-based (Philox-4x32-10) instead of the global `np.random` state; the Philox seed is drawn
    from `np.random` when `seed` is not given, so `np.random.seed(...)` still makes runs reproducible
    (reference: sample.py:398 draws `np.random.exponential`). An explicit `noise` array can be passed instead;
  * the arithmetic is float32 on the device; float64 inputs are accepted and the outputs cast back;
  * `fix_truncation=True` applies the intended top-`max_support_size` truncation; the default reproduces the
    reference literally (it masks the top entries OUT, sample.py:176-178).
"""
from __future__ import annotations

import dataclasses
import typing as typ

import numpy as np

from . import _lib
from .retrieval import RetrievalBatch
from .search import _current_stream_ptr


@dataclasses.dataclass(frozen=True)
class PrioritySampledSections:
    """A holder for the samples and the log-weights (sample.py:10-19)."""

    batch: typ.Any  # RetrievalBatch
    log_weights: np.ndarray
    max_sampling_id: np.ndarray
    lse_pos: np.ndarray
    lse_neg: np.ndarray
    raw_scores: dict[str, np.ndarray]


def _draw_seed() -> int:
    # two 32-bit draws from the legacy global RNG: `np.random.seed(s)` therefore fixes the Philox stream
    hi, lo = np.random.randint(0, 2**32, size=2, dtype=np.uint64)
    return int((int(hi) << 32) | int(lo))


def _device_sample(scores: np.ndarray, labels: np.ndarray | None, noise: np.ndarray | None, k_positive: int,
                   k_total: int, normalized: bool, temperature: float, max_support_size: int, quirks: int,
                   seed: int, offset: int, device: int):
    lib = _lib.load()
    _lib.require_gpu()
    B, K = scores.shape
    s32 = np.ascontiguousarray(scores, dtype=np.float32)
    lab8 = None if labels is None else np.ascontiguousarray(np.asarray(labels) > 0, dtype=np.uint8)
    nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float32)
    ids = np.empty((B, k_total), np.int64)
    logw = np.empty((B, k_total), np.float32)
    olab = np.empty((B, k_total), np.uint8)
    lse = np.zeros((B, 2), np.float32)
    rc = lib.vodb_sample(int(device), s32.ctypes.data, None if lab8 is None else lab8.ctypes.data,
                         None if nz is None else nz.ctypes.data, B, K, int(k_positive), int(k_total),
                         int(bool(normalized)), float(temperature), int(max_support_size), int(quirks),
                         int(seed) & (2**64 - 1), int(offset) & (2**64 - 1), ids.ctypes.data, logw.ctypes.data,
                         olab.ctypes.data, lse.ctypes.data, 0, _current_stream_ptr(device))
    _lib.check(rc, "vodb_sample")
    return ids, logw, olab.astype(np.bool_), lse


def labeled_priority_sampling(
    scores: np.ndarray,
    labels: np.ndarray,
    k_positive: int = 1,
    k_total: int = 2,
    normalized: bool = True,
    temperature: float = 1.0,
    max_support_size: None | int = None,
    *,
    seed: None | int = None,
    offset: int = 0,
    noise: None | np.ndarray = None,
    fix_truncation: bool = False,
    device: int = 0,
) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Sample search results using Priority Sampling for each label value {0, 1} (sample.py:87-157).

    Returns `(sample indices i64[..., k_total], log weights, labels bool, log-normalising constants [..., 2])`.
    A sample is positive if its label is > 0. Unused slots hold -1 / -inf / False.
    """
    max_support_size = max_support_size or -1
    if max_support_size >= 0:
        max_support_size = max(max_support_size, k_total)  # sample.py:133-135
    scores = np.asarray(scores)
    if scores.ndim not in (1, 2):
        raise ValueError(f"Expected a 1D or 2D array. Got {scores.ndim}D.")
    if k_positive > k_total:
        raise ValueError(f"k_positive={k_positive} > k_total={k_total} (the reference writes out of bounds here)")
    one_d = scores.ndim == 1
    s2 = scores[None] if one_d else scores
    l2 = None if labels is None else (np.asarray(labels)[None] if one_d else np.asarray(labels))
    n2 = None if noise is None else (np.asarray(noise)[None] if one_d else np.asarray(noise))
    if l2 is not None and l2.shape != s2.shape:
        raise ValueError(f"labels shape {l2.shape} != scores shape {s2.shape}")
    if seed is None and n2 is None:
        seed = _draw_seed()
    quirks = 0 if fix_truncation else _lib.QUIRK_INVERTED_SUPPORT
    ids, logw, olab, lse = _device_sample(s2, l2, n2, k_positive, k_total, normalized, temperature,
                                          max_support_size, quirks, seed or 0, offset, device)
    if scores.dtype != np.float32 and scores.dtype.kind == "f":
        logw, lse = logw.astype(scores.dtype), lse.astype(scores.dtype)
    if one_d:
        return ids[0], logw[0], olab[0], lse[0]
    return ids, logw, olab, lse


def priority_sampling_1d(scores: np.ndarray, k: int = 1, temperature: float = 1.0, max_support_size: int = -1, *,
                         seed: None | int = None, noise: None | np.ndarray = None, fix_truncation: bool = False,
                         device: int = 0) -> tuple[np.ndarray, np.ndarray]:
    """Sample from unnormalised log p(z) using priority sampling (sample.py:222-242). Unnormalised weights."""
    scores = np.asarray(scores)
    if scores.ndim > 1:
        raise ValueError("Expected a 1D array.")
    n = scores.shape[0]
    kk = min(int(k), n)
    # a single-label problem: everything is "negative", k_positive = 0, so the negative group gets all k draws
    ids, logw, _, _ = labeled_priority_sampling(scores, None, k_positive=0, k_total=kk, normalized=False,
                                                temperature=temperature,
                                                max_support_size=None if (max_support_size or -1) < 0 else max_support_size,
                                                seed=seed, noise=noise, fix_truncation=fix_truncation, device=device)
    return ids, logw


def sample_search_results(
    *,
    search_results: typ.Any,
    raw_scores: dict[str, np.ndarray],
    total: None | int,
    max_pos_sections: None | int,
    temperature: float = 1.0,
    max_support_size: None | int = None,
    seed: None | int = None,
    offset: int = 0,
    noise: None | np.ndarray = None,
    fix_truncation: bool = False,
    device: int = 0,
) -> PrioritySampledSections:
    """Sample the positive and negative sections using per-label priority sampling (sample.py:22-84).

    One `vodb_sample_results` call: the sampler kernel draws the picks, the gather kernel of the retrieve->sample
    chain (csrc/chain.cu) reads ids / scores at the picks and counts `max_sampling_id` on the device. Only the
    per-engine `raw_scores`, which never go to the GPU, are gathered here from the returned positions.
    """
    lib = _lib.load()
    _lib.require_gpu()
    retrieved = np.asarray(search_results.scores)
    if retrieved.ndim != 2:
        raise ValueError(f"Expected 2D scores, got {retrieved.ndim}D")
    B, K = retrieved.shape
    k_total = int(total or K)
    k_positive = int(max_pos_sections or k_total)
    if k_positive > k_total:
        raise ValueError(f"k_positive={k_positive} > k_total={k_total} (the reference writes out of bounds here)")
    support = max_support_size or -1
    s32 = np.ascontiguousarray(retrieved, dtype=np.float32)
    ids64 = np.ascontiguousarray(search_results.indices, dtype=np.int64)
    positives = None if search_results.labels is None else np.ascontiguousarray(np.asarray(search_results.labels) > 0, dtype=np.uint8)
    exp1 = None if noise is None else np.ascontiguousarray(noise, dtype=np.float32)
    if exp1 is not None and exp1.shape != retrieved.shape:
        raise ValueError(f"noise shape {exp1.shape} != scores shape {retrieved.shape}")
    if seed is None and exp1 is None:
        seed = _draw_seed()

    picked_ids = np.empty((B, k_total), np.int64)
    positions = np.empty((B, k_total), np.int64)
    picked_scores = np.empty((B, k_total), np.float32)
    log_weights = np.empty((B, k_total), np.float32)
    picked_labels = np.empty((B, k_total), np.uint8)
    lse = np.zeros((B, 2), np.float32)
    msid = np.zeros(B, np.float32)
    rc = lib.vodb_sample_results(int(device), s32.ctypes.data, ids64.ctypes.data,
                                 None if positives is None else positives.ctypes.data,
                                 None if exp1 is None else exp1.ctypes.data, B, K, k_positive, k_total,
                                 float(temperature), int(support), 0 if fix_truncation else _lib.QUIRK_INVERTED_SUPPORT,
                                 int(seed or 0) & (2**64 - 1), int(offset) & (2**64 - 1), picked_ids.ctypes.data,
                                 picked_scores.ctypes.data, log_weights.ctypes.data, picked_labels.ctypes.data,
                                 lse.ctypes.data, msid.ctypes.data, positions.ctypes.data, _current_stream_ptr(device))
    _lib.check(rc, "vodb_sample_results")

    if retrieved.dtype != np.float32 and retrieved.dtype.kind == "f":  # float64 callers get their own values back
        picked_scores = np.take_along_axis(retrieved, positions, axis=-1)
        log_weights, lse = log_weights.astype(retrieved.dtype), lse.astype(retrieved.dtype)
    per_engine = {name: np.take_along_axis(np.asarray(arr), positions, axis=-1) for name, arr in raw_scores.items()}
    batch_cls = type(search_results) if hasattr(type(search_results), "cast") else RetrievalBatch
    return PrioritySampledSections(
        batch=batch_cls(indices=picked_ids, scores=picked_scores, labels=picked_labels.astype(np.bool_)),
        max_sampling_id=msid,
        lse_pos=lse[..., 0],
        lse_neg=lse[..., 1],
        log_weights=log_weights,
        raw_scores=per_engine,
    )
