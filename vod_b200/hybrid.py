"""Hybrid merge: the step between search and sampling (normalise scores, weighted union by id, raw-score and
label gather), on the GPU.

Mirrors the reference functions, same names / arguments / return values:

    merge_search_results(search_results, weights) -> (RetrievalBatch, raw_scores)   core/merge.py:8-62
    normalize_search_scores_(search_results, offset)                               core/normalize.py:6-14
    _merge_search_results(search_results, weights)                                 core/search.py:79-125
    async_hybrid_search(text=..., shards=..., vector=..., ..., clients, weights)   core/search.py:20-76

The union / duplicate summation / gathers run in one CUDA kernel per call (`vodb_merge_results`,
vod_b200/csrc/merge_results.cu: sort by (id, position) + run heads + prefix-sum compaction instead of the
reference's O(K^2) numba probes). Scores are processed in the dtype of the first engine (float32 or float64), like
the reference's output buffer (`np.full(..., dtype=a_scores.dtype)`, merge.py:117-121); engines of another float
dtype are cast first (documented divergence: the reference multiplies by the weight in the engine's own dtype).
"""
from __future__ import annotations

import ctypes
import time
import typing as typ

import numpy as np

from . import _lib
from .retrieval import RetrievalBatch
from .search import SearchClient, _current_stream_ptr

FLOAT_INF_THRES = 3e12  # core/search.py:16
LOOKUP_CLIENT_NAME = "lookup"  # core/search.py:17


def _batch_like(template: typ.Any, **kw) -> typ.Any:
    cls = type(template) if hasattr(type(template), "cast") else RetrievalBatch
    return cls(**kw)


def _device_merge(results: dict[str, typ.Any], weights: dict[str, float], *, normalize: bool, offset: float,
                  zero: set[str], device: int):
    lib = _lib.load()
    _lib.require_gpu()
    keys = list(results.keys())
    first = results[keys[0]]
    fdtype = np.float64 if np.asarray(first.scores).dtype == np.float64 else np.float32
    B = len(first.scores)
    n = len(keys)
    scores = [np.ascontiguousarray(results[k].scores, dtype=fdtype) for k in keys]
    indices = [np.ascontiguousarray(results[k].indices, dtype=np.int64) for k in keys]
    labels = [None if results[k].labels is None else np.ascontiguousarray(results[k].labels, dtype=np.int64)
              for k in keys]
    for s, i in zip(scores, indices):
        if s.ndim != 2 or s.shape != i.shape or s.shape[0] != B:
            raise ValueError("All scores must have the same length.")  # merge.py:26-28
    label_engine = max((e for e, lab in enumerate(labels) if lab is not None), default=-1)  # last one wins, merge.py:52-60
    widths = (ctypes.c_int * n)(*[s.shape[1] for s in scores])
    out_width = max(1, sum(s.shape[1] for s in scores))
    vp = ctypes.c_void_p
    t_scores = (vp * n)(*[s.ctypes.data for s in scores])
    t_idx = (vp * n)(*[i.ctypes.data for i in indices])
    t_lab = (vp * n)(*[None if lab is None else lab.ctypes.data for lab in labels])
    w = (ctypes.c_double * n)(*[float(weights[k]) for k in keys])
    z = (ctypes.c_int * n)(*[1 if k in zero else 0 for k in keys])
    out_s = np.empty((B, out_width), fdtype)
    out_i = np.empty((B, out_width), np.int64)
    out_l = np.empty((B, out_width), np.int64) if label_engine >= 0 else None
    out_raw = np.empty((n, B, out_width), fdtype)
    counts = np.zeros(B, np.int32)
    rc = lib.vodb_merge_results(int(device), n, t_scores, t_idx, t_lab, widths, w, z, B, 3 if fdtype == np.float64 else 0,
                                int(normalize), float(offset), label_engine, out_width, out_s.ctypes.data,
                                out_i.ctypes.data, None if out_l is None else out_l.ctypes.data, out_raw.ctypes.data,
                                counts.ctypes.data, 0, _current_stream_ptr(device))
    _lib.check(rc, "vodb_merge_results")
    width = min(out_width, int(counts.max(initial=0)) + 1)  # merge.py:160-162 `[:, : max_cursor + 1]`
    merged = _batch_like(first, scores=out_s[:, :width].copy(), indices=out_i[:, :width].copy(),
                         labels=None if out_l is None else out_l[:, :width].copy())
    raw = {k: out_raw[e, :, :width].copy() for e, k in enumerate(keys)}
    return merged, raw


def merge_search_results(search_results: dict[str, typ.Any], weights: None | dict[str, float] = None, *,
                         device: int = 0) -> tuple[typ.Any, dict[str, np.ndarray]]:
    """Merge search results with weights (core/merge.py:8-62)."""
    if weights is None:
        weights = {k: 1.0 for k in search_results}
    elif not set(weights) >= set(search_results):
        raise ValueError(f"Expected weights to have keys {set(search_results)}. Found: {set(weights)}")
    if len(search_results) == 1:  # merge.py:18-22: no union, the single result is only scaled
        key = list(search_results.keys())[0]
        result = search_results[key]
        return result * weights[key], {key: search_results[key].scores}
    ulengths = {len(v.scores) for v in search_results.values()}
    if len(ulengths) != 1:
        raise ValueError(f"All scores must have the same length. Found: {ulengths}")
    return _device_merge(search_results, weights, normalize=False, offset=0.0, zero=set(), device=device)


def _subtract_min_score(scores: np.ndarray, offset: float = 0.0) -> np.ndarray:
    """core/normalize.py:17-20 (host glue for callers that only normalise; the merge kernel fuses this step)."""
    non_nan_scores = np.where(np.isinf(scores) | np.isnan(scores), np.inf, scores)
    min_score = np.amin(non_nan_scores, axis=-1, keepdims=True)
    return scores - min_score + offset


def normalize_search_scores_(search_results: dict[str, typ.Any], offset: float = 0.0) -> None:
    """Subtract the minimum score from all scores, in place (core/normalize.py:6-14)."""
    for key, result in search_results.items():
        if result.scores.size == 0:
            continue
        search_results[key].scores = _subtract_min_score(result.scores, offset=offset)


def _merge_search_results(search_results: dict[str, typ.Any], weights: dict[str, float], *,
                          device: int = 0) -> tuple[typ.Any, dict[str, np.ndarray]]:
    """`lookup` scores -> 0, every engine's scores -> s - row_min, weighted union, raw scores + labels gathered
    (core/search.py:79-125). One kernel launch instead of the reference's fill / normalise / merge / gather passes."""
    if LOOKUP_CLIENT_NAME not in search_results:
        raise ValueError(f"The `{LOOKUP_CLIENT_NAME}` client must be specified to lookup the golden/positive sections.")
    meta: dict[str, typ.Any] = {}
    for name, result in search_results.items():
        if name != LOOKUP_CLIENT_NAME:
            result.labels = None                     # core/search.py:93-96
    if "dense" in search_results:                    # core/search.py:99-100, 149-161
        r = search_results["dense"]
        is_inf = r.scores >= FLOAT_INF_THRES
        if is_inf.any():
            r.scores[is_inf] = np.nan
    for name, result in search_results.items():      # core/search.py:103-105
        for key, value in result.meta.items():
            meta[f"{name}_{key}"] = value
    all_weights = {LOOKUP_CLIENT_NAME: 0.0, **weights}
    if not set(all_weights) >= set(search_results):
        raise ValueError(f"Expected weights to have keys {set(search_results)}. Found: {set(all_weights)}")
    if len(search_results) == 1:
        key = LOOKUP_CLIENT_NAME
        res = search_results[key]
        res.scores.fill(0.0)
        combined, raw_scores = res * 0.0, {key: res.scores}
    else:
        combined, raw_scores = _device_merge(search_results, all_weights, normalize=True, offset=0.0,
                                             zero={LOOKUP_CLIENT_NAME}, device=device)
    raw_scores.pop(LOOKUP_CLIENT_NAME)
    combined.meta = meta
    return combined, raw_scores


def async_hybrid_search(*, text: list[str], shards: list[str], vector: None | np.ndarray = None,
                        subset_ids: None | list[list[str]] = None, section_ids: list[list[str]], top_k: int,
                        clients: dict[str, SearchClient], weights: dict[str, float],
                        lookup_engine_name: str = "sparse", device: int = 0):
    """Query every engine (plus the golden-section lookup on `lookup_engine_name`) and merge the results
    (core/search.py:20-76). The in-process GPU client needs no asyncio fan-out; calls are issued in order."""
    meta: dict[str, typ.Any] = {}
    if lookup_engine_name not in clients:
        raise ValueError(f"The `{lookup_engine_name}` client must be specified to lookup the golden/positive sections.")
    start = time.perf_counter()
    results: dict[str, typ.Any] = {}
    t0 = time.perf_counter()
    results[LOOKUP_CLIENT_NAME] = clients[lookup_engine_name].search(
        vector=vector, text=[""] * len(text), subset_ids=subset_ids, ids=section_ids, shard=shards, top_k=top_k)
    results[LOOKUP_CLIENT_NAME].meta["search_time"] = time.perf_counter() - t0
    for name, client in clients.items():
        t0 = time.perf_counter()
        results[name] = client.search(vector=vector, text=text, subset_ids=subset_ids, shard=shards, top_k=top_k)
        results[name].meta["search_time"] = time.perf_counter() - t0
    meta["search_time"] = time.perf_counter() - start
    combined, raw_scores = _merge_search_results(results, weights, device=device)
    combined.meta.update(meta)
    return combined, raw_scores
