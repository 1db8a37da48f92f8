"""Device-resident retrieve -> sample chain for dense-only flows (RealmCollate-style dynamic retrieval).

What `RealmCollate.__call__` does per training batch (src/vod_dataloaders/realm_collate.py:101-122): search the
top-`prefetch_n_sections` passages for every query, then `sample_search_results` draws `n_sections` of them with
importance weights. With the dense engine alone, both steps stay on the GPU behind ONE C-ABI call
(`vodb_retrieve_sample`): the scan leaves scores / ids in HBM, a label kernel matches them against an optional [B, P]
table of gold ids, the sampler reads them there, a gather kernel applies the picks, and only the [B, n_sections]
result crosses PCIe (one packed D2H) instead of the [B, top_k] lists (12 KB vs 384 KB at B=32, K=1000, k=8).
Scores handed to the sampler are the raw inner products: with a single engine the per-row shift of
`_subtract_min_score` cancels in the sampler's log-softmax.
"""
from __future__ import annotations

import typing as typ

import numpy as np

from . import _lib
from .retrieval import RetrievalBatch
from .sampling import PrioritySampledSections, _draw_seed
from .search import CorpusStore, _current_stream_ptr, _np_dtype_code, _torch_info


def sample_device(scores: typ.Any, labels: typ.Any | None, *, k_positive: int, k_total: int, normalized: bool = True,
                  temperature: float = 1.0, max_support_size: int | None = None, seed: int = 0, offset: int = 0,
                  fix_truncation: bool = False):
    """`labeled_priority_sampling` on CUDA tensors: scores f32 [B,K], labels uint8/bool [B,K] or None.
    Returns (ids i64 [B,k_total], log_weights f32, labels uint8, lse f32 [B,2]) CUDA tensors; only enqueues."""
    import torch

    if scores.dtype != torch.float32 or not scores.is_cuda or not scores.is_contiguous():
        raise ValueError("scores must be a contiguous float32 CUDA tensor")
    B, K = scores.shape
    dev = scores.device
    lab = None
    if labels is not None:
        lab = labels.to(torch.uint8).contiguous()
    ms = max_support_size or -1
    if ms >= 0:
        ms = max(ms, k_total)
    ids = torch.empty((B, k_total), dtype=torch.int64, device=dev)
    logw = torch.empty((B, k_total), dtype=torch.float32, device=dev)
    olab = torch.empty((B, k_total), dtype=torch.uint8, device=dev)
    lse = torch.empty((B, 2), dtype=torch.float32, device=dev)
    lib = _lib.load()
    rc = lib.vodb_sample(dev.index, scores.data_ptr(), None if lab is None else lab.data_ptr(), None, B, K,
                         int(k_positive), int(k_total), int(bool(normalized)), float(temperature), int(ms),
                         0 if fix_truncation else _lib.QUIRK_INVERTED_SUPPORT, int(seed) & (2**64 - 1),
                         int(offset) & (2**64 - 1), ids.data_ptr(), logw.data_ptr(), olab.data_ptr(), lse.data_ptr(), 1,
                         _current_stream_ptr(dev.index))
    _lib.check(rc, "vodb_sample")
    return ids, logw, olab, lse


class DenseRetrievalSampler:
    """search(top_k) -> labeled priority sampling(total) -> gathers, one C call (`vodb_retrieve_sample`): everything
    but the final [B, total] picks stays in HBM. Output equals `CorpusStore.search` -> labels from `positive_ids` ->
    `sample_search_results(...)` bit for bit, `max_sampling_id` and the wrap-around gather of unused slots included
    (core/sample.py:57-71)."""

    def __init__(self, store: CorpusStore, *, top_k: int = 1000, total: int = 8, max_pos_sections: int | None = None,
                 temperature: float = 1.0, max_support_size: int | None = None, mode: str | None = None,
                 fix_truncation: bool = False):
        self.store, self.top_k, self.total = store, top_k, total
        self.max_pos_sections = max_pos_sections or total
        self.temperature, self.max_support_size, self.mode = temperature, max_support_size, mode
        self.fix_truncation = fix_truncation

    def __call__(self, queries: typ.Any, positive_ids: typ.Any | None = None, *, seed: int | None = None,
                 offset: int = 0) -> PrioritySampledSections:
        st = self.store
        if hasattr(queries, "is_cuda"):  # torch tensor, host or device: handed over by pointer
            if queries.dim() != 2 or queries.shape[1] != st.dim:
                raise ValueError(f"expected queries of shape [B, {st.dim}], got {tuple(queries.shape)}")
            q_ptr, q_code, q_cuda, q_dev = _torch_info(queries)
            if q_cuda and q_dev != st.device:
                raise ValueError(f"queries live on cuda:{q_dev}, the store on cuda:{st.device}")
            keep, B = queries, int(queries.shape[0])
        else:
            q = np.ascontiguousarray(queries)
            if q.ndim != 2:
                raise ValueError(f"Expected 2D array, got {q.ndim}D array")  # server.py:82-83
            if q.shape[1] != st.dim:
                raise ValueError(f"query dimension {q.shape[1]} != index dimension {st.dim}")
            if q.dtype not in (np.float32, np.float16):
                q = q.astype(np.float32)
            keep, q_ptr, q_code, q_cuda, B = q, q.ctypes.data, _np_dtype_code(q), False, q.shape[0]
        gold, n_gold = None, 0
        if positive_ids is not None:
            gold = positive_ids.cpu().numpy() if hasattr(positive_ids, "cpu") else np.asarray(positive_ids)
            gold = np.ascontiguousarray(gold, dtype=np.int64).reshape(B, -1)
            n_gold = gold.shape[1]
        kt = int(self.total)
        if self.max_pos_sections > kt:
            raise ValueError(f"k_positive={self.max_pos_sections} > k_total={kt} (the reference writes out of bounds here)")
        ms = self.max_support_size or -1  # None / 0 -> no truncation, like sample.py:131
        out_idx, out_local = np.empty((B, kt), np.int64), np.empty((B, kt), np.int64)
        out_scores, out_logw = np.empty((B, kt), np.float32), np.empty((B, kt), np.float32)
        out_lab = np.empty((B, kt), np.uint8)
        out_lse, out_msid = np.empty((B, 2), np.float32), np.empty(B, np.float32)
        rc = st._lib.vodb_retrieve_sample(
            st.handle, q_ptr, q_code, int(q_cuda), B, int(self.top_k), st._mode(self.mode, q_code),
            None if n_gold == 0 else gold.ctypes.data, n_gold, int(self.max_pos_sections), kt,
            float(self.temperature), int(ms), 0 if self.fix_truncation else _lib.QUIRK_INVERTED_SUPPORT,
            (_draw_seed() if seed is None else int(seed)) & (2**64 - 1), int(offset) & (2**64 - 1),
            out_idx.ctypes.data, out_scores.ctypes.data, out_logw.ctypes.data, out_lab.ctypes.data, out_lse.ctypes.data,
            out_msid.ctypes.data, out_local.ctypes.data, _current_stream_ptr(st.device))
        del keep
        _lib.check(rc, "vodb_retrieve_sample")
        self.last_local_ids = out_local
        return PrioritySampledSections(
            batch=RetrievalBatch(indices=out_idx, scores=out_scores, labels=out_lab.astype(np.bool_)),
            log_weights=out_logw, max_sampling_id=out_msid, lse_pos=out_lse[:, 0], lse_neg=out_lse[:, 1],
            raw_scores={"dense": out_scores.copy()})
