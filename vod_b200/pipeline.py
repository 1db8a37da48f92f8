"""Device-resident retrieve -> sample chain for dense-only flows (RealmCollate-style dynamic retrieval).

What `RealmCollate.__call__` does per training batch (src/vod_dataloaders/realm_collate.py:101-122): search the
top-`prefetch_n_sections` passages for every query, then `sample_search_results` draws `n_sections` of them with
importance weights. With the dense engine alone, both steps can stay on the GPU: `vodb_search` leaves scores / ids in
HBM, `vodb_sample` reads them there, and only the [B, n_sections] picks cross PCIe (one D2H) instead of the
[B, top_k] lists (12 KB vs 384 KB at B=32, K=1000, k=8). Labels (gold sections) are matched on the device from an
optional [B, P] table of positive ids. Scores handed to the sampler are the raw inner products: with a single
engine the per-row shift of `_subtract_min_score` cancels in the sampler's log-softmax.
"""
from __future__ import annotations

import typing as typ

import numpy as np

from . import _lib
from .retrieval import RetrievalBatch
from .sampling import PrioritySampledSections, _draw_seed
from .search import CorpusStore, _current_stream_ptr


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
    """search(top_k) -> labeled priority sampling(total) with everything but the final picks resident in HBM."""

    def __init__(self, store: CorpusStore, *, top_k: int = 1000, total: int = 8, max_pos_sections: int | None = None,
                 temperature: float = 1.0, max_support_size: int | None = None, mode: str | None = None):
        self.store, self.top_k, self.total = store, top_k, total
        self.max_pos_sections = max_pos_sections or total
        self.temperature, self.max_support_size, self.mode = temperature, max_support_size, mode

    def __call__(self, queries: typ.Any, positive_ids: typ.Any | None = None, *, seed: int | None = None,
                 offset: int = 0) -> PrioritySampledSections:
        import torch

        dev = torch.device(f"cuda:{self.store.device}")
        q = queries if hasattr(queries, "is_cuda") else torch.from_numpy(np.ascontiguousarray(queries))
        q = q.to(dev, non_blocking=True)
        scores, ids = self.store.search_device(q, self.top_k, mode=self.mode)
        labels = None
        if positive_ids is not None:
            pos = positive_ids if hasattr(positive_ids, "is_cuda") else torch.from_numpy(np.ascontiguousarray(positive_ids))
            pos = pos.to(dev, non_blocking=True).to(torch.int64)
            labels = (ids.unsqueeze(-1) == pos.unsqueeze(1)).any(-1)  # tensor hand-off glue: gold-section match
        local, logw, olab, lse = sample_device(scores, labels, k_positive=self.max_pos_sections, k_total=self.total,
                                               temperature=self.temperature, max_support_size=self.max_support_size,
                                               seed=_draw_seed() if seed is None else seed, offset=offset)
        picked = local.clamp_min(0)  # -1 (unused slot) gathers the last column in the reference; keep ids as -1 instead
        out_ids = torch.where(local >= 0, torch.gather(ids, 1, picked), torch.full_like(local, -1))
        out_scores = torch.where(local >= 0, torch.gather(scores, 1, picked), torch.full_like(logw, float("-inf")))
        # one device -> host transfer of the [B, total] picks
        packed = [t.cpu() for t in (out_ids, out_scores, logw, olab, lse)]
        if self.store.check_async():  # a candidate list overflowed (adversarial order): redo on the safe schedule
            s_np, i_np = self.store.search(q.cpu().numpy(), self.top_k, mode=self.mode)
            return self.__class__._host_fallback(self, s_np, i_np, positive_ids, seed, offset)
        o_ids, o_scores, o_w, o_lab, o_lse = (t.numpy() for t in packed)
        return PrioritySampledSections(
            batch=RetrievalBatch(indices=o_ids, scores=o_scores, labels=o_lab.astype(np.bool_)),
            log_weights=o_w, max_sampling_id=np.full(len(o_ids), np.nan, np.float32), lse_pos=o_lse[:, 0],
            lse_neg=o_lse[:, 1], raw_scores={"dense": o_scores})

    def _host_fallback(self, scores, ids, positive_ids, seed, offset) -> PrioritySampledSections:
        from .sampling import sample_search_results

        labels = None
        if positive_ids is not None:
            pos = np.asarray(positive_ids.cpu() if hasattr(positive_ids, "cpu") else positive_ids)
            labels = (ids[:, :, None] == pos[:, None, :]).any(-1).astype(np.int64)
        batch = RetrievalBatch(scores=scores, indices=ids, labels=labels)
        return sample_search_results(search_results=batch, raw_scores={"dense": scores}, total=self.total,
                                     max_pos_sections=self.max_pos_sections, temperature=self.temperature,
                                     max_support_size=self.max_support_size, seed=seed, offset=offset,
                                     device=self.store.device)
