"""Host glue RealmCollate applies to the sampled sections before it fetches their text: in-batch-negative
flattening and the replacement of padding ids. Arrays here are [batch, n_sections] (32 x 8 at BASELINE configs[3]),
so this stays numpy on the host — it is listed for drop-in completeness, not as a kernel.

    flatten_samples(samples, padding=True)            src/vod_dataloaders/core/in_batch_negatives.py:10-52
    gather_values_by_indices(queries, indices, ...)   src/vod_dataloaders/core/numpy_ops.py:24-143
    replace_negative_indices_(indices, world_size)    src/vod_dataloaders/core/numpy_ops.py:257-263
"""
from __future__ import annotations

import math
import typing as typ

import numpy as np

from .retrieval import RetrievalBatch
from .sampling import PrioritySampledSections


def gather_values_by_indices(queries: np.ndarray, indices: np.ndarray, values: np.ndarray,
                             fill_value: typ.Optional[float | int] = None) -> np.ndarray:
    """out[..., u] = values[..., j] for the FIRST j with indices[..., j] == queries[..., u]; `fill_value` (NaN for
    float values, -1 otherwise) where no key matches. Shapes: ([U], [k], [k]), ([B,U], [k], [k]) or ([B,U], [B,k], [B,k])."""
    queries, indices, values = np.asarray(queries), np.asarray(indices), np.asarray(values)
    if queries.ndim not in (1, 2):
        raise ValueError(f"Expected queries to have ndim 1 or 2. Found: {queries.ndim}")
    if indices.ndim not in (1, 2) or indices.ndim > queries.ndim:
        raise ValueError(f"Expected indices to have ndim 1 or 2. Found: {indices.ndim}")
    if indices.ndim == 2 and (len(indices) != len(queries) or len(values) != len(queries)):
        raise ValueError(f"Expected keys and values to have length {len(queries)}.")
    if fill_value is None:
        fill_value = np.nan if values.dtype.kind == "f" else -1
    hit = queries[..., :, None] == indices[..., None, :]            # [..., U, k]
    first = hit.argmax(axis=-1)                                      # first matching key (0 when none)
    vals = values if indices.ndim == queries.ndim else np.broadcast_to(values, queries.shape[:-1] + values.shape)
    picked = np.take_along_axis(vals, first, axis=-1) if vals.shape[-1] else np.zeros(queries.shape, values.dtype)
    out = np.full(queries.shape, fill_value, dtype=values.dtype)
    found = hit.any(axis=-1)
    out[found] = picked[found]
    return out


def flatten_samples(samples: PrioritySampledSections, padding: bool = True) -> PrioritySampledSections:
    """Merge all sampled sections (positive and negative) of the batch into one flat pool shared by every query:
    `indices` becomes the 1-D array of unique section ids (padded with id 1 to batch * n_sections entries when
    `padding`, as the reference does for a static shape), and scores / labels / log-weights / raw scores become
    [batch, pool] with NaN (labels: 0) where a query did not sample that section."""
    indices = samples.batch.indices
    unique_indices = np.unique(indices)
    if padding:
        n_pad = math.prod(indices.shape) - unique_indices.shape[0]
        unique_indices = np.concatenate([unique_indices, np.ones((n_pad,), dtype=np.int64)])
    pool = unique_indices[None, :].repeat(indices.shape[0], axis=0)
    if samples.batch.labels is None:
        raise ValueError("The `search_results` must have labels.")
    batch_cls = type(samples.batch) if hasattr(type(samples.batch), "cast") else RetrievalBatch
    return PrioritySampledSections(
        batch=batch_cls(indices=unique_indices, scores=gather_values_by_indices(pool, indices, samples.batch.scores),
                        labels=gather_values_by_indices(pool, indices, samples.batch.labels, fill_value=0),
                        allow_unsafe=True),
        max_sampling_id=samples.max_sampling_id,
        raw_scores={k: gather_values_by_indices(pool, indices, v) for k, v in samples.raw_scores.items()},
        log_weights=gather_values_by_indices(pool, indices, samples.log_weights),
        lse_pos=samples.lse_pos, lse_neg=samples.lse_neg)


def replace_negative_indices_(indices: np.ndarray, world_size: int) -> None:
    """Replace negative (padding) ids by random valid ones, in place: `datasets.Dataset` cannot be indexed with -1."""
    is_negative = indices < 0
    n_negative = int(is_negative.sum())
    if n_negative:
        indices.setflags(write=True)
        indices[is_negative] = np.random.randint(0, world_size, size=n_negative)
