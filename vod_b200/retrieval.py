"""`RetrievalBatch` — the result container of the search boundary.

Mirror of the reference's `vod_types.RetrievalBatch` (src/vod_types/retrieval.py:18-58, 179-249):
same attributes (`scores`, `indices`, `labels`, `meta`), same shape validation, `cast`, `sorted`,
`__mul__`, `__add__`, `__getitem__`/`__iter__`, `to_dict` and `stack_samples` padding (-inf / -1,
retrieval.py:276-287). When the real `vod_types` package is importable the search client returns
the reference's own class instead (see `vod_b200.search._retrieval_batch_cls`).
"""
from __future__ import annotations

import copy
import math
import typing as typ
import warnings
from numbers import Number

import numpy as np


def _cast_to_numpy(x: typ.Any) -> np.ndarray:
    if hasattr(x, "detach") and hasattr(x, "cpu"):  # torch.Tensor without importing torch
        return x.detach().cpu().numpy()
    return np.asarray(x)


def _array_repr(x: typ.Any) -> str:
    return f"{type(x).__name__}(shape={x.shape}, dtype={x.dtype}))"


class RetrievalData:
    """Model search results (retrieval.py:18-133): `scores`, `indices`, optional `labels`, free-form `meta`."""

    __slots__ = ("scores", "indices", "labels", "allow_unsafe", "meta")
    _expected_dim: int = -1

    def __init__(self, scores, indices, labels=None, meta=None, allow_unsafe: bool = False):
        lead = len(indices.shape)  # scores may carry extra trailing dims; the leading ones must agree
        if not allow_unsafe and scores.shape[:lead] != indices.shape[:lead]:
            raise ValueError("The shapes of `scores` and `indices` must match up to the dimension of `indices`, "
                             f"but got {_array_repr(scores)} and {_array_repr(indices)}")
        if labels is not None and scores.shape[:lead] != labels.shape[:lead]:
            raise ValueError("The shapes of `scores` and `labels` must match up to the dimension of `indices`, ")
        if len(scores.shape) != self._expected_dim:
            raise ValueError(f"Scores must be {self._expected_dim}D, but got {_array_repr(scores)} and {_array_repr(indices)}")
        self.scores, self.indices, self.labels = scores, indices, labels
        self.meta = meta or {}
        self.allow_unsafe = allow_unsafe

    def _arrays(self) -> tuple:
        return self.scores, self.indices, self.labels

    def _derive(self, cls: type, fn: typ.Callable[[typ.Any], typ.Any], **kw: typ.Any) -> "RetrievalData":
        """New container of type `cls` with `fn` applied to scores, indices and (if present) labels."""
        s, i, lab = self._arrays()
        return cls(fn(s), fn(i), None if lab is None else fn(lab), **kw)

    @classmethod
    def cast(cls, scores, indices, labels=None, meta=None, allow_unsafe: bool = False):
        """Build from lists / numpy arrays / torch tensors (converted to numpy)."""
        as_np = [None if x is None else _cast_to_numpy(x) for x in (scores, indices, labels)]
        return cls(*as_np, meta=meta, allow_unsafe=allow_unsafe)

    def __len__(self) -> int:
        return len(self.scores)

    @property
    def shape(self) -> tuple[int, ...]:
        return self.scores.shape

    def __repr__(self) -> str:
        return (f"{type(self).__name__}[{type(self.scores).__name__}](scores={self.scores!r}, "
                f"indices={self.indices!r}, labels={self.labels!r}, meta={self.meta!r})")

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, type(self)):
            raise NotImplementedError(f"Cannot compare {type(self)} with {type(other)}")
        return bool(np.all(self.scores == other.scores) and np.all(self.indices == other.indices))

    def to_dict(self) -> dict[str, typ.Any]:
        return {name: None if arr is None else arr.tolist()
                for name, arr in zip(("scores", "indices", "labels"), self._arrays())}


class RetrievalTuple(RetrievalData):
    _expected_dim = 0


class _Indexable(RetrievalData):
    """Row access shared by samples (rows are tuples) and batches (rows are samples)."""

    _row_cls: type = RetrievalData

    def __getitem__(self, item: int):
        return self._derive(self._row_cls, lambda arr: arr[item])

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class RetrievalSample(_Indexable):
    _expected_dim = 1
    _row_cls = RetrievalTuple

    def __add__(self, other: "RetrievalSample") -> "RetrievalBatch":
        return stack_samples([self, other])


class RetrievalBatch(_Indexable):
    """A batch of search results: scores f32[B,K], indices i64[B,K] (retrieval.py:179-249)."""

    _expected_dim = 2
    _row_cls = RetrievalSample

    def __add__(self, other: "RetrievalBatch") -> "RetrievalBatch":
        """Concatenate along the batch axis; a side without labels contributes -1."""
        lab_a, lab_b = self.labels, other.labels
        if (lab_a is None) != (lab_b is None):
            lab_a = np.full_like(lab_b, -1) if lab_a is None else lab_a
            lab_b = np.full_like(lab_a, -1) if lab_b is None else lab_b
        return RetrievalBatch(np.concatenate([self.scores, other.scores]), np.concatenate([self.indices, other.indices]),
                              None if lab_a is None else np.concatenate([lab_a, lab_b]))

    def sorted(self) -> "RetrievalBatch":
        """Sort every row by score, descending (retrieval.py:211-220)."""
        order = np.flip(np.argsort(self.scores, axis=-1), axis=-1)
        return self._derive(RetrievalBatch, lambda arr: np.take_along_axis(arr, order, axis=-1), meta=copy.copy(self.meta))

    def __mul__(self, value: float) -> "RetrievalBatch":
        if not isinstance(value, Number):
            raise TypeError(f"Expected a number, but got `{type(value)}`")
        with warnings.catch_warnings():  # inf * 0 in padded slots is expected
            warnings.filterwarnings("ignore", category=RuntimeWarning)
            scaled = self.scores * value
        return RetrievalBatch(scaled, self.indices, self.labels, meta=copy.copy(self.meta))

    @classmethod
    def stack_samples(cls, samples: typ.Iterable[RetrievalSample]) -> "RetrievalBatch":
        return stack_samples(samples)

    @classmethod
    def concatenate_batches(cls, batches: typ.Iterable["RetrievalBatch"]) -> "RetrievalBatch":
        batches = list(batches)
        if not batches:
            raise ValueError("Cannot concatenate an empty list of batches")
        total = batches[0]
        for nxt in batches[1:]:
            total = total + nxt
        return total


def _pad_rows(arrays: list[np.ndarray], fill_value: typ.Any) -> np.ndarray:
    width = max(len(a) for a in arrays)
    out = np.full((len(arrays), width), fill_value, dtype=arrays[0].dtype)
    for j, a in enumerate(arrays):
        out[j, : len(a)] = a
    return out


def stack_samples(samples: typ.Iterable[RetrievalSample]) -> RetrievalBatch:
    """Stack ragged samples into a batch, padding with score -inf / index -1 / label -1 (retrieval.py:276-287)."""
    samples = list(samples)
    with_labels = all(s.labels is not None for s in samples)
    return RetrievalBatch(_pad_rows([s.scores for s in samples], -math.inf), _pad_rows([s.indices for s in samples], -1),
                          _pad_rows([s.labels for s in samples], -1) if with_labels else None)
