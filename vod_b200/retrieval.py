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
    """Model search results (retrieval.py:18-133)."""

    __slots__ = ("scores", "indices", "labels", "allow_unsafe", "meta")
    _expected_dim: int = -1

    def __init__(self, scores, indices, labels=None, meta=None, allow_unsafe: bool = False):
        dim = len(indices.shape)
        if not allow_unsafe and scores.shape[:dim] != indices.shape[:dim]:
            raise ValueError(
                "The shapes of `scores` and `indices` must match up to the dimension of `indices`, "
                f"but got {_array_repr(scores)} and {_array_repr(indices)}"
            )
        if labels is not None and (scores.shape[:dim] != labels.shape[:dim]):
            raise ValueError("The shapes of `scores` and `labels` must match up to the dimension of `indices`, ")
        if len(scores.shape) != self._expected_dim:
            raise ValueError(
                f"Scores must be {self._expected_dim}D, but got {_array_repr(scores)} and {_array_repr(indices)}"
            )
        self.allow_unsafe = allow_unsafe
        self.scores = scores
        self.indices = indices
        self.labels = labels
        self.meta = meta or {}

    @classmethod
    def cast(cls, scores, indices, labels=None, meta=None, allow_unsafe: bool = False):
        return cls(
            scores=_cast_to_numpy(scores),
            indices=_cast_to_numpy(indices),
            labels=_cast_to_numpy(labels) if labels is not None else None,
            meta=meta,
            allow_unsafe=allow_unsafe,
        )

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
        return {
            "scores": self.scores.tolist(),
            "indices": self.indices.tolist(),
            "labels": self.labels.tolist() if self.labels is not None else None,
        }


class RetrievalTuple(RetrievalData):
    _expected_dim = 0


class RetrievalSample(RetrievalData):
    _expected_dim = 1

    def __getitem__(self, item: int) -> RetrievalTuple:
        return RetrievalTuple(
            scores=self.scores[item],
            indices=self.indices[item],
            labels=self.labels[item] if self.labels is not None else None,
        )

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __add__(self, other: "RetrievalSample") -> "RetrievalBatch":
        return stack_samples([self, other])


class RetrievalBatch(RetrievalData):
    """A batch of search results: scores f32[B,K], indices i64[B,K] (retrieval.py:179-249)."""

    _expected_dim = 2

    def __getitem__(self, item: int) -> RetrievalSample:
        return RetrievalSample(
            scores=self.scores[item],
            indices=self.indices[item],
            labels=self.labels[item] if self.labels is not None else None,
        )

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __add__(self, other: "RetrievalBatch") -> "RetrievalBatch":
        return RetrievalBatch(
            scores=np.concatenate([self.scores, other.scores]),
            indices=np.concatenate([self.indices, other.indices]),
            labels=_merge_labels(self.labels, other.labels),
        )

    def sorted(self) -> "RetrievalBatch":
        """Sort by score, descending (retrieval.py:211-220)."""
        sort_ids = np.flip(np.argsort(self.scores, axis=-1), axis=-1)
        return RetrievalBatch(
            scores=np.take_along_axis(self.scores, sort_ids, axis=-1),
            indices=np.take_along_axis(self.indices, sort_ids, axis=-1),
            labels=np.take_along_axis(self.labels, sort_ids, axis=-1) if self.labels is not None else None,
            meta=copy.copy(self.meta),
        )

    def __mul__(self, value: float) -> "RetrievalBatch":
        if not isinstance(value, Number):
            raise TypeError(f"Expected a number, but got `{type(value)}`")
        with warnings.catch_warnings():
            warnings.filterwarnings("ignore", category=RuntimeWarning)
            return RetrievalBatch(scores=self.scores * value, indices=self.indices, labels=self.labels,
                                  meta=copy.copy(self.meta))

    @classmethod
    def stack_samples(cls, samples: typ.Iterable[RetrievalSample]) -> "RetrievalBatch":
        return stack_samples(samples)

    @classmethod
    def concatenate_batches(cls, batches: typ.Iterable["RetrievalBatch"]) -> "RetrievalBatch":
        output = None
        for batch in batches:
            output = batch if output is None else output + batch
        if output is None:
            raise ValueError("Cannot concatenate an empty list of batches")
        return output


def _stack_1d(arrays: list[np.ndarray], fill_value: typ.Any) -> np.ndarray:
    width = max(len(a) for a in arrays)
    out = np.full((len(arrays), width), fill_value, dtype=arrays[0].dtype)
    for j, a in enumerate(arrays):
        out[j, : len(a)] = a
    return out


def stack_samples(samples: typ.Iterable[RetrievalSample]) -> RetrievalBatch:
    """Stack ragged samples, padding with score -inf / index -1 / label -1 (retrieval.py:276-287)."""
    samples = list(samples)
    labels = [s.labels for s in samples]
    return RetrievalBatch(
        scores=_stack_1d([s.scores for s in samples], -math.inf),
        indices=_stack_1d([s.indices for s in samples], -1),
        labels=None if any(lbl is None for lbl in labels) else _stack_1d(labels, -1),
    )


def _merge_labels(a, b):
    if a is None and b is None:
        return None
    if a is None:
        a = np.full_like(b, fill_value=-1)
    if b is None:
        b = np.full_like(a, fill_value=-1)
    return np.concatenate([a, b])
