"""`ShardedSearchClient` — routes each query to the corpus named in `shard[i]` and re-bases the returned ids.

Mirror of the reference's router (src/vod_search/sharded_search.py:28-173, helpers :176-203). "Shard" here means a
different corpus (one search client per dataset), not a row partition of one corpus: results are NOT merged across
shards; row i of the output comes from the client named `shard[i]`, with `indices += offsets[shard[i]]` applied in
place (also to -1 padding, like the reference, sharded_search.py:103), rows padded to a common width with
score -inf / index -1 by `RetrievalBatch.stack_samples`.
"""
from __future__ import annotations

import collections

import numpy as np

from .retrieval import RetrievalBatch
from .search import SearchClient, SectionId, ShardName, SubsetId


class ShardedSearchClient(SearchClient):
    """A sharded search client (sharded_search.py:28-106)."""

    def __init__(self, shards: dict[ShardName, SearchClient], offsets: dict[ShardName, int]):
        self._shards = shards
        self._offsets = offsets
        if shards.keys() != offsets.keys():
            raise ValueError(
                f"Keys of `shards` and `offsets` must be the same. Found {shards.keys()} and {offsets.keys()}"
            )

    def __repr__(self) -> str:
        return f"{type(self).__name__}(shards={self._shards})"

    @property
    def shards(self) -> dict[ShardName, SearchClient]:
        return self._shards.copy()

    @property
    def offsets(self) -> dict[ShardName, int]:
        return self._offsets.copy()

    @property
    def requires_vectors(self) -> bool:  # type: ignore[override]
        return any(shard.requires_vectors for shard in self.shards.values())

    def ping(self) -> bool:
        return all(shard.ping() for shard in self.shards.values())

    def search(self, *, text: list[str], vector: None | np.ndarray = None,
               subset_ids: None | list[list[SubsetId]] = None, ids: None | list[list[SectionId]] = None,
               shard: None | list[ShardName] = None, top_k: int = 3) -> RetrievalBatch:
        if shard is None:
            raise ValueError("Must specify `shard`")
        if set(shard) > set(self.shards.keys()):
            raise ValueError(f"Invalid shard names {shard}. Valid names are {self.shards.keys()}")
        rows_of = _rows_by_shard(shard)
        # one sub-batch per corpus, rows in input order; absent optional fields stay absent ([] like the reference)
        where: dict[int, tuple[ShardName, int]] = {}
        results: dict[ShardName, RetrievalBatch] = {}
        for name, rows in rows_of.items():
            pick = lambda seq: [] if seq is None else [seq[i] for i in rows]  # noqa: E731
            result = self._shards[name].search(text=pick(text), ids=pick(ids), subset_ids=pick(subset_ids),
                                               vector=None if vector is None else np.stack([vector[i] for i in rows]),
                                               top_k=top_k)
            result.indices += self._offsets[name]  # in place, -1 padding included (sharded_search.py:103)
            results[name] = result
            where.update({row: (name, local) for local, row in enumerate(rows)})
        ordered = [results[name][local] for name, local in (where[i] for i in range(len(shard)))]
        cls = type(next(iter(results.values()))) if results else RetrievalBatch
        return cls.stack_samples(ordered)


def _rows_by_shard(shard: list[ShardName]) -> dict[ShardName, list[int]]:
    """Input rows of every corpus, in order of first appearance (the scatter step of sharded_search.py:176-194)."""
    groups: dict[ShardName, list[int]] = collections.OrderedDict()
    for row, name in enumerate(shard):
        groups.setdefault(name, []).append(row)
    return groups
