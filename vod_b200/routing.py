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
        queries, lookup = _scatter_queries(text=text, shard=shard, vector=vector, subset_ids=subset_ids, ids=ids)
        results_by_shard = {}
        for shard_name, query in queries.items():
            result = self.shards[shard_name].search(
                text=query["text"], ids=query["ids"], subset_ids=query["subset_ids"],
                vector=np.stack(query["vector"]) if vector is not None else None, top_k=top_k)
            result.indices += self.offsets[shard_name]  # in place, sharded_search.py:103
            results_by_shard[shard_name] = result
        gathered = [results_by_shard[name][j] for name, j in lookup]
        cls = type(next(iter(results_by_shard.values()))) if results_by_shard else RetrievalBatch
        return cls.stack_samples(gathered)


def _scatter_queries(text, shard, vector=None, subset_ids=None, ids=None):
    """sharded_search.py:176-194: group the rows by shard name, remember (shard, local row) per input row."""
    shards: dict = collections.defaultdict(lambda: collections.defaultdict(list))
    lookup = []
    for i, shard_name in enumerate(shard):
        shards[shard_name]["text"].append(text[i])
        shards[shard_name]["local_rank"].append(i)
        lookup.append((shard_name, len(shards[shard_name]["text"]) - 1))
        if subset_ids is not None:
            shards[shard_name]["subset_ids"].append(subset_ids[i])
        if ids is not None:
            shards[shard_name]["ids"].append(ids[i])
        if vector is not None:
            shards[shard_name]["vector"].append(vector[i])
    return dict(shards), lookup
