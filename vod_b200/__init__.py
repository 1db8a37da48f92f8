"""vod_b200 — B200-native dense retrieval for VOD's dynamic-retrieval hot path.

Exact maximum-inner-product top-k search over an HBM-resident passage-embedding store, followed by the
dataloader's labeled priority sampling, as hand-written CUDA (sm_100a) behind the reference's own
`vod_search` client interface and `vod_dataloaders` sampling functions. See DESIGN.md / INTEGRATION.md.
"""
from . import _lib
from ._lib import VodbError, VodbUnavailableError
from .pipeline import DenseRetrievalSampler, sample_device
from .retrieval import RetrievalBatch, RetrievalSample, RetrievalTuple
from .sampling import (PrioritySampledSections, labeled_priority_sampling, priority_sampling_1d,
                       sample_search_results)
from .search import (B200SearchClient, B200SearchMaster, CorpusStore, DoNotPickleError, SearchClient,
                     build_b200_index, merge_topk, merge_topk_device)
from .hybrid import async_hybrid_search, merge_search_results, normalize_search_scores_
from .collate import flatten_samples, gather_values_by_indices, replace_negative_indices_
from .routing import ShardedSearchClient
from .sharded import MultiGpuStore, ShardedCorpus, ShardedSearcher, shard_bounds
from .config import B200FactoryConfig, B200FactoryDiff, build_b200_search
from .zarr_io import ZarrV2Array, open_vectors, write_zarr_v2

__all__ = [
    "B200SearchClient", "B200SearchMaster", "CorpusStore", "DenseRetrievalSampler", "sample_device", "DoNotPickleError", "PrioritySampledSections",
    "RetrievalBatch", "RetrievalSample", "RetrievalTuple", "SearchClient", "MultiGpuStore", "ShardedCorpus", "ShardedSearchClient",
    "ShardedSearcher", "async_hybrid_search", "merge_search_results", "normalize_search_scores_",
    "VodbError", "VodbUnavailableError", "build_b200_index", "labeled_priority_sampling", "merge_topk",
    "merge_topk_device", "priority_sampling_1d", "sample_search_results", "shard_bounds",
    "B200FactoryConfig", "B200FactoryDiff", "build_b200_search", "ZarrV2Array", "open_vectors", "write_zarr_v2",
    "flatten_samples", "gather_values_by_indices", "replace_negative_indices_",
]
