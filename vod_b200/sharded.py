"""Row-sharded search across the GPUs of one box: one process per GPU, `torch.distributed` for the exchange.

The reference's analogue is faiss' `IndexShards` built by `faiss.index_cpu_to_all_gpus(index, co)` with
`co.shard = True` (src/vod_search/faiss_search/server.py:51-54, src/vod_configs/search.py:58,80): every GPU
scans its own rows, per-GPU top-k lists are merged on the host. Here every rank owns one `CorpusStore` holding
the contiguous row block `shard_bounds(n_total, world, rank)`; a search is

    local top-k on every rank (global ids = row_offset + local row, cf. sharded_search.py:103 `indices += offset`)
    -> one all-gather of the [B,k] scores and ids over NCCL/NVLink
    -> `vodb_merge_topk` (radix select + bitonic sort on the GPU) on every rank.

The only data-path collective is that all-gather: B*k*12 bytes per rank (77 KB at B=64, k=100).
"""
from __future__ import annotations

import typing as typ

import numpy as np


def shard_bounds(n_total: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition: rank r owns rows [r*ceil(N/G), min(N, (r+1)*ceil(N/G))).

    Blocks are rounded up to a multiple of 128 rows (one MMA tile) so that every shard start is tile aligned.
    Matches `add_with_ids(xs, arange(i0, i1))` id arithmetic (build_gpu.py:334): global id = offset + local row.
    """
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    per = -(-n_total // world_size)
    per = -(-per // 128) * 128
    lo = min(n_total, rank * per)
    hi = min(n_total, lo + per)
    return lo, hi


class ShardedSearcher:
    """Host-side composition: local search -> all-gather -> merge. Backend-agnostic (NCCL on GPUs, gloo in tests).

    local_search(queries, top_k) -> (scores [B,k] float32, ids [B,k] int64)   tensors on the group's device
    merge(scores [G,B,k], ids [G,B,k], k_out) -> (scores [B,k_out], ids [B,k_out])
    """

    def __init__(self, local_search: typ.Callable, merge: typ.Callable, group: typ.Any = None):
        self.local_search = local_search
        self.merge = merge
        self.group = group

    def world_size(self) -> int:
        import torch.distributed as dist

        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def search(self, queries: typ.Any, top_k: int):
        import torch
        import torch.distributed as dist

        scores, ids = self.local_search(queries, top_k)
        world = self.world_size()
        if world == 1:
            return scores, ids
        B = scores.shape[0]
        all_s = torch.empty((world, B, top_k), dtype=scores.dtype, device=scores.device)
        all_i = torch.empty((world, B, top_k), dtype=ids.dtype, device=ids.device)
        # concatenated layout [world*B, k]: accepted by both the NCCL and the gloo backend
        dist.all_gather_into_tensor(all_s.view(world * B, top_k), scores.contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_i.view(world * B, top_k), ids.contiguous(), group=self.group)
        return self.merge(all_s, all_i, top_k)


class ShardedCorpus:
    """This rank's shard of an `n_total x dim` corpus plus the cross-shard search."""

    def __init__(self, n_total: int, dim: int, dtype: str = "bfloat16", device: int = 0, group: typ.Any = None,
                 rank: int | None = None, world_size: int | None = None):
        import torch.distributed as dist

        from .search import CorpusStore, merge_topk_device

        if rank is None or world_size is None:
            if dist.is_available() and dist.is_initialized():
                rank, world_size = dist.get_rank(group), dist.get_world_size(group)
            else:
                rank, world_size = 0, 1
        self.rank, self.world = rank, world_size
        self.n_total, self.dim = n_total, dim
        self.lo, self.hi = shard_bounds(n_total, world_size, rank)
        self.store = CorpusStore(max(self.hi - self.lo, 0), dim, dtype=dtype, device=device, row_offset=self.lo)
        self.mode: str | None = None
        self._searcher = ShardedSearcher(lambda q, k: self.store.search_device(q, k, mode=self.mode),
                                         merge_topk_device, group)

    def fill_synthetic(self, seed: int, unit_norm: bool = False) -> None:
        """Every rank generates its own rows of the same global synthetic corpus (ids are global)."""
        self.store.fill_synthetic(seed, 0, self.hi - self.lo, unit_norm=unit_norm)

    def add_global(self, rows: np.ndarray, row0: int) -> None:
        """Add the part of global rows [row0, row0+len(rows)) that falls into this shard."""
        a, b = max(row0, self.lo), min(row0 + len(rows), self.hi)
        if a < b:
            self.store.add(rows[a - row0:b - row0], row0=a - self.lo)

    def search_device(self, queries: typ.Any, top_k: int, mode: str | None = None):
        self.mode = mode
        return self._searcher.search(queries, top_k)

    def close(self) -> None:
        self.store.close()
