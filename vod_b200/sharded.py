"""Row-sharded search across the GPUs of one box: one process per GPU, `torch.distributed` for the exchange.

The reference's analogue is faiss' `IndexShards` built by `faiss.index_cpu_to_all_gpus(index, co)` with
`co.shard = True` (src/vod_search/faiss_search/server.py:51-54, src/vod_configs/search.py:58,80): every GPU
scans its own rows, per-GPU top-k lists are merged on the host. Here every rank owns one `CorpusStore` holding
the contiguous row block `shard_bounds(n_total, world, rank)`; a search is

    local top-k on every rank (global ids = row_offset + local row, cf. sharded_search.py:103 `indices += offset`)
    -> exchange of the [B,k] scores and ids between the ranks
    -> merge (exact k-selection on the GPU, csrc/select.cu) on every rank.

Two exchange implementations, same results:
  * exchange="p2p" (default on GPUs): fused into the kernels. The final select kernel of every rank stores its list
    straight into every peer's gather buffer (CUDA-IPC peer-mapped memory, NVLink stores) as epoch-tagged 8-byte
    words (NCCL-LL style: tag and payload in one atomic store, no fence); the merge kernel spins on the tags of the
    entries it reads. No collective library call, no extra launch
    (`vodb_search_sharded`, include/vodb.h).
  * exchange="nccl": one `all_gather_into_tensor` per array over NCCL, then `vodb_merge_topk`.
The exchanged payload is B*k*12 bytes per rank (77 KB at B=64, k=100).
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
        return self.merge_gathered(scores, ids, top_k)

    def merge_gathered(self, scores: typ.Any, ids: typ.Any, top_k: int):
        """All-gather this rank's [B,k] list and merge the `world` lists (every rank gets the same result)."""
        import torch
        import torch.distributed as dist

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
                 rank: int | None = None, world_size: int | None = None, exchange: str = "p2p",
                 max_queries: int = 8192, max_k: int = 1000):
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
        self.group = group
        self.exchange = exchange if world_size > 1 else "none"
        self._xchg = None
        self._xchg_limits = (max_queries, max_k)
        if self.exchange == "p2p":
            self._setup_p2p(device, max_queries, max_k)

    def _setup_p2p(self, device: int, max_queries: int, max_k: int) -> None:
        """Create this rank's peer-mapped exchange buffer and connect to the peers' (CUDA IPC handles are
        all-gathered with torch.distributed)."""
        import ctypes

        import torch
        import torch.distributed as dist

        from . import _lib

        lib = _lib.load()
        handle = (ctypes.c_ubyte * 64)()
        x = ctypes.c_void_p()
        _lib.check(lib.vodb_xchg_create(ctypes.byref(x), device, self.rank, self.world, max_queries, max_k, handle),
                   "vodb_xchg_create")
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=f"cuda:{device}")
        everyone = torch.empty(self.world * 64, dtype=torch.uint8, device=f"cuda:{device}")
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        blob = bytes(everyone.cpu().tolist())
        _lib.check(lib.vodb_xchg_connect(x, blob), "vodb_xchg_connect")
        dist.barrier(group=self.group)
        self._xchg = x

    def fill_synthetic(self, seed: int, unit_norm: bool = False) -> None:
        """Every rank generates its own rows of the same global synthetic corpus (ids are global)."""
        self.store.fill_synthetic(seed, 0, self.hi - self.lo, unit_norm=unit_norm)

    def add_global(self, rows: np.ndarray, row0: int) -> None:
        """Add the part of global rows [row0, row0+len(rows)) that falls into this shard."""
        a, b = max(row0, self.lo), min(row0 + len(rows), self.hi)
        if a < b:
            self.store.add(rows[a - row0:b - row0], row0=a - self.lo)

    @property
    def ntotal(self) -> int:
        """Rows of the whole corpus (all shards), like `index.ntotal` of the reference's IndexShards."""
        return self.n_total

    @property
    def device(self) -> int:
        return self.store.device

    def search_device(self, queries: typ.Any, top_k: int, mode: str | None = None, safe: bool = False, out=None,
                      exchange: str | None = None):
        """Merged top-k over all shards for CUDA queries; returns (scores [B,k] f32, ids [B,k] i64) CUDA tensors,
        identical on every rank. Only enqueues work on torch's current stream. `exchange` overrides the corpus
        default for this call ("nccl": all-gather + merge kernel instead of the fused peer-store exchange)."""
        self.mode = mode
        how = self.exchange if exchange is None or self.world == 1 else exchange
        if how != "p2p":
            return self._searcher.search(queries, top_k)
        if self._xchg is None:
            raise ValueError("this corpus was created without the p2p exchange")
        import torch

        from . import _lib
        from .search import _current_stream_ptr, _torch_info

        ptr, code, is_cuda, dev = _torch_info(queries)
        if not is_cuda or dev != self.store.device:
            raise ValueError(f"search_device needs a tensor on cuda:{self.store.device}")
        B = int(queries.shape[0])
        if B * top_k > self._xchg_limits[0] * self._xchg_limits[1]:
            raise ValueError("batch x top_k exceeds the exchange buffer (raise max_queries / max_k)")
        if out is None:
            scores = torch.empty((B, top_k), dtype=torch.float32, device=queries.device)
            ids = torch.empty((B, top_k), dtype=torch.int64, device=queries.device)
        else:
            scores, ids = out
        lib = _lib.load()
        _lib.check(lib.vodb_search_sharded(self.store.handle, self._xchg, ptr, code, 1, B, int(top_k),
                                           self.store._mode(mode, code), int(safe), scores.data_ptr(), ids.data_ptr(), 1,
                                           _current_stream_ptr(self.store.device)), "vodb_search_sharded")
        return scores, ids

    def search(self, vectors: np.ndarray, top_k: int, mode: str | int | None = None) -> tuple[np.ndarray, np.ndarray]:
        """Host path with the `CorpusStore.search` signature, so that a `B200SearchMaster(store=corpus)` /
        `B200SearchClient` serves the sharded corpus unchanged: numpy [B, dim] in, merged (scores f32 [B,k], ids i64
        [B,k]) out, H2D / D2H inside the call. SPMD: every rank calls it with the same batch. A list overflow on any
        shard re-runs the batch on the overflow-proof schedule on all ranks (decided from the exchanged flags)."""
        from . import _lib
        from .search import _current_stream_ptr, _np_dtype_code

        q = np.ascontiguousarray(vectors)
        if q.ndim != 2:
            raise ValueError(f"Expected 2D array, got {q.ndim}D array")  # server.py:82-83
        if q.shape[1] != self.dim:
            raise ValueError(f"query dimension {q.shape[1]} != index dimension {self.dim}")
        if q.dtype not in (np.float32, np.float16):
            q = q.astype(np.float32)
        if self.exchange != "p2p":
            if self.world == 1:
                return self.store.search(q, top_k, mode=mode)
            import torch

            for safe in (False, True):  # unfused path: flags agreed on with an all-reduce
                if safe:
                    self.mode = mode
                    s_np, i_np = self.store.search(q, top_k, mode=mode)  # overflow-proof fallback inside
                    dev = f"cuda:{self.store.device}"
                    s, i = self._searcher.merge_gathered(torch.from_numpy(s_np).to(dev), torch.from_numpy(i_np).to(dev), top_k)
                else:
                    s, i = self.search_device(torch.from_numpy(q).to(f"cuda:{self.store.device}"), top_k, mode=mode)
                out = s.cpu().numpy(), i.cpu().numpy()
                if safe or not self.any_overflow():
                    return out
        B = q.shape[0]
        if B * top_k > self._xchg_limits[0] * self._xchg_limits[1]:
            raise ValueError("batch x top_k exceeds the exchange buffer (raise max_queries / max_k)")
        scores = np.empty((B, top_k), np.float32)
        ids = np.empty((B, top_k), np.int64)
        code = _np_dtype_code(q)
        lib = _lib.load()
        _lib.check(lib.vodb_search_sharded(self.store.handle, self._xchg, q.ctypes.data, code, 0, B, int(top_k),
                                           self.store._mode(mode, code), 0, scores.ctypes.data, ids.ctypes.data, 0,
                                           _current_stream_ptr(self.store.device)), "vodb_search_sharded")
        return scores, ids

    def any_overflow(self) -> bool:
        """True if a candidate list overflowed on ANY rank since the last check (then re-run with safe=True). With
        the fused exchange every rank already holds the OR of all shards' flags (they travel with the lists); the
        NCCL path agrees on it with an all-reduce."""
        local = self.store.check_async()
        if self.world == 1 or self.exchange == "p2p":
            return bool(local)
        import torch
        import torch.distributed as dist

        flag = torch.tensor([1 if local else 0], dtype=torch.int32, device=f"cuda:{self.store.device}")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        return bool(flag.item())

    def close(self) -> None:
        if self._xchg is not None:
            import torch.distributed as dist

            from . import _lib

            dist.barrier(group=self.group)  # nobody may still be storing into this rank's buffer
            _lib.load().vodb_xchg_destroy(self._xchg)
            self._xchg = None
        self.store.close()


class MultiGpuStore:
    """All GPUs of the box driven from ONE process: the drop-in analogue of the reference's
    `faiss.index_cpu_to_all_gpus(index, co)` with `co.shard = True` inside its single server process
    (src/vod_search/faiss_search/server.py:51-54). One `CorpusStore` per device holds the contiguous row block
    `shard_bounds(n, G, g)`; a search enqueues the G local scans back to back (they run concurrently, one stream
    per device), copies the G small [B,k] results to the first device over NVLink and merges them there
    (`vodb_merge_topk`). Same `search(vectors, top_k)` / `ntotal` / `close()` surface as `CorpusStore`.
    """

    def __init__(self, n_rows: int, dim: int, dtype: str = "bfloat16", devices: typ.Sequence[int] = (0,)):
        from .search import CorpusStore

        self.devices = list(devices)
        self.n_rows, self.dim = int(n_rows), int(dim)
        self.bounds = [shard_bounds(self.n_rows, len(self.devices), g) for g in range(len(self.devices))]
        self.stores = [CorpusStore(hi - lo, dim, dtype=dtype, device=d, row_offset=lo)
                       for d, (lo, hi) in zip(self.devices, self.bounds)]
        self.device = self.devices[0]
        self._bufs: dict = {}

    @property
    def ntotal(self) -> int:
        return sum(st.ntotal for st in self.stores)

    @property
    def dtype(self) -> str:
        return self.stores[0].dtype

    def add(self, rows: typ.Any, row0: int | None = None) -> None:
        """Write global rows [row0, row0+len(rows)) into the shards they belong to."""
        row0 = self.ntotal if row0 is None else int(row0)
        n = len(rows)
        for st, (lo, hi) in zip(self.stores, self.bounds):
            a, b = max(row0, lo), min(row0 + n, hi)
            if a < b:
                st.add(rows[a - row0:b - row0], row0=a - lo)

    def fill_synthetic(self, seed: int, unit_norm: bool = False) -> None:
        for st, (lo, hi) in zip(self.stores, self.bounds):
            st.fill_synthetic(seed, 0, hi - lo, unit_norm=unit_norm)

    def _buffers(self, B: int, k: int):
        """Per-(B, k) work buffers, allocated once: queries on every device, each device's [B,k] result, the gathered
        [G,B,k] lists and the merged result on the first device, and the pinned host mirror of the merged result."""
        import torch

        key = (B, k)
        buf = self._bufs.get(key)
        if buf is None:
            if len(self._bufs) > 8:
                self._bufs.clear()
            dev0 = torch.device(f"cuda:{self.device}")
            live = [st for st in self.stores if st.ntotal]
            buf = {
                "q": [torch.empty((B, self.dim), dtype=torch.float32, device=f"cuda:{st.device}") for st in live],
                "out": [(torch.empty((B, k), dtype=torch.float32, device=f"cuda:{st.device}"),
                         torch.empty((B, k), dtype=torch.int64, device=f"cuda:{st.device}")) for st in live],
                "all_s": torch.empty((len(live), B, k), dtype=torch.float32, device=dev0),
                "all_i": torch.empty((len(live), B, k), dtype=torch.int64, device=dev0),
                "host_s": torch.empty((B, k), dtype=torch.float32).pin_memory(),
                "host_i": torch.empty((B, k), dtype=torch.int64).pin_memory(),
                "done": [torch.cuda.Event() for _ in live],
            }
            self._bufs[key] = buf
        return buf

    def search(self, vectors: np.ndarray, top_k: int, mode: str | int | None = None) -> tuple[np.ndarray, np.ndarray]:
        """Host queries in, merged host results out. Everything is enqueued first — H2D of the queries and the local
        scan on every device (they run concurrently), peer copies of the [B,k] lists into the first device's gather
        buffer (ordered by events, no host wait), the merge kernel, one D2H into pinned memory — and the host waits
        once, at the end."""
        import torch

        from .search import merge_topk_device

        q = np.ascontiguousarray(vectors)
        if q.ndim != 2:
            raise ValueError(f"Expected 2D array, got {q.ndim}D array")
        if q.shape[1] != self.dim:
            raise ValueError(f"query dimension {q.shape[1]} != index dimension {self.dim}")
        if q.dtype != np.float32:
            q = q.astype(np.float32)
        B = q.shape[0]
        live = [st for st in self.stores if st.ntotal]
        buf = self._buffers(B, int(top_k))
        q_host = torch.from_numpy(q)
        dev0 = torch.device(f"cuda:{self.device}")
        for g, st in enumerate(live):  # enqueue everything first: the G scans overlap
            with torch.cuda.device(st.device):
                buf["q"][g].copy_(q_host, non_blocking=True)
                st.search_device(buf["q"][g], top_k, mode=mode, out=buf["out"][g])
                buf["done"][g].record()
        with torch.cuda.device(dev0):
            stream0 = torch.cuda.current_stream(dev0)
            for g in range(len(live)):
                stream0.wait_event(buf["done"][g])
                buf["all_s"][g].copy_(buf["out"][g][0], non_blocking=True)   # peer copy over NVLink
                buf["all_i"][g].copy_(buf["out"][g][1], non_blocking=True)
            ms, mi = merge_topk_device(buf["all_s"], buf["all_i"], top_k)
            buf["host_s"].copy_(ms, non_blocking=True)
            buf["host_i"].copy_(mi, non_blocking=True)
            stream0.synchronize()
        # read (and clear) EVERY store's sticky flag: a short-circuit would leave stale flags for the next search
        flags = [st.check_async() for st in live]
        if not any(flags):
            return buf["host_s"].numpy().copy(), buf["host_i"].numpy().copy()
        # rare: a list overflowed somewhere -> every shard again through the synchronous entry point, which falls
        # back to the overflow-proof schedule by itself
        parts = [st.search(q, top_k, mode=mode) for st in live]
        with torch.cuda.device(dev0):
            all_s = torch.stack([torch.from_numpy(s) for s, _ in parts]).to(dev0)
            all_i = torch.stack([torch.from_numpy(i) for _, i in parts]).to(dev0)
            ms, mi = merge_topk_device(all_s, all_i, top_k)
            return ms.cpu().numpy(), mi.cpu().numpy()

    def stats(self) -> dict[str, int]:
        return self.stores[0].stats()

    def close(self) -> None:
        for st in self.stores:
            st.close()
