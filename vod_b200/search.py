"""Dense search backend: HBM-resident corpus store + exact MIPS top-k behind VOD's search-client interface.

Mirrors the reference's dense backend, same names / argument meaning / error behaviour:

    SearchClient / SearchMaster      src/vod_search/base.py:32-77, 83-200
    FaissClient.search -> RetrievalBatch(scores f32[B,k], indices i64[B,k], labels=None, meta={"time"})
                                     src/vod_search/faiss_search/client.py:64-105
    FaissMaster (context manager, get_client(), unpicklable master)
                                     src/vod_search/faiss_search/client.py:108-186
    build_faiss_index(vectors, factory_string="Flat", ...)   src/vod_search/faiss_search/build.py:12-81

All arithmetic runs in libvodb.so (CUDA, sm_100a) through the C ABI of include/vodb.h; numpy is the
host buffer format, torch tensors are accepted for zero-copy device hand-off. No CPU fallback exists.
"""
from __future__ import annotations

import abc
import ctypes
import itertools
import os
import time
import typing as typ

import numpy as np

from . import _lib
from .retrieval import RetrievalBatch

ShardName = str
SubsetId = str
SectionId = str


def _retrieval_batch_cls():
    """The reference's own RetrievalBatch when `vod_types` is importable, else the local mirror."""
    try:  # pragma: no cover - vod_types is not installed in the build image
        import vod_types as vt  # type: ignore

        return vt.RetrievalBatch
    except Exception:
        return RetrievalBatch


class DoNotPickleError(Exception):
    """Raised when a master (which owns CUDA state) is pickled (base.py:24-29)."""

    def __init__(self, msg: None | str = None):
        super().__init__(msg or "This object cannot be pickled.")


class SearchClient(abc.ABC):
    """A client to interact with a search backend (base.py:32-77)."""

    requires_vectors: bool = True

    def __repr__(self) -> str:
        return f"{type(self).__name__}(requires_vectors={self.requires_vectors})"

    @abc.abstractmethod
    def ping(self) -> bool:
        raise NotImplementedError()

    @abc.abstractmethod
    def search(self, *, text: list[str], vector: None | np.ndarray = None,
               subset_ids: None | list[list[SubsetId]] = None, ids: None | list[list[SectionId]] = None,
               shard: None | list[ShardName] = None, top_k: int = 3) -> RetrievalBatch:
        raise NotImplementedError()

    async def async_search(self, *, text: list[str], vector: None | np.ndarray = None,
                           subset_ids: None | list[list[SubsetId]] = None,
                           ids: None | list[list[SectionId]] = None, shard: None | list[ShardName] = None,
                           top_k: int = 3) -> RetrievalBatch:
        return self.search(text=text, vector=vector, subset_ids=subset_ids, ids=ids, shard=shard, top_k=top_k)


# ---------------------------------------------------------------------------------------------------------
# corpus store
# ---------------------------------------------------------------------------------------------------------

_NP_DTYPES = {np.dtype(np.float32): _lib.F32, np.dtype(np.float16): _lib.F16}


def _np_dtype_code(a: np.ndarray) -> int:
    code = _NP_DTYPES.get(a.dtype)
    if code is None:
        raise TypeError(f"unsupported numpy dtype {a.dtype}; use float32 or float16")
    return code


def _torch_info(t: typ.Any) -> tuple[int, int, bool, int]:
    """(data_ptr, dtype code, is_cuda, device index) of a torch tensor, without importing torch at module load."""
    import torch

    codes = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}
    if t.dtype not in codes:
        raise TypeError(f"unsupported tensor dtype {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t.data_ptr(), codes[t.dtype], t.is_cuda, (t.device.index if t.is_cuda else -1)


def _current_stream_ptr(device: int) -> int:
    """torch's current CUDA stream on `device` (0 = legacy default stream when torch is not in use)."""
    try:
        import torch

        if torch.cuda.is_available():
            return int(torch.cuda.current_stream(device).cuda_stream)
    except Exception:
        pass
    return 0


class CorpusStore:
    """One row shard of passage embeddings resident in HBM (fp32 / bf16 / fp16), owned by this process.

    Replaces the faiss `IndexFlatIP` object (build.py:60: `faiss.index_factory(D, "Flat", METRIC_INNER_PRODUCT)`).
    """

    def __init__(self, n_rows: int, dim: int, dtype: str | int = "float32", device: int = 0, row_offset: int = 0):
        _lib.require_gpu()
        self._lib = _lib.load()
        code = dtype if isinstance(dtype, int) else _lib.DTYPE_NAMES[str(dtype).replace("torch.", "")]
        handle = ctypes.c_void_p()
        _lib.check(self._lib.vodb_store_create(ctypes.byref(handle), int(device), int(n_rows), int(dim), code,
                                              int(row_offset)), "vodb_store_create")
        self._h: ctypes.c_void_p | None = handle
        self._planes_fit: bool | None = None  # float32 store: do the bf16 planes of the tensor modes fit? (decided once)
        self.n_rows, self.dim, self.dtype_code, self.device, self.row_offset = int(n_rows), int(dim), code, int(device), int(row_offset)

    # -- lifecycle
    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            self._lib.vodb_store_destroy(self._h)
            self._h = None

    def __del__(self):  # noqa: D105
        try:
            self.close()
        except Exception:
            pass

    def __getstate__(self):
        raise DoNotPickleError("CorpusStore owns CUDA memory and cannot be pickled.")

    @property
    def handle(self) -> ctypes.c_void_p:
        if self._h is None:
            raise _lib.VodbError("the store has been closed")
        return self._h

    @property
    def ntotal(self) -> int:
        return int(self._lib.vodb_store_ntotal(self.handle))

    @property
    def nbytes(self) -> int:
        return int(self._lib.vodb_store_bytes(self.handle))

    @property
    def dtype(self) -> str:
        return {0: "float32", 1: "bfloat16", 2: "float16"}[self.dtype_code]

    # -- ingest
    def add(self, rows: typ.Any, row0: int | None = None) -> None:
        """Append (or write at `row0`) a [n, dim] block: numpy float32/float16 or a torch tensor (CPU or CUDA)."""
        row0 = self.ntotal if row0 is None else int(row0)
        stream = _current_stream_ptr(self.device)
        if isinstance(rows, np.ndarray):
            a = np.ascontiguousarray(rows)
            if a.ndim != 2 or a.shape[1] != self.dim:
                raise ValueError(f"expected rows of shape [n, {self.dim}], got {a.shape}")
            if a.dtype not in _NP_DTYPES:
                a = a.astype(np.float32)  # build.py:67 `np.asarray(batch).astype(np.float32)`
            _lib.check(self._lib.vodb_store_add(self.handle, a.ctypes.data, _np_dtype_code(a), 0, row0, a.shape[0],
                                               stream), "vodb_store_add")
            return
        ptr, code, is_cuda, dev = _torch_info(rows)
        if rows.dim() != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"expected rows of shape [n, {self.dim}], got {tuple(rows.shape)}")
        if is_cuda and dev != self.device:
            raise ValueError(f"tensor is on cuda:{dev} but the store is on cuda:{self.device}")
        _lib.check(self._lib.vodb_store_add(self.handle, ptr, code, int(is_cuda), row0, int(rows.shape[0]), stream),
                   "vodb_store_add")

    def fill_synthetic(self, seed: int, row0: int = 0, n: int | None = None, unit_norm: bool = False) -> None:
        n = self.n_rows - row0 if n is None else n
        _lib.check(self._lib.vodb_store_fill_synthetic(self.handle, int(seed), int(row0), int(n), int(unit_norm),
                                                      _current_stream_ptr(self.device)), "vodb_store_fill_synthetic")

    def read(self, row0: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dim), np.float32)
        _lib.check(self._lib.vodb_store_read(self.handle, int(row0), int(n), out.ctypes.data, 0,
                                            _current_stream_ptr(self.device)), "vodb_store_read")
        return out

    # -- search
    def prepare_tensor(self) -> bool:
        """Build what the tensor-core modes need ahead of the first search (float32 store: its bf16 planes, +6 bytes
        per element). False when the planes do not fit in HBM; the store then serves `mode="exact"` only."""
        rc = self._lib.vodb_store_prepare_tensor(self.handle, _current_stream_ptr(self.device))
        if rc == -3:  # VODB_ENOMEM
            return False
        _lib.check(rc, "vodb_store_prepare_tensor")
        return True

    def _mode(self, mode: str | int | None, q_code: int | None = None) -> int:
        """Scoring mode. "auto" keeps the float32 query exact on the tensor cores: a 16-bit store takes one term when
        the queries already are in the store dtype, else three 16-bit terms (full fp32 mantissa); a float32 store is
        searched through three bf16 planes of its rows x three query terms ("tensor3": IndexFlatIP parity at the
        fp32 tolerance, 3-4x faster than the fp32 CUDA-core kernel) when the planes fit in HBM (+6 bytes per stored
        element, built on first use), and by the CUDA-core kernel ("exact") otherwise. "tensor" forces one term
        (queries rounded to the store dtype)."""
        if mode is None or mode == "auto":
            if self.dtype_code == _lib.F32:
                if self._planes_fit is None:
                    self._planes_fit = self.prepare_tensor()
                return _lib.MODE_TENSOR_X3 if self._planes_fit else _lib.MODE_EXACT
            return _lib.MODE_TENSOR if q_code == self.dtype_code else _lib.MODE_TENSOR_X3
        if isinstance(mode, int):
            return mode
        return {"exact": _lib.MODE_EXACT, "fp32": _lib.MODE_EXACT, "tensor": _lib.MODE_TENSOR,
                "tensor2": _lib.MODE_TENSOR_X2, "tensor3": _lib.MODE_TENSOR_X3}[mode]

    def search(self, vectors: np.ndarray, top_k: int, mode: str | int | None = None) -> tuple[np.ndarray, np.ndarray]:
        """Host path: numpy [B, dim] in -> (scores f32 [B,k], ids i64 [B,k]); H2D / D2H inside the call."""
        q = np.ascontiguousarray(vectors)
        if q.ndim != 2:
            raise ValueError(f"Expected 2D array, got {q.ndim}D array")  # server.py:82-83
        if q.shape[1] != self.dim:
            raise ValueError(f"query dimension {q.shape[1]} != index dimension {self.dim}")
        if q.dtype not in _NP_DTYPES:
            q = q.astype(np.float32)
        scores = np.empty((q.shape[0], top_k), np.float32)
        ids = np.empty((q.shape[0], top_k), np.int64)
        _lib.check(self._lib.vodb_search(self.handle, q.ctypes.data, _np_dtype_code(q), 0, q.shape[0], int(top_k),
                                        self._mode(mode, _np_dtype_code(q)), scores.ctypes.data, ids.ctypes.data, 0,
                                        _current_stream_ptr(self.device)), "vodb_search")
        return scores, ids

    def search_device(self, vectors: typ.Any, top_k: int, mode: str | int | None = None, out: tuple | None = None):
        """Device path: CUDA torch tensor in, CUDA tensors out; only enqueues on torch's current stream."""
        import torch

        ptr, code, is_cuda, dev = _torch_info(vectors)
        if not is_cuda or dev != self.device:
            raise ValueError(f"search_device needs a tensor on cuda:{self.device}")
        if vectors.dim() != 2 or vectors.shape[1] != self.dim:
            raise ValueError(f"expected queries of shape [B, {self.dim}], got {tuple(vectors.shape)}")
        B = int(vectors.shape[0])
        if out is None:
            scores = torch.empty((B, top_k), dtype=torch.float32, device=vectors.device)
            ids = torch.empty((B, top_k), dtype=torch.int64, device=vectors.device)
        else:
            scores, ids = out
        _lib.check(self._lib.vodb_search(self.handle, ptr, code, 1, B, int(top_k), self._mode(mode, code), scores.data_ptr(),
                                        ids.data_ptr(), 1, _current_stream_ptr(self.device)), "vodb_search")
        return scores, ids

    def check_async(self) -> bool:
        """True if an asynchronous `search_device` since the last check overflowed a candidate list."""
        rc = self._lib.vodb_search_check(self.handle, _current_stream_ptr(self.device))
        _lib.check(rc, "vodb_search_check")
        return rc == 1

    def set_profiling(self, enable: bool) -> None:
        _lib.check(self._lib.vodb_store_set_profiling(self.handle, int(enable)), "vodb_store_set_profiling")

    def profile(self) -> dict[str, float]:
        """Summed CUDA-event timings of the scan kernels since profiling was enabled / last read."""
        arr = (ctypes.c_double * 4)()
        _lib.check(self._lib.vodb_store_profile(self.handle, arr), "vodb_store_profile")
        return {"score_ms": arr[0], "select_ms": arr[1], "score_launches": arr[2], "score_bytes": arr[3]}

    def stats(self) -> dict[str, int]:
        arr = (ctypes.c_int64 * 8)()
        _lib.check(self._lib.vodb_search_stats(self.handle, arr), "vodb_search_stats")
        return {"launches": arr[0], "segments": arr[1], "cap": arr[2], "safe_fallback": arr[3]}


def merge_topk(scores: np.ndarray, ids: np.ndarray, k_out: int, device: int = 0) -> tuple[np.ndarray, np.ndarray]:
    """Merge per-shard results [G, B, k] -> [B, k_out] on the GPU (host buffers)."""
    lib = _lib.load()
    s = np.ascontiguousarray(scores, np.float32)
    i = np.ascontiguousarray(ids, np.int64)
    if s.ndim != 3 or s.shape != i.shape:
        raise ValueError("expected scores/ids of shape [n_lists, B, k]")
    G, B, k_in = s.shape
    out_s = np.empty((B, k_out), np.float32)
    out_i = np.empty((B, k_out), np.int64)
    _lib.check(lib.vodb_merge_topk(int(device), s.ctypes.data, i.ctypes.data, G, B, k_in, int(k_out), out_s.ctypes.data,
                                   out_i.ctypes.data, 0, _current_stream_ptr(device)), "vodb_merge_topk")
    return out_s, out_i


def merge_topk_device(scores: typ.Any, ids: typ.Any, k_out: int):
    """Merge [G, B, k] CUDA tensors -> ([B, k_out] f32, [B, k_out] i64) on torch's current stream."""
    import torch

    lib = _lib.load()
    G, B, k_in = scores.shape
    dev = scores.device.index
    out_s = torch.empty((B, k_out), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((B, k_out), dtype=torch.int64, device=scores.device)
    _lib.check(lib.vodb_merge_topk(dev, scores.contiguous().data_ptr(), ids.contiguous().data_ptr(), G, B, k_in,
                                   int(k_out), out_s.data_ptr(), out_i.data_ptr(), 1, _current_stream_ptr(dev)),
               "vodb_merge_topk")
    return out_s, out_i


# ---------------------------------------------------------------------------------------------------------
# client / master
# ---------------------------------------------------------------------------------------------------------

_MASTERS: dict[int, "B200SearchMaster"] = {}
_master_ids = itertools.count(1)


class B200SearchClient(SearchClient):
    """Drop-in for `FaissClient` (client.py:18-105): `search(vector=[B,D] float32, top_k)` -> RetrievalBatch.

    The client is a light handle onto the master that owns the GPU state. In the master's process it calls the
    store directly; pickled into another process (DataLoader workers, like `FaissClient(host, port)`) it reaches
    the master through the Unix-socket transport of `vod_b200.transport` (address + auth key travel in the pickle).
    """

    requires_vectors: bool = True

    def __init__(self, master_id: int, pid: int | None = None, mode: str | None = None,
                 address: str | None = None, authkey: bytes | None = None):
        self.master_id = master_id
        self.pid = os.getpid() if pid is None else pid
        self.mode = mode
        self.address = address
        self.authkey = authkey
        self._remote = None

    def __repr__(self) -> str:
        return f"{type(self).__name__}[master={self.master_id}](requires_vectors={self.requires_vectors})"

    def __getstate__(self) -> dict:
        return {"master_id": self.master_id, "pid": self.pid, "mode": self.mode, "address": self.address,
                "authkey": self.authkey}

    def __setstate__(self, state: dict) -> None:
        self.__dict__.update(state)
        self._remote = None

    def _local_master(self) -> "B200SearchMaster | None":
        if self.pid == os.getpid():
            return _MASTERS.get(self.master_id)
        return None

    def _remote_search(self):
        if self.address is None or self.authkey is None:
            raise _lib.VodbError(
                "this B200SearchClient has no live master in this process and no transport address: enter the "
                "B200SearchMaster (`with master:`) before handing clients to other processes."
            )
        if self._remote is None:
            from .transport import RemoteSearch

            self._remote = RemoteSearch(self.address, self.authkey)
        return self._remote

    def ping(self, timeout: float = 120) -> bool:  # noqa: ARG002
        """True when the index is loaded and non-empty (server.py:59-66 health check)."""
        m = self._local_master()
        if m is not None:
            return m.store is not None and m.store.ntotal > 0
        if self.pid == os.getpid() or self.address is None:
            return False
        return self._remote_search().ping()

    def search(self, *, vector: np.ndarray, text: None | list[str] = None,  # noqa: ARG002
               subset_ids: None | list[list[SubsetId]] = None,  # noqa: ARG002
               ids: None | list[list[SectionId]] = None,  # noqa: ARG002
               shard: None | list[ShardName] = None,  # noqa: ARG002
               top_k: int = 3, timeout: float = 120) -> RetrievalBatch:  # noqa: ARG002
        """Search the index given a batch of vectors. `text`, `subset_ids`, `ids`, `shard` are accepted and
        ignored exactly like the faiss client does (client.py:68-71)."""
        start_time = time.time()
        m = self._local_master()
        if m is not None:
            if m.store is None:
                raise _lib.VodbError("the master has not been entered (`with master as m:`)")
            mode = self.mode or m.mode
            if hasattr(vector, "is_cuda"):  # torch tensor: queries straight from the encoder, no host round trip
                if vector.is_cuda and isinstance(m.store, CorpusStore) and vector.device.index == m.store.device:
                    scores, indices = _search_cuda_tensor(m.store, vector, top_k, mode)
                else:
                    scores, indices = m.store.search(vector.detach().float().cpu().numpy(), top_k, mode=mode)
            else:
                scores, indices = m.store.search(vector, top_k, mode=mode)
        elif self.pid == os.getpid():
            raise _lib.VodbError("the B200SearchMaster of this client is not active (`with master as m:`)")
        else:
            q = np.ascontiguousarray(vector.detach().float().cpu().numpy() if hasattr(vector, "is_cuda") else vector)
            if q.ndim != 2:
                raise ValueError(f"Expected 2D array, got {q.ndim}D array")  # server.py:82-83
            scores, indices = self._remote_search().search(q, top_k, self.mode)
        return _retrieval_batch_cls().cast(indices=indices, scores=scores, labels=None,
                                           meta={"time": time.time() - start_time})


def _search_cuda_tensor(store: CorpusStore, vector: typ.Any, top_k: int, mode: str | int | None):
    """CUDA tensor in (float32 / bfloat16 / float16, on the store's device), numpy results out; a list overflow on
    the fast schedule (reported by the device flag) falls back to the synchronous entry point."""
    if vector.dim() != 2:
        raise ValueError(f"Expected 2D array, got {vector.dim()}D array")  # server.py:82-83
    q = vector.detach().contiguous()
    scores, indices = store.search_device(q, top_k, mode=mode)
    host = scores.cpu().numpy(), indices.cpu().numpy()
    if store.check_async():
        return store.search(q.float().cpu().numpy(), top_k, mode=mode)
    return host


class B200SearchMaster:
    """Drop-in for `FaissMaster` (client.py:108-186) + `SearchMaster` (base.py:83-200).

        with B200SearchMaster(vectors, dtype="bfloat16") as master:
            client = master.get_client()
            result = client.search(vector=q, top_k=100)

    `__enter__` uploads the vectors into an HBM store on `device` (the analogue of the server process reading the
    index file, server.py:39-54); `__exit__` frees it. Masters refuse pickling (base.py:188-200).
    """

    def __init__(self, vectors: typ.Any = None, *, dtype: str = "float32", device: int = 0, mode: str | None = None,
                 row_offset: int = 0, add_batch_size: int = 1 << 18, skip_setup: bool = False,
                 free_resources: bool = False, store: CorpusStore | None = None, serve: bool = True,
                 devices: typ.Sequence[int] | None = None):
        self.vectors = vectors
        self.dtype = dtype
        self.device = device
        self.mode = mode
        self.row_offset = row_offset
        self.add_batch_size = add_batch_size  # build_gpu.py:294 `add_batch_size=2**18`
        self.skip_setup = skip_setup
        self.free_resources = free_resources
        self.store: CorpusStore | None = store
        self._owns_store = store is None
        self.master_id = next(_master_ids)
        self.devices = list(devices) if devices else [device]  # >1: row-shard over several GPUs in this process
        self.serve = serve  # expose the store to other processes (DataLoader workers) over a Unix socket
        self._server = None

    # -- context manager (base.py:100-116)
    def __enter__(self) -> "B200SearchMaster":
        if not self.skip_setup:
            self._setup()
        _MASTERS[self.master_id] = self
        if self.serve and self.store is not None:
            from .transport import SearchServer

            self._server = SearchServer(lambda v, k, mode: self.store.search(v, k, mode=mode or self.mode),
                                        lambda: self.store is not None and self.store.ntotal > 0)
            self._server.start()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb) -> None:  # noqa: ANN001
        _MASTERS.pop(self.master_id, None)
        if self._server is not None:
            self._server.stop()
            self._server = None
        if self.store is not None and self._owns_store:
            self.store.close()
            self.store = None

    def _setup(self) -> None:
        if self.store is not None:
            return
        if self.vectors is None:
            raise ValueError("B200SearchMaster needs `vectors` (or an existing `store`)")
        self.store = build_b200_index(self.vectors, dtype=self.dtype, device=self.device, row_offset=self.row_offset,
                                      add_batch_size=self.add_batch_size, devices=self.devices)

    def get_client(self) -> B200SearchClient:
        srv = self._server
        return B200SearchClient(self.master_id, mode=self.mode, address=srv.address if srv else None,
                                authkey=srv.authkey if srv else None)

    @property
    def service_name(self) -> str:
        return f"b200_search_master-{self.master_id}"

    @property
    def service_info(self) -> str:
        return f"B200Search[cuda:{self.device}]"

    def __getstate__(self):
        raise DoNotPickleError(f"{type(self).__name__} is not pickleable. To use in multiprocessing, "
                               "using a client instead (`server.get_client()`).")

    def __setstate__(self, state):  # noqa: ANN001
        raise DoNotPickleError(f"{type(self).__name__} is not pickleable.")


def _slice_rows(vectors: typ.Any, start: int, stop: int) -> np.ndarray:
    """`vt.slice_arrays_sequence(vectors, slice(i, j))` semantics (lazy_array.py:165-172): any sliceable sequence."""
    try:
        block = vectors[start:stop]
    except TypeError:
        block = [vectors[i] for i in range(start, stop)]
    if hasattr(block, "detach"):
        return block
    return np.asarray(block)


def build_b200_index(vectors: typ.Any, *, dtype: str = "float32", device: int = 0, row_offset: int = 0,
                     add_batch_size: int = 1 << 18, factory_string: str = "Flat",
                     devices: typ.Sequence[int] | None = None) -> typ.Any:
    """Build the HBM store from a sequence of 1-D vectors — the analogue of `build_faiss_index`
    (build.py:12-81) for `factory_string="Flat"` with the inner-product metric.

    Checks mirror the reference: only 1-D vectors (build.py:21-23); after the add loop the store must hold
    `len(vectors)` rows of the right width (build.py:75-79).
    """
    if factory_string != "Flat":
        raise ValueError(f"only the exact `Flat` (IndexFlatIP) factory is supported, got `{factory_string}`")
    n = len(vectors)
    if n == 0:
        raise ValueError("cannot build an index from an empty sequence of vectors")
    vector_shape = tuple(np.shape(vectors[0]))
    if len(vector_shape) > 1:
        raise ValueError(f"Only 1D vectors can be handled. Found shape `{vector_shape}`")
    dim = int(vector_shape[-1])
    if devices is not None and len(devices) > 1:  # `index_cpu_to_all_gpus(index, co.shard=True)` analogue, one process
        from .sharded import MultiGpuStore

        store = MultiGpuStore(n, dim, dtype=dtype, devices=devices)
    else:
        store = CorpusStore(n, dim, dtype=dtype, device=devices[0] if devices else device, row_offset=row_offset)
    for i in range(0, n, add_batch_size):
        batch = _slice_rows(vectors, i, min(n, i + add_batch_size))
        store.add(batch, row0=i)
    if store.ntotal != n or store.dim != dim:
        raise ValueError(f"Index size doesn't match the size of the vectors. Found vectors: `{dim}`, "
                         f"index: `{store.ntotal, store.dim}`")
    return store
