"""Cross-process access to the GPU-owning search master: a Unix-domain-socket server thread + client stub.

The reference hands picklable `FaissClient(host, port)` objects to DataLoader worker processes, which then reach
the single faiss server process over HTTP with base64-encoded `.npy` payloads inside JSON
(src/vod_search/faiss_search/client.py:64-105, server.py:76-91, src/vod_search/io.py:17-32; workers are started
with forkserver, src/vod_exps/train.py:15). Here the `B200SearchMaster` process owns the CUDA context and the HBM
store; it serves `search` / `ping` requests on an AF_UNIX socket (`multiprocessing.connection`, HMAC-authenticated)
with raw array buffers — no base64, no JSON, no TCP — and the pickled `B200SearchClient` connects to it from any
other process on the box. The store is single-threaded (like the reference's single uvicorn worker, server.py:98),
but instead of serialising the workers' requests one scan each, a dispatcher thread coalesces whatever requests
queued up while the previous scan ran (same top_k / mode / width) into ONE scan: a corpus scan costs the same HBM
traffic for 32 queries as for 256, so N workers share a pass instead of paying N passes.
"""
from __future__ import annotations

import collections
import os
import secrets
import tempfile
import threading
import time
import typing as typ
from multiprocessing.connection import Client, Connection, Listener

import numpy as np

SearchFn = typ.Callable[[np.ndarray, int, typ.Optional[str]], typ.Tuple[np.ndarray, np.ndarray]]


def _send_array(conn: Connection, a: np.ndarray) -> None:
    a = np.ascontiguousarray(a)
    conn.send((a.dtype.str, a.shape))
    conn.send_bytes(memoryview(a).cast("B"))


def _recv_array(conn: Connection) -> np.ndarray:
    dtype, shape = conn.recv()
    buf = conn.recv_bytes()
    return np.frombuffer(buf, dtype=np.dtype(dtype)).reshape(shape).copy()  # fresh, writable, caller-owned


class _Request:
    __slots__ = ("vectors", "top_k", "mode", "done", "result", "error")

    def __init__(self, vectors: np.ndarray, top_k: int, mode: str | None):
        self.vectors, self.top_k, self.mode = vectors, top_k, mode
        self.done = threading.Event()
        self.result: tuple[np.ndarray, np.ndarray] | None = None
        self.error: BaseException | None = None

    def key(self) -> tuple:
        return (self.top_k, self.mode, self.vectors.dtype.str, self.vectors.shape[1:])


class ScanCoalescer:
    """One dispatcher thread in front of the single-threaded store. `submit` blocks until the request's rows have
    been searched. The dispatcher takes the oldest request plus every compatible request already waiting (up to
    `max_queries` rows), runs ONE `search_fn` call over the concatenated rows and hands each caller its slice.
    It never waits for more requests to arrive: batches form only from what queued up during the previous scan, so
    a lone client sees no added latency.

    Width of a shared scan (cost model). A scan costs one pass over the corpus per query tile: up to `quantum` (128)
    queries ride on one pass of the multi-term tensor-core kernel, the 129th costs a second full pass. Queries per
    millisecond therefore peak at multiples of the tile, and a batch of 160 waiting queries is served faster as 128
    now + 32 with whatever arrives next than as one 160-wide scan (measured: 9.5 ms for 160 queries against 3.1 ms for
    32, bench.py `dataloader_workers`). So when more than one tile is waiting the batch is cut at the last request
    boundary that fits a multiple of `quantum`; the rest keeps its place at the head of the queue. `scan_log` keeps
    (queries, requests, milliseconds) of the most recent scans for diagnosis."""

    def __init__(self, search_fn: SearchFn, max_queries: int = 1024, quantum: int = 128):
        self.search_fn = search_fn
        self.max_queries = int(max_queries)
        self.quantum = max(1, int(quantum))
        self.scan_log: collections.deque[tuple[int, int, float]] = collections.deque(maxlen=256)
        self._pending: collections.deque[_Request] = collections.deque()
        self._cv = threading.Condition()
        self._stop = False
        self.n_scans = 0     # search_fn calls issued
        self.n_requests = 0  # requests served
        self._thread = threading.Thread(target=self._run, name="vodb-scan-coalescer", daemon=True)
        self._thread.start()

    def submit(self, vectors: np.ndarray, top_k: int, mode: str | None) -> tuple[np.ndarray, np.ndarray]:
        vectors = np.asarray(vectors)
        if vectors.ndim != 2:  # same check as the direct path (server.py:82-83); fail in the caller, not the batch
            raise ValueError(f"Expected a 2D array of query vectors, got shape {vectors.shape}")
        req = _Request(vectors, int(top_k), mode)
        with self._cv:
            if self._stop:
                raise RuntimeError("search server is shut down")
            self._pending.append(req)
            self._cv.notify()
        req.done.wait()
        if req.error is not None:
            raise req.error
        assert req.result is not None
        return req.result

    def close(self) -> None:
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        self._thread.join(timeout=5.0)

    def _take_batch(self) -> list[_Request]:
        with self._cv:
            while not self._pending and not self._stop:
                self._cv.wait()
            if not self._pending:
                return []
            head = self._pending.popleft()
            batch, rows, key = [head], len(head.vectors), head.key()
            keep: collections.deque[_Request] = collections.deque()
            while self._pending:
                r = self._pending.popleft()
                if r.key() == key and rows + len(r.vectors) <= self.max_queries:
                    batch.append(r)
                    rows += len(r.vectors)
                else:
                    keep.append(r)  # served by a later scan, arrival order preserved
            # cut at a multiple of the query tile: the requests past it would cost one more pass over the corpus
            if rows > self.quantum and rows % self.quantum:
                limit = rows // self.quantum * self.quantum
                at, n_keep = 0, 0
                for r in batch:
                    if at + len(r.vectors) > limit:
                        break
                    at += len(r.vectors)
                    n_keep += 1
                if n_keep >= 1 and at >= self.quantum:
                    for r in reversed(batch[n_keep:]):
                        keep.appendleft(r)  # back to the head of the queue, order preserved
                    batch = batch[:n_keep]
            self._pending = keep
            return batch

    def _run(self) -> None:
        while True:
            batch = self._take_batch()
            if not batch:
                for r in self._drain():
                    r.error = RuntimeError("search server is shut down")
                    r.done.set()
                return
            try:
                rows = batch[0].vectors if len(batch) == 1 else np.concatenate([r.vectors for r in batch], axis=0)
                t0 = time.perf_counter()
                scores, indices = self.search_fn(rows, batch[0].top_k, batch[0].mode)
                self.scan_log.append((len(rows), len(batch), (time.perf_counter() - t0) * 1e3))
                self.n_scans += 1
                self.n_requests += len(batch)
                at = 0
                for r in batch:
                    n = len(r.vectors)  # slices are copied: every caller owns and may mutate its result
                    r.result = (scores[at:at + n].copy(), indices[at:at + n].copy()) if len(batch) > 1 else (scores, indices)
                    at += n
            except BaseException as exc:  # noqa: BLE001 - every waiter of the failed scan gets the error
                for r in batch:
                    r.error = exc
            for r in batch:
                r.done.set()

    def _drain(self) -> list[_Request]:
        with self._cv:
            out = list(self._pending)
            self._pending.clear()
            return out


class SearchServer:
    """Serves `search_fn(vectors, top_k, mode) -> (scores, indices)` on a Unix domain socket. Concurrent requests
    are coalesced into shared scans (`ScanCoalescer`); `coalesce=False` executes them one by one under a lock."""

    def __init__(self, search_fn: SearchFn, ping_fn: typ.Callable[[], bool], address: str | None = None, *,
                 coalesce: bool = True, max_queries: int = 1024, quantum: int = 128):
        self.search_fn = search_fn
        self.ping_fn = ping_fn
        self.address = address or os.path.join(tempfile.gettempdir(), f"vodb-{os.getpid()}-{secrets.token_hex(4)}.sock")
        self.authkey = secrets.token_bytes(16)
        self._lock = threading.Lock()
        self._coalesce, self._max_queries, self._quantum = bool(coalesce), int(max_queries), int(quantum)
        self.coalescer: ScanCoalescer | None = None
        self._listener: Listener | None = None
        self._thread: threading.Thread | None = None
        self._stop = threading.Event()

    def start(self) -> None:
        if os.path.exists(self.address):
            os.unlink(self.address)
        if self._coalesce:
            self.coalescer = ScanCoalescer(self.search_fn, self._max_queries, self._quantum)
        self._listener = Listener(self.address, family="AF_UNIX", authkey=self.authkey)
        self._thread = threading.Thread(target=self._accept_loop, name="vodb-search-server", daemon=True)
        self._thread.start()

    def _search(self, vectors: np.ndarray, top_k: int, mode: str | None) -> tuple[np.ndarray, np.ndarray]:
        if self.coalescer is not None:
            return self.coalescer.submit(vectors, top_k, mode)
        with self._lock:
            return self.search_fn(vectors, top_k, mode)

    def stop(self) -> None:
        self._stop.set()
        if self.coalescer is not None:
            self.coalescer.close()
            self.coalescer = None
        try:  # unblock accept()
            Client(self.address, family="AF_UNIX", authkey=self.authkey).close()
        except Exception:
            pass
        if self._listener is not None:
            self._listener.close()
            self._listener = None
        if os.path.exists(self.address):
            try:
                os.unlink(self.address)
            except OSError:
                pass

    def _accept_loop(self) -> None:
        assert self._listener is not None
        while not self._stop.is_set():
            try:
                conn = self._listener.accept()
            except Exception:
                if self._stop.is_set():
                    return
                continue
            threading.Thread(target=self._serve, args=(conn,), daemon=True).start()

    def _serve(self, conn: Connection) -> None:
        try:
            while not self._stop.is_set():
                try:
                    op, args = conn.recv()
                except (EOFError, OSError):
                    return
                try:
                    if op == "ping":
                        conn.send(("ok", bool(self.ping_fn())))
                    elif op == "search":
                        vectors = _recv_array(conn)
                        scores, indices = self._search(vectors, int(args["top_k"]), args.get("mode"))
                        scores, indices = np.ascontiguousarray(scores), np.ascontiguousarray(indices)
                        # one framed reply: status + both array headers, then the two raw buffers. Nothing that can
                        # fail sits between "ok" and the payload; an I/O error past this point closes the connection
                        # (the client reconnects) instead of desynchronising the stream with an error tuple
                        conn.send(("ok", ((scores.dtype.str, scores.shape), (indices.dtype.str, indices.shape))))
                        try:
                            conn.send_bytes(memoryview(scores).cast("B"))
                            conn.send_bytes(memoryview(indices).cast("B"))
                        except (OSError, ValueError):
                            return
                    else:
                        conn.send(("error", f"unknown op {op!r}"))
                except Exception as exc:  # errors travel back like the reference's HTTP 500 + trace (server.py:89-91)
                    import traceback

                    conn.send(("error", f"{type(exc).__name__}: {exc}\n{traceback.format_exc()}"))
        finally:
            conn.close()


class RemoteSearch:
    """Client side: one lazily opened connection per process."""

    def __init__(self, address: str, authkey: bytes):
        self.address, self.authkey = address, authkey
        self._conn: Connection | None = None
        self._pid = -1

    def _connection(self) -> Connection:
        if self._conn is None or self._pid != os.getpid():
            self._conn = Client(self.address, family="AF_UNIX", authkey=self.authkey)
            self._pid = os.getpid()
        return self._conn

    def ping(self) -> bool:
        try:
            conn = self._connection()
            conn.send(("ping", None))
            status, value = conn.recv()
            return status == "ok" and bool(value)
        except Exception:
            self._conn = None
            return False

    def search(self, vectors: np.ndarray, top_k: int, mode: str | None) -> tuple[np.ndarray, np.ndarray]:
        conn = self._connection()
        try:
            conn.send(("search", {"top_k": int(top_k), "mode": mode}))
            _send_array(conn, np.asarray(vectors))
            status, msg = conn.recv()
            if status != "ok":
                raise RuntimeError(f"search server error: {msg}")
            out = []
            for dtype, shape in msg:  # fresh, writable, caller-owned arrays
                out.append(np.frombuffer(conn.recv_bytes(), dtype=np.dtype(dtype)).reshape(shape).copy())
            return out[0], out[1]
        except (EOFError, OSError):
            self._conn = None
            raise
