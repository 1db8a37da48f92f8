"""Cross-process access to the GPU-owning search master: a Unix-domain-socket server thread + client stub.

The reference hands picklable `FaissClient(host, port)` objects to DataLoader worker processes, which then reach
the single faiss server process over HTTP with base64-encoded `.npy` payloads inside JSON
(src/vod_search/faiss_search/client.py:64-105, server.py:76-91, src/vod_search/io.py:17-32; workers are started
with forkserver, src/vod_exps/train.py:15). Here the `B200SearchMaster` process owns the CUDA context and the HBM
store; it serves `search` / `ping` requests on an AF_UNIX socket (`multiprocessing.connection`, HMAC-authenticated)
with raw array buffers — no base64, no JSON, no TCP — and the pickled `B200SearchClient` connects to it from any
other process on the box. Requests are executed one at a time (the store is single-threaded, like the reference's
single uvicorn worker, server.py:98).
"""
from __future__ import annotations

import os
import secrets
import tempfile
import threading
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


class SearchServer:
    """Serves `search_fn(vectors, top_k, mode) -> (scores, indices)` on a Unix domain socket."""

    def __init__(self, search_fn: SearchFn, ping_fn: typ.Callable[[], bool], address: str | None = None):
        self.search_fn = search_fn
        self.ping_fn = ping_fn
        self.address = address or os.path.join(tempfile.gettempdir(), f"vodb-{os.getpid()}-{secrets.token_hex(4)}.sock")
        self.authkey = secrets.token_bytes(16)
        self._lock = threading.Lock()
        self._listener: Listener | None = None
        self._thread: threading.Thread | None = None
        self._stop = threading.Event()

    def start(self) -> None:
        if os.path.exists(self.address):
            os.unlink(self.address)
        self._listener = Listener(self.address, family="AF_UNIX", authkey=self.authkey)
        self._thread = threading.Thread(target=self._accept_loop, name="vodb-search-server", daemon=True)
        self._thread.start()

    def stop(self) -> None:
        self._stop.set()
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
                        with self._lock:
                            scores, indices = self.search_fn(vectors, int(args["top_k"]), args.get("mode"))
                        conn.send(("ok", None))
                        _send_array(conn, scores)
                        _send_array(conn, indices)
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
            return _recv_array(conn), _recv_array(conn)
        except (EOFError, OSError):
            self._conn = None
            raise
