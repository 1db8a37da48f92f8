"""The cross-process transport (Unix socket server thread + pickled client), with a numpy stand-in for the store so
that it runs without a GPU: a spawned worker process unpickles the client and searches through the socket."""
import multiprocessing as mp
import pickle

import numpy as np
import pytest

from vod_b200.search import B200SearchClient
from vod_b200.transport import SearchServer


def _fake_search(vectors, top_k, mode):
    s = np.tile(vectors.sum(axis=1, keepdims=True), (1, top_k)).astype(np.float32) - np.arange(top_k, dtype=np.float32)
    i = np.tile(np.arange(top_k, dtype=np.int64), (len(vectors), 1)) + (7 if mode == "exact" else 0)
    if vectors.shape[1] == 3:
        raise ValueError("query dimension 3 != index dimension 8")
    return s, i


def _worker(blob, q):
    client = pickle.loads(blob)
    out = client.search(vector=np.ones((4, 8), np.float32), top_k=5)
    ok_ping = client.ping()
    try:
        client.search(vector=np.ones((1, 3), np.float32), top_k=2)
        err = ""
    except RuntimeError as exc:
        err = str(exc)
    q.put((out.scores, out.indices, type(out).__name__, ok_ping, err, out.scores.flags.writeable))


@pytest.mark.timeout(120)
def test_pickled_client_searches_from_another_process():
    server = SearchServer(_fake_search, lambda: True)
    server.start()
    try:
        client = B200SearchClient(master_id=12345, mode="exact", address=server.address, authkey=server.authkey)
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        p = ctx.Process(target=_worker, args=(pickle.dumps(client), q))
        p.start()
        scores, indices, cls_name, ok_ping, err, writable = q.get(timeout=100)
        p.join(timeout=30)
        assert p.exitcode == 0
        exp_s, exp_i = _fake_search(np.ones((4, 8), np.float32), 5, "exact")
        assert np.array_equal(scores, exp_s) and np.array_equal(indices, exp_i)
        assert cls_name == "RetrievalBatch" and ok_ping and writable
        assert "query dimension 3" in err           # server-side errors travel back to the caller
    finally:
        server.stop()
    assert not B200SearchClient(1, pid=-1, address=server.address, authkey=server.authkey).ping()


def test_client_without_master_or_address_fails_loudly():
    import vod_b200

    c = B200SearchClient(master_id=999)
    assert c.ping() is False
    with pytest.raises(vod_b200.VodbError):
        c.search(vector=np.zeros((1, 8), np.float32), top_k=3)
    c2 = pickle.loads(pickle.dumps(B200SearchClient(master_id=999, pid=-1)))
    with pytest.raises(vod_b200.VodbError):
        c2.search(vector=np.zeros((1, 8), np.float32), top_k=3)


@pytest.mark.timeout(60)
def test_concurrent_requests_share_scans():
    """Requests that queue up while a scan runs are served by ONE following scan, each caller gets its own rows."""
    import threading
    import time

    from vod_b200.transport import ScanCoalescer

    seen = []

    def slow_search(vectors, top_k, mode):
        seen.append(len(vectors))
        time.sleep(0.05)
        return _fake_search(vectors, top_k, mode)

    co = ScanCoalescer(slow_search, max_queries=64)
    results = {}

    def caller(t):
        v = np.full((3 + t % 2, 8), float(t), np.float32)
        k = 5 if t != 6 else 7                       # one request with another top_k: never batched with the rest
        results[t] = (v, k, co.submit(v, k, "tensor"))

    threads = [threading.Thread(target=caller, args=(t,)) for t in range(10)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    co.close()
    assert co.n_requests == 10 and co.n_scans < 10 and max(seen) > 4
    for t, (v, k, (s, i)) in results.items():
        exp_s, exp_i = _fake_search(v, k, "tensor")
        assert np.array_equal(s, exp_s) and np.array_equal(i, exp_i) and s.flags.writeable
    with pytest.raises(ValueError):
        ScanCoalescer(slow_search).submit(np.zeros(8, np.float32), 3, None)
    with pytest.raises(RuntimeError):
        co.submit(np.zeros((1, 8), np.float32), 3, None)


@pytest.mark.timeout(60)
def test_shared_scans_are_cut_at_query_tile_multiples():
    """Cost model of the coalescer: a scan costs one corpus pass per `quantum` queries, so when more than one tile
    is waiting the batch stops at the last request that fits a multiple of the tile; nobody is dropped or reordered."""
    import threading
    import time

    from vod_b200.transport import ScanCoalescer

    gate = threading.Event()
    widths = []

    def search(vectors, top_k, mode):
        widths.append(len(vectors))
        if len(widths) == 1:
            gate.wait(5)            # hold the first scan until the other requests have queued up
        return _fake_search(vectors, top_k, mode)

    co = ScanCoalescer(search, max_queries=1024, quantum=128)
    out = {}

    def caller(t):
        v = np.full((32, 8), float(t), np.float32)
        out[t] = (v, co.submit(v, 4, None))

    first = threading.Thread(target=caller, args=(0,))
    first.start()
    while not widths:
        time.sleep(0.005)
    rest = [threading.Thread(target=caller, args=(t,)) for t in range(1, 6)]   # 5 x 32 = 160 queries wait
    for th in rest:
        th.start()
        time.sleep(0.01)            # deterministic arrival order
    while len(co._pending) < 5:
        time.sleep(0.005)
    gate.set()
    for th in [first, *rest]:
        th.join()
    co.close()
    assert widths == [32, 128, 32], widths   # 160 waiting -> one full tile now, the remaining request next
    assert [n for n, _, _ in co.scan_log] == widths and all(ms >= 0 for _, _, ms in co.scan_log)
    for t, (v, (s, i)) in out.items():
        exp_s, exp_i = _fake_search(v, 4, None)
        assert np.array_equal(s, exp_s) and np.array_equal(i, exp_i)


@pytest.mark.timeout(60)
def test_failed_scan_reaches_every_waiter_and_server_keeps_running():
    from vod_b200.transport import ScanCoalescer

    co = ScanCoalescer(_fake_search)
    with pytest.raises(ValueError, match="query dimension 3"):
        co.submit(np.ones((2, 3), np.float32), 2, None)
    s, i = co.submit(np.ones((2, 8), np.float32), 2, None)
    assert s.shape == (2, 2) and i.shape == (2, 2)
    co.close()
