"""The IndexFlatIP restatement (oracle/flat_ip.py) against first principles — the reference holds no golden
vectors for the search path (parity unpinned), so the oracle is checked against a float64 brute
force, exact integer data and the documented faiss conventions."""
import numpy as np
import pytest

from oracle import flat_ip
from tests.helpers import int_valued


def brute(xb, xq, k):
    ip = xq.astype(np.float64) @ xb.astype(np.float64).T
    out_s = np.full((len(xq), k), -flat_ip.FLT_MAX)
    out_i = np.full((len(xq), k), -1, np.int64)
    for q in range(len(xq)):
        order = sorted(range(xb.shape[0]), key=lambda r: (-ip[q, r], r))[:k]
        out_s[q, :len(order)] = ip[q, order]
        out_i[q, :len(order)] = order
    return out_s, out_i


@pytest.mark.parametrize("n,d,q,k", [(1, 8, 3, 4), (50, 16, 5, 10), (3000, 32, 21, 100), (70000, 8, 2, 7)])
def test_exact_on_integer_data(n, d, q, k):
    rng = np.random.default_rng(n)
    xb, xq = int_valued(rng, (n, d)), int_valued(rng, (q, d))
    s, i = flat_ip.search(xb, xq, k)
    bs, bi = brute(xb, xq, k)
    assert np.array_equal(i, bi)          # ties broken by (score desc, id asc)
    assert np.array_equal(s.astype(np.float64), bs)
    assert s.dtype == np.float32 and i.dtype == np.int64


def test_descending_and_padding_conventions():
    rng = np.random.default_rng(0)
    xb, xq = rng.standard_normal((5, 12), dtype=np.float32), rng.standard_normal((2, 12), dtype=np.float32)
    s, i = flat_ip.search(xb, xq, 8)
    assert (np.diff(s[:, :5], axis=1) <= 0).all()
    assert (i[:, 5:] == -1).all() and (s[:, 5:] == -np.finfo(np.float32).max).all()  # faiss: -FLT_MAX / -1
    with pytest.raises(ValueError):
        flat_ip.search(xb, xq[0], 3)  # server.py:82-83: 2D input required


def test_float_data_within_north_star_tolerance():
    rng = np.random.default_rng(1)
    xb, xq = rng.standard_normal((20000, 96), dtype=np.float32), rng.standard_normal((33, 96), dtype=np.float32)
    s, i = flat_ip.search(xb, xq, 100)
    rs, ri = flat_ip.search_f64(xb, xq, 100)
    rep = flat_ip.compare_topk(xb, xq, s, i, rs.astype(np.float32), ri)
    assert rep["ok"], rep
    assert flat_ip.recall_at_k(i, ri) > 0.999


def test_row_offset_and_merge():
    rng = np.random.default_rng(2)
    xb, xq = int_valued(rng, (900, 16)), int_valued(rng, (4, 16))
    s, i = flat_ip.search(xb, xq, 20)
    sa, ia = flat_ip.search(xb[:512], xq, 20)
    sb, ib = flat_ip.search(xb[512:], xq, 20, row_offset=512)
    ms, mi = flat_ip.merge_sorted(sa, ia, sb, ib, 20)
    assert np.array_equal(mi, i) and np.array_equal(ms, s)
