"""BASELINE.json's full sizes, checked through size-independent properties (no CPU scan of 10M rows):
planted nearest neighbours, float64 re-scoring of the returned ids from the counter-based generator, sortedness,
shard-and-merge == unsharded, tensor-core path == fp32 CUDA-core path, batch-size independence."""
import numpy as np
import pytest

import vod_b200
from oracle import flat_ip

pytestmark = pytest.mark.gpu

N, D, K = 10_000_000, 768, 100
SEED = 1234


def _queries(nq, dtype, seed=5678):
    import torch

    g = torch.Generator().manual_seed(seed)
    q = torch.randn((nq, D), generator=g, dtype=torch.float32)
    td = torch.bfloat16 if dtype == "bfloat16" else torch.float16
    return q.to(td).to(torch.float32).numpy()


@pytest.fixture(scope="module")
def big_store():
    st = vod_b200.CorpusStore(N, D, dtype="bfloat16")
    st.fill_synthetic(SEED)
    yield st
    st.close()


def _rescore(twin, ids, xq, dtype_code):
    """float64 scores of the returned ids, rows regenerated on the CPU from the counter-based generator."""
    out = np.zeros(ids.shape, np.float64)
    for q in range(ids.shape[0]):
        rows = np.stack([twin.synth_rows(SEED, int(r), 1, D, dtype=dtype_code)[0] for r in ids[q]])
        out[q] = rows.astype(np.float64) @ xq[q].astype(np.float64)
    return out


def test_config2_10m_bf16_q64_properties(big_store, twin):
    xq = _queries(64, "bfloat16")
    s, i = big_store.search(xq, K, mode="tensor")
    assert big_store.stats()["safe_fallback"] == 0
    assert (np.diff(s, axis=1) <= 0).all()
    assert i.min() >= 0 and i.max() < N
    assert all(len(set(row)) == K for row in i)
    # scores are the true inner products of the returned rows (checked for 8 queries: 800 regenerated rows)
    true = _rescore(twin, i[:8], xq[:8], 1)
    assert np.abs(s[:8] - true).max() <= 1e-5 * np.abs(true).max()
    # the fp32 CUDA-core kernel over the same store returns the same neighbours (independent code path)
    s2, i2 = big_store.search(xq, K, mode="exact")
    assert flat_ip.recall_at_k(i, i2) >= 0.999
    assert np.abs(s - s2).max() <= 2e-5 * np.abs(s2).max()
    # batch-size independence: the same queries inside a 300-query batch (query tile 256 path) give the same result
    xq_big = np.concatenate([xq, _queries(236, "bfloat16", seed=99)])
    s3, i3 = big_store.search(xq_big, K, mode="tensor")
    assert np.array_equal(i3[:64], i) and np.array_equal(s3[:64], s)


def test_config2_planted_neighbours(twin):
    """Rows overwritten with 4*q (exact in bf16) must come back first, with score 4*|q|^2."""
    import torch

    st = vod_b200.CorpusStore(N, D, dtype="bfloat16")
    st.fill_synthetic(SEED)
    xq = _queries(64, "bfloat16", seed=7)
    rng = np.random.default_rng(0)
    rows = np.sort(rng.choice(N, size=64, replace=False))
    rows[0], rows[-1] = 0, N - 1                       # first and last row of the store
    for q, r in enumerate(rows):
        st.add(4.0 * xq[q:q + 1], row0=int(r))
    st.add(torch.from_numpy(2.0 * xq[:1]), row0=int(rows[1]) + 1)   # a runner-up for query 0
    s, i = st.search(xq, K, mode="tensor")
    assert np.array_equal(i[:, 0], rows)
    expect = 4.0 * (xq.astype(np.float64) ** 2).sum(axis=1)
    assert np.abs(s[:, 0] - expect).max() <= 1e-5 * expect.max()
    assert i[0, 1] == rows[1] + 1
    st.close()


def test_shard_and_merge_equals_unsharded(big_store):
    xq = _queries(64, "bfloat16", seed=11)
    s, i = big_store.search(xq, K, mode="tensor")
    parts_s, parts_i = [], []
    for rank in range(2):
        lo, hi = vod_b200.shard_bounds(N, 2, rank)
        st = vod_b200.CorpusStore(hi - lo, D, dtype="bfloat16", row_offset=lo)
        st.fill_synthetic(SEED)
        ps, pi = st.search(xq, K, mode="tensor")
        parts_s.append(ps)
        parts_i.append(pi)
        st.close()
    ms, mi = vod_b200.merge_topk(np.stack(parts_s), np.stack(parts_i), K)
    assert np.array_equal(mi, i) and np.array_equal(ms, s)


def test_large_batch_8192_matches_small_batch(big_store):
    import torch

    xq = _queries(8192, "bfloat16", seed=21)
    ds, di = big_store.search_device(torch.from_numpy(xq).cuda(), K, mode="tensor")
    torch.cuda.synchronize()
    assert not big_store.check_async()
    ds, di = ds.cpu().numpy(), di.cpu().numpy()
    assert (np.diff(ds, axis=1) <= 0).all()
    for lo in (0, 4096, 8128):
        s, i = big_store.search(xq[lo:lo + 64], K, mode="tensor")
        assert np.array_equal(di[lo:lo + 64], i) and np.array_equal(ds[lo:lo + 64], s)


def test_config3_fp16_top1000(twin):
    """Per-GPU slice of config 3 (100M x 768 fp16 over 8 GPUs = 12.5M rows per shard), top-1000."""
    n = 12_500_000
    st = vod_b200.CorpusStore(n, D, dtype="float16", row_offset=3 * n)
    st.fill_synthetic(SEED)
    xq = _queries(64, "float16", seed=31)
    s, i = st.search(xq, 1000, mode="tensor")
    assert st.stats()["safe_fallback"] == 0
    assert (np.diff(s, axis=1) <= 0).all() and i.min() >= 3 * n and i.max() < 4 * n
    assert all(len(set(row)) == 1000 for row in i)
    true = np.zeros((2, 1000))
    for q in range(2):
        rows = np.stack([twin.synth_rows(SEED, int(r), 1, D, dtype=2)[0] for r in i[q]])
        true[q] = rows.astype(np.float64) @ xq[q].astype(np.float64)
    assert np.abs(s[:2] - true).max() <= 1e-5 * np.abs(true).max()
    s2, i2 = st.search(xq[:8], 1000, mode="exact")
    assert flat_ip.recall_at_k(i[:8], i2) >= 0.999
    st.close()


def test_config5_ingest_then_fp32_exact_search():
    """Index refresh path at reduced row count: stream fp32 host vectors (D=1024) into a bf16 store in chunks,
    then fp32-exact top-100; compared with the oracle over the same rounded values."""
    from tests.helpers import round_to

    n, d = 400_000, 1024
    rng = np.random.default_rng(5)
    st = vod_b200.CorpusStore(n, d, dtype="bfloat16")
    chunks = []
    for lo in range(0, n, 100_000):
        block = rng.standard_normal((100_000, d), dtype=np.float32)
        st.add(block)                                   # host fp32 -> device bf16 (RNE), like build.py:67-73
        chunks.append(round_to(block, "bfloat16"))
    xb = np.concatenate(chunks)
    assert st.ntotal == n
    xq = rng.standard_normal((64, d), dtype=np.float32)   # fp32 queries, NOT rounded: exact mode keeps them
    s, i = st.search(xq, 100, mode="exact")
    rs, ri = flat_ip.search(xb, xq, 100)
    rep = flat_ip.compare_topk(xb, xq, s, i, rs, ri, rtol=1e-5)
    assert rep["ok"], rep
    st.close()
