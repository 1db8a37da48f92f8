"""Parity of the CUDA search path (through the C ABI) with the IndexFlatIP oracle. Needs a B200."""
import numpy as np
import pytest

import vod_b200
from oracle import flat_ip
from tests.helpers import int_valued, round_to

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north star: fp32-exact mode, scores within 1e-5 relative, index swaps only between <=1e-5 near-ties


def _store(xb, dtype="float32", **kw):
    st = vod_b200.CorpusStore(len(xb), xb.shape[1], dtype=dtype, **kw)
    st.add(xb)
    return st


def test_config1_exact_100k_x_768_q256_k100():
    """BASELINE config 1: faiss IndexFlatIP exact top-100, 100k x 768 fp32, 256 queries."""
    rng = np.random.default_rng(1234)
    xb = rng.standard_normal((100_000, 768), dtype=np.float32)
    xq = rng.standard_normal((256, 768), dtype=np.float32)
    st = _store(xb)
    s, i = st.search(xq, 100, mode="exact")
    rs, ri = flat_ip.search(xb, xq, 100)
    rep = flat_ip.compare_topk(xb, xq, s, i, rs, ri, rtol=RTOL)
    assert rep["ok"], rep
    assert (np.diff(s, axis=1) <= 0).all()
    assert st.stats()["safe_fallback"] == 0
    st.close()


@pytest.mark.parametrize("dtype", ["float32", "bfloat16", "float16"])
@pytest.mark.parametrize("n,d,nq,k", [(1, 8, 1, 1), (5, 100, 3, 8), (129, 64, 33, 100), (1000, 130, 65, 7),
                                      (5000, 768, 130, 2048), (40000, 96, 257, 100), (300_000, 64, 9, 10)])
def test_exact_mode_bit_exact_on_integer_data(dtype, n, d, nq, k):
    """Integer-valued vectors: every float32 dot product is exact whatever the summation order, so scores, ids,
    tie-breaking (score desc, id asc) and the -FLT_MAX / -1 padding must match the oracle bit for bit."""
    rng = np.random.default_rng(n + d)
    xb, xq = int_valued(rng, (n, d)), int_valued(rng, (nq, d))
    st = _store(xb, dtype)
    s, i = st.search(xq, k, mode="exact")
    rs, ri = flat_ip.search(xb, xq, k)
    assert np.array_equal(i, ri)
    assert np.array_equal(s, rs)
    st.close()


@pytest.mark.parametrize("nq", [7, 600])
@pytest.mark.parametrize("k", [100, 127, 128, 129, 192, 256, 257, 384, 385, 511, 512, 513, 1000])
def test_selection_paths_bit_exact(nq, k):
    """Every selection path of `block_select` against the oracle: thread-maximum bound + rank sort (2k <= CTA size:
    1024 threads for the few-query lists, 256 for many queries), radix select + rank sort (<= 384 selected), radix
    select + bitonic sort. Integers in [-100, 100] are exact in bf16 and their 64-term dot products (< 2^24) exact in
    fp32, and with scores spread over +-640000 ties are rare, so the bound really filters (the +-3 data of the tests
    above ties so often that most of its lists fall through to the radix path)."""
    rng = np.random.default_rng(1000 * nq + k)
    xb = rng.integers(-100, 101, size=(30_000, 64)).astype(np.float32)
    xq = rng.integers(-100, 101, size=(nq, 64)).astype(np.float32)
    st = _store(xb, "bfloat16")
    s, i = st.search(xq, k, mode="tensor")
    rs, ri = flat_ip.search(xb, xq, k)
    assert np.array_equal(i, ri)
    assert np.array_equal(s, rs)
    assert st.stats()["safe_fallback"] == 0
    st.close()


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
@pytest.mark.parametrize("n,d,nq,k", [(1, 8, 1, 1), (130, 64, 5, 100), (1000, 130, 64, 7), (5000, 768, 65, 100),
                                      (40000, 96, 129, 100), (70000, 768, 300, 1000), (300_000, 64, 9, 10)])
def test_tensor_mode_bit_exact_on_integer_data(dtype, n, d, nq, k):
    """Same exactness argument for the tcgen05 path: bf16/fp16 hold small integers exactly, products are exact,
    fp32 accumulation of integers < 2^24 is exact."""
    rng = np.random.default_rng(n + d + 1)
    xb, xq = int_valued(rng, (n, d)), int_valued(rng, (nq, d))
    st = _store(xb, dtype)
    s, i = st.search(xq, k, mode="tensor")
    rs, ri = flat_ip.search(xb, xq, k)
    assert np.array_equal(i, ri)
    assert np.array_equal(s, rs)
    st.close()


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
def test_tensor_mode_recall_and_score_error(dtype):
    """bf16/fp16 mode: recall@k vs the fp32 oracle on the same stored (rounded) values >= 0.999; score error
    <= 2e-3 relative (observed ~1e-6: products are exact in fp32, only the accumulation order differs)."""
    rng = np.random.default_rng(7)
    xb = round_to(rng.standard_normal((200_000, 768), dtype=np.float32), dtype)
    xq = round_to(rng.standard_normal((64, 768), dtype=np.float32), dtype)
    st = _store(xb, dtype)
    s, i = st.search(xq, 100, mode="tensor")
    rs, ri = flat_ip.search(xb, xq, 100)
    assert flat_ip.recall_at_k(i, ri) >= 0.999
    rep = flat_ip.compare_topk(xb, xq, s, i, rs, ri, rtol=2e-3)
    assert rep["ok"], rep
    assert rep["max_score_rel_err"] < 1e-4, rep
    # exact mode over the same 16-bit store agrees with the oracle at the fp32 tolerance
    s2, i2 = st.search(xq, 100, mode="exact")
    rep2 = flat_ip.compare_topk(xb, xq, s2, i2, rs, ri, rtol=RTOL)
    assert rep2["ok"], rep2
    st.close()


def test_unit_norm_embeddings_stress_near_ties():
    rng = np.random.default_rng(3)
    xb = rng.standard_normal((50_000, 256), dtype=np.float32)
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    xq = rng.standard_normal((40, 256), dtype=np.float32)
    xq /= np.linalg.norm(xq, axis=1, keepdims=True)
    st = _store(xb)
    s, i = st.search(xq, 100, mode="exact")
    rs, ri = flat_ip.search(xb, xq, 100)
    rep = flat_ip.compare_topk(xb, xq, s, i, rs, ri, rtol=RTOL)
    assert rep["ok"], rep
    st.close()


def test_duplicate_rows_and_zero_rows():
    rng = np.random.default_rng(4)
    base = int_valued(rng, (50, 32))
    xb = np.concatenate([base[rng.integers(0, 50, 20000)], np.zeros((3000, 32), np.float32)])
    xq = int_valued(rng, (17, 32))
    for dtype, mode in [("float32", "exact"), ("bfloat16", "tensor")]:
        st = _store(xb, dtype)
        s, i = st.search(xq, 500, mode=mode)
        rs, ri = flat_ip.search(xb, xq, 500)
        assert np.array_equal(i, ri) and np.array_equal(s, rs)
        st.close()


@pytest.mark.parametrize("dtype,mode", [("float32", "exact"), ("bfloat16", "tensor")])
def test_adversarial_order_triggers_overflow_proof_fallback(dtype, mode):
    """Rows sorted by increasing score: every row beats the running threshold, lists overflow, the call must fall
    back to the overflow-proof schedule and still be exact."""
    n, d = 150_000, 64
    xb = np.zeros((n, d), np.float32)
    xb[:, 0] = (np.arange(n) // 64) % 256   # blocks of 64 tied rows, values 0..255 (exact in bf16)
    xb[:, 1] = np.arange(n) // (64 * 256)   # slow counter 0..9
    xq = np.zeros((3, d), np.float32)
    xq[:, 0] = 1.0
    xq[:, 1] = 256.0                        # score = slow*256 + fast: non-decreasing in the row id
    st = _store(xb, dtype)
    s, i = st.search(xq, 100, mode=mode)
    rs, ri = flat_ip.search(xb, xq, 100)
    assert np.array_equal(i, ri) and np.array_equal(s, rs)
    assert st.stats()["safe_fallback"] == 1
    st.close()


def test_host_and_device_entry_points_agree():
    import torch

    rng = np.random.default_rng(5)
    xb, xq = int_valued(rng, (30000, 128)), int_valued(rng, (64, 128))
    st = vod_b200.CorpusStore(len(xb), 128, dtype="bfloat16")
    st.add(torch.from_numpy(xb).cuda())                       # device-side ingest (build_gpu.py:294-380 analogue)
    s, i = st.search(xq, 50)
    ds, di = st.search_device(torch.from_numpy(xq).cuda(), 50)
    torch.cuda.synchronize()
    assert not st.check_async()
    assert np.array_equal(ds.cpu().numpy(), s) and np.array_equal(di.cpu().numpy(), i)
    dsb, dib = st.search_device(torch.from_numpy(xq).cuda().to(torch.bfloat16), 50, mode="exact")
    assert np.array_equal(dib.cpu().numpy(), i)
    st.close()


def test_ingest_variants_and_readback():
    import torch

    rng = np.random.default_rng(6)
    x = rng.standard_normal((1000, 100), dtype=np.float32)
    for dtype in ("float32", "bfloat16", "float16"):
        st = vod_b200.CorpusStore(1000, 100, dtype=dtype)
        st.add(x[:300])
        st.add(x[300:600].astype(np.float16))                  # fp16 source
        st.add(torch.from_numpy(x[600:800]).to(torch.bfloat16))  # CPU bf16 tensor
        st.add(torch.from_numpy(x[800:]).cuda())              # CUDA fp32 tensor
        assert st.ntotal == 1000
        got = st.read(0, 1000)
        exp = np.concatenate([
            round_to(x[:300], dtype),
            round_to(x[300:600].astype(np.float16).astype(np.float32), dtype),
            round_to(round_to(x[600:800], "bfloat16"), dtype),
            round_to(x[800:], dtype)])
        assert np.array_equal(got, exp)
        st.close()


def test_synthetic_fill_matches_cpu_twin(twin):
    for dtype, code in (("float32", 0), ("bfloat16", 1), ("float16", 2)):
        for unit in (False, True):
            st = vod_b200.CorpusStore(700, 200, dtype=dtype, row_offset=12345)
            st.fill_synthetic(99, unit_norm=unit)
            got = st.read(0, 700)
            exp = twin.synth_rows(99, 12345, 700, 200, dtype=code, unit_norm=unit)
            assert np.array_equal(got, exp), (dtype, unit)
            st.close()


def test_row_offset_gives_global_ids_and_merge_matches_unsharded():
    rng = np.random.default_rng(8)
    xb, xq = int_valued(rng, (5000, 64)), int_valued(rng, (20, 64))
    parts = [(0, 2048), (2048, 4096), (4096, 5000)]
    all_s, all_i = [], []
    for lo, hi in parts:
        st = vod_b200.CorpusStore(hi - lo, 64, dtype="float32", row_offset=lo)
        st.add(xb[lo:hi])
        s, i = st.search(xq, 300)
        all_s.append(s)
        all_i.append(i)
        st.close()
    ms, mi = vod_b200.merge_topk(np.stack(all_s), np.stack(all_i), 300)
    rs, ri = flat_ip.search(xb, xq, 300)
    assert np.array_equal(mi, ri) and np.array_equal(ms, rs)
    # a shard smaller than k contributes -1 padded slots that must sort last
    st = vod_b200.CorpusStore(10, 64, row_offset=7)
    st.add(xb[:10])
    s, i = st.search(xq, 300)
    assert (i[:, 10:] == -1).all() and (i[:, :10] >= 7).all()
    ms2, mi2 = vod_b200.merge_topk(np.stack([s, s]), np.stack([i, i + np.where(i >= 0, 100, 0)]), 25)
    assert (mi2[:, 20:] == -1).all() and (ms2[:, 20:] == -np.finfo(np.float32).max).all()
    st.close()


def test_client_master_drop_in():
    rng = np.random.default_rng(9)
    vectors = int_valued(rng, (3000, 48))
    with vod_b200.B200SearchMaster(vectors, dtype="float32") as master:
        client = master.get_client()
        assert client.ping()
        q = int_valued(rng, (10, 48))
        out = client.search(vector=q, text=["ignored"] * 10, subset_ids=None, ids=None, shard=None, top_k=5)
        assert type(out).__name__ == "RetrievalBatch"
        assert out.scores.dtype == np.float32 and out.indices.dtype == np.int64 and out.labels is None
        assert out.scores.shape == (10, 5) and "time" in out.meta
        rs, ri = flat_ip.search(vectors, q, 5)
        assert np.array_equal(out.indices, ri) and np.array_equal(out.scores, rs)
        out.indices += 100                      # callers mutate results in place (sharded_search.py:103)
        assert out.scores.flags.writeable
        with pytest.raises(ValueError):
            client.search(vector=q[0], top_k=5)  # server.py:82-83
        with pytest.raises(ValueError):
            client.search(vector=q[:, :10], top_k=5)
    assert not client.ping()


def test_error_paths():
    st = vod_b200.CorpusStore(10, 8)
    with pytest.raises(vod_b200.VodbError):
        st.search(np.zeros((1, 8), np.float32), 3)        # empty store (server.py:59-62 health check)
    st.add(np.ones((10, 8), np.float32))
    with pytest.raises(vod_b200.VodbError):
        st.search(np.zeros((1, 8), np.float32), 0)
    with pytest.raises(vod_b200.VodbError):
        st.search(np.zeros((1, 8), np.float32), 4096)
    with pytest.raises(vod_b200.VodbError):
        st.add(np.ones((5, 8), np.float32), row0=8)
    s, i = st.search(np.zeros((0, 8), np.float32), 3)
    assert s.shape == (0, 3)
    st.close()
    with pytest.raises(vod_b200.VodbError):
        st.search(np.zeros((1, 8), np.float32), 3)


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
def test_tensor_multi_term_queries_keep_fp32_query_precision(dtype):
    """float32 (unrounded) queries against a 16-bit store. One term rounds the query to the store dtype; two / three
    terms split it into hi + lo (+ lo2) parts accumulated in the same TMEM accumulator. Three bf16 terms carry the
    full 24-bit mantissa: every product is exact in fp32, only the accumulation order differs from the oracle, so
    the fp32-exact tolerance (1e-5 relative, near-tie swaps only) applies — IndexFlatIP parity on tensor cores."""
    rng = np.random.default_rng(17)
    xb = round_to(rng.standard_normal((300_000, 768), dtype=np.float32), dtype)
    xq = rng.standard_normal((64, 768), dtype=np.float32)          # NOT representable in 16 bits
    st = _store(xb, dtype)
    rs, ri = flat_ip.search(xb, xq, 100)
    recalls = {}
    for mode in ("tensor", "tensor2", "tensor3"):
        s, i = st.search(xq, 100, mode=mode)
        recalls[mode] = flat_ip.recall_at_k(i, ri)
        if mode == "tensor3":
            rep = flat_ip.compare_topk(xb, xq, s, i, rs, ri, rtol=RTOL)
            assert rep["ok"], rep
    assert recalls["tensor2"] >= 0.999, recalls
    assert recalls["tensor3"] >= 0.9995, recalls
    assert recalls["tensor"] >= 0.95, recalls      # documented: rounding the query costs recall on near-ties
    # larger batches use the 128-query tile variants
    xq2 = rng.standard_normal((200, 768), dtype=np.float32)
    rs2, ri2 = flat_ip.search(xb, xq2, 100)
    for mode in ("tensor2", "tensor3"):
        s, i = st.search(xq2, 100, mode=mode)
        assert flat_ip.recall_at_k(i, ri2) >= 0.999
    s, i = st.search(xq2, 100, mode="tensor3")
    rep = flat_ip.compare_topk(xb, xq2, s, i, rs2, ri2, rtol=RTOL)
    assert rep["ok"], rep
    st.close()


@pytest.mark.parametrize("mode", ["tensor2", "tensor3"])
def test_multi_term_modes_bit_exact_on_integer_data(mode):
    rng = np.random.default_rng(23)
    xb, xq = int_valued(rng, (50_000, 200)), int_valued(rng, (130, 200))
    st = _store(xb, "bfloat16")
    s, i = st.search(xq, 64, mode=mode)
    rs, ri = flat_ip.search(xb, xq, 64)
    assert np.array_equal(i, ri) and np.array_equal(s, rs)
    st.close()


def test_fp32_store_on_tensor_cores_meets_the_fp32_exact_tolerance():
    """A float32 store searched on the tensor cores: the rows are kept as three bf16 planes (row = p0 + p1 + p2
    exactly), the float32 queries as three bf16 terms, and the six products c_p * q_t with p + t <= 2 accumulate in
    fp32 — IndexFlatIP parity at the fp32-exact tolerance on arbitrary float32 data (BASELINE config 1 shape), also
    after rows are appended or overwritten (the planes follow the store)."""
    rng = np.random.default_rng(1234)
    xb = rng.standard_normal((100_000, 768), dtype=np.float32)
    xq = rng.standard_normal((256, 768), dtype=np.float32)
    st = vod_b200.CorpusStore(len(xb), 768, dtype="float32")
    st.add(xb[:60_000])
    s, i = st.search(xq[:64], 100, mode="tensor3")          # planes built for the first 60k rows
    rs, ri = flat_ip.search(xb[:60_000], xq[:64], 100)
    rep = flat_ip.compare_topk(xb[:60_000], xq[:64], s, i, rs, ri, rtol=RTOL)
    assert rep["ok"], rep
    st.add(xb[60_000:], row0=60_000)                        # planes extended
    rs, ri = flat_ip.search(xb, xq, 100)
    recalls = {}
    for mode in ("tensor", "tensor2", "tensor3"):
        s, i = st.search(xq, 100, mode=mode)
        recalls[mode] = flat_ip.recall_at_k(i, ri)
        assert (np.diff(s, axis=1) <= 0).all()
        if mode == "tensor3":
            rep = flat_ip.compare_topk(xb, xq, s, i, rs, ri, rtol=RTOL)
            assert rep["ok"], rep
            # leading product and corrections are accumulated apart (the tensor core truncates every accumulation):
            # the scores sit well inside the tolerance, not at its edge
            assert rep["max_score_rel_err"] <= 4e-6, rep
    assert recalls["tensor3"] >= 0.9995 and recalls["tensor2"] >= 0.999 and recalls["tensor"] >= 0.9, recalls
    se, ie = st.search(xq, 100, mode="exact")               # the CUDA-core kernel agrees on the same store
    assert flat_ip.recall_at_k(ie, i) >= 0.9995
    xb[:1000] = rng.standard_normal((1000, 768), dtype=np.float32) * 3   # overwrite rows: planes redone from row 0
    st.add(xb[:1000], row0=0)
    s, i = st.search(xq[:9], 100, mode="tensor3")
    rs, ri = flat_ip.search(xb, xq[:9], 100)
    rep = flat_ip.compare_topk(xb, xq[:9], s, i, rs, ri, rtol=RTOL)
    assert rep["ok"], rep
    assert np.isin(i, np.arange(1000)).mean() > 0.5         # the rescaled rows now dominate the top-100
    st.close()


@pytest.mark.parametrize("mode", ["tensor", "tensor2", "tensor3"])
@pytest.mark.parametrize("n,d,nq,k", [(5, 100, 3, 8), (1000, 130, 65, 7), (40_000, 96, 257, 100)])
def test_fp32_store_tensor_modes_bit_exact_on_integer_data(mode, n, d, nq, k):
    rng = np.random.default_rng(n + nq)
    xb, xq = int_valued(rng, (n, d)), int_valued(rng, (nq, d))
    st = _store(xb, "float32")
    s, i = st.search(xq, k, mode=mode)
    rs, ri = flat_ip.search(xb, xq, k)
    assert np.array_equal(i, ri) and np.array_equal(s, rs)
    st.close()


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
def test_empty_correction_terms_are_skipped_without_changing_results(dtype):
    """float32 queries that are exact in the store dtype have empty correction terms: the multi-term modes then run
    the one-term computation (decided on the device, per batch) and must return exactly what `tensor` returns; one
    inexact query in the batch switches the full computation back on."""
    rng = np.random.default_rng(31)
    xb = round_to(rng.standard_normal((60_000, 320), dtype=np.float32), dtype)
    st = _store(xb, dtype)
    for nq in (40, 100, 200):
        xq = round_to(rng.standard_normal((nq, 320), dtype=np.float32), dtype)
        s1, i1 = st.search(xq, 50, mode="tensor")
        import torch

        xq_dev = torch.from_numpy(xq).cuda()
        for mode in ("tensor2", "tensor3"):
            s, i = st.search(xq, 50, mode=mode)              # host queries: classified on the host (one-term kernels)
            assert np.array_equal(i, i1) and np.array_equal(s, s1), (nq, mode)
            ds, di = st.search_device(xq_dev, 50, mode=mode)  # device queries: prepare_kernel's term masks decide
            torch.cuda.synchronize()
            assert not st.check_async()
            assert np.array_equal(di.cpu().numpy(), i1) and np.array_equal(ds.cpu().numpy(), s1), (nq, mode, "device")
        xq[nq // 2] = rng.standard_normal(320, dtype=np.float32)       # not representable: corrections needed again
        s3, i3 = st.search(xq, 50, mode="tensor3")
        rs, ri = flat_ip.search(xb, xq, 50)
        rep = flat_ip.compare_topk(xb, xq, s3, i3, rs, ri, rtol=RTOL)
        assert rep["ok"], rep
    st.close()


def test_stale_query_staging_rows_never_reach_the_filter():
    """Regression: a 9-query tensor-mode search right after a 64-query exact-mode search. The staging rows 9..63
    still hold the float32 bit patterns of the earlier batch, which read as bf16 include inf / NaN; they must be
    cleared, or unused accumulator columns produce +inf scores that pass the threshold filter."""
    rng = np.random.default_rng(5)
    xb = rng.standard_normal((50_000, 256), dtype=np.float32)
    xb[:500] *= 3                                            # strong rows first: tight thresholds early
    xq = rng.standard_normal((64, 256), dtype=np.float32)
    for dtype in ("bfloat16", "float32"):
        st = _store(round_to(xb, dtype) if dtype != "float32" else xb, dtype)
        ref = round_to(xb, dtype) if dtype != "float32" else xb
        st.search(xq * 1e30, 10, mode="exact")               # leaves huge float32 values in the staging buffer
        s, i = st.search(xq[:9], 100, mode="tensor3")
        rs, ri = flat_ip.search(ref, xq[:9], 100)
        rep = flat_ip.compare_topk(ref, xq[:9], s, i, rs, ri, rtol=RTOL)
        assert rep["ok"], rep
        assert st.stats()["safe_fallback"] == 0
        st.close()


def _remote_worker(blob, q):
    import pickle

    client = pickle.loads(blob)
    rng = np.random.default_rng(9)
    xq = int_valued(rng, (10, 48))
    out = client.search(vector=xq, top_k=5)
    q.put((client.ping(), out.scores, out.indices))


@pytest.mark.timeout(180)
def test_pickled_client_reaches_the_gpu_master_from_a_worker_process():
    """DataLoader workers get pickled clients (reference: FaissClient(host, port) over HTTP); here they reach the
    GPU-owning master over the Unix-socket transport and get the same bits as the in-process call."""
    import multiprocessing as mp
    import pickle

    rng = np.random.default_rng(9)
    vectors = int_valued(np.random.default_rng(1), (3000, 48))
    xq = int_valued(rng, (10, 48))
    with vod_b200.B200SearchMaster(vectors, dtype="bfloat16") as master:
        client = master.get_client()
        local = client.search(vector=xq, top_k=5)
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        p = ctx.Process(target=_remote_worker, args=(pickle.dumps(client), q))
        p.start()
        ok, s, i = q.get(timeout=150)
        p.join(timeout=30)
    assert ok and np.array_equal(s, local.scores) and np.array_equal(i, local.indices)


def test_pair_kernel_overflow_fallback_and_odd_tile_counts():
    """The 2-CTA kernel (batches > 128 queries): adversarial row order -> list overflow -> overflow-proof schedule;
    segment / shard sizes that leave the follower CTA of the last pair without rows."""
    n, d = 150_000 + 128, 64   # odd number of 128-row tiles
    xb = np.zeros((n, d), np.float32)
    xb[:, 0] = (np.arange(n) // 64) % 256
    xb[:, 1] = np.arange(n) // (64 * 256)
    xq = np.zeros((130, d), np.float32)
    xq[:, 0] = 1.0
    xq[:, 1] = 256.0
    xq[:, 2] = np.arange(130)  # distinct queries, same ranking
    st = _store(xb, "bfloat16")
    s, i = st.search(xq, 100, mode="tensor")
    rs, ri = flat_ip.search(xb, xq, 100)
    assert np.array_equal(i, ri) and np.array_equal(s, rs)
    assert st.stats()["safe_fallback"] == 1
    st.close()
    rng = np.random.default_rng(31)
    for rows in (129, 257, 128 * 7 + 5):
        xb2, xq2 = int_valued(rng, (rows, 128)), int_valued(rng, (200, 128))
        st = _store(xb2, "float16")
        s, i = st.search(xq2, 100, mode="tensor")
        rs, ri = flat_ip.search(xb2, xq2, 100)
        assert np.array_equal(i, ri) and np.array_equal(s, rs)
        st.close()


@pytest.mark.timeout(300)
def test_host_threads_share_a_store():
    """Several host threads search one store concurrently (the store serialises them); every result is exact."""
    import threading

    rng = np.random.default_rng(77)
    xb = int_valued(rng, (30_000, 96))
    st = _store(xb, "bfloat16")
    errors = []

    def work(t):
        try:
            r = np.random.default_rng(t)
            for _ in range(6):
                nq, k = int(r.choice([3, 40, 130])), int(r.choice([5, 100, 700]))
                xq = int_valued(r, (nq, 96))
                s, i = st.search(xq, k, mode=str(r.choice(["tensor", "tensor3", "exact"])))
                rs, ri = flat_ip.search(xb, xq, k)
                assert np.array_equal(i, ri) and np.array_equal(s, rs)
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[0]
    st.close()


@pytest.mark.parametrize("dtype,d", [("bfloat16", 1024), ("float16", 2048), ("float32", 512), ("float32", 1024)])
def test_power_of_two_row_strides(dtype, d):
    """Rows of 2 KB / 4 KB (dim 1024 bf16, 2048 fp16, 512 / 1024 fp32): every mode, both ingest paths and the
    read-back agree with the oracle (the K loop covers ceil(dim / 64) chunks whatever the pitch is)."""
    rng = np.random.default_rng(d)
    n = 3000
    xb, xq = int_valued(rng, (n, d)), int_valued(rng, (70, d))
    st = vod_b200.CorpusStore(n, d, dtype=dtype)
    st.add(xb[:1000])
    import torch

    st.add(torch.from_numpy(xb[1000:]).cuda(), row0=1000)              # device ingest
    assert np.array_equal(st.read(0, n), xb)
    rs, ri = flat_ip.search(xb, xq, 50)
    for mode in ("exact", "tensor", "tensor2", "tensor3"):
        s, i = st.search(xq, 50, mode=mode)
        assert np.array_equal(i, ri) and np.array_equal(s, rs), mode
    st.close()


def test_client_accepts_cuda_tensors():
    """`search(vector=<CUDA tensor>)`: encoder outputs go to the search without a host round trip of the queries;
    results equal the numpy path (bf16 tensor, float32 tensor, and a CPU tensor)."""
    import torch

    rng = np.random.default_rng(41)
    xb, xq = int_valued(rng, (20_000, 128)), int_valued(rng, (33, 128))
    with vod_b200.B200SearchMaster(xb, dtype="bfloat16") as master:
        client = master.get_client()
        ref = client.search(vector=xq, top_k=40)
        for t in (torch.from_numpy(xq).cuda(), torch.from_numpy(xq).cuda().to(torch.bfloat16), torch.from_numpy(xq)):
            out = client.search(vector=t, top_k=40)
            assert np.array_equal(out.indices, ref.indices) and np.array_equal(out.scores, ref.scores)
            assert out.scores.flags.writeable and isinstance(out.scores, np.ndarray)
        with pytest.raises(ValueError):
            client.search(vector=torch.zeros(128).cuda(), top_k=3)


@pytest.mark.parametrize("mode", ["tensor", "tensor3"])
def test_every_query_tile_width_is_bit_exact(mode):
    """The MMA width follows the item's query count (multiples of 32) and multi-term batches run on CTA pairs
    (<64,T> up to 64 queries, <128,T> above, the wide one-term kernel when the correction terms are empty): every
    tile width, the last tile of multi-tile batches, and both sides of each kernel switch, against the oracle."""
    rng = np.random.default_rng(17)
    n, d, k = 9_000, 192, 50     # > 2 x 4096 rows: dump segment + filtered segments; odd number of 256-row pair tiles
    xb = int_valued(rng, (n, d))
    st = _store(xb, "bfloat16")
    for nq in (1, 31, 32, 33, 63, 64, 65, 96, 97, 127, 128, 129, 160, 191, 193, 224, 255, 256, 257, 288, 300, 383, 385,
               511, 513, 700):
        xq = int_valued(rng, (nq, d))
        if mode == "tensor3":
            # 18 significant bits per element: needs all three bf16 terms. Only 16 nonzero dimensions per query, so every
            # partial sum is a multiple of 2^-16 below 2^8 and float32 accumulation stays exact in any order.
            xq = xq + int_valued(rng, (nq, d)) / 65536.0
            keep = np.zeros((nq, d), bool)
            for r in range(nq):
                keep[r, rng.choice(d, size=16, replace=False)] = True
            xq = np.where(keep, xq, 0.0).astype(np.float32)
        s, i = st.search(xq, k, mode=mode)
        rs, ri = flat_ip.search(xb, xq, k)
        assert np.array_equal(i, ri), (mode, nq)
        assert np.array_equal(s, rs), (mode, nq)
        if mode == "tensor3" and nq in (64, 128, 300):   # correction terms empty: the one-term launch takes the batch
            xq0 = int_valued(rng, (nq, d))
            s, i = st.search(xq0, k, mode=mode)
            rs, ri = flat_ip.search(xb, xq0, k)
            assert np.array_equal(i, ri) and np.array_equal(s, rs), (mode, nq, "empty terms")
    st.close()


def test_auto_mode_serves_a_float32_store_from_the_tensor_cores():
    """`auto` on a float32 store = three bf16 planes x three query terms (fp32-exact tolerance), planes built by
    `prepare_tensor` on first use; `exact` stays available as the CUDA-core cross-check."""
    rng = np.random.default_rng(23)
    xb = rng.standard_normal((50_000, 256), dtype=np.float32)
    xq = rng.standard_normal((40, 256), dtype=np.float32)
    st = _store(xb, "float32")
    assert st._planes_fit is None
    s, i = st.search(xq, 100)                       # auto
    assert st._planes_fit is True and st._mode(None, 0) == vod_b200._lib.MODE_TENSOR_X3
    rs, ri = flat_ip.search(xb, xq, 100)
    rep = flat_ip.compare_topk(xb, xq, s, i, rs, ri, rtol=RTOL)
    assert rep["ok"], rep
    s2, i2 = st.search(xq, 100, mode="exact")
    rep2 = flat_ip.compare_topk(xb, xq, s2, i2, rs, ri, rtol=RTOL)
    assert rep2["ok"], rep2
    st.close()


def test_index_built_from_a_zarr_store_on_disk(tmp_path):
    """The reference's predict step leaves embeddings in a zarr-v2 store (ts_factory.py:54-90); `open_vectors` reads
    it lazily and `build_b200_index` / `ingest` stream it into HBM in blocks."""
    rng = np.random.default_rng(29)
    xb = int_valued(rng, (2_345, 96))
    path = vod_b200.write_zarr_v2(tmp_path / "vectors", xb, chunk_size=100)
    vectors = vod_b200.open_vectors(path)
    xq = int_valued(rng, (9, 96))
    rs, ri = flat_ip.search(xb, xq, 20)
    with vod_b200.B200SearchMaster(vectors, dtype="bfloat16", add_batch_size=700, serve=False) as master:
        out = master.get_client().search(vector=xq, top_k=20)
        assert np.array_equal(out.indices, ri) and np.array_equal(out.scores, rs)
    from vod_b200 import zarr_io

    st = vod_b200.CorpusStore(len(xb), 96, dtype="float16")
    assert zarr_io.ingest(st, vectors, batch_rows=512) == len(xb)
    s, i = st.search(xq, 20, mode="tensor")
    assert np.array_equal(i, ri) and np.array_equal(s, rs)
    with pytest.raises(vod_b200.VodbError):          # append-only: a block that would leave a gap is refused
        st2 = vod_b200.CorpusStore(100, 96, dtype="float16")
        st2.add(xb[:10], row0=50)
    st.close()
