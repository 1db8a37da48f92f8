"""Parity of the CUDA hybrid-merge kernel (vodb_merge_results through vod_b200.hybrid) with the reference's numba
merge (golden vectors) and with the oracle; ports of the reference's merge / normalise unit tests; the retrieve ->
merge -> sample chain end to end."""
import pathlib

import numpy as np
import pytest

import vod_b200
from oracle import merge_ref
from vod_b200 import hybrid
from vod_b200.retrieval import RetrievalBatch

pytestmark = pytest.mark.gpu


def _same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


@pytest.fixture(scope="module")
def golden_merge():
    return np.load(pathlib.Path(__file__).parent / "golden" / "merge_ref.npz")


def test_bit_identical_to_reference_golden(golden_merge):
    n = 0
    for row in golden_merge["meta"]:
        cid, ne = int(row[0]), int(row[1])
        p = f"c{cid:03d}_"
        keys = [f"e{e}" for e in range(ne)]
        batches = {k: RetrievalBatch(scores=golden_merge[p + k + "_s"].copy(), indices=golden_merge[p + k + "_i"].copy(),
                                     labels=golden_merge[p + k + "_l"].copy() if k == "e0" else None) for k in keys}
        weights = {k: float(row[2 + e]) for e, k in enumerate(keys)}
        merged, raw = hybrid.merge_search_results(batches, weights)
        assert _same(merged.indices, golden_merge[p + "out_i"]), p
        assert _same(merged.scores, golden_merge[p + "out_s"]), p
        assert _same(merged.labels, golden_merge[p + "out_l"]), p
        for k in keys:
            assert _same(raw[k], golden_merge[p + "raw_" + k]), (p, k)
        n += 1
    assert n == 96


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_hybrid_merge_with_lookup_matches_oracle(dtype):
    """`_merge_search_results` (core/search.py:79-125): lookup scores zeroed, per-engine row-min subtraction, weighted
    union, raw scores / labels gathered — one kernel launch, compared with the oracle chain."""
    rng = np.random.default_rng(11)
    B = 4  # the oracle is a pure-Python O(B*K^2) loop
    res = {}
    for name, K in (("lookup", 8), ("dense", 1000), ("sparse", 1000)):
        idx = np.stack([rng.choice(5000, size=K, replace=False) for _ in range(B)]).astype(np.int64)
        sc = (rng.normal(size=(B, K)) * 5 + 80).astype(dtype)
        lab = (rng.uniform(size=(B, K)) < 0.5).astype(np.int64)
        if name != "dense":
            pad = rng.uniform(size=(B, K)) < 0.1
            idx[pad], sc[pad], lab[pad] = -1, -np.inf, -1
        res[name] = (sc, idx, lab)
    weights = {"dense": 1.0, "sparse": 0.35}
    exp_s, exp_i, exp_l, exp_raw = merge_ref.merge_hybrid(res, weights)
    batches = {k: RetrievalBatch(scores=v[0].copy(), indices=v[1].copy(), labels=v[2].copy(), meta={"time": 0.1})
               for k, v in res.items()}
    merged, raw = hybrid._merge_search_results(batches, weights)
    assert _same(merged.indices, exp_i) and _same(merged.scores, exp_s) and _same(merged.labels, exp_l)
    assert set(raw) == {"dense", "sparse"}
    for k in raw:
        assert _same(raw[k], exp_raw[k])
    assert merged.meta["dense_time"] == 0.1
    with pytest.raises(ValueError):
        hybrid._merge_search_results({"dense": batches["dense"]}, weights)


# ---- port of src/vod_dataloaders/tests/test_merge_search_results.py:52-81 ---------------------------------
@pytest.mark.parametrize("seed", list(range(10)))
@pytest.mark.parametrize("seq_length", [10, 30])
@pytest.mark.parametrize("n_values", [300, 1000])
def test_merge_search_results(seed, seq_length, n_values):
    rgn = np.random.default_rng(seed)
    alen = seq_length // 2
    blen = seq_length - alen
    labels = {i: rgn.choice([False, True], p=[0.5, 0.5]) for i in range(n_values)}
    a_i = rgn.choice(n_values, size=(alen,), replace=False)
    b_i = rgn.choice(n_values, size=(blen,), replace=False)
    search_results = {
        "a": RetrievalBatch.cast(indices=a_i[None, :], labels=[[labels[i] for i in a_i]], scores=rgn.uniform(0.0, 10.0, size=(1, alen))),
        "b": RetrievalBatch.cast(indices=b_i[None, :], labels=[[labels[i] for i in b_i]], scores=rgn.uniform(0.0, 10.0, size=(1, blen))),
    }
    weights = {"a": rgn.uniform(0.0, 1.0), "b": rgn.uniform(0.0, 1.0)}
    merged, raw_scores = hybrid.merge_search_results(search_results, weights)
    lookups = {key: dict(zip(v.indices[0], v.scores[0])) for key, v in search_results.items()}
    indices = merged.indices[0]
    for key, raw_key in raw_scores.items():
        for i, s_raw in zip(indices, raw_key[0]):
            s_input = lookups[key].get(i, np.nan)
            assert (np.isnan(s_input) and np.isnan(s_raw)) or (s_raw == s_input)
    for i, merged_s in zip(indices, merged.scores[0]):
        if i < 0:
            assert merged_s == -np.inf
            continue
        assert merged_s == sum(lookups[key].get(i, 0.0) * weight for key, weight in weights.items())


def test_single_engine_and_argument_checks():
    b = RetrievalBatch(scores=np.ones((2, 3), np.float32), indices=np.arange(6).reshape(2, 3))
    merged, raw = hybrid.merge_search_results({"dense": b}, {"dense": 0.5})
    assert np.array_equal(merged.scores, np.full((2, 3), 0.5, np.float32)) and raw["dense"] is b.scores
    with pytest.raises(ValueError):
        hybrid.merge_search_results({"a": b, "b": b}, {"a": 1.0})
    with pytest.raises(ValueError):
        hybrid.merge_search_results({"a": b, "b": RetrievalBatch(scores=np.ones((3, 3), np.float32), indices=np.zeros((3, 3), np.int64))})


class _StubSparse(vod_b200.SearchClient):
    """Stands for the Elasticsearch client: BM25-like scores, and gold-section labels when `ids` are given."""
    requires_vectors = False

    def __init__(self, n):
        self.n = n

    def ping(self):
        return True

    def search(self, *, text, vector=None, subset_ids=None, ids=None, shard=None, top_k=3):
        rng = np.random.default_rng(len(text) + top_k + (0 if ids is None else 1))
        B = len(text)
        if ids is not None:  # lookup query: return the gold sections first, labelled 1
            idx = np.full((B, top_k), -1, np.int64)
            sc = np.full((B, top_k), -np.inf, np.float32)
            lab = np.full((B, top_k), -1, np.int64)
            for b, gold in enumerate(ids):
                g = [int(x) for x in gold][:top_k]
                idx[b, :len(g)], sc[b, :len(g)], lab[b, :len(g)] = g, 1.0, 1
            return RetrievalBatch(scores=sc, indices=idx, labels=lab)
        idx = np.stack([rng.choice(self.n, size=top_k, replace=False) for _ in range(B)]).astype(np.int64)
        return RetrievalBatch(scores=np.sort(rng.uniform(1, 30, size=(B, top_k)).astype(np.float32))[:, ::-1].copy(),
                              indices=idx)


def test_retrieve_merge_sample_chain(twin):
    """RealmCollate-style flow (realm_collate.py:101-122): dense GPU search + sparse stub + gold lookup -> hybrid
    merge -> labeled priority sampling; every stage equals its oracle."""
    from oracle import flat_ip

    rng = np.random.default_rng(0)
    n, d, B, K = 20000, 64, 8, 200
    vectors = rng.integers(-3, 4, size=(n, d)).astype(np.float32)
    queries = rng.integers(-3, 4, size=(B, d)).astype(np.float32)
    gold = [[str(int(x)) for x in rng.choice(n, size=2, replace=False)] for _ in range(B)]
    with vod_b200.B200SearchMaster(vectors, dtype="bfloat16") as master:
        clients = {"dense": master.get_client(), "sparse": _StubSparse(n)}
        merged, raw = hybrid.async_hybrid_search(text=["q"] * B, shards=["s"] * B, vector=queries, section_ids=gold,
                                                 top_k=K, clients=clients, weights={"dense": 1.0, "sparse": 0.5})
    ds, di = flat_ip.search(vectors, queries, K)
    sp, lk = _StubSparse(n).search(text=["q"] * B, top_k=K), _StubSparse(n).search(text=[""] * B, ids=gold, top_k=K)
    exp_s, exp_i, exp_l, exp_raw = merge_ref.merge_hybrid(
        {"lookup": (lk.scores, lk.indices, lk.labels), "dense": (ds, di, None), "sparse": (sp.scores, sp.indices, None)},
        {"dense": 1.0, "sparse": 0.5})
    assert _same(merged.indices, exp_i) and _same(merged.scores, exp_s) and _same(merged.labels, exp_l)
    assert _same(raw["dense"], exp_raw["dense"]) and "search_time" in merged.meta
    out = vod_b200.sample_search_results(search_results=merged, raw_scores=raw, total=8, max_pos_sections=2, seed=7)
    t_ids, t_w, t_lab, _ = twin.sample(merged.scores, merged.labels > 0, k_positive=2, k_total=8, seed=7)
    assert np.array_equal(out.batch.indices, np.take_along_axis(merged.indices, t_ids, axis=-1))
    assert np.array_equal(out.log_weights.view(np.uint32), t_w.view(np.uint32))
    assert (out.batch.labels[:, :2]).all() and not out.batch.labels[:, 2:].any()  # the two gold sections come first
