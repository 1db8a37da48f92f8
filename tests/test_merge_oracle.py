"""The merge / normalise oracle (oracle/merge_ref.py) pinned against the reference's numba code (golden file +
live), the reference's own unit tests ported to it, and the sharded-client router against the reference class."""
import numpy as np
import pytest

from oracle import merge_ref, ref_shim


def _cases(golden_merge, max_entries=None):
    for row in golden_merge["meta"]:
        cid, n = int(row[0]), int(row[1])
        p = f"c{cid:03d}_"
        keys = [f"e{e}" for e in range(n)]
        inp = {k: (golden_merge[p + k + "_s"], golden_merge[p + k + "_i"], golden_merge[p + k + "_l"] if k == "e0" else None)
               for k in keys}
        if max_entries and sum(v[0].shape[1] for v in inp.values()) > max_entries:
            continue
        weights = {k: float(row[2 + e]) for e, k in enumerate(keys)}
        yield p, inp, weights


def _same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


@pytest.fixture(scope="module")
def golden_merge():
    import pathlib

    return np.load(pathlib.Path(__file__).parent / "golden" / "merge_ref.npz")


def test_oracle_matches_reference_golden(golden_merge):
    n = 0
    for p, inp, weights in _cases(golden_merge, max_entries=200):
        s, i, lab, raw = merge_ref.merge(inp, weights)
        assert _same(i, golden_merge[p + "out_i"]), p
        assert _same(s, golden_merge[p + "out_s"]), p
        assert _same(lab, golden_merge[p + "out_l"]), p
        for k in inp:
            assert _same(raw[k], golden_merge[p + "raw_" + k]), (p, k)
        assert _same(merge_ref.subtract_min_score(inp["e1"][0], 0.5), golden_merge[p + "norm_e1"]), p
        n += 1
    assert n >= 48


def test_oracle_matches_live_reference():
    if not ref_shim.available():
        pytest.skip("reference tree not present (GPU box)")
    mods = ref_shim.load()
    RB, merge = mods["retrieval"].RetrievalBatch, mods["merge"]
    rng = np.random.default_rng(3)
    inp = {}
    for k, K in (("lookup", 5), ("dense", 20), ("sparse", 15)):
        idx = np.stack([rng.choice(40, size=K, replace=False) for _ in range(4)]).astype(np.int64)
        inp[k] = (rng.normal(size=(4, K)).astype(np.float32), idx, (rng.uniform(size=(4, K)) < 0.5).astype(np.int64))
    w = {"lookup": 0.0, "dense": 1.0, "sparse": 0.3}
    merged, raw = merge.merge_search_results({k: RB(scores=v[0].copy(), indices=v[1].copy(), labels=v[2] if k == "lookup" else None)
                                              for k, v in inp.items()}, w)
    s, i, lab, r = merge_ref.merge({k: (v[0], v[1], v[2] if k == "lookup" else None) for k, v in inp.items()}, w)
    assert _same(s, merged.scores) and _same(i, merged.indices) and _same(lab, merged.labels)
    for k in inp:
        assert _same(r[k], raw[k])


# ---- port of src/vod_dataloaders/tests/test_merge_search_results.py:52-81 ---------------------------------
@pytest.mark.parametrize("seed", list(range(5)))
@pytest.mark.parametrize("seq_length", [10, 30])
@pytest.mark.parametrize("n_values", [300, 1000])
def test_merge_is_weighted_sum(seed, seq_length, n_values):
    rgn = np.random.default_rng(seed)
    alen = seq_length // 2
    blen = seq_length - alen
    a_i = rgn.choice(n_values, size=(alen,), replace=False)
    b_i = rgn.choice(n_values, size=(blen,), replace=False)
    res = {"a": (rgn.uniform(0, 10, size=(1, alen)), a_i[None].astype(np.int64), None),
           "b": (rgn.uniform(0, 10, size=(1, blen)), b_i[None].astype(np.int64), None)}
    weights = {"a": rgn.uniform(0, 1), "b": rgn.uniform(0, 1)}
    s, i, _, raw = merge_ref.merge(res, weights)
    lookups = {k: dict(zip(v[1][0], v[0][0])) for k, v in res.items()}
    for key, rk in raw.items():
        for idx, s_raw in zip(i[0], rk[0]):
            s_in = lookups[key].get(idx, np.nan)
            assert (np.isnan(s_in) and np.isnan(s_raw)) or s_raw == s_in
    for idx, ms in zip(i[0], s[0]):
        if idx < 0:
            assert ms == -np.inf
            continue
        assert ms == sum(lookups[k].get(idx, 0.0) * w for k, w in weights.items())


# ---- port of src/vod_dataloaders/tests/test_normalize.py:15-32 -------------------------------------------
@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("nan_prob", [0.0, 0.3])
@pytest.mark.parametrize("inf_prob", [0.0, 0.3])
@pytest.mark.parametrize("offset", [0, 1.0, -10.0])
def test_subtract_min_score(seed, nan_prob, inf_prob, offset, n=100):
    rgn = np.random.default_rng(seed)
    scores = rgn.uniform(0.0, 10.0, size=(n,))
    scores = np.where(rgn.uniform(size=n) < nan_prob, np.nan, scores)
    scores = np.where(rgn.uniform(size=n) < inf_prob, -np.inf, scores)
    out = merge_ref.subtract_min_score(scores, offset=offset)
    finite = [s for s in scores if np.isfinite(s)]
    mn = min(finite) if finite else np.nan
    for o, v in zip(scores, out):
        if np.isnan(o):
            assert np.isnan(v)
        elif np.isinf(o):
            assert np.isinf(v)
        else:
            assert v == o - mn + offset


# ---- the corpus router -----------------------------------------------------------------------------------
class _FakeClient:
    requires_vectors = True

    def __init__(self, cls, n, seed):
        self.cls, self.n, self.seed, self.calls = cls, n, seed, []

    def ping(self):
        return True

    def search(self, *, text, vector=None, subset_ids=None, ids=None, shard=None, top_k=3):
        self.calls.append((list(text), None if vector is None else np.array(vector), subset_ids, ids, top_k))
        rng = np.random.default_rng(self.seed + len(text))
        k = min(top_k, self.n)
        return self.cls(scores=rng.normal(size=(len(text), k)).astype(np.float32),
                        indices=rng.integers(0, self.n, size=(len(text), k)).astype(np.int64))


def test_sharded_client_matches_reference_router():
    from vod_b200.retrieval import RetrievalBatch
    from vod_b200.routing import ShardedSearchClient

    shard = ["wiki", "pubmed", "wiki", "wiki", "pubmed"]
    text = [f"q{i}" for i in range(5)]
    vec = np.arange(5 * 4, dtype=np.float32).reshape(5, 4)
    mine = ShardedSearchClient({"wiki": _FakeClient(RetrievalBatch, 100, 1), "pubmed": _FakeClient(RetrievalBatch, 2, 2)},
                               {"wiki": 0, "pubmed": 1000})
    out = mine.search(text=text, vector=vec, shard=shard, top_k=4)
    assert out.scores.shape == (5, 4) and out.indices.dtype == np.int64
    assert (out.indices[[1, 4], :2] >= 1000).all() and (out.indices[[1, 4], 2:] == -1).all()  # ragged rows padded
    assert np.isneginf(out.scores[[1, 4], 2:]).all()
    assert mine.requires_vectors and mine.ping()
    with pytest.raises(ValueError):
        mine.search(text=text, vector=vec, top_k=4)
    with pytest.raises(ValueError):
        ShardedSearchClient({"a": None}, {"b": 0})
    if not ref_shim.available():
        return
    # the reference's own router over the same fake clients gives the same batch
    import importlib.util
    import sys
    import types

    mods = ref_shim.load()
    import vod_b200.search as vsearch

    base = types.ModuleType("vod_search.base")
    for name in ("SearchClient", "SectionId", "ShardName", "SubsetId"):
        setattr(base, name, getattr(vsearch, name))
    base.SearchMaster = type("SearchMaster", (), {"__class_getitem__": classmethod(lambda cls, item: cls)})
    pkg = types.ModuleType("vod_search")
    pkg.__path__ = []
    sys.modules.setdefault("vod_search", pkg)
    sys.modules["vod_search.base"] = base
    spec = importlib.util.spec_from_file_location("vod_search.sharded_search",
                                                  ref_shim.REFERENCE_ROOT / "src/vod_search/sharded_search.py")
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    RB = mods["retrieval"].RetrievalBatch
    ref_client = ref_mod.ShardedSearchClient({"wiki": _FakeClient(RB, 100, 1), "pubmed": _FakeClient(RB, 2, 2)},
                                             {"wiki": 0, "pubmed": 1000})
    ref_out = ref_client.search(text=text, vector=vec, shard=shard, top_k=4)
    assert np.array_equal(ref_out.indices, out.indices) and np.array_equal(ref_out.scores, out.scores)
    a, b = mine.shards["wiki"].calls[0], ref_client.shards["wiki"].calls[0]
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2:] == b[2:]
