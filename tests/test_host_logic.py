"""Host-side mirror of the reference interface (no GPU): RetrievalBatch, masters/clients, shard arithmetic."""
import pickle

import numpy as np
import pytest

import vod_b200
from vod_b200 import retrieval as rt
from vod_b200.search import B200SearchClient, B200SearchMaster, DoNotPickleError, SearchClient


def test_retrieval_batch_shape_checks():
    s, i = np.zeros((2, 3), np.float32), np.zeros((2, 3), np.int64)
    b = rt.RetrievalBatch(scores=s, indices=i)
    assert b.shape == (2, 3) and len(b) == 2 and b.labels is None and b.meta == {}
    with pytest.raises(ValueError):
        rt.RetrievalBatch(scores=np.zeros((2, 4), np.float32), indices=i)
    with pytest.raises(ValueError):
        rt.RetrievalBatch(scores=np.zeros(3, np.float32), indices=np.zeros(3, np.int64))
    with pytest.raises(ValueError):
        rt.RetrievalBatch(scores=s, indices=i, labels=np.zeros((2, 2)))


def test_retrieval_batch_ops_match_reference_semantics():
    b = rt.RetrievalBatch.cast(scores=[[1.0, 3.0, 2.0]], indices=[[7, 8, 9]], labels=[[0, 1, 0]])
    sb = b.sorted()
    assert sb.scores.tolist() == [[3.0, 2.0, 1.0]] and sb.indices.tolist() == [[8, 9, 7]] and sb.labels.tolist() == [[1, 0, 0]]
    assert (b * 2.0).scores.tolist() == [[2.0, 6.0, 4.0]]
    with pytest.raises(TypeError):
        b * "x"
    cat = b + b
    assert cat.shape == (2, 3)
    sample = b[0]
    assert isinstance(sample, rt.RetrievalSample) and sample.scores.shape == (3,)
    stacked = rt.RetrievalBatch.stack_samples([
        rt.RetrievalSample(scores=np.array([1.0, 2.0]), indices=np.array([4, 5])),
        rt.RetrievalSample(scores=np.array([3.0]), indices=np.array([6])),
    ])
    assert stacked.indices.tolist() == [[4, 5], [6, -1]]
    assert stacked.scores[1, 1] == -np.inf  # retrieval.py:283-285 padding
    assert b.to_dict()["indices"] == [[7, 8, 9]]
    import torch

    c = rt.RetrievalBatch.cast(scores=torch.zeros(1, 2), indices=torch.zeros(1, 2, dtype=torch.int64))
    assert isinstance(c.scores, np.ndarray)


def test_reference_batch_class_is_equivalent_when_available():
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ref_cls = ref_shim.load()["retrieval"].RetrievalBatch
    s = np.array([[0.5, 2.0, 1.0], [3.0, -1.0, 0.0]], np.float32)
    i = np.array([[1, 2, 3], [4, 5, 6]], np.int64)
    a, b = ref_cls(scores=s, indices=i).sorted(), rt.RetrievalBatch(scores=s, indices=i).sorted()
    assert np.array_equal(a.scores, b.scores) and np.array_equal(a.indices, b.indices)
    assert np.array_equal((ref_cls(scores=s, indices=i) * 0.5).scores, (rt.RetrievalBatch(scores=s, indices=i) * 0.5).scores)


def test_master_is_unpicklable_and_client_is_a_handle():
    m = B200SearchMaster(np.zeros((4, 8), np.float32), skip_setup=True)
    with pytest.raises(DoNotPickleError):
        pickle.dumps(m)
    c = m.get_client()
    assert isinstance(c, SearchClient) and c.requires_vectors is True
    c2 = pickle.loads(pickle.dumps(c))
    assert isinstance(c2, B200SearchClient) and c2.master_id == c.master_id
    assert c2.ping() is False  # master not entered
    with pytest.raises(vod_b200.VodbError):
        c2.search(vector=np.zeros((1, 8), np.float32), top_k=3)


def test_search_signature_is_keyword_only_like_the_reference():
    import inspect

    sig = inspect.signature(B200SearchClient.search)
    for name in ("vector", "text", "subset_ids", "ids", "shard", "top_k"):
        assert sig.parameters[name].kind is inspect.Parameter.KEYWORD_ONLY
    assert sig.parameters["top_k"].default == 3


def test_shard_bounds_cover_and_align():
    for n, g in [(10_000_000, 8), (100_000_000, 4), (1000, 2), (130, 8), (0, 2), (128, 1)]:
        prev = 0
        for r in range(g):
            lo, hi = vod_b200.shard_bounds(n, g, r)
            assert lo == prev and lo <= hi <= n
            assert lo % 128 == 0 or lo == n
            prev = hi
        assert prev == n
    with pytest.raises(ValueError):
        vod_b200.shard_bounds(10, 2, 2)


def test_sampling_argument_checks_do_not_need_a_gpu():
    with pytest.raises(ValueError):
        vod_b200.labeled_priority_sampling(np.zeros((2, 2, 2), np.float32), None)
    with pytest.raises(ValueError):
        vod_b200.labeled_priority_sampling(np.zeros((2, 5), np.float32), np.zeros((2, 5), bool), k_positive=4, k_total=2)
    with pytest.raises(ValueError):
        vod_b200.priority_sampling_1d(np.zeros((2, 5), np.float32))


def test_build_rejects_non_flat_and_matrix_vectors():
    with pytest.raises(ValueError):
        vod_b200.build_b200_index(np.zeros((4, 8), np.float32), factory_string="IVF16,Flat")
    with pytest.raises(ValueError):
        vod_b200.build_b200_index(np.zeros((4, 8, 2), np.float32))


def _plan(n_rows, nq, k, safe=False):
    import ctypes

    from vod_b200 import _lib

    lib = _lib.load()
    cap = ctypes.c_int()
    bounds = (ctypes.c_int64 * 65536)()
    n = lib.vodb_plan_scan(n_rows, nq, k, int(safe), ctypes.byref(cap), bounds, 65536)
    assert 2 <= n <= 65536, _lib.last_error()
    return cap.value, list(bounds[:n])


@pytest.mark.parametrize("n_rows", [1, 100, 4096, 5000, 1_250_000, 10_000_000, 100_000_001])
@pytest.mark.parametrize("nq,k", [(1, 1), (64, 100), (64, 1000), (300, 7), (8192, 100), (8192, 2048)])
def test_scan_schedule_invariants(n_rows, nq, k):
    """The C++ planner behind vodb_search (no GPU involved): segments tile the shard, start on 128-row tile
    boundaries, the dump segment fits a list, and the safe schedule can never overflow one."""
    for safe in (False, True):
        cap, b = _plan(n_rows, nq, k, safe)
        assert cap >= 4 * k and cap & (cap - 1) == 0 and nq * cap * 8 <= max(4 << 30, nq * 8192 * 8)
        assert b[0] == 0 and b[-1] == n_rows and all(x < y for x, y in zip(b, b[1:]))
        assert all(x % 128 == 0 for x in b[:-1])
        assert b[1] - b[0] <= cap                      # dump mode: slot = row - row_begin
        if safe:
            assert all(y - x <= cap - k for x, y in zip(b, b[1:]))   # k kept + every row of a segment still fits
    _, small = _plan(n_rows, 64, k)
    _, large = _plan(n_rows, 8192, k)
    assert len(large) >= len(small)                    # large batches refresh the thresholds more often


def test_scan_schedule_headline_shapes():
    cap, b = _plan(10_000_000, 64, 100)
    assert cap == 65536 and len(b) - 1 == 3 and b[1] == 16384           # BASELINE configs[1], 64 queries
    cap, b = _plan(1_250_000, 64, 100)
    assert len(b) - 1 == 2 and b[1] == 16384                            # its 8-GPU shard: dump segment + one scan
    cap, b = _plan(10_000_000, 8192, 100)
    assert len(b) - 1 == 7
    cap, b = _plan(12_500_000, 64, 1000)
    assert cap == 65536


def test_b200_factory_config_mirrors_the_faiss_config_surface():
    """`backend: "b200"` config + diff + fingerprint + factory function (src/vod_configs/search.py:110-153,
    src/vod_search/factory.py:131-190), host logic only."""
    import pydantic
    import pytest

    import vod_b200

    cfg = vod_b200.B200FactoryConfig(dtype="float16", devices=[0, 1], metric="inner_product")
    assert cfg.backend == "b200" and cfg.factory == "Flat" and cfg.metric == 0 and cfg.add_batch_size == 2**18
    merged = cfg + vod_b200.B200FactoryDiff(mode="tensor3", add_batch_size=1024)
    assert merged.mode == "tensor3" and merged.add_batch_size == 1024 and merged.dtype == "float16"
    assert (cfg + None) is cfg
    assert cfg.fingerprint() != merged.fingerprint()                                   # mode changes the results
    assert cfg.fingerprint() == cfg.model_copy(update={"serve": False}).fingerprint()  # serving details do not
    with pytest.raises(pydantic.ValidationError):
        vod_b200.B200FactoryConfig(metric="l2")
    with pytest.raises(pydantic.ValidationError):
        vod_b200.B200FactoryConfig(nprobe=16)  # strict model: unknown keys are rejected like StrictModel does
    master = vod_b200.build_b200_search([[0.0, 1.0]], config={"backend": "b200", "dtype": "float32"})
    assert isinstance(master, vod_b200.B200SearchMaster) and master.dtype == "float32" and master.store is None
    with pytest.raises(ValueError):
        vod_b200.build_b200_search([[0.0]], config={"factory": "IVF100,Flat"})


def test_scan_schedule_invariants_for_random_shapes():
    """Property check of the C++ planner (vodb_plan_scan): for any shard size / batch / k the segments tile the shard,
    start on tile boundaries, the dump segment fits a list, the expected survivors of every later segment
    (k * rows(segment) / rows(before), rows in random order) stay within an eighth of the list, and the safe schedule
    cannot overflow."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=200, deadline=None)
    @hyp.given(n_rows=st.integers(1, 200_000_000), nq=st.integers(1, 8192), k=st.integers(1, 2048))
    def check(n_rows, nq, k):
        cap, b = _plan(n_rows, nq, k)
        assert b[0] == 0 and b[-1] == n_rows and all(x < y for x, y in zip(b, b[1:]))
        assert all(x % 128 == 0 for x in b[:-1]) and b[1] <= cap and cap >= 4 * k
        for lo, hi in zip(b[1:], b[2:]):
            # growth rule: k * seg / before <= cap / 8; a tail shorter than a quarter segment is folded into the last
            # one (x1.25) and segment lengths are rounded to 128 rows
            assert k * (hi - lo - 128) / lo <= 1.25 * cap / 8 + 1e-9, (b, cap)
        cap_s, bs = _plan(n_rows, nq, k, True)
        assert all(y - x <= cap_s - k for x, y in zip(bs, bs[1:]))

    check()
