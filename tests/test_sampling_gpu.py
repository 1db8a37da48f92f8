"""Parity of the fused CUDA sampling kernel (through the C ABI): bit-identical to the CPU twin, faithful to the
reference's numba sampler (golden vectors), and statistically unbiased (ports of the reference's tests)."""
import collections

import numpy as np
import pytest

import vod_b200
from tests.helpers import assert_faithful_to_reference, golden_cases

pytestmark = pytest.mark.gpu


def _same_bits(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def _run_both(twin, scores, labels, **kw):
    noise = kw.pop("noise", None)
    fix = kw.pop("fix_truncation", False)
    g = vod_b200.labeled_priority_sampling(scores, labels, noise=noise, fix_truncation=fix, **kw)
    t = twin.sample(scores, labels, k_positive=kw.get("k_positive", 1), k_total=kw.get("k_total", 2),
                    normalized=kw.get("normalized", True), temperature=kw.get("temperature", 1.0),
                    max_support=(max(kw["max_support_size"], kw.get("k_total", 2)) if kw.get("max_support_size") else -1),
                    quirks=0 if fix else 1, noise=noise, seed=kw.get("seed", 0) or 0, offset=kw.get("offset", 0))
    return g, t


@pytest.mark.parametrize("K", [1, 7, 100, 1000, 4096, 8192])
@pytest.mark.parametrize("k_total,k_positive", [(1, 0), (8, 3), (32, 8)])
def test_bit_identical_to_cpu_twin_philox(twin, K, k_total, k_positive):
    rng = np.random.default_rng(K * 31 + k_total)
    B = 5
    scores = (rng.normal(size=(B, K)) * 3).astype(np.float32)
    scores[rng.uniform(size=(B, K)) < 0.1] = -np.inf
    if K > 10:
        scores[0, 3] = np.nan
        scores[1, :] = -np.inf          # an all -inf row
    labels = rng.uniform(size=(B, K)) < 0.05
    for temperature in (1.0, 0.0):
        for normalized in (True, False):
            for ms in (None, 50):
                (gi, gw, gl, glse), (ti, tw, tl, tlse) = _run_both(
                    twin, scores, labels, k_positive=k_positive, k_total=k_total, normalized=normalized,
                    temperature=temperature, max_support_size=ms, seed=1234 + K, offset=77)
                assert np.array_equal(gi, ti), (temperature, normalized, ms)
                assert np.array_equal(gl, tl)
                assert _same_bits(gw, tw), (temperature, normalized, ms, gw, tw)
                assert _same_bits(glse, tlse)


def test_bit_identical_with_injected_noise_and_fixed_truncation(twin):
    rng = np.random.default_rng(0)
    scores = np.sort(rng.normal(size=(32, 1000)).astype(np.float32) * 4, axis=1)[:, ::-1].copy()
    scores -= scores.min(axis=1, keepdims=True)
    labels = np.zeros((32, 1000), bool)
    labels[:, :3] = True
    noise = rng.exponential(size=(32, 1000)).astype(np.float32)
    for fix in (False, True):
        (gi, gw, gl, glse), (ti, tw, tl, tlse) = _run_both(twin, scores, labels, k_positive=3, k_total=8,
                                                          max_support_size=100, noise=noise, fix_truncation=fix)
        assert np.array_equal(gi, ti) and np.array_equal(gl, tl) and _same_bits(gw, tw) and _same_bits(glse, tlse)


def test_faithful_to_reference_golden(golden):
    n = 0
    for case in golden_cases(golden):
        ms = case["max_support"]
        ids, logw, olab, lse = vod_b200.labeled_priority_sampling(
            case["scores"], case["labels"], k_positive=case["k_positive"], k_total=case["k_total"], normalized=True,
            temperature=case["temperature"], max_support_size=None if ms < 0 else ms, noise=case["noise"])
        assert_faithful_to_reference(case, ids, logw, olab, lse)
        n += 1
    assert n == 144


def test_config4_realm_collate_shapes(twin):
    """BASELINE config 4: 32-query batches, top-K=1000 retrieval, priority sampling of k=8 with importance weights."""
    rng = np.random.default_rng(42)
    scores = np.sort(rng.normal(size=(32, 1000)).astype(np.float32) * 5 + 100, axis=1)[:, ::-1].copy()
    indices = rng.permutation(10_000_000)[:32 * 1000].reshape(32, 1000).astype(np.int64)
    labels = np.zeros((32, 1000), np.int64)
    for b in range(32):
        labels[b, : b % 4] = 1
    batch = vod_b200.RetrievalBatch(scores=scores - scores.min(axis=1, keepdims=True), indices=indices, labels=labels)
    out = vod_b200.sample_search_results(search_results=batch, raw_scores={"dense": scores}, total=8,
                                         max_pos_sections=3, temperature=1.0, seed=42, offset=0)
    assert out.batch.indices.shape == (32, 8) and out.log_weights.shape == (32, 8)
    assert out.batch.labels.dtype == np.bool_ and out.raw_scores["dense"].shape == (32, 8)
    assert out.max_sampling_id.shape == (32,) and out.lse_pos.shape == (32,) and out.lse_neg.shape == (32,)
    ti, tw, tl, tlse = twin.sample(batch.scores, labels > 0, k_positive=3, k_total=8, seed=42, offset=0)
    assert np.array_equal(out.batch.indices, np.take_along_axis(indices, ti, axis=-1))
    assert _same_bits(out.log_weights, tw) and np.array_equal(out.batch.labels, tl)
    # per-label self-normalisation: weights of each label group sum to one
    w = np.exp(out.log_weights.astype(np.float64))
    for b in range(32):
        for lab in (True, False):
            sel = out.batch.labels[b] == lab
            if sel.any():
                assert abs(w[b][sel].sum() - 1.0) < 1e-5
    # same seed/offset -> same sample; another offset -> another sample
    again = vod_b200.sample_search_results(search_results=batch, raw_scores={}, total=8, max_pos_sections=3, seed=42)
    other = vod_b200.sample_search_results(search_results=batch, raw_scores={}, total=8, max_pos_sections=3, seed=42, offset=1)
    assert np.array_equal(again.batch.indices, out.batch.indices)
    assert not np.array_equal(other.batch.indices, out.batch.indices)


def test_numpy_global_seed_controls_default_noise():
    scores = np.random.default_rng(0).normal(size=(4, 64)).astype(np.float32)
    np.random.seed(123)
    a = vod_b200.labeled_priority_sampling(scores, np.zeros_like(scores, bool), k_positive=0, k_total=4)
    np.random.seed(123)
    b = vod_b200.labeled_priority_sampling(scores, np.zeros_like(scores, bool), k_positive=0, k_total=4)
    assert np.array_equal(a[0], b[0]) and _same_bits(a[1], b[1])


def test_float64_and_1d_inputs():
    s = np.random.default_rng(1).normal(size=50)
    ids, logw, lab, lse = vod_b200.labeled_priority_sampling(s, np.zeros(50, bool), k_positive=0, k_total=5, seed=3)
    assert ids.shape == (5,) and logw.dtype == np.float64 and lse.shape == (2,)
    z, lw = vod_b200.priority_sampling_1d(s.astype(np.float32), k=5, seed=3)
    assert z.dtype == np.int64 and lw.dtype == np.float32 and len(z) == 5


# ---- ports of the reference's statistical tests (src/vod_dataloaders/tests/test_priority_sampling.py) ---------

def _softmax(x):
    x = np.where(np.isnan(x), -np.inf, x).astype(np.float64)
    m = x.max() if np.isfinite(x.max()) else 0.0
    e = np.exp(x - m)
    return e / e.sum()


@pytest.mark.parametrize("seed", list(range(10)))
@pytest.mark.parametrize("n_trials,n,k,inf_frac", [(100, 100, 10, 0), (1000, 100, 10, 0), (100, 100, 100, 0),
                                                  (1000, 100, 10, 0.5), (1000, 100, 10, 95)])
def test_priority_sampling_1d_unbiased(seed, n_trials, n, k, inf_frac):
    rgn = np.random.default_rng(seed)
    f = rgn.normal(size=n).astype(np.float32)
    unorm_log_p = rgn.uniform(size=n).astype(np.float32)
    if inf_frac > 0:
        unorm_log_p[rgn.uniform(size=n) < inf_frac] = -np.inf
    if np.all(unorm_log_p == -np.inf):
        m = rgn.uniform(size=n) < (1 - inf_frac)
        unorm_log_p = np.where(m, unorm_log_p, rgn.normal(size=len(unorm_log_p))).astype(np.float32)
    mu = np.sum(_softmax(unorm_log_p) * f)
    # the reference loops n_trials calls of priority_sampling_1d; one batched call draws the same number of samples
    z, log_w, _, _ = vod_b200.labeled_priority_sampling(np.repeat(unorm_log_p[None], n_trials, 0), None, k_positive=0,
                                                        k_total=min(k, n), normalized=False, seed=seed + 1000)
    assert np.all(~np.isnan(log_w)) and z.dtype == np.int64 and log_w.dtype == unorm_log_p.dtype
    mu_hats = [np.sum(_softmax(log_w[i]) * np.take(f, z[i])) for i in range(n_trials)]
    assert np.isclose(mu, np.mean(mu_hats), atol=10.0 / np.sqrt(n_trials * k))


@pytest.mark.parametrize("seed", list(range(10)))
@pytest.mark.parametrize("label_thres", [0.5, 0, 1])
def test_labeled_priority_sampling_unbiased(seed, label_thres, n_trials=3000, n=32, k_positive=4, k_total=8):
    rgn = np.random.default_rng(seed)
    f = rgn.normal(size=n).astype(np.float32)
    unorm_log_p = rgn.uniform(size=n).astype(np.float32)
    unorm_log_p[unorm_log_p < 0.2] = -np.inf
    labels = np.where(rgn.normal(size=n) > label_thres, 1, 0)
    mu_a = np.sum(_softmax(unorm_log_p[labels == 1]) * f[labels == 1]) if np.sum(labels == 1) > 0 else None
    mu_b = np.sum(_softmax(unorm_log_p[labels == 0]) * f[labels == 0]) if np.sum(labels == 0) > 0 else None
    z_, log_w_, ls_, _ = vod_b200.labeled_priority_sampling(
        unorm_log_p[None].repeat(n_trials, axis=0), labels[None].repeat(n_trials, axis=0), k_positive=k_positive,
        k_total=k_total, normalized=False, seed=seed)
    assert np.all(~np.isnan(log_w_))
    mu_a_hats, mu_b_hats = [], []
    for i in range(n_trials):
        z, log_w, ls = z_[i], log_w_[i], ls_[i]
        counts = collections.Counter(z[z >= 0])
        assert max(counts.values()) == 1
        if mu_a is not None:
            mu_a_hats.append(np.sum(_softmax(log_w[ls == 1]) * np.take(f, z[ls == 1])))
        if mu_b is not None:
            sel = (ls == 0) & (z >= 0)
            mu_b_hats.append(np.sum(_softmax(log_w[sel]) * np.take(f, z[sel])))
    if mu_a is not None:
        assert np.isclose(mu_a, np.mean(mu_a_hats), atol=10.0 / np.sqrt(n_trials * min(k_positive, np.sum(labels == 1))))
    if mu_b is not None:
        assert np.isclose(mu_b, np.mean(mu_b_hats), atol=10.0 / np.sqrt(n_trials * min(k_total - k_positive, np.sum(labels == 0))))


def test_device_resident_retrieve_then_sample_equals_host_path():
    """Config 4 chain (32 queries, top-K=1000, sample 8): the device-resident pipeline returns the same picks and
    weights as search -> numpy -> sample_search_results, with a single [B,8] device->host transfer."""
    from tests.helpers import int_valued

    rng = np.random.default_rng(2)
    n, d, B = 50_000, 128, 32
    st = vod_b200.CorpusStore(n, d, dtype="bfloat16")
    st.add(int_valued(rng, (n, d)))
    xq = int_valued(rng, (B, d))
    s, i = st.search(xq, 1000, mode="tensor")
    gold = np.stack([i[b, rng.choice(50, size=2, replace=False)] for b in range(B)])  # two retrieved ids per row are "gold"
    labels = (i[:, :, None] == gold[:, None, :]).any(-1).astype(np.int64)
    host = vod_b200.sample_search_results(search_results=vod_b200.RetrievalBatch(scores=s, indices=i, labels=labels),
                                          raw_scores={"dense": s}, total=8, max_pos_sections=3, seed=5, offset=2)
    pipe = vod_b200.DenseRetrievalSampler(st, top_k=1000, total=8, max_pos_sections=3, mode="tensor")
    dev = pipe(xq, positive_ids=gold, seed=5, offset=2)
    assert np.array_equal(dev.batch.indices, host.batch.indices)
    assert np.array_equal(dev.batch.labels, host.batch.labels)
    assert _same_bits(dev.log_weights, host.log_weights)
    assert np.array_equal(dev.batch.scores, host.batch.scores)
    assert np.array_equal(dev.max_sampling_id, host.max_sampling_id) and dev.max_sampling_id.dtype == np.float32
    assert _same_bits(dev.lse_pos, host.lse_pos) and _same_bits(dev.lse_neg, host.lse_neg)
    assert np.array_equal(dev.raw_scores["dense"], host.raw_scores["dense"])
    assert dev.batch.labels[:, :2].all()
    # no gold ids, deterministic top-`total` (temperature 0), device-resident fp16-free torch queries
    import torch

    host0 = vod_b200.sample_search_results(search_results=vod_b200.RetrievalBatch(scores=s, indices=i, labels=None),
                                           raw_scores={"dense": s}, total=8, max_pos_sections=3, temperature=0.0, seed=1)
    pipe0 = vod_b200.DenseRetrievalSampler(st, top_k=1000, total=8, max_pos_sections=3, mode="tensor", temperature=0.0)
    dev0 = pipe0(torch.from_numpy(xq).cuda(), seed=1)
    assert np.array_equal(dev0.batch.indices, host0.batch.indices) and _same_bits(dev0.log_weights, host0.log_weights)
    assert np.array_equal(dev0.max_sampling_id, host0.max_sampling_id)
    assert np.array_equal(pipe0.last_local_ids, np.tile(np.arange(8), (B, 1)))
    st.close()


def test_chain_with_fewer_rows_than_picks_wraps_like_numpy():
    """total > finite candidates: unused sampler slots (-1) gather the last retrieved column, as
    np.take_along_axis does in the reference (core/sample.py:57-58); max_sampling_id counts finite negatives."""
    from tests.helpers import int_valued

    rng = np.random.default_rng(3)
    st = vod_b200.CorpusStore(5, 64, dtype="float32")
    st.add(int_valued(rng, (5, 64)))
    xq = int_valued(rng, (3, 64))
    s, i = st.search(xq, 8, mode="exact")
    gold = i[:, :1].copy()
    labels = (i[:, :, None] == gold[:, None, :]).any(-1).astype(np.int64)
    host = vod_b200.sample_search_results(search_results=vod_b200.RetrievalBatch(scores=s, indices=i, labels=labels),
                                          raw_scores={"dense": s}, total=8, max_pos_sections=2, seed=9, offset=1)
    dev = vod_b200.DenseRetrievalSampler(st, top_k=8, total=8, max_pos_sections=2, mode="exact")(xq, gold, seed=9, offset=1)
    assert np.array_equal(dev.batch.indices, host.batch.indices)
    assert np.array_equal(dev.batch.scores, host.batch.scores)
    assert np.array_equal(dev.batch.labels, host.batch.labels)
    assert _same_bits(dev.log_weights, host.log_weights)
    assert np.array_equal(dev.max_sampling_id, host.max_sampling_id)
    with pytest.raises(ValueError):
        vod_b200.DenseRetrievalSampler(st, top_k=8, total=4, max_pos_sections=5)(xq)
    st.close()


def test_sample_search_results_matches_the_reference_function():
    """vod_b200.sample_search_results (sampler kernel + gather kernel, one C call) against the outputs of the
    reference's own sample_search_results on the same inputs and the same Exp(1) noise (50 golden cases)."""
    import pathlib

    from tests.helpers import assert_results_match_reference, results_cases

    npz = np.load(pathlib.Path(__file__).parent / "golden" / "sample_results_ref.npz")
    n = 0
    for case in results_cases(npz):
        batch = vod_b200.RetrievalBatch(scores=case["scores"], indices=case["indices"], labels=case["labels"])
        out = vod_b200.sample_search_results(search_results=batch, raw_scores={"dense": case["scores"], "sparse": case["sparse"]},
                                             total=case["total"], max_pos_sections=case["k_positive"],
                                             temperature=case["temperature"], max_support_size=case["support"],
                                             noise=case["noise"])
        assert_results_match_reference(case, out.batch.indices, out.batch.scores, out.batch.labels, out.log_weights,
                                       out.max_sampling_id, out.lse_pos, out.lse_neg, out.raw_scores)
        n += 1
    assert n == 50
