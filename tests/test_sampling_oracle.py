"""The CPU twin (oracle/sample_twin.c) pinned against the reference's own numba sampler.

  * golden vectors generated from the reference (tests/golden/make_golden.py);
  * the live reference when /root/reference is present (build container only);
  * the reference's statistical unit tests (src/vod_dataloaders/tests/test_priority_sampling.py) ported to the twin.
"""
import collections

import numpy as np
import pytest

from tests.helpers import assert_faithful_to_reference, golden_cases


def test_twin_matches_reference_golden(golden, twin):
    n = 0
    for case in golden_cases(golden):
        ids, logw, olab, lse = twin.sample(case["scores"], case["labels"], k_positive=case["k_positive"],
                                           k_total=case["k_total"], normalized=True, temperature=case["temperature"],
                                           max_support=case["max_support"], quirks=1, noise=case["noise"])
        assert_faithful_to_reference(case, ids, logw, olab, lse)
        n += 1
    assert n == 144


def test_twin_matches_live_reference(twin):
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference tree not present (GPU box)")
    import warnings

    warnings.filterwarnings("ignore")
    ref = ref_shim.load()["sample"]
    rng = np.random.default_rng(11)
    for K, kt, kp, temp in [(64, 8, 2, 1.0), (333, 16, 4, 1.0), (1000, 8, 3, 0.0), (17, 32, 5, 1.0)]:
        B = 3
        scores = (rng.normal(size=(B, K)) * 2).astype(np.float32)
        labels = rng.uniform(size=(B, K)) < 0.05
        noise = rng.exponential(size=(B, K)).astype(np.float32)
        s = np.full((B, kt), -1, np.int64)
        w = np.full((B, kt), -np.inf, np.float32)
        l = np.zeros((B, kt), np.bool_)
        c = np.zeros((B, 2), np.float32)
        ref._labeled_priority_sampling_2d_(scores.copy(), labels.copy(), noise.copy(), kp, kt, s, w, l, c, True, temp, -1)
        ids, logw, olab, lse = twin.sample(scores, labels, k_positive=kp, k_total=kt, temperature=temp, noise=noise)
        assert np.array_equal(ids, s)
        assert np.array_equal(olab, l)
        fin = np.isfinite(w)
        assert np.array_equal(fin, np.isfinite(logw))
        assert np.abs(w[fin] - logw[fin]).max() < 3e-5


def test_unused_slots_and_short_rows(twin):
    scores = np.array([[0.5, 0.1, -np.inf]], np.float32)
    labels = np.array([[True, False, False]])
    ids, logw, olab, lse = twin.sample(scores, labels, k_positive=1, k_total=8, temperature=0.0)
    # k_total is clipped to K=3 (sample.py:267); only 1 finite negative -> k_positive grows to 2 (sample.py:277-278)
    assert ids.shape == (1, 8)
    assert list(ids[0, :3]) == [0, 1, 2] and (ids[0, 3:] == -1).all()
    assert list(olab[0, :3]) == [True, False, False]
    assert np.isneginf(logw[0, 3:]).all()
    assert logw[0, 0] == 0.0  # single positive, self-normalised
    assert logw[0, 2] == -np.inf  # -inf scored pick keeps weight -inf (reference intent)


def test_temperature_zero_is_deterministic_topk(twin):
    rng = np.random.default_rng(5)
    scores = rng.normal(size=(4, 200)).astype(np.float32)
    ids, logw, _, _ = twin.sample(scores, None, k_positive=0, k_total=10, temperature=0.0, normalized=False)
    for b in range(4):
        assert list(ids[b]) == list(np.argsort(-scores[b], kind="stable")[:10])


def test_inverted_truncation_quirk(twin):
    """max_support masks the TOP entries out by default (reference sample.py:176-178); quirks=0 is the fixed version."""
    scores = -np.arange(300, dtype=np.float32)[None] / 10
    ids_q, _, _, _ = twin.sample(scores, None, k_positive=0, k_total=5, temperature=0.0, max_support=100, quirks=1)
    ids_f, _, _, _ = twin.sample(scores, None, k_positive=0, k_total=5, temperature=0.0, max_support=100, quirks=0)
    assert list(ids_q[0]) == [100, 101, 102, 103, 104]
    assert list(ids_f[0]) == [0, 1, 2, 3, 4]


# ---- ports of the reference's statistical tests (test_priority_sampling.py:8-110) -----------------------

def _softmax(x):
    x = np.where(np.isnan(x), -np.inf, x).astype(np.float64)
    m = x.max() if np.isfinite(x.max()) else 0.0
    e = np.exp(x - m)
    return e / e.sum()


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("n_trials,n,k,inf_frac", [(1000, 100, 10, 0), (100, 100, 100, 0), (1000, 100, 10, 0.5)])
def test_priority_sampling_unbiased(twin, seed, n_trials, n, k, inf_frac):
    rgn = np.random.default_rng(seed)
    f = rgn.normal(size=n).astype(np.float32)
    unorm_log_p = rgn.uniform(size=n).astype(np.float32)
    if inf_frac > 0:
        unorm_log_p[rgn.uniform(size=n) < inf_frac] = -np.inf
    mu = np.sum(_softmax(unorm_log_p) * f)
    z, log_w, _, _ = twin.sample(np.repeat(unorm_log_p[None], n_trials, 0), None, k_positive=0, k_total=min(k, n),
                                 normalized=False, seed=seed, offset=17)
    assert not np.isnan(log_w).any()
    assert z.dtype == np.int64
    mu_hats = [np.sum(_softmax(log_w[i]) * np.take(f, z[i])) for i in range(n_trials)]
    atol = 10.0 / np.sqrt(n_trials * k)
    assert np.isclose(mu, np.mean(mu_hats), atol=atol)


@pytest.mark.parametrize("seed", [0, 1, 2, 9])
@pytest.mark.parametrize("label_thres", [0.5, 0, 1])
def test_labeled_priority_sampling_unbiased(twin, seed, label_thres, n_trials=3000, n=32, k_positive=4, k_total=8):
    rgn = np.random.default_rng(seed)
    f = rgn.normal(size=n).astype(np.float32)
    unorm_log_p = rgn.uniform(size=n).astype(np.float32)
    unorm_log_p[unorm_log_p < 0.2] = -np.inf
    labels = np.where(rgn.normal(size=n) > label_thres, 1, 0)
    mu_a = np.sum(_softmax(unorm_log_p[labels == 1]) * f[labels == 1]) if np.sum(labels == 1) > 0 else None
    mu_b = np.sum(_softmax(unorm_log_p[labels == 0]) * f[labels == 0]) if np.sum(labels == 0) > 0 else None
    z_, log_w_, ls_, _ = twin.sample(np.repeat(unorm_log_p[None], n_trials, 0), np.repeat(labels[None], n_trials, 0),
                                     k_positive=k_positive, k_total=k_total, normalized=False, seed=seed)
    assert not np.isnan(log_w_).any()  # the 2 cases numba 0.65 fails pass here: -inf, not NaN
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


def test_twin_plus_numpy_glue_reproduces_the_reference_sample_search_results(twin):
    """The whole reference function (sample.py:22-84), not just its numba core: 50 cases generated by
    tests/golden/make_golden_results.py from the reference's own sample_search_results with recorded noise."""
    import pathlib

    from tests.helpers import assert_results_match_reference, results_cases

    npz = np.load(pathlib.Path(__file__).parent / "golden" / "sample_results_ref.npz")
    n = 0
    for case in results_cases(npz):
        lab = None if case["labels"] is None else case["labels"] > 0
        ms = case["support"]
        local, logw, olab, lse = twin.sample(case["scores"], lab, k_positive=case["k_positive"], k_total=case["total"],
                                             temperature=case["temperature"],
                                             max_support=-1 if ms is None else max(ms, case["total"]), noise=case["noise"])
        take = lambda a: np.take_along_axis(a, local, axis=-1)  # noqa: E731
        picked = take(case["scores"])
        neg_ref = np.ones_like(case["scores"], bool) if lab is None else ~lab
        floor = np.amin(np.where(~olab & np.isfinite(picked), picked, np.inf), axis=-1, keepdims=True)
        msid = (neg_ref & np.isfinite(case["scores"]) & (case["scores"] >= floor)).astype(np.float32).sum(-1)
        assert_results_match_reference(case, take(case["indices"]), picked, olab, logw, msid, lse[:, 0], lse[:, 1],
                                       {"dense": picked, "sparse": take(case["sparse"])})
        n += 1
    assert n == 50
