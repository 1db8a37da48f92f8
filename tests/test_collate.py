"""Host glue after sampling (in-batch negatives, padding ids) against the reference's own functions, loaded from
/root/reference when present (build container) and against fixed expectations otherwise."""
import numpy as np
import pytest

import vod_b200
from vod_b200 import collate


def _samples(rng, B=6, k=5, pool=12):
    idx = np.stack([rng.choice(pool, size=k, replace=False) for _ in range(B)]).astype(np.int64)
    idx[0, -1] = -1
    scores = rng.standard_normal((B, k)).astype(np.float32)
    labels = rng.random((B, k)) < 0.3
    return vod_b200.PrioritySampledSections(
        batch=vod_b200.RetrievalBatch(scores=scores, indices=idx, labels=labels),
        log_weights=rng.standard_normal((B, k)).astype(np.float32), max_sampling_id=np.arange(B, dtype=np.float32),
        lse_pos=np.zeros(B, np.float32), lse_neg=np.ones(B, np.float32),
        raw_scores={"dense": scores * 2, "sparse": rng.standard_normal((B, k)).astype(np.float32)})


def test_gather_values_first_match_and_fill():
    q = np.array([[3, 9, 4], [1, 1, 7]])
    keys = np.array([[4, 3, 3], [7, 2, 1]])
    vals = np.array([[0.5, 1.5, 2.5], [10.0, 20.0, 30.0]], np.float32)
    out = collate.gather_values_by_indices(q, keys, vals)
    assert np.array_equal(out, np.array([[1.5, np.nan, 0.5], [30.0, 30.0, 10.0]], np.float32), equal_nan=True)
    lab = collate.gather_values_by_indices(q, keys, vals > 1, fill_value=0)
    assert lab.dtype == np.bool_ and np.array_equal(lab, [[True, False, False], [True, True, True]])
    assert np.array_equal(collate.gather_values_by_indices(q[0], keys[0], np.array([7, 8, 9])), [8, -1, 7])
    assert np.array_equal(collate.gather_values_by_indices(q, keys[1], vals[1]),
                          np.array([[np.nan, np.nan, np.nan], [30, 30, 10]], np.float32), equal_nan=True)
    with pytest.raises(ValueError):
        collate.gather_values_by_indices(q, keys[:1], vals[:1])


@pytest.mark.parametrize("padding", [True, False])
def test_flatten_samples_shapes_and_content(padding):
    s = _samples(np.random.default_rng(0))
    flat = collate.flatten_samples(s, padding=padding)
    B, k = s.batch.indices.shape
    uniq = np.unique(s.batch.indices)
    U = B * k if padding else len(uniq)
    assert flat.batch.indices.shape == (U,) and flat.batch.scores.shape == (B, U) and flat.log_weights.shape == (B, U)
    assert np.array_equal(flat.batch.indices[: len(uniq)], uniq) and (flat.batch.indices[len(uniq):] == 1).all()
    for b in range(B):
        for j in range(k):
            u = int(np.flatnonzero(flat.batch.indices == s.batch.indices[b, j])[0])
            assert flat.batch.scores[b, u] == s.batch.scores[b, j]
            assert flat.raw_scores["sparse"][b, u] == s.raw_scores["sparse"][b, j]
            assert flat.batch.labels[b, u] == s.batch.labels[b, j]
    assert np.isnan(flat.batch.scores).sum() >= B * (len(uniq) - k)
    with pytest.raises(ValueError):
        s.batch.labels = None
        collate.flatten_samples(s)


def test_replace_negative_indices_in_place():
    a = np.array([[3, -1, 5], [-1, -1, 2]])
    np.random.seed(0)
    collate.replace_negative_indices_(a, world_size=10)
    assert (a >= 0).all() and (a < 10).all() and a[0, 0] == 3 and a[1, 2] == 2


def test_against_the_reference_functions():
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference tree not present (GPU box)")
    import importlib.util
    import sys

    mods = ref_shim.load()
    path = ref_shim.REFERENCE_ROOT / "src" / "vod_dataloaders" / "core" / "in_batch_negatives.py"
    spec = importlib.util.spec_from_file_location("vod_dataloaders.core.in_batch_negatives", path)
    ibn = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = ibn
    spec.loader.exec_module(ibn)
    rng = np.random.default_rng(7)
    for trial in range(5):
        s = _samples(rng, B=4 + trial, k=3 + trial, pool=10 + 3 * trial)
        ref_in = mods["sample"].PrioritySampledSections(
            batch=mods["retrieval"].RetrievalBatch(scores=s.batch.scores.copy(), indices=s.batch.indices.copy(),
                                                   labels=s.batch.labels.copy()),
            log_weights=s.log_weights.copy(), max_sampling_id=s.max_sampling_id, lse_pos=s.lse_pos, lse_neg=s.lse_neg,
            raw_scores={k: v.copy() for k, v in s.raw_scores.items()})
        for padding in (True, False):
            ref = ibn.flatten_samples(ref_in, padding=padding)
            got = collate.flatten_samples(s, padding=padding)
            assert np.array_equal(got.batch.indices, ref.batch.indices)
            assert np.array_equal(got.batch.scores, ref.batch.scores, equal_nan=True)
            assert np.array_equal(got.batch.labels, ref.batch.labels)
            assert np.array_equal(got.log_weights, ref.log_weights, equal_nan=True)
            for key in s.raw_scores:
                assert np.array_equal(got.raw_scores[key], ref.raw_scores[key], equal_nan=True)
    np.random.seed(3)
    a = np.array([[3, -1, 5], [-1, -1, 2]])
    b = a.copy()
    collate.replace_negative_indices_(a, 100)
    np.random.seed(3)
    mods["numpy_ops"].replace_negative_indices_(b, 100)
    assert np.array_equal(a, b)
