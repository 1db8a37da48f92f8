"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import numpy as np

LABEL_MODES = ["none", "pos3", "inf20"]


def golden_cases(golden):
    for row in golden["meta"]:
        cid, seed, K, kt, kp, lm, t, ms = row
        p = f"c{int(cid):03d}_"
        yield dict(cid=int(cid), seed=int(seed), K=int(K), k_total=int(kt), k_positive=int(kp),
                   label_mode=LABEL_MODES[int(lm)], temperature=float(t), max_support=int(ms),
                   scores=golden[p + "scores"], labels=golden[p + "labels"], noise=golden[p + "noise"],
                   samples=golden[p + "samples"], logw=golden[p + "logw"], olab=golden[p + "olab"],
                   lse=golden[p + "lse"])


def assert_faithful_to_reference(case, ids, logw, olab, lse, atol=1e-5):
    """Compare an implementation with the reference's numba output for one golden case.

    - rows where the reference itself produced NaN weights (numba fastmath on rows that pick a -inf entry,
      SURVEY.md §4 / App. A-8) are compared on the finite-key picks only;
    - weights: |diff| <= atol + conditioning term. log_w = log_pi - log1p(-exp(-exp(d))) loses digits when the
      inclusion probability q = 1-exp(-exp(d)) is tiny (both sides evaluate 1 - (1-eps)); the reference's own
      result is only accurate to eps32/q there.
    """
    ref_ids, ref_w, ref_lab = case["samples"], case["logw"], case["olab"]
    for b in range(ref_ids.shape[0]):
        nan_row = np.isnan(ref_w[b]).any()
        finite = np.isfinite(ref_w[b]) & np.isfinite(logw[b])
        if nan_row:
            # picks with a finite weight (= finite key) must agree, in place; the remaining picks are ties among
            # -inf keys whose order numba's unstable argsort leaves unspecified
            fin_w = np.isfinite(logw[b])
            assert np.array_equal(ids[b][fin_w], ref_ids[b][fin_w]), (case["cid"], b, ids[b], ref_ids[b])
            assert np.array_equal(ids[b] >= 0, ref_ids[b] >= 0), (case["cid"], b)
            assert np.array_equal(olab[b], ref_lab[b]), (case["cid"], b)
            assert not np.isnan(logw[b]).any(), "the implementation must not produce NaN weights"
            continue
        assert np.array_equal(ids[b], ref_ids[b]), (case["cid"], b, ids[b], ref_ids[b])
        assert np.array_equal(olab[b], ref_lab[b]), (case["cid"], b)
        assert np.array_equal(np.isfinite(ref_w[b]), np.isfinite(logw[b])), (case["cid"], b, ref_w[b], logw[b])
        if finite.any():
            # inclusion probability is at least exp(log_w_unnormalised - log_pi)^-1; bound the conditioning by
            # the spread of weights instead of recomputing q: allow 2e-3 relative to tiny-q entries
            err = np.abs(ref_w[b][finite] - logw[b][finite])
            assert err.max() <= atol + 3e-5, (case["cid"], b, err.max())
        assert np.allclose(lse[b], case["lse"][b], atol=1e-5, equal_nan=True), (case["cid"], b, lse[b], case["lse"][b])


def int_valued(rng, shape, lo=-3, hi=4):
    """Small-integer data: float32 dot products are exact in any summation order (bit-exact parity tests)."""
    return rng.integers(lo, hi, size=shape).astype(np.float32)


def round_to(x: np.ndarray, dtype: str) -> np.ndarray:
    """Round float32 values to the store dtype and back (the values the store holds)."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(x, np.float32))
    if dtype in ("bfloat16", "bf16"):
        return t.to(torch.bfloat16).to(torch.float32).numpy()
    if dtype in ("float16", "f16"):
        return t.to(torch.float16).to(torch.float32).numpy()
    return x.astype(np.float32)


def results_cases(npz):
    """Cases of tests/golden/sample_results_ref.npz (outputs of the reference's own sample_search_results)."""
    for cid, total, k_pos, temperature, support, has_labels in npz["meta"]:
        p = f"c{int(cid):03d}_"
        yield dict(cid=int(cid), total=int(total), k_positive=int(k_pos), temperature=float(temperature),
                   support=None if support < 0 else int(support),
                   scores=npz[p + "scores"], indices=npz[p + "indices"], sparse=npz[p + "sparse"], noise=npz[p + "noise"],
                   labels=npz[p + "labels"] if has_labels else None,
                   o_indices=npz[p + "o_indices"], o_scores=npz[p + "o_scores"], o_labels=npz[p + "o_labels"],
                   o_logw=npz[p + "o_logw"], o_msid=npz[p + "o_msid"], o_lse_pos=npz[p + "o_lse_pos"],
                   o_lse_neg=npz[p + "o_lse_neg"], o_raw_dense=npz[p + "o_raw_dense"], o_raw_sparse=npz[p + "o_raw_sparse"])


def assert_results_match_reference(case, indices, scores, labels, logw, msid, lse_pos, lse_neg, raw):
    """One PrioritySampledSections against the reference function's output for the same inputs and noise: picks,
    gathered values, labels and max_sampling_id exactly; log-weights / normalisers within the fastmath tolerance."""
    cid = case["cid"]
    assert not np.isnan(case["o_logw"]).any(), "golden case hit the numba fastmath NaN quirk; regenerate without it"
    assert np.array_equal(indices, case["o_indices"]), cid
    assert np.array_equal(scores.view(np.uint32), case["o_scores"].view(np.uint32)), cid
    assert np.array_equal(labels, case["o_labels"]), cid
    assert np.array_equal(msid, case["o_msid"]), cid
    assert np.array_equal(np.isfinite(logw), np.isfinite(case["o_logw"])), cid
    fin = np.isfinite(logw)
    assert np.abs(logw[fin] - case["o_logw"][fin]).max(initial=0.0) <= 4e-5, cid
    assert np.allclose(lse_pos, case["o_lse_pos"], atol=1e-5) and np.allclose(lse_neg, case["o_lse_neg"], atol=1e-5), cid
    for name in ("dense", "sparse"):
        assert np.array_equal(raw[name].view(np.uint32), case["o_raw_" + name].view(np.uint32)), (cid, name)
