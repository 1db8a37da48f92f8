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
