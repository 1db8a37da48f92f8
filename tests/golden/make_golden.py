"""tests/golden/make_golden.py — regenerates tests/golden/sampling_ref.npz.

Runs the REFERENCE's own numba sampler (loaded from /root/reference by oracle/ref_shim.py)
on seeded inputs with injected Exp(1) noise and stores inputs + outputs, so that the CPU twin
and the CUDA kernel can be checked against the reference on a box where /root/reference does
not exist. Run from the repo root in the build container:

    python tests/golden/make_golden.py

Reference entry point: `_labeled_priority_sampling_2d_` (src/vod_dataloaders/core/sample.py:323-352),
called exactly as `labeled_priority_sampling_2d` does (sample.py:388-418) but with the noise
passed in instead of drawn from the global np.random state (sample.py:398).
Environment this file was produced with: numba 0.65.0, numpy 2.3.5 (the reference pins numba ^0.57.1;
see SURVEY.md §4 for the fastmath caveat on rows containing -inf).
"""
from __future__ import annotations

import itertools
import pathlib
import sys
import warnings

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_shim  # noqa: E402


def make_case(seed: int, K: int, k_total: int, label_mode: str, temperature: float, max_support: int | None):
    rng = np.random.default_rng([seed, K, k_total, int(temperature), max_support or 0, len(label_mode)])
    B = 2
    # descending "retrieval-like" scores minus the row minimum (core/search.py:79-125 hands the sampler that)
    scores = np.sort(rng.normal(size=(B, K)).astype(np.float32) * 3.0, axis=1)[:, ::-1].copy()
    scores -= scores.min(axis=1, keepdims=True)
    labels = np.zeros((B, K), np.bool_)
    if label_mode == "pos3":
        for b in range(B):
            labels[b, rng.choice(K, size=3, replace=False)] = True
    elif label_mode == "inf20":
        for b in range(B):
            labels[b, rng.choice(K, size=5, replace=False)] = True
        scores[rng.uniform(size=(B, K)) < 0.2] = -np.inf
    noise = rng.exponential(size=(B, K)).astype(np.float32)
    return scores, labels, noise


def main() -> None:
    warnings.filterwarnings("ignore")
    ref = ref_shim.load()["sample"]
    out: dict[str, np.ndarray] = {}
    meta = []
    cid = 0
    for seed, K, k_total, label_mode, temperature, max_support in itertools.product(
        (0, 1, 2), (100, 1000), (8, 32), ("none", "pos3", "inf20"), (0.0, 1.0), (None, 50)
    ):
        scores, labels, noise = make_case(seed, K, k_total, label_mode, temperature, max_support)
        k_positive = 3
        B = scores.shape[0]
        samples = np.full((B, k_total), -1, np.int64)
        logw = np.full((B, k_total), -np.inf, np.float32)
        olab = np.zeros((B, k_total), np.bool_)
        lse = np.zeros((B, 2), np.float32)
        ms = max_support or -1
        if ms >= 0:
            ms = max(ms, k_total)  # sample.py:133-135
        ref._labeled_priority_sampling_2d_(scores.copy(), labels.copy(), noise.copy(), k_positive, k_total,
                                           samples, logw, olab, lse, True, temperature, ms)
        p = f"c{cid:03d}_"
        out[p + "scores"], out[p + "labels"], out[p + "noise"] = scores, labels, noise
        out[p + "samples"], out[p + "logw"], out[p + "olab"], out[p + "lse"] = samples, logw, olab, lse
        meta.append((cid, seed, K, k_total, k_positive, label_mode, temperature, ms))
        cid += 1
    out["meta"] = np.array([(c, s, K, kt, kp, ["none", "pos3", "inf20"].index(lm), t, ms)
                            for c, s, K, kt, kp, lm, t, ms in meta], np.float64)
    path = pathlib.Path(__file__).with_name("sampling_ref.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size/1e6:.2f} MB, {cid} cases)")


if __name__ == "__main__":
    main()
