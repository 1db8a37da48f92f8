"""tests/golden/make_golden_results.py — regenerates tests/golden/sample_results_ref.npz.

Runs the REFERENCE's own `sample_search_results` (src/vod_dataloaders/core/sample.py:22-84, loaded from
/root/reference by oracle/ref_shim.py) — the whole function: labeled priority sampling, the gathers of ids / scores /
raw scores at the picks and `max_sampling_id` — on seeded inputs, and stores inputs + outputs + the Exp(1) noise the
reference drew, so that `vod_b200.sample_search_results(noise=...)` can be checked against it on a box where
/root/reference does not exist. The reference draws its noise as the first `np.random.exponential(size=shape)` of
the call (sample.py:398): the script seeds `np.random`, replays that draw to record it, re-seeds and calls the
reference. Run from the repo root in the build container:

    python tests/golden/make_golden_results.py
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


def make_inputs(seed: int, B: int, K: int, label_mode: str):
    rng = np.random.default_rng([seed, B, K, len(label_mode)])
    scores = np.sort(rng.normal(size=(B, K)).astype(np.float32) * 3.0, axis=1)[:, ::-1].copy()
    scores -= scores.min(axis=1, keepdims=True)
    indices = rng.permutation(10_000_000)[: B * K].reshape(B, K).astype(np.int64)
    labels = None
    if label_mode != "none":
        labels = np.zeros((B, K), np.int64)
        for b in range(B):
            labels[b, rng.choice(K, size=b % 5, replace=False)] = 1 + (b % 2)  # any value > 0 is positive
    if label_mode == "pad":  # ragged rows: trailing padding as the merge leaves it (id -1, score -inf)
        for b in range(B):
            n_pad = (b * 7) % (K // 2)
            if n_pad:
                scores[b, -n_pad:] = -np.inf
                indices[b, -n_pad:] = -1
                labels[b, -n_pad:] = -1
    sparse = rng.normal(size=(B, K)).astype(np.float32)
    return scores, indices, labels, sparse


def main() -> None:
    warnings.filterwarnings("ignore")
    mods = ref_shim.load()
    ref_sample, retrieval = mods["sample"], mods["retrieval"]
    out: dict[str, np.ndarray] = {}
    meta = []
    cid = 0
    grid = list(itertools.product((0,), ((4, 64), (16, 256)), (8, 16), ("none", "pos", "pad"), (0.0, 1.0), (None, 100)))
    grid += [(1, (32, 1000), 8, "pos", 1.0, None), (1, (32, 1000), 8, "pad", 1.0, 100)]  # BASELINE configs[3] shape
    for seed, (B, K), total, label_mode, temperature, support in grid:
        scores, indices, labels, sparse = make_inputs(seed, B, K, label_mode)
        batch = retrieval.RetrievalBatch(scores=scores.copy(), indices=indices.copy(),
                                         labels=None if labels is None else labels.copy())
        np.random.seed(1000 + cid)
        noise = np.random.exponential(size=scores.shape).astype(scores.dtype)
        np.random.seed(1000 + cid)
        res = ref_sample.sample_search_results(search_results=batch, raw_scores={"dense": scores, "sparse": sparse},
                                               total=total, max_pos_sections=3, temperature=temperature,
                                               max_support_size=support)
        p = f"c{cid:03d}_"
        out[p + "scores"], out[p + "indices"], out[p + "sparse"], out[p + "noise"] = scores, indices, sparse, noise
        if labels is not None:
            out[p + "labels"] = labels
        out[p + "o_indices"], out[p + "o_scores"] = res.batch.indices, res.batch.scores
        out[p + "o_labels"], out[p + "o_logw"] = res.batch.labels, res.log_weights
        out[p + "o_msid"], out[p + "o_lse_pos"], out[p + "o_lse_neg"] = res.max_sampling_id, res.lse_pos, res.lse_neg
        out[p + "o_raw_dense"], out[p + "o_raw_sparse"] = res.raw_scores["dense"], res.raw_scores["sparse"]
        meta.append((cid, total, 3, temperature, support or -1, labels is not None))
        cid += 1
    out["meta"] = np.array(meta, np.float64)
    path = pathlib.Path(__file__).with_name("sample_results_ref.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size/1e6:.2f} MB, {cid} cases)")


if __name__ == "__main__":
    main()
