"""tests/golden/make_golden_merge.py — regenerates tests/golden/merge_ref.npz from the REFERENCE's own code
(`merge.merge_search_results`, `normalize._subtract_min_score`: src/vod_dataloaders/core/{merge,normalize}.py loaded
from /root/reference by oracle/ref_shim.py). Run in the build container: `python tests/golden/make_golden_merge.py`."""
from __future__ import annotations

import pathlib
import sys
import warnings

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref_shim  # noqa: E402


def make_inputs(seed, B, widths, n_values, dtype, with_pads, with_nan):
    rng = np.random.default_rng([seed, B, n_values, len(widths)])
    out = {}
    for e, K in enumerate(widths):
        idx = np.stack([rng.choice(n_values, size=K, replace=False) for _ in range(B)]).astype(np.int64)
        sc = rng.uniform(0.0, 10.0, size=(B, K)).astype(dtype)
        lab = (rng.uniform(size=(B, K)) < 0.3).astype(np.int64)
        if with_pads:
            pad = rng.uniform(size=(B, K)) < 0.15
            idx[pad] = -1
            sc[pad] = -np.inf
            lab[pad] = -1
        if with_nan:
            sc[rng.uniform(size=(B, K)) < 0.05] = np.nan
        out[f"e{e}"] = (sc, idx, lab)
    return out


def main():
    warnings.filterwarnings("ignore")
    mods = ref_shim.load()
    RB, merge, normalize = mods["retrieval"].RetrievalBatch, mods["merge"], mods["normalize"]
    blob, meta, cid = {}, [], 0
    for seed in range(4):
        for widths in ((7, 9), (30, 50), (40, 40, 25), (300, 300, 200)):
            for dtype in (np.float32, np.float64):
                for with_pads, with_nan in ((False, False), (True, False), (True, True)):
                    B, n_values = 3, max(60, 2 * max(widths))
                    inp = make_inputs(seed, B, widths, n_values, dtype, with_pads, with_nan)
                    rng = np.random.default_rng(seed + 100)
                    weights = {k: float(rng.uniform(0.0, 1.0)) for k in inp}
                    # labels only on the first engine (like the lookup engine in core/search.py)
                    batches = {k: RB(scores=v[0].copy(), indices=v[1].copy(), labels=v[2].copy() if k == "e0" else None)
                               for k, v in inp.items()}
                    merged, raw = merge.merge_search_results(batches, weights)
                    p = f"c{cid:03d}_"
                    for k, v in inp.items():
                        blob[p + k + "_s"], blob[p + k + "_i"], blob[p + k + "_l"] = v
                    blob[p + "out_s"], blob[p + "out_i"], blob[p + "out_l"] = merged.scores, merged.indices, merged.labels
                    for k, v in raw.items():
                        blob[p + "raw_" + k] = v
                    blob[p + "norm_e1"] = normalize._subtract_min_score(inp["e1"][0], offset=0.5)
                    meta.append((cid, len(widths), [weights[k] for k in inp] + [0.0] * (3 - len(widths))))
                    cid += 1
    blob["meta"] = np.array([[c, n, *w] for c, n, w in meta], np.float64)
    path = pathlib.Path(__file__).with_name("merge_ref.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path} ({path.stat().st_size / 1e6:.2f} MB, {cid} cases)")


if __name__ == "__main__":
    main()
