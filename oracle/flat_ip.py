"""oracle/flat_ip.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  *** parity unpinned ***

CPU restatement of the search the reference delegates to faiss:

    faiss_index.search(query_vec, k)            src/vod_search/faiss_search/server.py:72,84
    index = faiss.index_factory(D, "Flat", METRIC_INNER_PRODUCT); index.add(f32 rows)
                                                src/vod_search/faiss_search/build.py:60,67-73
    RetrievalBatch.cast(indices, scores)        src/vod_search/faiss_search/client.py:97-102

The arithmetic lives in a third-party dependency that is NOT vendored under /root/reference:
faiss, pinned `faiss-cpu==1.7.4` (requirements.txt:42, poetry.lock:708-709). faiss is not
installable here (no wheel, no network) and the reference has no test, fixture or golden
vector that touches the search path, so this oracle is pinned only by the
published definition of IndexFlatIP, restated here:

  * `IndexFlat::search` -> `knn_inner_product`: for nq >= 20 the database is processed in
    blocks (4096 queries x 1024 rows), each block's inner products come from one BLAS
    `sgemm_` in float32, and the k best per query are collected (heap for k < 100,
    reservoir otherwise); for nq < 20 plain float32 dot products are used. Results are
    returned sorted by descending score; unfilled slots hold id -1 / score -FLT_MAX.
  * the order of float32 accumulation inside sgemm is unspecified, and so is the order of
    exactly-tied scores. This restatement pins ties to (score desc, id asc).

Because float32 accumulation order differs between BLAS builds, comparisons against this
oracle use the north-star tolerance (scores within 1e-5 relative; index mismatches allowed
only between entries whose float64 scores differ by <= 1e-5 relative) — see
`compare_topk` below, which re-scores in float64 to classify mismatches.
"""
from __future__ import annotations

import numpy as np

FLT_MAX = float(np.finfo(np.float32).max)
DB_BLOCK = 1024  # faiss distance_compute_blas_database_bs
Q_BLOCK = 4096   # faiss distance_compute_blas_query_bs


def topk_desc_stable(scores: np.ndarray, base: int, k: int) -> tuple[np.ndarray, np.ndarray]:
    """k best of each row of `scores` [Q, n] in (score desc, id asc) order; ids offset by `base`."""
    nq, n = scores.shape
    kk = min(k, n)
    if kk < n:
        # k-th largest value per row, then a (score desc, column asc) sort of the kk selected columns. A row whose
        # k-th value is tied across the selection boundary must take the tied entries with the smallest columns:
        # those rows (rare) are redone from every entry >= the k-th value.
        cand = np.argpartition(scores, n - kk, axis=1)[:, n - kk:]
        vals = np.take_along_axis(scores, cand, axis=1)
        order = np.lexsort((cand, -vals), axis=1)
        cols = np.take_along_axis(cand, order, axis=1)
        kth = vals.min(axis=1)
        tied = np.nonzero((scores >= kth[:, None]).sum(axis=1) > kk)[0]
        for q in tied:
            c = np.nonzero(scores[q] >= kth[q])[0]
            cols[q] = c[np.argsort(-scores[q, c], kind="stable")[:kk]]
    else:
        cols = np.argsort(-scores, axis=1, kind="stable")[:, :kk]
    return np.take_along_axis(scores, cols, axis=1), cols + base


def merge_sorted(s_a, i_a, s_b, i_b, k):
    """Merge two per-query candidate sets, keep k best in (score desc, id asc) order."""
    s = np.concatenate([s_a, s_b], axis=1)
    i = np.concatenate([i_a, i_b], axis=1)
    # lexsort: last key is primary. Empty slots (id -1) must sort last among equal scores.
    idkey = np.where(i < 0, np.iinfo(np.int64).max, i)
    order = np.lexsort((idkey, -s.astype(np.float64)), axis=1)[:, :k]
    return np.take_along_axis(s, order, axis=1), np.take_along_axis(i, order, axis=1)


def _survivors(ip: np.ndarray, tau: np.ndarray, base: int) -> tuple[np.ndarray, np.ndarray]:
    """Entries of `ip` [Q, n] with score >= tau[q], as padded per-query lists (score -FLT_MAX / id -1 padding).
    This is the collection step of faiss' heap / reservoir result handlers: a score only enters a query's result
    set if it beats the current k-th best (`>=` keeps exact ties in play so that (score desc, id asc) stays exact)."""
    mask = ip >= tau[:, None]
    counts = mask.sum(axis=1)
    width = int(counts.max(initial=0))
    out_s = np.full((ip.shape[0], width), -FLT_MAX, np.float32)
    out_i = np.full((ip.shape[0], width), -1, np.int64)
    if width:
        rows, cols = np.nonzero(mask)  # row-major: grouped by query, columns ascending
        slot = np.arange(len(rows)) - np.repeat(np.cumsum(counts) - counts, counts)
        out_s[rows, slot] = ip[rows, cols]
        out_i[rows, slot] = cols + base
    return out_s, out_i


_POOL = None


def _collect(ip, cur_s, cur_i, base, k, threads):
    """Collection step for one block of scores, query rows spread over `threads` host threads (numpy releases the
    GIL in the comparisons / sorts; faiss runs the same step under `#pragma omp parallel for` over the queries)."""
    def one(lo, hi):
        if (cur_i[lo:hi, -1] < 0).any():   # result sets not full yet: plain top-k of the block
            s, i = topk_desc_stable(ip[lo:hi], base, k)
        else:                              # full: only scores that reach the current k-th best can enter
            s, i = _survivors(ip[lo:hi], cur_s[lo:hi, -1], base)
        return merge_sorted(cur_s[lo:hi], cur_i[lo:hi], s, i, k) if s.shape[1] else (cur_s[lo:hi], cur_i[lo:hi])

    nq = ip.shape[0]
    if threads <= 1 or nq < 2 * threads:
        return one(0, nq)
    global _POOL
    if _POOL is None or _POOL._max_workers < threads:
        import concurrent.futures

        _POOL = concurrent.futures.ThreadPoolExecutor(threads, thread_name_prefix="flat-ip")
    step = -(-nq // threads)
    parts = list(_POOL.map(lambda lo: one(lo, min(nq, lo + step)), range(0, nq, step)))
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])


def search(xb: np.ndarray, xq: np.ndarray, k: int, *, row_offset: int = 0, state=None, threads: int = 1):
    """IndexFlatIP.search restated: returns (scores f32 [Q,k], ids i64 [Q,k]).

    `state` = (scores, ids) of an earlier call over other rows of the same index: the scan continues from it (used
    by the benchmark to scan a corpus block by block without holding it in memory); ids in it are global.
    `threads`: host threads for the result collection (the sgemm uses the BLAS library's own thread pool).
    """
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    xq = np.ascontiguousarray(xq, dtype=np.float32)
    if xq.ndim != 2:
        raise ValueError(f"Expected 2D array, got {xq.ndim}D array")  # server.py:82-83
    if xb.ndim != 2 or xb.shape[1] != xq.shape[1]:
        raise ValueError("dimension mismatch")
    nq, n = xq.shape[0], xb.shape[0]
    out_s = np.full((nq, k), -FLT_MAX, np.float32) if state is None else np.array(state[0], np.float32)
    out_i = np.full((nq, k), -1, np.int64) if state is None else np.array(state[1], np.int64)
    for q0 in range(0, nq, Q_BLOCK):
        q1 = min(nq, q0 + Q_BLOCK)
        cur_s, cur_i = out_s[q0:q1], out_i[q0:q1]
        for b0 in range(0, n, DB_BLOCK * 64):  # 64 faiss blocks per sgemm call: same math, fewer python trips
            b1 = min(n, b0 + DB_BLOCK * 64)
            ip = xq[q0:q1] @ xb[b0:b1].T  # float32 sgemm
            cur_s, cur_i = _collect(ip, cur_s, cur_i, b0 + row_offset, k, threads)
        out_s[q0:q1], out_i[q0:q1] = cur_s, cur_i
    return out_s, out_i


def search_f64(xb: np.ndarray, xq: np.ndarray, k: int) -> tuple[np.ndarray, np.ndarray]:
    """Same search with float64 accumulation (tie classifier / ground truth)."""
    ip = xq.astype(np.float64) @ xb.astype(np.float64).T
    s, i = topk_desc_stable(ip, 0, k)
    i = i.astype(np.int64)
    if s.shape[1] < k:
        pad = k - s.shape[1]
        s = np.concatenate([s, np.full((s.shape[0], pad), -FLT_MAX)], axis=1)
        i = np.concatenate([i, np.full((i.shape[0], pad), -1, np.int64)], axis=1)
    return s, i


def compare_topk(xb, xq, got_s, got_i, ref_s, ref_i, *, rtol=1e-5):
    """North-star comparison. Returns a dict of diagnostics; `ok` is the verdict.

    - every returned score must match the float64 re-score of the returned id within rtol (relative);
    - ids must equal the oracle's except where the float64 scores of the two ids differ by <= rtol relative
      (near-ties whose order float32 accumulation may legitimately flip);
    - as a set, every returned id must have a float64 score >= (1 - rtol) * the oracle's k-th score.
    """
    xb64 = np.asarray(xb, np.float64)
    xq64 = np.asarray(xq, np.float64)
    got_i = np.asarray(got_i)
    ref_i = np.asarray(ref_i)
    valid = got_i >= 0
    same_valid = bool(np.array_equal(valid, ref_i >= 0))
    gi = np.where(valid, got_i, 0)
    ri = np.where(ref_i >= 0, ref_i, 0)
    true_got = np.einsum("qd,qkd->qk", xq64, xb64[gi])
    true_ref = np.einsum("qd,qkd->qk", xq64, xb64[ri])
    scale = np.maximum(np.abs(true_ref), 1e-30)
    score_err = np.where(valid, np.abs(np.asarray(got_s, np.float64) - true_got) / np.maximum(np.abs(true_got), 1e-30), 0)
    mism = (got_i != ref_i) & valid
    tie_gap = np.where(mism, np.abs(true_got - true_ref) / scale, 0.0)
    kth = np.where(ref_i >= 0, true_ref, np.inf).min(axis=1, keepdims=True)
    below = valid & (true_got < kth - rtol * np.abs(kth))
    out = {
        "same_valid_mask": same_valid,
        "max_score_rel_err": float(score_err.max(initial=0.0)),
        "n_index_mismatch": int(mism.sum()),
        "max_tie_gap": float(tie_gap.max(initial=0.0)),
        "n_below_kth": int(below.sum()),
        "n": int(valid.sum()),
    }
    out["ok"] = bool(same_valid and out["max_score_rel_err"] <= rtol and out["max_tie_gap"] <= rtol
                     and out["n_below_kth"] == 0)
    return out


def recall_at_k(got_i: np.ndarray, ref_i: np.ndarray) -> float:
    hits = 0
    total = 0
    for g, r in zip(np.asarray(got_i), np.asarray(ref_i)):
        r = r[r >= 0]
        hits += len(np.intersect1d(g[g >= 0], r))
        total += len(r)
    return hits / max(total, 1)
