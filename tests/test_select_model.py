"""The thread-maximum bound of `block_select` (vod_b200/csrc/select.cu), restated in numpy so that its claims can be
checked without a GPU: the bound never exceeds the k-th best key, so every member of the exact top-k survives it, and
ranking the survivors by (score desc, id asc) gives exactly what the oracle's k-selection gives. The GPU parity tests
(`tests/test_search_gpu.py::test_selection_paths_bit_exact`) check the kernel itself; this file checks the argument."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import flat_ip


def ord_u32(x: np.ndarray) -> np.ndarray:
    """Order-preserving uint32 image of float32 (common.cuh `ord_u32`): all bits of negatives flipped, the sign bit of
    the rest set; NaN -> 0 (sorts last), -0.0 == +0.0."""
    b = np.asarray(x, np.float32).view(np.uint32).copy()
    mag = b & 0x7FFFFFFF
    b[mag == 0] = 0
    out = np.where(b & 0x80000000, ~b, b | 0x80000000).astype(np.uint32)
    out[mag > 0x7F800000] = 0
    return out


def thread_of_entry(n: int, threads: int, vectorised: bool) -> np.ndarray:
    """Which thread reads entry i in the first sweep: 16-byte loads (float4 v -> thread v % threads) for the first
    n - n % 4 entries of an aligned list, one entry per thread and round after that (and for unaligned lists)."""
    owner = np.arange(n) % threads
    if vectorised:
        n4 = n // 4
        owner[: n4 * 4] = np.repeat(np.arange(n4) % threads, 4)
        owner[n4 * 4:] = np.arange(n - n4 * 4) % threads   # scalar tail: entry n4*4 + tid
    return owner


def kth_largest_top16(values: np.ndarray, k: int) -> int:
    """Two 8-bit radix passes over one value per thread: the top 16 bits of the k-th largest, low 16 bits zero."""
    top = values >> 24
    hist = np.bincount(top, minlength=256)
    cum, bin_a = 0, 0
    for b in range(255, -1, -1):
        if cum + hist[b] >= k:
            bin_a = b
            break
        cum += hist[b]
    need = k - cum
    second = (values[top == bin_a] >> 16) & 255
    hist = np.bincount(second, minlength=256)
    cum, bin_b = 0, 0
    for b in range(255, -1, -1):
        if cum + hist[b] >= need:
            bin_b = b
            break
        cum += hist[b]
    return (bin_a << 24) | (bin_b << 16)


def select_model(scores: np.ndarray, ids: np.ndarray, k: int, threads: int, vectorised: bool = True):
    """Returns (top-k scores, top-k ids, number of survivors) the way the fast path computes them, or None when the
    kernel would fall through to the radix select (more survivors than threads)."""
    n = len(scores)
    assert n > k and 2 * k <= threads
    keys = ord_u32(scores)
    owner = thread_of_entry(n, threads, vectorised)
    tmax = np.zeros(threads, np.uint32)
    np.maximum.at(tmax, owner, keys)
    bound = kth_largest_top16(tmax, k)
    surv = np.nonzero(keys >= bound)[0]
    assert len(surv) >= k, "the k largest thread maxima are k different entries"
    if len(surv) > threads:
        return None
    order = np.lexsort((ids[surv], -keys[surv].astype(np.int64)))   # (score desc, id asc)
    pick = surv[order[:k]]
    return scores[pick], ids[pick], len(surv)


def oracle_topk(scores: np.ndarray, ids: np.ndarray, k: int):
    s, pos = flat_ip.topk_desc_stable(scores[None, :], 0, k)   # ties: smaller position first
    return s[0], pos[0]


@pytest.mark.parametrize("n,k,threads", [(16384, 100, 1024), (7600, 100, 1024), (700, 100, 256), (4096, 100, 256),
                                         (101, 100, 256), (16384, 500, 1024), (1003, 7, 64), (1200, 128, 256)])
def test_bound_keeps_the_exact_top_k(n, k, threads):
    rng = np.random.default_rng(n + k)
    scores = rng.standard_normal(n).astype(np.float32) * 7
    ids = rng.permutation(n).astype(np.int64)              # list order is not id order (filtered lists)
    got = select_model(scores, ids, k, threads)
    assert got is not None
    order = np.lexsort((ids, -ord_u32(scores).astype(np.int64)))[:k]
    assert np.array_equal(got[0], scores[order]) and np.array_equal(got[1], ids[order])
    if n >= 4 * threads:                                   # every thread has entries: the bound is tight
        assert got[2] <= 1.3 * (-threads * np.log(1 - k / threads)) + 32, got[2]


def test_dump_list_matches_the_oracle_selection():
    """A dump list is in row order (position == id), which is the oracle's tie order."""
    rng = np.random.default_rng(5)
    scores = rng.integers(-40, 41, size=16384).astype(np.float32)       # many ties
    ids = np.arange(16384, dtype=np.int64)
    got = select_model(scores, ids, 100, 1024)
    ref_s, ref_i = oracle_topk(scores, ids, 100)
    if got is not None:                                                  # else: the kernel takes the radix path
        assert np.array_equal(got[0], ref_s) and np.array_equal(got[1], ref_i)


def test_all_equal_scores_fall_through_to_the_general_path():
    scores = np.zeros(5000, np.float32)
    assert select_model(scores, np.arange(5000, dtype=np.int64), 100, 1024) is None


def test_short_lists_read_with_16_byte_loads_leave_threads_empty():
    """300 entries on 256 threads: 75 threads hold four entries each, fewer than k = 128 have any — the k-th largest
    thread maximum is the empty threads' 0, everything survives, and the kernel falls through to the radix select;
    read one entry per thread (unaligned list) the bound filters."""
    rng = np.random.default_rng(3)
    scores = rng.standard_normal(300).astype(np.float32)
    ids = np.arange(300, dtype=np.int64)
    assert select_model(scores, ids, 128, 256, vectorised=True) is None
    got = select_model(scores, ids, 128, 256, vectorised=False)
    assert got is not None and np.array_equal(got[1], oracle_topk(scores, ids, 128)[1])


def test_survivor_count_does_not_grow_with_the_list():
    rng = np.random.default_rng(9)
    counts = []
    for n in (4096, 16384, 32768):
        s = rng.standard_normal(n).astype(np.float32)
        counts.append(select_model(s, np.arange(n, dtype=np.int64), 100, 1024)[2])
    assert max(counts) < 160 and min(counts) >= 100, counts               # ~ -1024 ln(1 - 100/1024) = 105, +5% for the 16-bit bound


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(1, 128), st.sampled_from([256, 512, 1024]), st.booleans(),
       st.sampled_from(["normal", "few_values", "sorted", "clustered", "negative"]))
def test_bound_is_a_lower_bound_for_any_list(seed, k, threads, vectorised, shape):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(k + 1, 6000))
    if shape == "normal":
        scores = rng.standard_normal(n)
    elif shape == "few_values":
        scores = rng.integers(-2, 3, size=n).astype(np.float64)
    elif shape == "sorted":
        scores = np.sort(rng.standard_normal(n))[::-1]
    elif shape == "clustered":                              # the best entries all land in a few threads
        scores = rng.standard_normal(n)
        scores[:: threads] += 50
    else:
        scores = -np.abs(rng.standard_normal(n)) * 1e-3
    scores = scores.astype(np.float32)
    ids = rng.permutation(n).astype(np.int64)
    keys = ord_u32(scores)
    owner = thread_of_entry(n, threads, vectorised)
    tmax = np.zeros(threads, np.uint32)
    np.maximum.at(tmax, owner, keys)
    bound = kth_largest_top16(tmax, k)
    kth_key = np.sort(keys)[::-1][k - 1]
    assert bound <= kth_key
    got = select_model(scores, ids, k, threads, vectorised)
    if got is not None:
        order = np.lexsort((ids, -keys.astype(np.int64)))[:k]
        assert np.array_equal(got[1], ids[order]) and np.array_equal(got[0], scores[order])
