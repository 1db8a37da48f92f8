"""Small pass over every kernel for compute-sanitizer (memcheck / racecheck): exact + tensor search (1-3 query terms,
fp32 store through bf16 planes, dump + filtered segments, safe fallback), retrieve->sample chain, merge, hybrid merge, sampling. Sizes are tiny: the sanitizer is ~100x slower."""
import sys

import numpy as np

sys.path.insert(0, ".")
import vod_b200
from vod_b200 import hybrid

small = len(sys.argv) > 1
rng = np.random.default_rng(0)
n = 6000 if small else 20000
xb = rng.integers(-3, 4, size=(n, 96)).astype(np.float32)
xq = rng.integers(-3, 4, size=(9, 96)).astype(np.float32)
ref = None
for dtype, modes in (("float32", ["exact", "tensor", "tensor2", "tensor3"]), ("bfloat16", ["exact", "tensor", "tensor2", "tensor3"])):
    st = vod_b200.CorpusStore(n, 96, dtype=dtype)
    st.add(xb)
    for mode in modes:
        s, i = st.search(xq, 20, mode=mode)
        if ref is None:
            ref = (s, i)
        assert np.array_equal(i, ref[1]) and np.array_equal(s, ref[0]), (dtype, mode)
    if dtype == "bfloat16":  # retrieve -> label -> sample -> gather chain in one call
        chain = vod_b200.DenseRetrievalSampler(st, top_k=20, total=4, max_pos_sections=2, mode="tensor")(xq, ref[1][:, :2].copy(), seed=1)
        assert chain.batch.indices.shape == (9, 4)
    st.close()
# every scoring-kernel shape of round 2: resident pair kernels (<= 64 / <= 128 queries), the wide one-term pair kernel
# with a narrow last query tile, multi-term pair kernels (<64,3>, <128,3> and the two-launch policy above 128
# queries), blocked item order (2 query tiles), 19-row last corpus tile, k > rows in the dump segment
st = vod_b200.CorpusStore(n + 19, 96, dtype="float16")
st.add(np.concatenate([xb, xb[:19]]))
for nq in ((33, 130) if small else (33, 100, 130, 300)):
    q = rng.integers(-3, 4, size=(nq, 96)).astype(np.float32)
    a = st.search(q, 10, mode="tensor")
    b3 = st.search(q + np.float32(1.0 / 4096), 10, mode="tensor3")   # non-empty correction terms
    c = st.search(q, 10, mode="tensor3")                             # empty correction terms: skipped on the device
    assert np.array_equal(a[1], c[1]) and np.array_equal(a[0], c[0]) and b3[1].shape == (nq, 10)
st.close()
# selection paths: continuous scores (the thread-maximum bound filters and the rank sort finishes the job), top-100 and
# top-300 (general radix path on a 256-thread CTA), many queries (small CTAs)
xc = rng.standard_normal((n, 96)).astype(np.float32)
st = vod_b200.CorpusStore(n, 96, dtype="bfloat16")
st.add(xc)
for nq, k in ((9, 100), (9, 300), (530, 100)):
    s_c, i_c = st.search(rng.standard_normal((nq, 96)).astype(np.float32), k, mode="tensor")
    assert (np.diff(s_c, axis=1) <= 0).all() and (i_c >= 0).all() and len(set(i_c[0].tolist())) == k
st.close()
ms, mi = vod_b200.merge_topk(np.stack([ref[0], ref[0]]), np.stack([ref[1], ref[1] + 100000]), 20)
b = vod_b200.RetrievalBatch(scores=ref[0], indices=ref[1])
m, raw = hybrid.merge_search_results({"a": b, "b": vod_b200.RetrievalBatch(scores=ref[0] * 2, indices=ref[1][:, ::-1].copy())},
                                     {"a": 1.0, "b": 0.5})
out = vod_b200.sample_search_results(search_results=m, raw_scores=raw, total=8, max_pos_sections=2, seed=3, max_support_size=10)
print("sanitizer probe ok", out.batch.indices.shape)
