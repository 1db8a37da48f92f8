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
ms, mi = vod_b200.merge_topk(np.stack([ref[0], ref[0]]), np.stack([ref[1], ref[1] + 100000]), 20)
b = vod_b200.RetrievalBatch(scores=ref[0], indices=ref[1])
m, raw = hybrid.merge_search_results({"a": b, "b": vod_b200.RetrievalBatch(scores=ref[0] * 2, indices=ref[1][:, ::-1].copy())},
                                     {"a": 1.0, "b": 0.5})
out = vod_b200.sample_search_results(search_results=m, raw_scores=raw, total=8, max_pos_sections=2, seed=3, max_support_size=10)
print("sanitizer probe ok", out.batch.indices.shape)
