"""One pass over the kernels beside the tensor-core scan, for an `ncu --set full` capture: fp32-exact scan (CUDA
cores), the selects of both paths, cross-shard merge, hybrid merge, and the retrieve -> sample chain (label match,
sampler, gather). Sizes: 1M x 768 rows (exact scan ~ 3 GB of fp32), C4-shaped lists (32 x 1000)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import vod_b200
from vod_b200 import hybrid

rng = np.random.default_rng(0)
n, d = 1_000_000, 768
xq = rng.standard_normal((64, d), dtype=np.float32)

st = vod_b200.CorpusStore(n, d, dtype="float32")
st.fill_synthetic(1234)
s_exact, i_exact = st.search(xq, 100, mode="exact")            # score_exact_kernel + select_kernel
s_pl, i_pl = st.search(xq, 100)                                # auto on a float32 store: split_planes + score_tc_kernel<64,3,3>
assert float(np.abs(s_exact - s_pl).max()) < 1e-3
st.close()

st = vod_b200.CorpusStore(n, d, dtype="bfloat16")
st.fill_synthetic(1234)
s_t3, i_t3 = st.search(xq, 100, mode="tensor3")                # 3-term tensor scan of the same values
s1k, i1k = st.search(xq[:32], 1000, mode="tensor3")
gold = i1k[:, :2].copy()
pipe = vod_b200.DenseRetrievalSampler(st, top_k=1000, total=8, max_pos_sections=3, mode="tensor3")
picks = pipe(xq[:32], gold, seed=42, offset=0)                  # match_labels + sample_kernel + gather_picks
st.close()

parts = np.stack([s_exact] * 8), np.stack([i_exact + 1_000_000 * g for g in range(8)])
vod_b200.merge_topk(parts[0], parts[1], 100)                    # merge_kernel (8 shards)

lookup = vod_b200.RetrievalBatch(scores=np.zeros((32, 4), np.float32), indices=i1k[:, :4].copy(),
                                 labels=np.ones((32, 4), np.int64))
dense = vod_b200.RetrievalBatch(scores=s1k, indices=i1k)
sparse = vod_b200.RetrievalBatch(scores=(s1k[:, ::-1] * 0.5).copy(), indices=(i1k[:, ::-1] + (np.arange(1000) % 2) * 7).copy())
merged, raw = hybrid._merge_search_results({"lookup": lookup, "dense": dense, "sparse": sparse}, {"dense": 1.0, "sparse": 0.7})
print("aux probe ok", merged.scores.shape, picks.batch.indices.shape, float(np.abs(s_exact - s_t3).max()))
