"""BASELINE configs 3 and 5 on one GPU's share of the work (run under gpurun): timings for profiles/."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import vod_b200

out = {}
def timed(fn, n):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

# ---- config 3: 100M x 768 fp16 over 8 GPUs -> 12.5M rows per GPU, top-1000 ----
n3 = 12_500_000
st = vod_b200.CorpusStore(n3, 768, dtype="float16"); st.fill_synthetic(1234)
g = torch.Generator().manual_seed(3)
for nq, reps in ((64, 10), (8192, 2)):
    q = torch.randn((nq, 768), generator=g).to(torch.float16).cuda()
    ms = timed(lambda: st.search_device(q, 1000, mode="tensor"), reps)
    assert not st.check_async()
    stats = st.stats()
    out[f"c3_q{nq}_k1000_ms"] = ms
    out[f"c3_q{nq}_segments"] = stats["segments"]; out[f"c3_q{nq}_cap"] = stats["cap"]
    if nq == 64: out["c3_q64_GBps"] = n3 * 768 * 2 / ms / 1e6
    else: out["c3_q8192_TFLOPs"] = 2.0 * nq * n3 * 768 / ms / 1e9
st.close()

# ---- config 5: 50M x 1024 bf16 over 8 GPUs -> 6.25M rows per GPU: host fp32 ingest, then fp32-exact top-100 ----
n5, d5 = 6_250_000, 1024
st = vod_b200.CorpusStore(n5, d5, dtype="bfloat16")
chunk = torch.randn((262_144, d5), dtype=torch.float32).pin_memory()   # add_batch_size = 2**18 (build_gpu.py:294)
torch.cuda.synchronize()
t0 = time.perf_counter()
row = 0
while row < n5:
    m = min(len(chunk), n5 - row)
    st.add(chunk[:m], row0=row)
    row += m
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out["c5_ingest_s"] = dt
out["c5_ingest_host_GBps"] = n5 * d5 * 4 / dt / 1e9
dev_chunk = chunk.cuda()
t0 = time.perf_counter()
row = 0
while row < n5:
    m = min(len(dev_chunk), n5 - row)
    st.add(dev_chunk[:m], row0=row)
    row += m
torch.cuda.synchronize()
out["c5_ingest_from_device_s"] = time.perf_counter() - t0
q = torch.randn((64, d5), generator=g).cuda()
for mode in ("tensor3", "exact", "tensor"):
    out[f"c5_search_{mode}_ms"] = timed(lambda: st.search_device(q, 100, mode=mode), 5 if mode != "exact" else 2)
a, b = st.search_device(q, 100, mode="tensor3"), st.search_device(q, 100, mode="exact")
torch.cuda.synchronize()
out["c5_tensor3_vs_exact_recall"] = float(np.mean([len(np.intersect1d(x, y)) / 100 for x, y in zip(a[1].cpu().numpy(), b[1].cpu().numpy())]))
out["c5_tensor3_vs_exact_max_rel_score_diff"] = float(((a[0] - b[0]).abs() / b[0].abs()).max())
print(json.dumps(out))
