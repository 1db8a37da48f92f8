"""Development probe: k = 1000 on one 12.5M-row shard (the 8-GPU shard of BASELINE configs[2]) + merge kernel timing."""
import json, sys, time
sys.path.insert(0, ".")
import torch, vod_b200
from vod_b200.search import merge_topk_device
rows = 12_500_000
st = vod_b200.CorpusStore(rows, 768, dtype="float16"); st.fill_synthetic(1234)
g = torch.Generator().manual_seed(1)
for nq, k in ((64, 100), (64, 1000), (8192, 1000)):
    n = 12 if nq <= 512 else 4
    qs = torch.randn((n, nq, 768), generator=g).to(torch.float16).to(torch.float32).cuda()
    for i in range(2): st.search_device(qs[i], k, mode="tensor")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(2, n): st.search_device(qs[i], k, mode="tensor")
    e1.record(); torch.cuda.synchronize()
    st.set_profiling(True)
    for i in range(2, n): st.search_device(qs[i], k, mode="tensor")
    p = st.profile(); st.set_profiling(False)
    print(json.dumps({"nq": nq, "k": k, "ms": e0.elapsed_time(e1) / (n - 2), "score_ms": p["score_ms"] / (n - 2), "select_ms": p["select_ms"] / (n - 2),
                      "segments": st.stats()["segments"], "cap": st.stats()["cap"], "overflow": st.check_async()}), flush=True)
for world, nq, k in ((8, 64, 100), (8, 64, 1000), (8, 8192, 1000), (2, 64, 1000)):
    s = torch.randn((world, nq, k), device="cuda").sort(dim=-1, descending=True).values.contiguous()
    i = torch.randint(0, 10**8, (world, nq, k), device="cuda")
    for _ in range(3): merge_topk_device(s, i, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): merge_topk_device(s, i, k)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"merge_world": world, "nq": nq, "k": k, "us": e0.elapsed_time(e1) / 10 * 1e3}), flush=True)
