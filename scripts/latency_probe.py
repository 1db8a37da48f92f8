"""Why is an isolated 64-query search slower than one of a back-to-back stream? Per-call CUDA-event time and the
library's per-kernel event sums (vodb_store_set_profiling) under different submission patterns. Run on the GPU box."""
import json
import time

import torch

import vod_b200

ROWS, DIM, NQ, K = 10_000_000, 768, 64, 100
dev = torch.device("cuda:0")
store = vod_b200.CorpusStore(ROWS, DIM, dtype="bfloat16")
store.fill_synthetic(1234)
g = torch.Generator(device=dev).manual_seed(1)
queries = torch.randn((64, NQ, DIM), device=dev, generator=g).to(torch.bfloat16)
out = (torch.empty((NQ, K), dtype=torch.float32, device=dev), torch.empty((NQ, K), dtype=torch.int64, device=dev))
for i in range(10):
    store.search_device(queries[i], K, mode="tensor", out=out)
torch.cuda.synchronize()


def pct(xs):
    xs = sorted(xs)
    return {"p10": xs[len(xs) // 10], "p50": xs[len(xs) // 2], "p90": xs[(len(xs) * 9) // 10]}


res = {}
# 1. back to back
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(40):
    store.search_device(queries[i % 64], K, mode="tensor", out=out)
e1.record()
torch.cuda.synchronize()
res["back_to_back_ms"] = e0.elapsed_time(e1) / 40


def isolated(n, gap_s=0.0, group=1, profile=False):
    store.set_profiling(profile)
    if profile:
        store.profile()
    tot, score, select = [], [], []
    for i in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for j in range(group):
            store.search_device(queries[(i * group + j) % 64], K, mode="tensor", out=out)
        b.record()
        torch.cuda.synchronize()
        tot.append(a.elapsed_time(b) / group)
        if profile:
            p = store.profile()
            score.append(p["score_ms"] / group)
            select.append(p["select_ms"] / group)
        if gap_s:
            time.sleep(gap_s)
    store.set_profiling(False)
    r = {"call_ms": pct(tot)}
    if profile:
        r["score_ms"], r["select_ms"] = pct(score), pct(select)
    return r


res["isolated"] = isolated(40)
res["isolated_profiled"] = isolated(40, profile=True)
res["isolated_gap_5ms"] = isolated(40, gap_s=0.005)
res["isolated_gap_50ms"] = isolated(20, gap_s=0.05)
res["pairs"] = isolated(20, group=2)
res["quads"] = isolated(10, group=4)
# wall-clock view of one isolated call (host side)
w = []
for i in range(30):
    t0 = time.perf_counter()
    store.search_device(queries[i], K, mode="tensor", out=out)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    w.append(((t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3))
res["host_enqueue_ms"] = pct([x[0] for x in w])
res["host_total_ms"] = pct([x[1] for x in w])
print(json.dumps(res))
