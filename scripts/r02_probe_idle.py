"""Development probe: duration of one 64-query search (CUDA events around it) against the idle time before it.
Back-to-back searches take 2.2 ms each, isolated ones 2.4-2.5 ms: is it the gap?"""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, vod_b200
import bench
st = vod_b200.CorpusStore(10_000_000, 768, dtype="bfloat16"); st.fill_synthetic(1234)
q = torch.from_numpy(bench.make_queries(np, 40, 64, "bfloat16")).cuda()
for i in range(5): st.search_device(q[i], 100, mode="tensor")
torch.cuda.synchronize()
out = {}
for gap_us in (0, 20, 100, 500, 2000, 20000):
    ts = []
    for i in range(5, 40):
        torch.cuda.synchronize()
        if gap_us:
            t0 = time.perf_counter()
            while (time.perf_counter() - t0) * 1e6 < gap_us:
                pass
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); st.search_device(q[i], 100, mode="tensor"); a1.record()
        torch.cuda.synchronize()
        ts.append(a0.elapsed_time(a1))
    ts.sort()
    out[f"idle_{gap_us}us"] = {"p10": ts[3], "p50": ts[len(ts) // 2], "p90": ts[-4]}
# two searches per measurement: the second one starts on a busy GPU
ts = []
for i in range(5, 39):
    torch.cuda.synchronize()
    a0, a1, a2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a0.record(); st.search_device(q[i], 100, mode="tensor"); a1.record(); st.search_device(q[i + 1], 100, mode="tensor"); a2.record()
    torch.cuda.synchronize()
    ts.append((a0.elapsed_time(a1), a1.elapsed_time(a2)))
out["pair_first_p50"] = sorted(t[0] for t in ts)[len(ts) // 2]
out["pair_second_p50"] = sorted(t[1] for t in ts)[len(ts) // 2]
print(json.dumps(out))
