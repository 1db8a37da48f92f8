"""Development probe 2: per-search durations inside one unsynchronised loop (events between searches), after a
stream-level wait instead of a device sync, and with the clocks sampled around isolated searches."""
import json, subprocess, sys, threading, time
sys.path.insert(0, ".")
import numpy as np, torch, vod_b200
import bench
st = vod_b200.CorpusStore(10_000_000, 768, dtype="bfloat16"); st.fill_synthetic(1234)
q = torch.from_numpy(bench.make_queries(np, 40, 64, "bfloat16")).cuda()
for i in range(5): st.search_device(q[i], 100, mode="tensor")
torch.cuda.synchronize()
out = {}
# (a) 30 searches enqueued without any host wait, an event between each
ev = [torch.cuda.Event(enable_timing=True) for _ in range(31)]
ev[0].record()
for i in range(30):
    st.search_device(q[5 + i], 100, mode="tensor"); ev[i + 1].record()
torch.cuda.synchronize()
d = [ev[i].elapsed_time(ev[i + 1]) for i in range(30)]
out["unsynchronised_loop_per_search_ms"] = [round(x, 3) for x in d]
# (b) isolated searches, host waits on the end event (event.synchronize) instead of the whole device
ts = []
for i in range(5, 40):
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(); st.search_device(q[i], 100, mode="tensor"); a1.record()
    a1.synchronize()
    ts.append(a0.elapsed_time(a1))
ts.sort(); out["isolated_event_sync_ms"] = {"p10": ts[3], "p50": ts[len(ts) // 2], "p90": ts[-4]}
# (c) isolated searches with a 3 ms sleep (not spin) before each
ts = []
for i in range(5, 40):
    torch.cuda.synchronize(); time.sleep(0.003)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(); st.search_device(q[i], 100, mode="tensor"); a1.record(); torch.cuda.synchronize()
    ts.append(a0.elapsed_time(a1))
ts.sort(); out["isolated_after_3ms_sleep_ms"] = {"p10": ts[3], "p50": ts[len(ts) // 2], "p90": ts[-4]}
# (d) host path (vodb_search with host buffers) back to back, wall clock per call
qh = torch.from_numpy(bench.make_queries(np, 40, 64, "bfloat16")).pin_memory()
ts = []
for i in range(5, 40):
    t0 = time.perf_counter(); st.search(qh[i].numpy(), 100, mode="tensor"); ts.append((time.perf_counter() - t0) * 1e3)
ts.sort(); out["host_call_ms"] = {"p10": ts[3], "p50": ts[len(ts) // 2], "p90": ts[-4]}
ts = []
for i in range(5, 40):
    time.sleep(0.003)
    t0 = time.perf_counter(); st.search(qh[i].numpy(), 100, mode="tensor"); ts.append((time.perf_counter() - t0) * 1e3)
ts.sort(); out["host_call_after_3ms_sleep_ms"] = {"p10": ts[3], "p50": ts[len(ts) // 2], "p90": ts[-4]}
print(json.dumps(out))
