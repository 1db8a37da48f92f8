"""Development probe (run under gpurun): host-path vs device-path time of one 64-query search over 10M x 768 bf16 for
the query-precision modes, with queries that are / are not exact in the store dtype."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, vod_b200
import bench

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
st = vod_b200.CorpusStore(rows, 768, dtype="bfloat16"); st.fill_synthetic(1234)
out = []
for label, sd in (("bf16_exact", "bfloat16"), ("full_f32", None)):
    q = torch.from_numpy(bench.make_queries(np, 30, 64, sd)).pin_memory()
    qd = q.cuda()
    for mode in ("tensor", "tensor3"):
        for i in range(5): st.search(q[i].numpy(), 100, mode=mode)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(5, 30): st.search(q[i].numpy(), 100, mode=mode)
        host_ms = (time.perf_counter() - t0) / 25 * 1e3
        for i in range(5): st.search_device(qd[i], 100, mode=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(5, 30): st.search_device(qd[i], 100, mode=mode)
        e1.record(); torch.cuda.synchronize()
        lat = []
        for i in range(5, 30):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); st.search_device(qd[i], 100, mode=mode); a1.record(); torch.cuda.synchronize()
            lat.append(a0.elapsed_time(a1))
        st.set_profiling(True)
        for i in range(5, 30): st.search_device(qd[i], 100, mode=mode)
        p = st.profile(); st.set_profiling(False)
        out.append({"queries": label, "mode": mode, "host_ms": host_ms, "device_back_to_back_ms": e0.elapsed_time(e1) / 25,
                    "device_single_ms_p50": sorted(lat)[12], "score_ms": p["score_ms"] / 25, "select_ms": p["select_ms"] / 25})
        print(json.dumps(out[-1]), flush=True)
