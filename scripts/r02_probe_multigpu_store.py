"""Development probe (multi-GPU box): the single-process drop-in `B200SearchMaster(devices=[...])` (MultiGpuStore) —
the reference's deployment shape, one server process owning every GPU (server.py:51-54) — on BASELINE configs[1]."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, vod_b200
import bench

n_dev = torch.cuda.device_count()
rows = 10_000_000
store = vod_b200.MultiGpuStore(rows, 768, dtype="bfloat16", devices=list(range(n_dev)))
store.fill_synthetic(1234)
for d in range(n_dev):
    torch.cuda.synchronize(d)
out = {"devices": n_dev, "rows": rows}
with vod_b200.B200SearchMaster(store=store, serve=False) as master:
    client = master.get_client()
    for label, sd in (("bf16_exact_queries", "bfloat16"), ("full_f32_queries", None)):
        q = torch.from_numpy(bench.make_queries(np, 30, 64, sd)).pin_memory()
        for i in range(5):
            client.search(vector=q[i].numpy(), top_k=100)
        ts = []
        for i in range(5, 30):
            t0 = time.perf_counter()
            res = client.search(vector=q[i].numpy(), top_k=100)
            ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        out[label] = {"ms_p50": ts[len(ts) // 2], "ms_p10": ts[2], "ms_p90": ts[-3], "queries_per_s": 64 / (sum(ts) / len(ts) * 1e-3)}
    # same answer as one store holding everything (device 0)
    single = vod_b200.CorpusStore(rows, 768, dtype="bfloat16", device=0)
    single.fill_synthetic(1234)
    s1, i1 = single.search(q[7].numpy(), 100)
    r = client.search(vector=q[7].numpy(), top_k=100)
    out["equals_single_store"] = bool(np.array_equal(r.indices, i1) and np.array_equal(r.scores, s1))
print(json.dumps(out))
