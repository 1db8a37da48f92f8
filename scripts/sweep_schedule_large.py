"""Development probe: schedule of large batches (8192 queries) — first-segment rows x growth, for k=100 on the 10M-row
bf16 shard and k=1000 on a 12.5M-row fp16 shard (BASELINE configs[2], one GPU's share). Run under gpurun."""
import json, os, subprocess, sys
sys.path.insert(0, ".")
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, vod_b200
    rows, k, dtype = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    st = vod_b200.CorpusStore(rows, 768, dtype=dtype); st.fill_synthetic(1234)
    g = torch.Generator().manual_seed(1)
    tdt = torch.bfloat16 if dtype == "bfloat16" else torch.float16
    qs = torch.randn((5, 8192, 768), generator=g).to(tdt).cuda()
    for i in range(2): st.search_device(qs[i], k, mode="tensor")
    torch.cuda.synchronize()
    assert not st.check_async()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(2, 5): st.search_device(qs[i], k, mode="tensor")
    e1.record(); torch.cuda.synchronize()
    ovf = st.check_async()
    st.set_profiling(True); st.profile()
    for i in range(2, 5): st.search_device(qs[i], k, mode="tensor")
    p = st.profile()
    print(json.dumps({"rows": rows, "k": k, "first": os.environ.get("VODB_FIRST_ROWS_LARGE"), "growth": os.environ.get("VODB_GROWTH_LARGE"),
                      "ms": e0.elapsed_time(e1) / 3, "score_ms": p["score_ms"] / 3, "select_ms": p["select_ms"] / 3,
                      "segments": st.stats()["segments"], "cap": st.stats()["cap"], "overflow": bool(ovf)}))
else:
    for rows, k, dtype in ((12_500_000, 1000, "float16"), (10_000_000, 100, "bfloat16")):
        for first in ("4096", "16384"):
            for growth in ("2", "3", "5"):
                env = dict(os.environ, VODB_FIRST_ROWS_LARGE=first, VODB_GROWTH_LARGE=growth)
                r = subprocess.run([sys.executable, __file__, "child", str(rows), str(k), dtype], env=env, capture_output=True, text=True)
                print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
