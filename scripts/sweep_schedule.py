"""Development probe: time one 64-query search for several scan schedules and shard sizes (run under gpurun)."""
import json, os, subprocess, sys
sys.path.insert(0, ".")
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, vod_b200
    rows = int(sys.argv[2])
    st = vod_b200.CorpusStore(rows, 768, dtype="bfloat16"); st.fill_synthetic(1234)
    g = torch.Generator().manual_seed(1)
    qs = torch.randn((30, 64, 768), generator=g).to(torch.bfloat16).to(torch.float32).cuda()
    for i in range(5): st.search_device(qs[i], 100, mode="tensor")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5, 30): st.search_device(qs[i], 100, mode="tensor")
    e1.record(); torch.cuda.synchronize()
    st.set_profiling(True)
    for i in range(5, 30): st.search_device(qs[i], 100, mode="tensor")
    p = st.profile()
    print(json.dumps({"rows": rows, "first": os.environ.get("VODB_FIRST_ROWS"), "growth": os.environ.get("VODB_GROWTH"),
                      "ms": e0.elapsed_time(e1) / 25, "score_ms": p["score_ms"] / 25, "select_ms": p["select_ms"] / 25,
                      "segments": st.stats()["segments"]}))
else:
    for rows in (10_000_000, 1_250_000):
        for first in ("1024", "2048", "4096", "8192"):
            for growth in ("8", "20", "32"):
                env = dict(os.environ, VODB_FIRST_ROWS=first, VODB_GROWTH=growth)
                r = subprocess.run([sys.executable, __file__, "child", str(rows)], env=env, capture_output=True, text=True)
                print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
