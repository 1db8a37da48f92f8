"""Development probe (run under gpurun): one search per (shard size, query count, first-segment rows, growth, list
capacity) through the VODB_FIRST_ROWS / VODB_GROWTH / VODB_CAP knobs."""
import json, os, subprocess, sys
sys.path.insert(0, ".")
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, vod_b200
    rows, nq, k = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    st = vod_b200.CorpusStore(rows, 768, dtype="bfloat16"); st.fill_synthetic(1234)
    g = torch.Generator().manual_seed(1)
    n = 30 if nq <= 512 else 8
    qs = torch.randn((n, nq, 768), generator=g).to(torch.bfloat16).to(torch.float32).cuda()
    for i in range(3): st.search_device(qs[i], k, mode="tensor")
    torch.cuda.synchronize()
    ovf = st.check_async()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3, n): st.search_device(qs[i], k, mode="tensor")
    e1.record(); torch.cuda.synchronize()
    st.set_profiling(True)
    for i in range(3, n): st.search_device(qs[i], k, mode="tensor")
    p = st.profile()
    print(json.dumps({"rows": rows, "nq": nq, "k": k, "first": os.environ.get("VODB_FIRST_ROWS"), "growth": os.environ.get("VODB_GROWTH"),
                      "cap_env": os.environ.get("VODB_CAP"), "ms": e0.elapsed_time(e1) / (n - 3), "score_ms": p["score_ms"] / (n - 3),
                      "select_ms": p["select_ms"] / (n - 3), "segments": st.stats()["segments"], "cap": st.stats()["cap"],
                      "overflow": bool(ovf or st.check_async())}))
else:
    grid = []
    for rows in (1_250_000, 10_000_000):
        for nq in (64,):
            for first, growth, cap in (("4096", "20", "16384"), ("4096", "32", "32768"), ("8192", "64", "65536"), ("16384", "40", "32768"),
                                       ("16384", "80", "65536"), ("16384", "160", "131072"), ("32768", "40", "65536"), ("8192", "160", "131072")):
                grid.append((rows, nq, 100, first, growth, cap))
    for nq in (128, 192, 256, 320, 384, 512, 1024):
        for first, growth, cap in (("4096", "3", "16384"), ("4096", "8", "16384"), ("4096", "20", "16384"), ("8192", "32", "32768"), ("16384", "80", "65536")):
            grid.append((10_000_000, nq, 100, first, growth, cap))
    for rows, nq, k, first, growth, cap in grid:
        env = dict(os.environ, VODB_FIRST_ROWS=first, VODB_GROWTH=growth, VODB_CAP=cap, VODB_FIRST_ROWS_LARGE=first, VODB_GROWTH_LARGE=growth)
        r = subprocess.run([sys.executable, __file__, "child", str(rows), str(nq), str(k)], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
