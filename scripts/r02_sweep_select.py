"""Development probe: select / merge kernel time against the CTA size (VODB_SEL_THREADS) and the selection path
(VODB_FAST_SELECT=0: radix select only) on the 8-GPU shard shape. `ab` as the first argument: fast path on / off only."""
import json, os, subprocess, sys
sys.path.insert(0, ".")
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, vod_b200
    from vod_b200.search import merge_topk_device
    st = vod_b200.CorpusStore(1_250_000, 768, dtype="bfloat16"); st.fill_synthetic(1234)
    g = torch.Generator().manual_seed(1)
    qs = torch.randn((40, 64, 768), generator=g).to(torch.bfloat16).to(torch.float32).cuda()
    for i in range(5): st.search_device(qs[i], 100, mode="tensor")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5, 40): st.search_device(qs[i], 100, mode="tensor")
    e1.record(); torch.cuda.synchronize()
    st.set_profiling(True)
    for i in range(5, 40): st.search_device(qs[i], 100, mode="tensor")
    p = st.profile()
    s = torch.randn((8, 64, 100), device="cuda").sort(dim=-1, descending=True).values.contiguous()
    ids = torch.randint(0, 10**7, (8, 64, 100), device="cuda")
    for _ in range(5): merge_topk_device(s, ids, 100)
    torch.cuda.synchronize()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record()
    for _ in range(50): merge_topk_device(s, ids, 100)
    m1.record(); torch.cuda.synchronize()
    print(json.dumps({"threads": os.environ.get("VODB_SEL_THREADS", "default"), "fast_select": os.environ.get("VODB_FAST_SELECT", "1"), "search_ms": e0.elapsed_time(e1) / 35,
                      "select_ms_per_search": p["select_ms"] / 35, "merge8_us": m0.elapsed_time(m1) / 50 * 1e3}))
else:
    ab = len(sys.argv) > 1 and sys.argv[1] == "ab"
    for th, fast in (((None, "0"), (None, "1"), ("512", "1"), ("256", "1")) if ab else [(t, "1") for t in (None, "128", "256", "512", "1024")]):
        env = dict(os.environ)
        env["VODB_FAST_SELECT"] = fast
        if th: env["VODB_SEL_THREADS"] = th
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
