"""Development probe: scan bandwidth (64 queries, bf16, `tensor`) against the row length, with the dense pitch
(VODB_PITCH_PAD=0) and with one extra 64-element chunk per row (VODB_PITCH_PAD=2). Run under gpurun."""
import json, os, subprocess, sys
sys.path.insert(0, ".")
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, vod_b200
    dim = int(sys.argv[2])
    rows = int(float(sys.argv[3]) / (dim * 2)) // 128 * 128
    st = vod_b200.CorpusStore(rows, dim, dtype="bfloat16"); st.fill_synthetic(1234)
    g = torch.Generator().manual_seed(1)
    qs = torch.randn((14, 64, dim), generator=g).to(torch.bfloat16).cuda()
    for i in range(4): st.search_device(qs[i], 100, mode="tensor")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(4, 14): st.search_device(qs[i], 100, mode="tensor")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"dim": dim, "rows": rows, "pad": os.environ.get("VODB_PITCH_PAD"), "ms": ms, "GBps": rows * dim * 2 / ms / 1e6}))
else:
    cases = [(d, 8e9) for d in (256, 384, 512, 640, 768, 1024, 1280, 1536, 2048)] if len(sys.argv) < 2 else \
            [(512, 12.8e9), (1024, 12.8e9), (1024, 25.6e9), (2048, 12.8e9), (768, 15.36e9), (768, 25.6e9), (1024, 4e9)]
    for dim, nbytes in cases:
        for pad in ("0", "2"):
            r = subprocess.run([sys.executable, __file__, "child", str(dim), str(nbytes)], env=dict(os.environ, VODB_PITCH_PAD=pad), capture_output=True, text=True)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
