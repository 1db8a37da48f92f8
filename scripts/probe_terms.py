import sys, json
sys.path.insert(0, '.')
import numpy as np, torch
import vod_b200
st = vod_b200.CorpusStore(10_000_000, 768, dtype="bfloat16"); st.fill_synthetic(1234)
g = torch.Generator().manual_seed(1)
out = {}
for nq in (64, 8192):
    q = torch.randn((nq, 768), generator=g).cuda()
    for mode in ("tensor", "tensor2", "tensor3", "exact"):
        if mode == "exact" and nq > 64: continue
        for _ in range(2): st.search_device(q, 100, mode=mode)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        n = 5 if nq == 64 else 2
        e0.record()
        for _ in range(n): r = st.search_device(q, 100, mode=mode)
        e1.record(); torch.cuda.synchronize()
        out[f"{mode}_q{nq}_ms"] = e0.elapsed_time(e1) / n
        if nq == 64: out[f"{mode}_ids"] = r[1].cpu().numpy()
# float32 queries that are exact in bf16 (what a bf16 encoder hands over): the 3-term mode skips the empty terms
qe = torch.randn((64, 768), generator=g).to(torch.bfloat16).to(torch.float32).cuda()
for mode in ("tensor", "tensor3"):
    for _ in range(2): st.search_device(qe, 100, mode=mode)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): r = st.search_device(qe, 100, mode=mode)
    e1.record(); torch.cuda.synchronize()
    out[f"{mode}_q64_bf16_exact_queries_ms"] = e0.elapsed_time(e1) / 5
    out[f"{mode}_bf16q"] = (r[0].cpu().numpy(), r[1].cpu().numpy())
a, b = out.pop("tensor_bf16q"), out.pop("tensor3_bf16q")
out["tensor3_equals_tensor_on_bf16_exact_queries"] = bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]))
ex = out.pop("exact_ids")
for mode in ("tensor", "tensor2", "tensor3"):
    ids = out.pop(f"{mode}_ids")
    out[f"{mode}_recall_vs_exact"] = float(np.mean([len(np.intersect1d(a, b)) / 100 for a, b in zip(ids, ex)]))
print(json.dumps(out))
