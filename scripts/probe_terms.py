import sys, time, json
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
ex = out.pop("exact_ids")
for mode in ("tensor", "tensor2", "tensor3"):
    ids = out.pop(f"{mode}_ids")
    out[f"{mode}_recall_vs_exact"] = float(np.mean([len(np.intersect1d(a, b)) / 100 for a, b in zip(ids, ex)]))
print(json.dumps(out))
