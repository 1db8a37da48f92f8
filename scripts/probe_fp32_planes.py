"""float32 store: CUDA-core exact kernel vs the tensor-core path over bf16 planes (tensor / tensor2 / tensor3).
Times whole searches with CUDA events (device-resident queries and results). Run on the GPU box."""
import json
import sys

import torch

sys.path.insert(0, ".")
import vod_b200

ROWS, DIM, K = 4_000_000, 768, 100
dev = torch.device("cuda:0")
store = vod_b200.CorpusStore(ROWS, DIM, dtype="float32")
store.fill_synthetic(1234)
g = torch.Generator(device=dev).manual_seed(1)
res = {"rows": ROWS, "dim": DIM, "k": K, "store_gb": ROWS * DIM * 4 / 1e9}
ref = {}
for nq in (64, 256, 1024):
    q = torch.randn((8, nq, DIM), device=dev, generator=g)
    for mode in ("exact", "tensor3", "tensor2", "tensor"):
        for i in range(2):
            out = store.search_device(q[i], K, mode=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3 if mode == "exact" else 6
        e0.record()
        for i in range(n):
            out = store.search_device(q[2 + i], K, mode=mode)
        e1.record()
        out = store.search_device(q[7], K, mode=mode)  # same batch for every mode: recall against the exact kernel
        torch.cuda.synchronize()
        assert not store.check_async()
        ms = e0.elapsed_time(e1) / n
        res[f"q{nq}_{mode}_ms"] = ms
        res[f"q{nq}_{mode}_GBps_of_fp32_bytes"] = ROWS * DIM * 4 / ms / 1e6
        ids = out[1]
        if mode == "exact":
            ref[nq] = ids.clone()
        else:
            same = (ids.unsqueeze(2) == ref[nq].unsqueeze(1)).any(2).float().mean().item()
            res[f"q{nq}_{mode}_recall_vs_exact"] = same
print(json.dumps(res))
