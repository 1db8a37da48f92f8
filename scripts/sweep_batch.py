"""Roofline curve over the query batch size: 10M x 768 bf16, top-100, `tensor` mode, device-resident queries and
results, CUDA events over back-to-back searches. Roofline time = max(corpus bytes / HBM peak, flops / bf16 peak) with
the measured peaks of MEASURED_PEAKS.json. Run on the GPU box: PYTHONPATH=. python scripts/sweep_batch.py"""
import json
import pathlib
import sys

import torch

sys.path.insert(0, ".")
import vod_b200

ROWS, DIM, K = 10_000_000, 768, 100
peaks = {"hbm_gbs": 6534.8, "bf16_tflops": 1671.7}
try:
    mp = json.loads((pathlib.Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())
    peaks["hbm_gbs"] = float(mp.get("hbm_gbs", mp.get("hbm_copy_gbs", peaks["hbm_gbs"])))
except Exception:
    pass
dev = torch.device("cuda:0")
store = vod_b200.CorpusStore(ROWS, DIM, dtype="bfloat16")
store.fill_synthetic(1234)
g = torch.Generator(device=dev).manual_seed(1)
rows = []
for nq in (1, 8, 32, 64, 96, 128, 192, 256, 384, 512, 1024, 2048, 4096, 8192):
    reps = 12 if nq <= 512 else (6 if nq <= 2048 else 3)
    q = torch.randn((reps + 2, nq, DIM), device=dev, generator=g).to(torch.bfloat16)
    for i in range(2):
        store.search_device(q[i], K, mode="tensor")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        store.search_device(q[2 + i], K, mode="tensor")
    e1.record()
    torch.cuda.synchronize()
    assert not store.check_async()
    ms = e0.elapsed_time(e1) / reps
    t_hbm = ROWS * DIM * 2 / (peaks["hbm_gbs"] * 1e9) * 1e3
    t_tc = 2.0 * nq * ROWS * DIM / (peaks["bf16_tflops"] * 1e12) * 1e3
    st = store.stats()
    rows.append({"nq": nq, "ms": ms, "qps": nq / ms * 1e3, "roofline_ms": max(t_hbm, t_tc), "frac": max(t_hbm, t_tc) / ms,
                 "bound": "hbm" if t_hbm >= t_tc else "tensor", "segments": st["segments"], "cap": st["cap"]})
    print(json.dumps(rows[-1]), flush=True)
