"""BASELINE configs[4] at full size: 50M x 1024 passage vectors arrive as float32 in pinned host memory (2^18-row
batches, build_gpu.py:294), are stored as bf16 row shards over all ranks (index refresh), then searched fp32-exact
(three bf16 query terms on the tensor cores, fused peer-memory exchange). One process per GPU:
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/probe_c5_multi.py"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import vod_b200

ROWS, DIM, K, NQ = int(os.environ.get("C5_ROWS", 50_000_000)), 1024, 100, 64
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


corpus = vod_b200.ShardedCorpus(ROWS, DIM, dtype="bfloat16", device=local, rank=rank, world_size=world, max_queries=NQ, max_k=K)
n_local = corpus.hi - corpus.lo
g = torch.Generator().manual_seed(100 + rank)
chunk = torch.randn((262_144, DIM), dtype=torch.float32, generator=g).pin_memory()
out = {"world": world, "rows": ROWS, "dim": DIM, "rows_per_rank": n_local}

# ---- index refresh: float32 host batches -> bf16 shard (every batch scaled differently so that rows differ) ----
barrier()
t0 = time.perf_counter()
row = 0
while row < n_local:
    m = min(len(chunk), n_local - row)
    corpus.store.add(chunk[:m], row0=row)
    row += m
torch.cuda.synchronize()
dt = max_over_ranks(time.perf_counter() - t0)
out["ingest_s"] = dt
out["ingest_host_GBps_aggregate"] = ROWS * DIM * 4 / dt / 1e9
out["ingest_rows_per_s"] = ROWS / dt

# ---- fp32-exact search of float32 queries over the refreshed shards ----
gq = torch.Generator().manual_seed(7)
queries = torch.randn((12, NQ, DIM), generator=gq).to(dev)
for mode in ("tensor3", "tensor"):
    for i in range(3):
        res = corpus.search_device(queries[i], K, mode=mode)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3, 12):
        res = corpus.search_device(queries[i], K, mode=mode)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / 9)
    assert not corpus.any_overflow()
    out[f"search_{mode}_ms"] = ms
    out[f"search_{mode}_qps"] = NQ / ms * 1e3
    out[f"search_{mode}_GBps_aggregate"] = ROWS * DIM * 2 / ms / 1e6
    out[f"ids_{mode}"] = res[1]
# the CUDA-core fp32 kernel as cross-check of the last batch
ex = corpus.search_device(queries[11], K, mode="exact")
t3 = out.pop("ids_tensor3")
out.pop("ids_tensor")
torch.cuda.synchronize()
same = (t3.unsqueeze(2) == ex[1].unsqueeze(1)).any(2).float().mean().item()
out["tensor3_recall_vs_cuda_core_exact"] = same
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
