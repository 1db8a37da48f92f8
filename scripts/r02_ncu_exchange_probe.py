"""Two-rank search of a 2 x 1.25M-row corpus through the fused exchange, arranged so that rank 0 can run under ncu:

  rank 1 (never profiled) enqueues its search first; its final select stores the tagged words into rank 0's gather
  buffer and its merge kernel then waits for rank 0's words. Rank 0 starts 0.3 s later, so every word it waits for is
  already there when ncu saves the memory image it replays from; the words rank 0's select stores to rank 1 are the
  same in every replay pass.

Started by scripts/r02_ncu_exchange.sh (RANK / WORLD_SIZE / MASTER_* in the environment; no torchrun so that only
rank 0's command line carries ncu)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import vod_b200

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
rows, dim = 2_500_000, 768
corpus = vod_b200.ShardedCorpus(rows, dim, dtype="bfloat16", device=rank, max_queries=64, max_k=1000)
corpus.fill_synthetic(1234)
rng = np.random.default_rng(7)
q = torch.from_numpy(rng.standard_normal((64, dim), dtype=np.float32)).to(torch.bfloat16).cuda()


def one_search(k: int):
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        time.sleep(0.3)
    s, i = corpus.search_device(q, k)
    torch.cuda.synchronize()
    return s, i


results = {}
for name, k in (("warm", 100), ("k100", 100), ("k1000", 1000)):
    s, i = one_search(k)
    results[name] = (s.cpu().numpy(), i.cpu().numpy())
assert not corpus.any_overflow()
assert np.array_equal(results["warm"][1], results["k100"][1])
assert np.array_equal(results["k1000"][1][:, :100], results["k100"][1])
gathered = [None] * world
dist.all_gather_object(gathered, results["k1000"][1][:, :8].tolist())
assert all(g == gathered[0] for g in gathered)
print(f"rank {rank}: exchange probe ok", flush=True)
corpus.close()
dist.destroy_process_group()
