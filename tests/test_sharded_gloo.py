"""World-size-2 `gloo` test (CPU) of the row-sharded search composition: local search -> all-gather -> merge.

The product wires `ShardedSearcher` to the CUDA store and the CUDA merge kernel; here the same host logic runs
over gloo with the oracle standing in for the two GPU calls, and must reproduce the unsharded oracle result.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import int_valued


def _worker(rank: int, world: int, port: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import flat_ip
        from vod_b200.sharded import ShardedSearcher, shard_bounds

        rng = np.random.default_rng(0)
        n, d, nq, k = 1000, 16, 5, 37
        xb, xq = int_valued(rng, (n, d)), int_valued(rng, (nq, d))
        lo, hi = shard_bounds(n, world, rank)

        def local_search(queries, top_k):
            s, i = flat_ip.search(xb[lo:hi], queries.numpy(), top_k, row_offset=lo)
            return torch.from_numpy(s), torch.from_numpy(i)

        def merge(all_s, all_i, k_out):
            s, i = all_s[0].numpy(), all_i[0].numpy()
            for g in range(1, all_s.shape[0]):
                s, i = flat_ip.merge_sorted(s, i, all_s[g].numpy(), all_i[g].numpy(), k_out)
            return torch.from_numpy(s), torch.from_numpy(i)

        searcher = ShardedSearcher(local_search, merge)
        s, i = searcher.search(torch.from_numpy(xq), k)
        ref_s, ref_i = flat_ip.search(xb, xq, k)
        ok = bool(np.array_equal(i.numpy(), ref_i) and np.array_equal(s.numpy(), ref_s))
        q.put((rank, ok, (lo, hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_search_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] for r in results), results
    assert sorted(r[2] for r in results) == [(0, 512), (512, 1000)]
