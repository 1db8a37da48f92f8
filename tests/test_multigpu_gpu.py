"""Row-sharded search on >= 2 GPUs of one box (skipped on single-GPU boxes): the fused peer-memory exchange and the
NCCL all-gather path must both reproduce the single-shard result bit for bit."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank: int, world: int, port: int, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    ok, msg = True, ""
    try:
        import vod_b200

        n, d, k = 1_000_003, 256, 100
        g = torch.Generator().manual_seed(5)
        xq = torch.randn((64, d), generator=g).to(torch.bfloat16).to(torch.float32)
        results = {}
        for exchange in ("p2p", "nccl"):
            corpus = vod_b200.ShardedCorpus(n, d, dtype="bfloat16", device=rank, rank=rank, world_size=world,
                                            exchange=exchange, max_queries=256, max_k=1000)
            corpus.fill_synthetic(77)
            for rep in range(3):  # several epochs: exercises the parity double-buffering of the exchange
                s, i = corpus.search_device(xq.cuda(), k, mode="tensor")
            s2, i2 = corpus.search_device(xq[:7].cuda(), 1000, mode="exact")
            torch.cuda.synchronize()
            assert not corpus.any_overflow()
            results[exchange] = (s.cpu().numpy(), i.cpu().numpy(), s2.cpu().numpy(), i2.cpu().numpy())
            corpus.close()
        a, b = results["p2p"], results["nccl"]
        for x, y in zip(a, b):
            if not np.array_equal(x, y):
                ok, msg = False, "p2p and nccl exchange disagree"
        if rank == 0 and ok:
            single = vod_b200.CorpusStore(n, d, dtype="bfloat16", device=0)
            single.fill_synthetic(77)
            ss, si = single.search(xq.numpy(), k, mode="tensor")
            ss2, si2 = single.search(xq[:7].numpy(), 1000, mode="exact")
            single.close()
            if not (np.array_equal(si, a[1]) and np.array_equal(ss, a[0])):
                ok, msg = False, "sharded tensor-mode result differs from the single-shard result"
            if not (np.array_equal(si2, a[3]) and np.array_equal(ss2, a[2])):
                ok, msg = False, "sharded exact-mode result differs from the single-shard result"
        # the drop-in client over the sharded corpus (host buffers, every rank in lockstep): the last shard of an
        # adversarially ordered corpus overflows its lists on the fast schedule; the overflow flag travels with the
        # exchanged lists, so ALL ranks re-run on the overflow-proof schedule inside the call, in step
        if ok:
            from oracle import flat_ip

            n2 = 150_000 * world
            adv = np.zeros((n2, 64), np.float32)
            adv[:, 0] = (np.arange(n2) // 64) % 256
            adv[:, 1] = np.arange(n2) // (64 * 256)
            # only the last shard climbs (-> its lists overflow); the others hold ordinary rows with low, varied scores
            arng = np.random.default_rng(11)
            adv[: n2 - 150_000, :2] = 0.0
            adv[: n2 - 150_000, 2:] = arng.integers(-3, 4, size=(n2 - 150_000, 62)).astype(np.float32)
            aq = np.zeros((3, 64), np.float32)
            aq[:, 0], aq[:, 1] = 1.0, 256.0
            aq[:, 2:] = arng.integers(-2, 3, size=(3, 62)).astype(np.float32)
            corpus = vod_b200.ShardedCorpus(n2, 64, dtype="bfloat16", device=rank, rank=rank, world_size=world,
                                            exchange="p2p", max_queries=64, max_k=100)
            corpus.add_global(adv, 0)
            with vod_b200.B200SearchMaster(store=corpus, serve=False, mode="tensor") as master:
                for rep in range(2):
                    res = master.get_client().search(vector=aq, top_k=100)
                    fell_back = corpus.store.stats()["safe_fallback"] == 1
                    rs, ri = flat_ip.search(adv, aq, 100)
                    if not (np.array_equal(res.indices, ri) and np.array_equal(res.scores, rs)):
                        ok, msg = False, f"sharded client result differs from the oracle on the adversarial corpus (rep {rep})"
                    if not fell_back:
                        ok, msg = False, f"rank {rank} did not re-run on the overflow-proof schedule (rep {rep})"
                ds, di = corpus.search_device(torch.from_numpy(aq).cuda(), 100, mode="tensor")
                torch.cuda.synchronize()
                if not corpus.any_overflow():  # asynchronous path: every rank must learn about the overflow
                    ok, msg = False, f"rank {rank}: any_overflow() missed another shard's overflow"
                # ranks that disagree on the batch shape must not hang in the merge: the shape travels with the
                # overflow flag, every rank sees the mismatch and the next check raises
                bad_nq = 8 if rank == 0 else 16
                corpus.search_device(torch.from_numpy(np.tile(aq, (6, 1))[:bad_nq].copy()).cuda(), 100, mode="tensor")
                torch.cuda.synchronize()
                try:
                    corpus.any_overflow()
                    ok, msg = False, f"rank {rank}: mismatched batch shapes were not reported"
                except vod_b200.VodbError as exc:
                    if "batch shapes" not in str(exc):
                        ok, msg = False, f"rank {rank}: unexpected error {exc}"
                res = master.get_client().search(vector=aq, top_k=100)   # and the corpus keeps working afterwards
                if not np.array_equal(res.indices, ri):
                    ok, msg = False, f"rank {rank}: search after a reported mismatch is wrong"
            corpus.close()
        dist.barrier()
    except Exception as exc:  # report instead of hanging the other rank
        ok, msg = False, f"{type(exc).__name__}: {exc}"
    q.put((rank, ok, msg))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_search_two_gpus():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in results), results


@pytest.mark.timeout(300)
def test_single_process_multi_gpu_master():
    """`B200SearchMaster(vectors, devices=[0, 1, ...])`: all GPUs of the box behind ONE drop-in master / client, the
    analogue of faiss' index_cpu_to_all_gpus(shard=True) in the reference's single server process. Includes the
    adversarial-order corpus, which forces the overflow-proof fallback on every shard."""
    import torch

    import vod_b200
    from oracle import flat_ip
    from tests.helpers import int_valued

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    devices = list(range(min(torch.cuda.device_count(), 4)))
    rng = np.random.default_rng(3)
    vectors, xq = int_valued(rng, (30_011, 96)), int_valued(rng, (33, 96))
    rs, ri = flat_ip.search(vectors, xq, 50)
    with vod_b200.B200SearchMaster(vectors, dtype="bfloat16", devices=devices) as master:
        assert master.store.ntotal == len(vectors)
        out = master.get_client().search(vector=xq, top_k=50)
        assert np.array_equal(out.indices, ri) and np.array_equal(out.scores, rs)
    n = 400_000
    adv = np.zeros((n, 64), np.float32)
    adv[:, 0] = (np.arange(n) // 64) % 256
    adv[:, 1] = np.arange(n) // (64 * 256)
    q = np.zeros((3, 64), np.float32)
    q[:, 0], q[:, 1] = 1.0, 256.0
    rs, ri = flat_ip.search(adv, q, 100)
    with vod_b200.B200SearchMaster(adv, dtype="bfloat16", devices=devices[:2], mode="tensor") as master:
        out = master.get_client().search(vector=q, top_k=100)
        assert np.array_equal(out.indices, ri) and np.array_equal(out.scores, rs)
