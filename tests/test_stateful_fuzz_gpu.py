"""Stateful differential test: three stores (fp32 / bf16 / fp16) live side by side and receive a random sequence of
appends, overwrites and searches (every mode, host / device / one-call chain entry points, batch sizes on both
sides of every tile boundary, k from 1 to 2048). Workspaces, staging buffers, bf16 planes and scratch are reused
from call to call, so anything one call leaves behind meets the next one. Data are small integers: every mode must
reproduce the oracle bit for bit, ties included."""
import numpy as np
import pytest

import vod_b200
from oracle import flat_ip
from tests.helpers import int_valued

pytestmark = pytest.mark.gpu


def _same_bits(a, b):
    return np.array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("seed", [2024, 7])
def test_random_call_sequences_on_shared_workspaces(seed):
    import torch

    rng = np.random.default_rng(seed)
    d, cap = 200, 40_000
    rows = {dt: int_valued(rng, (cap, d)) for dt in ("float32", "bfloat16", "float16")}
    stores = {dt: vod_b200.CorpusStore(cap, d, dtype=dt) for dt in rows}
    filled = {}
    for dt, st in stores.items():
        st.add(rows[dt][:15_000])
        filled[dt] = 15_000
    n_checked = 0
    for step in range(110):
        dt = str(rng.choice(list(stores)))
        st, n = stores[dt], filled[dt]
        u = rng.random()
        if u < 0.10 and n < cap:                                   # append a block
            m = min(int(rng.integers(1, 7000)), cap - n)
            st.add(rows[dt][n:n + m], row0=n)
            filled[dt] = n + m
            continue
        if u < 0.16:                                               # overwrite a block in place
            r0 = int(rng.integers(0, n - 300))
            rows[dt][r0:r0 + 300] = int_valued(rng, (300, d), lo=-3, hi=4) * (1 + (step % 3))
            rows[dt][r0:r0 + 300] = np.clip(rows[dt][r0:r0 + 300], -6, 6)
            st.add(rows[dt][r0:r0 + 300], row0=r0)
            continue
        nq = int(rng.choice([1, 2, 9, 31, 64, 65, 128, 129, 200, 256, 257, 300]))
        k = int(rng.choice([1, 7, 100, 1000, 2048]))
        mode = str(rng.choice(["exact", "tensor", "tensor2", "tensor3"]))
        entry = str(rng.choice(["host", "device", "chain"]))
        xq = int_valued(rng, (nq, d))
        ref_s, ref_i = flat_ip.search(rows[dt][:n], xq, k)
        tag = (seed, step, dt, n, nq, k, mode, entry)
        if entry == "host":
            s, i = st.search(xq, k, mode=mode)
        elif entry == "device":
            ts, ti = st.search_device(torch.from_numpy(xq).cuda(), k, mode=mode)
            s, i = ts.cpu().numpy(), ti.cpu().numpy()
            if st.check_async():                                   # overflow on the fast schedule: redo synchronously
                s, i = st.search(xq, k, mode=mode)
        else:
            total = min(8, k)
            gold = ref_i[:, :1].copy()
            picks = vod_b200.DenseRetrievalSampler(st, top_k=k, total=total, max_pos_sections=min(2, total), mode=mode)(
                xq, gold, seed=step, offset=nq)
            labels = (ref_i[:, :, None] == gold[:, None, :]).any(-1).astype(np.int64)
            host = vod_b200.sample_search_results(
                search_results=vod_b200.RetrievalBatch(scores=ref_s, indices=ref_i, labels=labels),
                raw_scores={"dense": ref_s}, total=total, max_pos_sections=min(2, total), seed=step, offset=nq)
            assert np.array_equal(picks.batch.indices, host.batch.indices), tag
            assert np.array_equal(picks.batch.scores, host.batch.scores), tag
            assert _same_bits(picks.log_weights, host.log_weights), tag
            assert np.array_equal(picks.max_sampling_id, host.max_sampling_id), tag
            n_checked += 1
            continue
        assert np.array_equal(i, ref_i), tag
        assert np.array_equal(s, ref_s), tag
        n_checked += 1
    assert n_checked > 60
    for st in stores.values():
        st.close()
