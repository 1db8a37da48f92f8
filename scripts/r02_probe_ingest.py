"""Development probe: host -> HBM ingest rate of `CorpusStore.add` (pinned float32 source, bf16 store; 2^18-row blocks
like build_gpu.py:294) and of `zarr_io.ingest` from an uncompressed zarr-v2 store on local disk."""
import json, sys, tempfile, time
sys.path.insert(0, ".")
import numpy as np, torch, vod_b200
from vod_b200 import zarr_io

dim, block, n_blocks = 1024, 1 << 18, 12
src = torch.randn((block, dim), dtype=torch.float32).pin_memory()
st = vod_b200.CorpusStore(block * n_blocks, dim, dtype="bfloat16")
st.add(src, row0=0)
torch.cuda.synchronize()
t0 = time.perf_counter()
for b in range(1, n_blocks):
    st.add(src, row0=b * block)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gb = (n_blocks - 1) * block * dim * 4 / 1e9
out = {"pinned_f32_to_bf16_store_gb_per_s": gb / dt, "rows_per_s": (n_blocks - 1) * block / dt, "gb": gb}
# one big add (many 64 MiB chunks inside one call: the double-buffered path)
big = torch.randn((block * 4, dim), dtype=torch.float32).pin_memory()
st2 = vod_b200.CorpusStore(block * 4, dim, dtype="bfloat16")
st2.add(big[:1024], row0=0)
torch.cuda.synchronize(); t0 = time.perf_counter()
st2.add(big, row0=0)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
out["one_call_4gb_gb_per_s"] = big.numel() * 4 / 1e9 / dt
with tempfile.TemporaryDirectory() as d:
    x = np.random.default_rng(0).standard_normal((200_000, 768), dtype=np.float32)
    p = zarr_io.write_zarr_v2(d + "/emb", x, chunk_size=100)
    arr = zarr_io.open_vectors(p)
    st3 = vod_b200.CorpusStore(len(x), 768, dtype="bfloat16")
    t0 = time.perf_counter(); zarr_io.ingest(st3, arr, batch_rows=1 << 16); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out["zarr_uncompressed_chunks100_gb_per_s"] = x.nbytes / 1e9 / dt
    ok = np.array_equal(st3.read(199_990, 10), vod_b200.search.np.asarray(torch.from_numpy(x[199_990:]).to(torch.bfloat16).to(torch.float32)))
    out["zarr_roundtrip_ok"] = bool(ok)
print(json.dumps(out))
