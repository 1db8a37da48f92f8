"""ncu target: a few 64-query searches over a 1.25M x 768 bf16 shard (the 8-GPU shard of BASELINE configs[1])."""
import sys
sys.path.insert(0, ".")
import torch, vod_b200
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 64
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
mode = sys.argv[4] if len(sys.argv) > 4 else "tensor"
st = vod_b200.CorpusStore(rows, 768, dtype="bfloat16"); st.fill_synthetic(1234)
g = torch.Generator().manual_seed(1)
qs = torch.randn((4, nq, 768), generator=g)
if mode == "tensor":
    qs = qs.to(torch.bfloat16).to(torch.float32)
qs = qs.cuda()
for i in range(4):
    st.search_device(qs[i], k, mode=mode)
torch.cuda.synchronize()
print(st.stats())
