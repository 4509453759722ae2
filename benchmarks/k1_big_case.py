"""One large batch-1 call on 768-byte rows (C5's row shape): ncu target for the K1 ring kernel."""
import sys, torch
sys.path.insert(0, ".")
import instacart_next_order_recommendation_b200 as icr
N, D = 5_208_333, 384
g = torch.Generator(device="cuda").manual_seed(0)
rows = torch.empty(N, D, dtype=torch.bfloat16, device="cuda")
for s in range(0, N, 1 << 20):
    e = min(N, s + (1 << 20))
    rows[s:e] = torch.nn.functional.normalize(torch.randn(e - s, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
cat = icr.DeviceCatalog(rows, dtype=torch.bfloat16)
q = torch.nn.functional.normalize(torch.randn(1, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
for _ in range(4):
    v, i = cat.topk(q, 100)
torch.cuda.synchronize()
print("ok", v[0, :3].tolist())
