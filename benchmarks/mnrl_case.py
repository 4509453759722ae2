import sys, torch
sys.path.insert(0, ".")
import instacart_next_order_recommendation_b200 as icr
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dt = torch.bfloat16 if len(sys.argv) > 2 and sys.argv[2] == "bf16" else torch.float32
a = torch.randn(B, 384, device="cuda").to(dt).requires_grad_(True)
p = torch.randn(B, 384, device="cuda").to(dt).requires_grad_(True)
for _ in range(5):
    a.grad = p.grad = None
    icr.mnrl_loss(a, p, 20.0).backward()
torch.cuda.synchronize()
print("ok")
