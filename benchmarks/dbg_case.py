import sys, torch
sys.path.insert(0, ".")
from instacart_next_order_recommendation_b200 import ops
Q, N, D, k = [int(x) for x in sys.argv[1:5]]
dt = torch.float32 if len(sys.argv) < 6 or sys.argv[5] == "f32" else torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
c = torch.randn(N, D, device="cuda", generator=g).to(dt)
q = torch.randn(Q, D, device="cuda", generator=g).to(dt)
v, i = ops.cos_topk(q, c, k, path=ops.PATH_GEMV)
torch.cuda.synchronize()
ref = torch.nn.functional.normalize(q.double(), dim=1) @ torch.nn.functional.normalize(c.double(), dim=1).T
rv, ri = ref.topk(k, dim=1)
print("ok", ((v.double() - rv).abs().max()).item(), (i != ri).float().mean().item())
