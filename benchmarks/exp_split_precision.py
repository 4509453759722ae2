"""Experiment (not product): does hi*hi + hi*lo + lo*hi on fp16 tensor cores with fp32 accumulation reach
fp32-level accuracy for cosine scores? Uses cuBLAS via torch only to answer the numerics question."""
import torch

from instacart_next_order_recommendation_b200 import ops

torch.manual_seed(0)
N, D, Q = 49688, 384, 256
c = torch.nn.functional.normalize(torch.randn(N, D, device="cuda"), dim=1)
q = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda"), dim=1)
ref = q.double() @ c.double().T
f32 = q @ c.T
pc, pq = ops.split_f16_planes(c), ops.split_f16_planes(q)
Dp = pc.shape[1] // 2
ch, cl, qh, ql = pc[:, :D], pc[:, Dp:Dp + D], pq[:, :D], pq[:, Dp:Dp + D]
# check the split itself (exact arithmetic on the planes, in fp64)
rec = (qh.double() + ql.double()) / 256
print("split reconstruction rel err", ((rec - q.double()).abs().max() / q.abs().max()).item())
exact3 = (qh.double() @ ch.double().T + qh.double() @ cl.double().T + ql.double() @ ch.double().T) / 65536
print("3-term exact-arith abs err vs f64", (exact3 - ref).abs().max().item())
def mm32(a, b):
    try:
        return torch.mm(a, b.T.contiguous() if False else b.T, out_dtype=torch.float32)
    except Exception as e:
        print("out_dtype mm unavailable:", type(e).__name__, str(e)[:100])
        return None
t = mm32(qh.contiguous(), ch.contiguous())
if t is not None:
    s3 = (t + mm32(qh.contiguous(), cl.contiguous()) + mm32(ql.contiguous(), ch.contiguous())) / 65536
    top = ref.topk(100, dim=1)
    got = s3.double().gather(1, top.indices)
    print("tensor-core 3-term: max abs err all", (s3.double() - ref).abs().max().item(), " max rel err top100", ((got - top.values).abs() / top.values.abs()).max().item())
    got32 = f32.double().gather(1, top.indices)
    print("plain fp32 mm     : max abs err all", (f32.double() - ref).abs().max().item(), " max rel err top100", ((got32 - top.values).abs() / top.values.abs()).max().item())
    s1 = t / 65536
    got1 = s1.double().gather(1, top.indices)
    print("tensor-core 1-term (hi only): max rel err top100", ((got1 - top.values).abs() / top.values.abs()).max().item())
