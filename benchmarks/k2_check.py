"""Quick GEMM-path (K2) correctness probe against torch on the same GPU (development aid, not a test)."""
import sys, time
import torch
from instacart_next_order_recommendation_b200 import ops

def check(Q, N, D, k, dtype, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    c = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1).to(dtype)
    q = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(dtype)
    t0 = time.time()
    v, i = ops.cos_topk(q, c, k, path=ops.PATH_GEMM)
    torch.cuda.synchronize()
    dt = time.time() - t0
    ref = torch.nn.functional.normalize(q.double(), dim=1) @ torch.nn.functional.normalize(c.double(), dim=1).T
    rv, ri = ref.topk(k, dim=1)
    err = ((v.double() - rv).abs() / rv.abs().clamp_min(0.05)).max().item()
    idm = (i != ri).float().mean().item()
    print(f"Q={Q} N={N} D={D} k={k} {dtype}: max rel err {err:.3e}, id mismatch frac {idm:.4f}, {dt*1e3:.1f} ms", flush=True)
    return err

if __name__ == "__main__":
    torch.cuda.set_device(0)
    check(128, 1024, 64, 10, torch.float32)
    check(128, 1024, 64, 10, torch.bfloat16)
    check(100, 5000, 384, 100, torch.float32)
    check(300, 49688, 384, 100, torch.float32)
    check(300, 49688, 384, 100, torch.bfloat16)
    check(1000, 20000, 768, 100, torch.bfloat16)
    check(16, 300000, 384, 100, torch.float32)
    check(10000, 49688, 384, 100, torch.float32)
    print("k2_check done")
