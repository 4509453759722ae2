"""C3 B-sweep: fused MNRL fwd+bwd (CUDA-core kernels for small batches, tcgen05 path above) vs PyTorch eager on the same GPU."""
import sys, torch
sys.path.insert(0, ".")
import instacart_next_order_recommendation_b200 as icr

def eager(a, p, scale):
    an = torch.nn.functional.normalize(a.float(), dim=1); pn = torch.nn.functional.normalize(p.float(), dim=1)
    s = an @ pn.T * scale
    return torch.nn.functional.cross_entropy(s, torch.arange(a.shape[0], device=a.device))

def timeit(fn, a, p, iters):
    for _ in range(5):
        a.grad = p.grad = None; fn(a, p, 20.0).backward()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a.grad = p.grad = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(a, p, 20.0).backward(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]

Bs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [64, 128, 256, 512, 1024, 2048, 4096, 8192]
print("| B | dtype | ours us (median / min) | eager us (median / min) | speed-up | algorithmic TFLOP/s (6 B^2 D) |")
print("|---|---|---|---|---|---|")
for dt in (torch.bfloat16, torch.float32):
    for B in Bs:
        a = torch.randn(B, 384, device="cuda").to(dt).requires_grad_(True)
        p = torch.randn(B, 384, device="cuda").to(dt).requires_grad_(True)
        it = 30 if B <= 2048 else 10
        o = timeit(icr.mnrl_loss, a, p, it); e = timeit(eager, a, p, it)
        print(f"| {B} | {str(dt).split('.')[-1]} | {o[0]:.1f} / {o[1]:.1f} | {e[0]:.1f} / {e[1]:.1f} | {e[0] / o[0]:.2f}x | {6.0 * B * B * 384 / (o[0] * 1e-6) / 1e12:.1f} |")
