"""C3 (MNRL fwd+bwd, B=256, D=384): wall clock per step of the autograd path, the one-call path and the CUDA-graph step, vs eager."""
import sys
import time

import torch

sys.path.insert(0, ".")
import instacart_next_order_recommendation_b200 as icr  # noqa: E402
from instacart_next_order_recommendation_b200 import ops  # noqa: E402

B, D, scale = (int(sys.argv[1]) if len(sys.argv) > 1 else 256), 384, 20.0
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(7)
a0 = torch.nn.functional.normalize(torch.randn(B, D, device=dev, generator=g), dim=1)
p0 = torch.nn.functional.normalize(a0 + 0.3 * torch.randn(B, D, device=dev, generator=g), dim=1)


def wall(fn, n=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def dev_time(fn, n=100):
    for _ in range(10):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
    return ts[len(ts) // 2]


for dt in (torch.bfloat16, torch.float16, torch.float32):
    a = a0.to(dt).requires_grad_(True)
    p = p0.to(dt).requires_grad_(True)

    def autograd_path():
        a.grad = p.grad = None
        icr.mnrl_loss(a, p, scale).backward()

    def one_call():
        ops.mnrl_forward_backward(a.detach(), p.detach(), scale)

    def eager():
        a.grad = p.grad = None
        an = torch.nn.functional.normalize(a.float(), dim=1)
        pn = torch.nn.functional.normalize(p.float(), dim=1)
        torch.nn.functional.cross_entropy(an @ pn.T * scale, torch.arange(B, device=dev)).backward()

    step = icr.mnrl_step_graph(B, D, dt, scale)
    ad, pd = a.detach(), p.detach()
    loss, ga, gp = step(ad, pd)
    autograd_path()
    err = max((ga.float() - a.grad.float()).abs().max().item(), (gp.float() - p.grad.float()).abs().max().item())
    print(f"{str(dt):16s} B={B}: autograd {wall(autograd_path):7.1f} us | one call {wall(one_call):7.1f} us | graph step wall {wall(lambda: step(ad, pd)):6.1f} us, "
          f"device {dev_time(lambda: step.graph.replay()):6.1f} us | eager {wall(eager):7.1f} us | graph-vs-autograd grad diff {err:.2e}")
