"""IR-evaluator tail: device metric kernel vs the host numpy path vs the per-query Python loops (oracle), C2-sized."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import instacart_next_order_recommendation_b200 as icr
from instacart_next_order_recommendation_b200 import ops
from oracle import oracle

Q, N, K = 10_000, 49_688, 100
rng = np.random.default_rng(0)
ids = np.stack([rng.permutation(N)[:K] for _ in range(Q)]).astype(np.int64)
rel = [set(int(x) for x in rng.choice(N, size=8, replace=False)) | {int(ids[q, rng.integers(0, 20)])} for q in range(Q)]
specs = ([(ops.METRIC_ACCURACY, k) for k in (1, 3, 5, 10)] + [(ops.METRIC_PRECISION, k) for k in (1, 3, 5, 10)]
         + [(ops.METRIC_RECALL, k) for k in (1, 3, 5, 10)] + [(ops.METRIC_MRR, 10), (ops.METRIC_NDCG, 10), (ops.METRIC_MAP, 100)])
table = ops.RelevanceTable([sorted(r) for r in rel], device="cuda")
ids_d = torch.from_numpy(ids).cuda()
for _ in range(3):
    means, _ = ops.ir_metrics(ids_d, table, specs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    means, _ = ops.ir_metrics(ids_d, table, specs)
e1.record(); torch.cuda.synchronize()
print(f"device metric kernels: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per {Q} x {K} id matrix ({len(specs)} metrics)")
t0 = time.perf_counter(); got = means.tolist(); t_read = time.perf_counter() - t0
corpus = {str(i): "" for i in range(N)}
ev = icr.InformationRetrievalEvaluator({str(q): "" for q in range(Q)}, corpus, {str(q): {str(r) for r in rel[q]} for q in range(Q)})
t0 = time.perf_counter(); host = ev.compute_metrics_from_ids(ids); t_np = time.perf_counter() - t0
t0 = time.perf_counter(); want = oracle.st_ir_metrics([list(r) for r in ids[:1000]], rel[:1000]); t_loop = (time.perf_counter() - t0) * Q / 1000
print(f"host numpy path: {t_np * 1e3:.1f} ms (+ {ids.nbytes / 1e6:.0f} MB of ids to the host); per-query Python loops (as upstream): ~{t_loop * 1e3:.0f} ms")
print("max |device - numpy|:", max(abs(a - b) for a, b in zip(got, host.values())))
