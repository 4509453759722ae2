"""Index -> HBM load: DeviceCatalog.from_index (memmap -> pinned staging -> HBM) vs the np.load + upload the reference-style path does."""
import sys, time, tempfile
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, ".")
import instacart_next_order_recommendation_b200 as icr
from instacart_next_order_recommendation_b200.index import EmbeddingIndex

N, D = (int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000), 384
d = Path(tempfile.mkdtemp(dir="/tmp"))
corpus = d / "corpus.json"; corpus.write_text("{}")
ids = [str(i) for i in range(N)]
emb = np.random.default_rng(0).standard_normal((N, D), dtype=np.float32)
idx = EmbeddingIndex(corpus, "m"); idx.save(ids, emb); idx.save_bf16_sidecar()
torch.zeros(1, device="cuda"); gb = N * D * 4 / 1e9

def t(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0); del r
    return best

a = t(lambda: icr.DeviceCatalog(idx.load(ids)))
b = t(lambda: icr.DeviceCatalog.from_index(idx, ids))
c = t(lambda: icr.DeviceCatalog.from_index(idx, ids, dtype=torch.bfloat16, use_sidecar=False))
e = t(lambda: icr.DeviceCatalog.from_index(idx, ids, dtype=torch.bfloat16))
f = t(lambda: icr.DeviceCatalog.from_index(idx, ids, rows=(0, N // 8)))
print(f"{N} x {D} fp32 index ({gb:.2f} GB on disk, page cache warm)")
print(f"np.load + DeviceCatalog(host array)        : {a*1e3:7.1f} ms  ({gb/a:.1f} GB/s)")
print(f"from_index fp32 (streamed, + planes)        : {b*1e3:7.1f} ms  ({gb/b:.1f} GB/s)")
print(f"from_index bf16 (fp32 up, convert on device): {c*1e3:7.1f} ms  ({gb/c:.1f} GB/s of fp32)")
print(f"from_index bf16 from the sidecar            : {e*1e3:7.1f} ms  ({gb/2/e:.1f} GB/s of bf16)")
print(f"from_index fp32, one of 8 row shards        : {f*1e3:7.1f} ms")
