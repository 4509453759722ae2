"""Embedding index: the reference's on-disk catalog cache plus its device-resident form.

``EmbeddingIndex`` keeps the reference's format byte for byte
(src/inference/serve_recommendations.py:66-130; file names from src/constants.py:88-92):

    <corpus dir>/.embedding_index/<sha256(model_dir|corpus_path)[:16]>/
        manifest.json   {"corpus_path", "model_dir", "corpus_mtime", "n_products"}
        embeddings.npy  float32 [N, D]
        product_ids.json

``DeviceCatalog`` is what the kernels read: the same rows resident in HBM (fp32, or bf16
on request), plus — for fp32 — the fp16 (hi|lo) operand planes of the tensor-core path,
built once at load instead of re-normalising the catalog on every request as
``sentence_transformers.util.cos_sim`` does (serve_recommendations.py:214).
"""

from __future__ import annotations

import hashlib
import json
import logging
from pathlib import Path

import numpy as np
import torch

from . import ops
from .similarity import default_device, to_device_matrix

logger = logging.getLogger(__name__)

INDEX_SUBDIR = ".embedding_index"
MANIFEST_FILENAME = "manifest.json"
EMBEDDINGS_FILENAME = "embeddings.npy"
PRODUCT_IDS_FILENAME = "product_ids.json"


class EmbeddingIndex:
    """Disk cache of product embeddings, keyed by corpus path + model dir + corpus mtime."""

    def __init__(self, corpus_path: Path, model_dir: Path | str):
        self.corpus_path = Path(corpus_path).resolve()
        self.model_dir = model_dir
        self._dir = self._index_dir()

    def _index_dir(self) -> Path:
        key = hashlib.sha256(f"{self.model_dir!s}|{self.corpus_path!s}".encode()).hexdigest()[:16]
        return self.corpus_path.parent / INDEX_SUBDIR / key

    @property
    def directory(self) -> Path:
        return self._dir

    def _corpus_mtime(self):
        try:
            return self.corpus_path.stat().st_mtime
        except OSError:
            return None

    def load(self, product_ids: list[str]) -> np.ndarray | None:
        """Embeddings if every validity check of the reference passes, else None."""
        try:
            meta = json.loads((self._dir / MANIFEST_FILENAME).read_text())
        except (OSError, json.JSONDecodeError):
            return None
        if meta.get("corpus_path") != str(self.corpus_path) or meta.get("model_dir") != str(self.model_dir):
            return None
        mtime = self._corpus_mtime()
        if mtime is None or meta.get("corpus_mtime") != mtime:
            return None
        try:
            embeddings = np.load(self._dir / EMBEDDINGS_FILENAME)
            cached_ids = json.loads((self._dir / PRODUCT_IDS_FILENAME).read_text())
        except (OSError, ValueError):
            return None
        if cached_ids != product_ids or len(embeddings) != len(product_ids):
            return None
        return embeddings

    def save(self, product_ids: list[str], embeddings: np.ndarray) -> None:
        self._dir.mkdir(parents=True, exist_ok=True)
        mtime = self._corpus_mtime()
        manifest = {
            "corpus_path": str(self.corpus_path),
            "model_dir": str(self.model_dir),
            "corpus_mtime": 0 if mtime is None else mtime,
            "n_products": len(product_ids),
        }
        (self._dir / MANIFEST_FILENAME).write_text(json.dumps(manifest, indent=2))
        np.save(self._dir / EMBEDDINGS_FILENAME, np.asarray(embeddings).astype(np.float32))
        (self._dir / PRODUCT_IDS_FILENAME).write_text(json.dumps(product_ids))
        logger.info("Saved embedding index to %s (%d products)", self._dir, len(product_ids))


class DeviceCatalog:
    """Catalog rows resident in HBM, with whatever the kernels want precomputed.

    dtype float32 keeps the reference's numerics (scores within 1e-5 of the fp32 oracle);
    dtype bfloat16 halves the bytes a batch-1 request streams (scores within 2e-3).
    """

    def __init__(self, embeddings, *, device: torch.device | None = None, dtype: torch.dtype = torch.float32,
                 row_offset: int = 0, build_planes: bool | None = None):
        if dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("catalog dtype must be float32 or bfloat16")
        dev = device if device is not None else (embeddings.device if isinstance(embeddings, torch.Tensor) and embeddings.is_cuda else default_device())
        self.rows = to_device_matrix(embeddings, device=dev, dtype=dtype)
        self.rows = ops._rows(self.rows)
        self.row_offset = int(row_offset)
        self.planes: torch.Tensor | None = None
        if build_planes is None:
            build_planes = dtype == torch.float32
        if build_planes and dtype == torch.float32 and self.rows.shape[0] > 0:
            self.planes = ops.split_f16_planes(self.rows)
        self.inv_norms: torch.Tensor | None = None
        if dtype == torch.bfloat16 and self.rows.shape[0] > 0:
            self.inv_norms = ops.row_inv_norms(self.rows)

    @property
    def device(self) -> torch.device:
        return self.rows.device

    @property
    def dtype(self) -> torch.dtype:
        return self.rows.dtype

    def __len__(self) -> int:
        return self.rows.shape[0]

    @property
    def dim(self) -> int:
        return self.rows.shape[1]

    @property
    def nbytes(self) -> int:
        return self.rows.numel() * self.rows.element_size() + (0 if self.planes is None else self.planes.numel() * 2)

    # ---- request-sized calls: one CUDA graph per (Q, k) -------------------------------------------------
    def _graph_for(self, Q: int, k: int, path: int):
        """Capture prep + scoring + select of a fixed-shape request once; replays cost one graph launch.

        The serve path is launch-latency-bound (a batch-1 request streams 76 MB in ~12 µs of HBM time), so the
        Python/ctypes work and the gaps between the 2-3 kernels of a call matter more than the kernels.
        """
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        key = (Q, k, path)
        entry = self._graphs.get(key)
        if entry is None:
            static_q = torch.zeros(Q, self.rows.shape[1], dtype=self.dtype, device=self.device)
            kw = dict(cat_planes=self.planes, cat_inv_norms=self.inv_norms, row_offset=self.row_offset, path=path)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):  # warm-up outside capture: one-time attribute setting, lazy module load
                ops.cos_topk(static_q, self.rows, k, **kw)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                vals, ids = ops.cos_topk(static_q, self.rows, k, **kw)
            entry = (graph, static_q, vals, ids)
            self._graphs[key] = entry
        return entry

    def topk_small(self, queries, k: int, *, path: int = ops.PATH_AUTO, copy: bool = True):
        """topk() for request-sized batches (Q <= 8) through a cached CUDA graph.

        `queries` may live on the host (numpy / CPU tensor, as SentenceTransformer.encode returns it) or on the
        device. With copy=False the returned tensors are the graph's static outputs, overwritten by the next call.
        """
        q = queries if isinstance(queries, torch.Tensor) else torch.as_tensor(queries)
        if q.dim() == 1:
            q = q.unsqueeze(0)
        k = min(int(k), len(self))
        if k < 1 or q.shape[0] == 0:
            return self.topk(q, max(k, 1))
        graph, static_q, vals, ids = self._graph_for(q.shape[0], k, path)
        D = min(q.shape[1], static_q.shape[1])
        static_q[:, :D].copy_(q[:, :D], non_blocking=True)  # H2D (or D2D) + dtype conversion in one op
        graph.replay()
        return (vals.clone(), ids.clone()) if copy else (vals, ids)

    def topk_host(self, queries: torch.Tensor, k: int, *, out: tuple[torch.Tensor, torch.Tensor] | None = None,
                  n_chunks: int = 2, path: int = ops.PATH_AUTO):
        """Host-to-host top-k: CPU query matrix in, CPU (values [Q,k] f32, ids [Q,k] i64) out.

        The batch is cut into `n_chunks` pieces that go round-robin over two side streams, each doing
        H2D copy -> fused top-k -> D2H copy, so the copies of one piece overlap the kernels of another
        (the B200 has separate copy engines for each direction). Pass pinned tensors (and pinned `out`)
        for truly asynchronous copies. Returns after enqueueing; the current stream waits on the side
        streams, so `torch.cuda.current_stream().synchronize()` makes the outputs valid.
        """
        if queries.is_cuda:
            raise ValueError("topk_host takes host tensors; use topk() for device-resident queries")
        if queries.dim() == 1:
            queries = queries.unsqueeze(0)
        Q = queries.shape[0]
        k = min(int(k), len(self))
        if out is None:
            out = (torch.empty(Q, k, dtype=torch.float32).pin_memory(), torch.empty(Q, k, dtype=torch.int64).pin_memory())
        vals_h, ids_h = out
        if Q == 0 or k < 1:
            return vals_h, ids_h
        if not hasattr(self, "_side_streams"):
            self._side_streams = [torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)]
        cur = torch.cuda.current_stream(self.device)
        n_chunks = max(1, min(n_chunks, (Q + 255) // 256))
        per = -(-Q // n_chunks)
        start = torch.cuda.Event()
        start.record(cur)
        for c in range(n_chunks):
            lo, hi = c * per, min(Q, (c + 1) * per)
            if lo >= hi:
                break
            st = self._side_streams[c % 2]
            st.wait_event(start)
            with torch.cuda.stream(st):
                qd = queries[lo:hi].to(self.device, non_blocking=True)
                if qd.dtype != self.dtype:
                    qd = qd.to(self.dtype)
                v, i = ops.cos_topk(qd, self.rows, k, cat_planes=self.planes, cat_inv_norms=self.inv_norms,
                                    row_offset=self.row_offset, path=path)
                vals_h[lo:hi].copy_(v, non_blocking=True)
                ids_h[lo:hi].copy_(i, non_blocking=True)
        for st in self._side_streams:
            cur.wait_stream(st)
        return vals_h, ids_h

    def topk(self, queries, k: int, *, exclude_mask: torch.Tensor | None = None, path: int = ops.PATH_AUTO):
        """(values [Q,k], global ids [Q,k]) of the k most cosine-similar rows per query."""
        q = to_device_matrix(queries, device=self.device, dtype=self.dtype)
        k = min(int(k), len(self))
        if k < 1:
            return (torch.empty(q.shape[0], 0, device=self.device), torch.empty(q.shape[0], 0, dtype=torch.int64, device=self.device))
        return ops.cos_topk(q, self.rows, k, cat_planes=self.planes, cat_inv_norms=self.inv_norms, exclude_mask=exclude_mask, row_offset=self.row_offset, path=path)
