"""Drop-ins for ``sentence_transformers.util.cos_sim`` and the fused cosine top-k.

Reference call sites replaced (the import line is the only change a caller makes):
  src/inference/serve_recommendations.py:30,214,250 ; src/baselines/content_based.py:13,54 ;
  scripts/compare_untrained_vs_trained.py:22,74.
"""

from __future__ import annotations

import numpy as np
import torch

from . import ops


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "instacart_next_order_recommendation_b200 needs a CUDA device (B200, sm_100a); "
            "there is no CPU fallback for the retrieval path"
        )
    return torch.device("cuda", torch.cuda.current_device())


# Opt-in (ICR_CACHE_HOST_OPERANDS=1): keep the device copy of LARGE host operands between calls. With the import swap of
# INTEGRATION.md §1 the reference hands its numpy catalog to cos_sim on every request (serve_recommendations.py:214); the
# 76 MB upload (~1.4 ms over PCIe) then dwarfs the kernel. Entries are keyed by the array's buffer address, shape, strides
# and dtype and validated by a sample of its values, so an array that was rewritten in place is uploaded again (a change
# confined to unsampled elements would go unnoticed: hence opt-in, for catalogs that are immutable once loaded).
_HOST_CACHE: dict = {}
_HOST_CACHE_MIN_BYTES = 1 << 20
_HOST_CACHE_MAX_ENTRIES = 4


def _host_fingerprint(a: np.ndarray):
    flat = a.reshape(-1)
    step = max(1, flat.size // 1024)
    return (float(flat[::step].astype(np.float64).sum()), float(flat[:64].astype(np.float64).sum()), float(flat[-64:].astype(np.float64).sum()))


def _cached_upload(a: np.ndarray, device: torch.device, dtype: torch.dtype | None):
    import os

    if os.environ.get("ICR_CACHE_HOST_OPERANDS", "0") != "1" or a.nbytes < _HOST_CACHE_MIN_BYTES or not a.flags.c_contiguous:
        return None
    key = (a.ctypes.data, a.shape, a.strides, str(a.dtype), str(device), str(dtype))
    fp = _host_fingerprint(a)
    hit = _HOST_CACHE.get(key)
    if hit is not None and hit[0] == fp:
        return hit[1]
    t = torch.as_tensor(a)
    if not t.is_floating_point() or t.dtype in (torch.float64, torch.float16):
        t = t.to(torch.float32)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    t = t.to(device)
    if len(_HOST_CACHE) >= _HOST_CACHE_MAX_ENTRIES:
        _HOST_CACHE.pop(next(iter(_HOST_CACHE)))
    _HOST_CACHE[key] = (fp, t)
    return t


def to_device_matrix(x, device: torch.device | None = None, dtype: torch.dtype | None = None) -> torch.Tensor:
    """Accepts what ST's cos_sim accepts (Tensor | ndarray | list; 1-D -> [1, D]) and uploads it."""
    if isinstance(x, np.ndarray) and x.ndim == 2 and x.nbytes >= _HOST_CACHE_MIN_BYTES:
        cached = _cached_upload(x, device if device is not None else default_device(), dtype)
        if cached is not None:
            return cached
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.dim() != 2:
        raise ValueError(f"expected a vector or a [rows, dim] matrix, got shape {tuple(t.shape)}")
    if not t.is_floating_point() or t.dtype in (torch.float64, torch.float16):
        t = t.to(torch.float32)
    if device is None:
        device = t.device if t.is_cuda else default_device()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.device != device:
        if not t.is_cuda and t.numel() * t.element_size() >= (1 << 20):
            t = t.pin_memory() if not t.is_pinned() else t
        t = t.to(device, non_blocking=True)
    return t


def cos_sim(a, b) -> torch.Tensor:
    """cos_sim(a, b)[i, j] = cosine(a[i], b[j]) as an f32 CUDA tensor [len(a), len(b)].

    Same signature and conversions as ``sentence_transformers.util.cos_sim`` (5.2.2); both
    operands are L2-normalised with eps 1e-12 inside the kernel.
    """
    bt = to_device_matrix(b)
    at = to_device_matrix(a, device=bt.device)
    return ops.cos_sim_dense(at, bt)


def cos_topk(queries, catalog, k: int, *, sorted: bool = True, **kw):
    """(values f32 [Q,k], indices int64 [Q,k]) == torch.topk(cos_sim(queries, catalog), k, dim=1).

    The [Q, N] score matrix is never written to HBM. Results are always sorted descending with
    ties broken by the lower row (a deterministic instance of what ``sorted=False`` permits).
    """
    del sorted
    ct = to_device_matrix(catalog)
    qt = to_device_matrix(queries, device=ct.device)
    k = int(k)
    if k < 1:
        raise ValueError("k must be >= 1")
    k_eff = min(k, ct.shape[0])
    if k_eff == 0:
        return (torch.empty(qt.shape[0], 0, device=ct.device), torch.empty(qt.shape[0], 0, dtype=torch.int64, device=ct.device))
    return ops.cos_topk(qt, ct, k_eff, **kw)
