"""``MultipleNegativesRankingLoss`` drop-in backed by the fused K3 forward/backward kernels.

Constructed exactly as the reference does (src/training/train_sbert.py:182-185):
``MultipleNegativesRankingLoss(model, scale=...)`` and called by the trainer as
``loss(sentence_features, labels)``. ``similarity_fct`` is accepted for signature
compatibility; only cosine similarity (the default, and what the reference uses) is fused.
"""

from __future__ import annotations

from typing import Any, Iterable

import torch
from torch import nn

from . import ops


class _FusedMNRL(torch.autograd.Function):
    """loss = mean_i CE(scale * cos_sim(A, P)[i], i) with one kernel forward, one backward."""

    @staticmethod
    def forward(ctx, anchors: torch.Tensor, positives: torch.Tensor, scale: float):
        a = anchors.detach()
        p = positives.detach()
        loss, saved = ops.mnrl_forward(a, p, scale)
        ctx.save_for_backward(a, p, saved)
        ctx.scale = float(scale)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        a, p, saved = ctx.saved_tensors
        ga, gp = ops.mnrl_backward(a, p, ctx.scale, saved, grad_out)
        return ga, gp, None


def mnrl_loss(anchors: torch.Tensor, positives: torch.Tensor, scale: float = 20.0) -> torch.Tensor:
    """Functional form on embeddings [B, D] (float32 or bfloat16 CUDA tensors)."""
    if anchors.dtype != positives.dtype:
        positives = positives.to(anchors.dtype)
    if anchors.dtype not in (torch.float32, torch.bfloat16):
        anchors, positives = anchors.float(), positives.float()
    return _FusedMNRL.apply(anchors, positives, scale)


class MultipleNegativesRankingLoss(nn.Module):
    def __init__(self, model, scale: float = 20.0, similarity_fct=None, gather_across_devices: bool = False) -> None:
        super().__init__()
        self.model = model
        self.scale = scale
        if similarity_fct is not None and getattr(similarity_fct, "__name__", "") != "cos_sim":
            raise NotImplementedError("the fused kernel implements cosine similarity (the reference's choice) only")
        if gather_across_devices:
            raise NotImplementedError("cross-device negatives are not part of the reference's configuration (train_sbert.py:184-185)")

    def forward(self, sentence_features: Iterable[dict[str, torch.Tensor]], labels: torch.Tensor | None = None) -> torch.Tensor:
        embeddings = [self.model(f)["sentence_embedding"] for f in sentence_features]
        return self.compute_loss_from_embeddings(embeddings, labels)

    def compute_loss_from_embeddings(self, embeddings: list[torch.Tensor], labels: torch.Tensor | None = None) -> torch.Tensor:
        if len(embeddings) != 2:
            # hard-negative columns would make the candidate matrix [k*B, D]; the reference feeds (anchor, positive) pairs only
            raise NotImplementedError("fused MNRL expects (anchor, positive) pairs, as prepared by src/data/prepare_instacart_sbert.py")
        return mnrl_loss(embeddings[0], embeddings[1], self.scale)

    def get_config_dict(self) -> dict[str, Any]:
        return {"scale": self.scale, "similarity_fct": "cos_sim"}
