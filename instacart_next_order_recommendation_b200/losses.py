"""``MultipleNegativesRankingLoss`` drop-in backed by the fused K3 forward/backward kernels.

Constructed exactly as the reference does (src/training/train_sbert.py:182-185):
``MultipleNegativesRankingLoss(model, scale=...)`` and called by the trainer as
``loss(sentence_features, labels)``. ``similarity_fct`` is accepted for signature
compatibility; only cosine similarity (the default, and what the reference uses) is fused.
"""

from __future__ import annotations

from typing import Any, Callable, Iterable

import torch
import torch.distributed as dist
from torch import nn

from . import ops


_KERNEL_DTYPES = (torch.float32, torch.bfloat16, torch.float16)  # read natively by icr_mnrl_* (no eager up-cast)


class _FusedMNRL(torch.autograd.Function):
    """loss = mean_i CE(scale * cos_sim(A, P)[i], i) with one kernel forward, one backward."""

    @staticmethod
    def forward(ctx, anchors: torch.Tensor, positives: torch.Tensor, scale: float):
        a = anchors.detach()
        p = positives.detach()
        ctx.scale = float(scale)
        ctx.fused = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        if ctx.fused:
            # a training step wants the gradients anyway: one library call computes the loss and d loss / d (A, P) for
            # dL/dloss = 1 (one prep and one host round trip instead of two); backward() only scales them
            loss, grads = ops.mnrl_forward_backward(a, p, scale)
            ctx.save_for_backward(grads)
            ctx.dim = a.shape[1]
            return loss
        loss, saved = ops.mnrl_forward(a, p, scale)
        ctx.save_for_backward(a, p, saved)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.fused:
            (grads,) = ctx.saved_tensors
            out = ops.mnrl_scale_grads(grads, grad_out)  # the only arithmetic of backward: one launch for both gradients
            if out.shape[2] != ctx.dim:  # embedding dim was padded to the vector width
                return out[0][:, : ctx.dim], out[1][:, : ctx.dim], None
            return out[0], out[1], None
        a, p, saved = ctx.saved_tensors
        ga, gp = ops.mnrl_backward(a, p, ctx.scale, saved, grad_out)
        return ga, gp, None


class _FusedMNRLGathered(torch.autograd.Function):
    """Cross-device in-batch negatives (sentence-transformers' ``gather_across_devices=True``).

    Every rank all-gathers the positives ([G*B, D]), scores its B local anchors against all of them with the
    rectangular kernels (label of anchor i = rank*B + i) and, in backward, reduce-scatters its [G*B, D] candidate
    gradient so that each rank ends with the sum over ranks of d loss_r / d (its own positives) — the gradient
    ``all_gather`` with autograd support produces upstream. `kernels` is a test seam (the gloo tests of this host
    logic plug a CPU function pair in); product code leaves it None.
    """

    @staticmethod
    def forward(ctx, anchors, positives, scale, group, kernels):
        a, p = anchors.detach(), positives.detach().contiguous()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        B = p.shape[0]
        gathered = torch.empty(world * B, p.shape[1], dtype=p.dtype, device=p.device)
        dist.all_gather_into_tensor(gathered, p, group=group)
        fwd, _ = kernels if kernels is not None else (ops.mnrl_forward_rect, ops.mnrl_backward_rect)
        loss, saved = fwd(a, gathered, scale, rank * B)
        ctx.save_for_backward(a, gathered, saved)
        ctx.scale, ctx.group, ctx.kernels, ctx.offset, ctx.B = float(scale), group, kernels, rank * B, B
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        a, gathered, saved = ctx.saved_tensors
        _, bwd = ctx.kernels if ctx.kernels is not None else (ops.mnrl_forward_rect, ops.mnrl_backward_rect)
        ga, gc = bwd(a, gathered, ctx.scale, ctx.offset, saved, grad_out)
        gp = torch.empty(ctx.B, gc.shape[1], dtype=gc.dtype, device=gc.device)
        dist.reduce_scatter_tensor(gp, gc.contiguous(), op=dist.ReduceOp.SUM, group=ctx.group)
        return ga, gp, None, None, None


def mnrl_loss_gathered(anchors: torch.Tensor, positives: torch.Tensor, scale: float = 20.0, group=None,
                       _kernels: tuple[Callable, Callable] | None = None) -> torch.Tensor:
    """This rank's MNRL over its anchors against the positives of EVERY rank of `group` (B_eff = G * B candidates)."""
    if not dist.is_initialized():
        raise RuntimeError("cross-device negatives need an initialised torch.distributed process group")
    if anchors.dtype != positives.dtype:
        positives = positives.to(anchors.dtype)
    if anchors.dtype not in _KERNEL_DTYPES:
        anchors, positives = anchors.float(), positives.float()
    return _FusedMNRLGathered.apply(anchors, positives, scale, group if group is not None else dist.group.WORLD, _kernels)


class MnrlStepGraph:
    """Loss + both gradients of MNRL for a FIXED batch shape as ONE CUDA-graph replay.

    A training step at the reference's batch size (256 x 384, src/training/train_sbert.py:182-185, configs/train.yaml:15)
    is 151 MFLOP: the kernels take ~20 us, the Python / ctypes / autograd glue around them several times that. The graph
    holds static input buffers; ``__call__`` copies the step's embeddings in, replays, and returns the static outputs
    (loss f32 scalar, d loss / d anchors, d loss / d positives - for dL/dloss = 1), which the next call overwrites.
    Use ``backward_into(anchors, positives)`` inside a training loop to feed the gradients to autograd.
    """

    def __init__(self, B: int, D: int, dtype: torch.dtype = torch.bfloat16, scale: float = 20.0, device: torch.device | None = None):
        if dtype not in _KERNEL_DTYPES:
            raise TypeError(f"MNRL kernels read float32, bfloat16 or float16 embeddings, got {dtype}")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.scale = float(scale)
        self.a = torch.zeros(B, D, dtype=dtype, device=dev)
        self.p = torch.zeros(B, D, dtype=dtype, device=dev)
        self.a[:, 0] = 1.0  # any non-degenerate rows for the warm-up
        self.p[:, 0] = 1.0
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside capture: lazy module load, one-time attribute setting
            ops.mnrl_forward_backward(self.a, self.p, self.scale)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.grads = ops.mnrl_forward_backward(self.a, self.p, self.scale)

    def __call__(self, anchors: torch.Tensor, positives: torch.Tensor):
        if anchors.shape != self.a.shape or positives.shape != self.p.shape:
            raise ValueError(f"this graph was captured for batches of shape {tuple(self.a.shape)}")
        self.a.copy_(anchors, non_blocking=True)
        self.p.copy_(positives, non_blocking=True)
        self.graph.replay()
        D = self.a.shape[1]
        return self.loss, self.grads[0][:, :D], self.grads[1][:, :D]

    def backward_into(self, anchors: torch.Tensor, positives: torch.Tensor) -> torch.Tensor:
        """One training step: returns the loss and pushes its gradients into the graph that produced the embeddings."""
        loss, ga, gp = self(anchors.detach(), positives.detach())
        torch.autograd.backward([anchors, positives], [ga.to(anchors.dtype), gp.to(positives.dtype)])
        return loss


def mnrl_step_graph(B: int, D: int, dtype: torch.dtype = torch.bfloat16, scale: float = 20.0, device=None) -> MnrlStepGraph:
    return MnrlStepGraph(B, D, dtype, scale, device)


def mnrl_loss(anchors: torch.Tensor, positives: torch.Tensor, scale: float = 20.0) -> torch.Tensor:
    """Functional form on embeddings [B, D]: float32, bfloat16 or float16 CUDA tensors. float16 is what the reference's
    training produces under ``fp16=True`` autocast (src/training/train_sbert.py:210,232); the kernels read it natively and
    return float16 gradients (internal math is fp32 for every input type)."""
    if anchors.dtype != positives.dtype:
        positives = positives.to(anchors.dtype)
    if anchors.dtype not in _KERNEL_DTYPES:
        anchors, positives = anchors.float(), positives.float()
    return _FusedMNRL.apply(anchors, positives, scale)


class MultipleNegativesRankingLoss(nn.Module):
    def __init__(self, model, scale: float = 20.0, similarity_fct=None, gather_across_devices: bool = False) -> None:
        super().__init__()
        self.model = model
        self.scale = scale
        if similarity_fct is not None and getattr(similarity_fct, "__name__", "") != "cos_sim":
            raise NotImplementedError("the fused kernel implements cosine similarity (the reference's choice) only")
        # upstream option; the reference leaves it off (train_sbert.py:184-185). On: negatives come from every rank's batch.
        self.gather_across_devices = bool(gather_across_devices)

    def forward(self, sentence_features: Iterable[dict[str, torch.Tensor]], labels: torch.Tensor | None = None) -> torch.Tensor:
        embeddings = [self.model(f)["sentence_embedding"] for f in sentence_features]
        return self.compute_loss_from_embeddings(embeddings, labels)

    def compute_loss_from_embeddings(self, embeddings: list[torch.Tensor], labels: torch.Tensor | None = None) -> torch.Tensor:
        if len(embeddings) != 2:
            # hard-negative columns would make the candidate matrix [k*B, D]; the reference feeds (anchor, positive) pairs only
            raise NotImplementedError("fused MNRL expects (anchor, positive) pairs, as prepared by src/data/prepare_instacart_sbert.py")
        if self.gather_across_devices and dist.is_initialized() and dist.get_world_size() > 1:
            return mnrl_loss_gathered(embeddings[0], embeddings[1], self.scale)
        return mnrl_loss(embeddings[0], embeddings[1], self.scale)

    def get_config_dict(self) -> dict[str, Any]:
        return {"scale": self.scale, "similarity_fct": "cos_sim", "gather_across_devices": self.gather_across_devices}
