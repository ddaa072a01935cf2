"""Fused head optimiser.

`FusedHeadAdamW` is `torch.optim.AdamW` for ArcFace head weights -- the reference trains the head with its own
optimiser, `AdamW(model.classifier.parameters(), lr=1e-2)` (nlp_classifier_train.py:94-97, multimodal_classifier_train.py:
161-164) -- whose step is ONE kernel per weight (`ops.adamw_normalize`) that also writes the next forward's
L2-normalised bf16 rows and inverse norms.  A head registered with `attach(head)` then skips the weight half of K1:
its forward runs the cosine GEMM straight from those rows (`engine.forward_eager(..., w_cache=...)`), as long as
nothing else has modified the weight since (tensor version check).  After the fusion the step's HBM traffic per
weight element drops from 28 (AdamW) + 6 (K1) to 30 bytes, and the forward kernel from 6 to 2.

Same hyper-parameters, update rule, `param_groups` and state names (`step`, `exp_avg`, `exp_avg_sq`) as
torch.optim.AdamW (amsgrad / maximize / capturable are not supported), so LR schedulers and optimiser checkpoints
interoperate.
"""
from __future__ import annotations

import torch

from . import engine, ops


class FusedHeadAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, heads=()):
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._heads = {}
        for h in heads:
            self.attach(h)

    @classmethod
    def for_head(cls, head, **kw):
        """Optimiser over `head.parameters()` that also feeds the head's normalised-weight cache."""
        return cls(head.parameters(), heads=(head,), **kw)

    def attach(self, head) -> None:
        """Register a head (ArcMarginProduct / ShardedArcMarginProduct) whose `weight` this optimiser updates."""
        self._heads[id(head.weight)] = head

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.dim() != 2 or p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                    raise RuntimeError("FusedHeadAdamW updates contiguous fp32 [C, D] CUDA head weights only")
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                head = self._heads.get(id(p))
                what = inv_nw = None
                if head is not None:
                    cache = engine._W_CACHE.get(head)
                    if cache is not None and cache[0].shape == p.shape and cache[0].device == p.device:
                        what, inv_nw = cache[0], cache[1]
                    else:
                        what = torch.empty(p.shape, dtype=torch.bfloat16, device=p.device)
                        inv_nw = torch.empty(p.shape[0], dtype=torch.float32, device=p.device)
                ops.adamw_normalize(p, g, st["exp_avg"], st["exp_avg_sq"], float(group["lr"]), float(beta1), float(beta2),
                                    float(group["eps"]), float(group["weight_decay"]), int(st["step"].item()), what, inv_nw)
                torch.autograd.graph.increment_version(p)  # the kernel wrote through the raw pointer
                if head is not None:
                    torch.autograd.graph.increment_version(what)
                    # valid while the weight is exactly what this step left behind
                    engine._W_CACHE[head] = (what, inv_nw, p._version, p.data_ptr())
        return loss
