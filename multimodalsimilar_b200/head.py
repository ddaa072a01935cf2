"""Drop-in ArcFace head: the host-side mirror of the reference's `arcface.ArcMarginProduct`
(/root/reference/arcface.py:17-67) over the sm_100a kernels.

Same constructor, attributes, `weight` parameter, `forward(x, label)`, `forward_test(x)` and
`update_m(delta)`; what changes is that the B x C logit matrix is never built.  `forward` returns a
`FusedLogits` handle that answers exactly what the reference's training loops ask of `preds`
(nlp_classifier_train.py:120-123):

    loss = nn.CrossEntropyLoss()(preds, labels)   -> fused mean cross-entropy (autograd-connected)
    loss.backward()                                -> fused dX / dW
    torch.argmax(preds, dim=-1)                    -> fused argmax (first maximum, like torch)

and materialises real logits (through the logits kernel, detached) only if anything else touches it.
New code can call `head.loss(x, label) -> (loss, argmax)` directly.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import Parameter

from . import engine, ops


ArcFaceCEFunction = engine.ArcFaceCEFunction   # the autograd.Function behind `loss` (engine.run_step applies it)


class FusedLogits:
    """What `ArcMarginProduct.forward` returns: the logits of arcface.py:61 as a lazy handle.

    `F.cross_entropy(handle, label)` / `nn.CrossEntropyLoss()(handle, label)` and
    `torch.argmax(handle, dim=-1)` are answered from the fused kernels.  `handle.materialize()` (also
    triggered by any other torch function) runs the logits kernel and returns a detached fp32 [B, C]
    tensor -- the debug / small-C path.
    """

    def __init__(self, head, x, label):
        self._head = head
        self._x = x
        self._label = label
        self._fused = None
        self._dense = None
        self.shape = torch.Size((x.shape[0], head.out_feature))

    # -- fused answers
    def _run(self):
        if self._fused is None:
            self._fused = self._head.loss(self._x, self._label)
        return self._fused

    def cross_entropy(self):
        return self._run()[0]

    def argmax(self, dim=-1, keepdim=False):
        if dim not in (-1, 1):
            return self.materialize().argmax(dim=dim, keepdim=keepdim)
        a = self._run()[1]
        return a.unsqueeze(-1) if keepdim else a

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 2

    @property
    def device(self):
        return self._x.device

    @property
    def dtype(self):
        return torch.float32

    def materialize(self) -> torch.Tensor:
        if self._dense is None:
            self._dense = self._head.logits(self._x, self._label)
        return self._dense

    def detach(self):
        return self.materialize()

    def cpu(self):
        return self.materialize().cpu()

    def float(self):
        return self.materialize()

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __repr__(self):
        return "FusedLogits(shape=%s, device=%s)" % (tuple(self.shape), self.device)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is F.cross_entropy:
            return cls._cross_entropy(*args, **kwargs)
        if func in (torch.argmax, torch.Tensor.argmax) and isinstance(args[0], FusedLogits):
            return args[0].argmax(*args[1:], **kwargs)
        args = tuple(a.materialize() if isinstance(a, FusedLogits) else a for a in args)
        kwargs = {k: (v.materialize() if isinstance(v, FusedLogits) else v) for k, v in kwargs.items()}
        return func(*args, **kwargs)

    @staticmethod
    def _cross_entropy(input, target, weight=None, size_average=None, ignore_index=-100, reduce=None,
                       reduction="mean", label_smoothing=0.0):
        self = input
        fused_ok = (
            isinstance(self, FusedLogits) and weight is None and reduction == "mean" and label_smoothing == 0.0
            and size_average is None and reduce is None and torch.is_tensor(target)
            and target.dtype in (torch.int64, torch.int32) and target.numel() == self._label.numel()
            and (target.data_ptr() == self._label.data_ptr() or bool(torch.equal(target.view(-1).to(self._label.dtype),
                                                                                  self._label.view(-1))))
        )
        if fused_ok:
            return self.cross_entropy()
        if not isinstance(self, FusedLogits):
            return F.cross_entropy(input, target.materialize() if isinstance(target, FusedLogits) else target,
                                   weight=weight, size_average=size_average, ignore_index=ignore_index,
                                   reduce=reduce, reduction=reduction, label_smoothing=label_smoothing)
        if torch.is_grad_enabled() and (self._x.requires_grad or self._head.weight.requires_grad):
            raise NotImplementedError(
                "FusedLogits: only the reference's loss (mean cross-entropy against the labels given to forward, "
                "no class weights / smoothing) is differentiable without the B x C logits; got other options")
        dense = self.materialize() if isinstance(self, FusedLogits) else input
        return F.cross_entropy(dense, target, weight=weight, size_average=size_average, ignore_index=ignore_index,
                               reduce=reduce, reduction=reduction, label_smoothing=label_smoothing)


class ArcMarginProduct(nn.Module):
    """Additive angular margin head (arcface.py:17-67), B200-native.

    Constructor and attributes follow the reference: positional `(in_feature, out_feature, s, m,
    easy_margin)` (nlp_classifier.py:15, cv_classifier.py:38) or the keywords `in_feature=` /
    `out_feature=` (multimodal_classifier.py:22); `in_features=` / `out_features=` are accepted as aliases.
    """

    # class-level defaults: a module unpickled from a REFERENCE checkpoint (checkpoint.install_reference_shim)
    # has only the reference's attributes in its __dict__
    validate_labels = False
    use_cuda_graph = True
    precision = "bf16"
    sample_rate = 1.0
    sample_seed = None
    sparse_grad = False

    def __init__(self, in_feature=128, out_feature=10575, s=64.0, m=0.40, easy_margin=False, *,
                 in_features=None, out_features=None, validate_labels=False, use_cuda_graph=True, precision="bf16",
                 sample_rate=1.0, sample_seed=None, sparse_grad=False):
        super().__init__()
        if in_features is not None:
            in_feature = in_features
        if out_features is not None:
            out_feature = out_features
        self.in_feature = in_feature
        self.out_feature = out_feature
        self.s = s
        self.m = m
        self.weight = Parameter(torch.empty(out_feature, in_feature))
        nn.init.xavier_uniform_(self.weight)  # arcface.py:24-25
        self.easy_margin = easy_margin
        self.validate_labels = validate_labels
        self.use_cuda_graph = use_cuda_graph
        # 'bf16' (throughput: bf16 operands, fp32 accumulation) or 'bf16x3' (parity: hi/lo bf16 pairs, three
        # tensor-core products per cosine -- cosines within 1e-5 of the reference's fp32 head, ~3x the GEMM work)
        engine.precision_code(precision)
        self.precision = precision
        # PartialFC-style class sampling (the reference always trains every class): in training
        # mode each step runs on the batch's label classes plus uniformly drawn negatives, round(sample_rate * C) rows
        # in all -- GEMM work and weight traffic scale with the rate; the rows that were not drawn get a zero gradient
        # (dense `weight.grad` by default, a torch.sparse_coo gradient with sparse_grad=True).  Eval paths see all classes.
        if not 0.0 < float(sample_rate) <= 1.0:
            raise ValueError("sample_rate must be in (0, 1], got %r" % (sample_rate,))
        self.sample_rate = float(sample_rate)
        self.sample_seed = sample_seed
        self.sparse_grad = bool(sparse_grad)
        self.cos_m, self.sin_m, self.th, self.mm = ops.margin_constants(m)

    def update_m(self, delta):
        """arcface.py:35-42: accepted only while 1e-6 <= m + delta <= 1.0, silently ignored otherwise."""
        updated = self.m + delta
        if updated >= 1e-6 and updated <= 1.0:
            self.m = updated
            self.cos_m, self.sin_m, self.th, self.mm = ops.margin_constants(self.m)

    # ------------------------------------------------------------------ training path
    def _prep(self, x, label):
        if not x.is_cuda or not self.weight.is_cuda:
            raise RuntimeError("multimodalsimilar_b200.ArcMarginProduct runs on a B200 only (module and inputs must be "
                               "on a CUDA device); there is no CPU path")
        x = x.to(torch.float32).contiguous()
        label = label.reshape(-1).to(device=x.device, dtype=torch.int64).contiguous()
        if x.dim() != 2 or x.shape[1] != self.in_feature or label.numel() != x.shape[0]:
            raise ValueError("expected x [B, %d] and B labels, got %s and %s" % (self.in_feature, tuple(x.shape),
                                                                               tuple(label.shape)))
        return x, label

    def loss(self, x, label):
        """Fused head + mean softmax cross-entropy.  Returns (loss [], argmax int64 [B])."""
        x, label = self._prep(x, label)
        w = self.weight if self.weight.is_contiguous() else self.weight.contiguous()
        cfg = engine.StepConfig(float(self.s), float(self.m), bool(self.easy_margin), 0, w.shape[0],
                                engine.precision_code(self.precision))
        return engine.run_step(self, ops, None, x, w, label, cfg, bool(self.validate_labels))

    def forward(self, x, label):
        return FusedLogits(self, x, label)

    def _operands(self, x):
        """(xhat, inv_nx, what, inv_nw) in the head's precision mode: bf16 [., D], or the 3 D wide hi/lo rows."""
        w = self.weight.detach().contiguous()
        if engine.precision_code(self.precision):
            xhat, inv_nx, _ = ops.normalize_cast3(x, 0)
            what, inv_nw, _ = ops.normalize_cast3(w, 1)
        else:
            xhat, inv_nx, _ = ops.normalize_cast(x)
            what, inv_nw, _ = ops.normalize_cast(w)
        return xhat, inv_nx, what, inv_nw

    @torch.no_grad()
    def logits(self, x, label=None):
        """Materialised logits (detached): arcface.py:61 when `label` is given, s * cos otherwise."""
        if label is None:
            x = x.to(torch.float32).contiguous()
        else:
            x, label = self._prep(x, label)
        w = self.weight.detach().contiguous()
        xhat, inv_nx, what, inv_nw = self._operands(x)
        if label is None:
            return ops.logits(xhat, what, None, None, float(self.s))
        lm = ops.label_margin(x, w, inv_nx, inv_nw, label, 0, w.shape[0], float(self.s), float(self.m),
                              bool(self.easy_margin))
        return ops.logits(xhat, what, lm.z_label, lm.label_local, float(self.s))

    # ------------------------------------------------------------------ eval path
    @torch.no_grad()
    def forward_test(self, x):
        """arcface.py:65-67: bare cosines [B, C] (no margin, no scale), materialised like the reference."""
        x = x.to(torch.float32).contiguous()
        xhat, _, what, _ = self._operands(x)
        return ops.logits(xhat, what, None, None, 1.0)

    @torch.no_grad()
    def predict(self, x):
        """argmax_c cos[b, c] without materialising the cosines (what the eval loops do with
        forward_test's output, nlp_classifier_train.py:143-156).  Returns (argmax int64 [B], max cosine [B])."""
        x = x.to(torch.float32).contiguous()
        xhat, _, what, _ = self._operands(x)
        outs = [ops.forward_rows(xhat[lo:lo + ops.MAX_BATCH], what, None, 1.0, 0)     # MAX_BATCH rows per launch
                for lo in range(0, xhat.shape[0], ops.MAX_BATCH)]
        return torch.cat([o[2] for o in outs]), torch.cat([o[0] for o in outs])

    @torch.no_grad()
    def predict_topk(self, x, k: int):
        """The k best classes per row by cosine, descending, without the B x C matrix: (cosines fp32 [B, k],
        class ids int64 [B, k]).  What top-k over `forward_test(x)` returns (arcface.py:65-67)."""
        x = x.to(torch.float32).contiguous()
        xhat, _, what, _ = self._operands(x)
        return ops.cosine_topk(xhat, what, k)

    def last_sample_index(self):
        """Sorted class ids (int64) the most recent sampled training step ran on, or None (sample_rate == 1)."""
        return engine.last_sample(self)

    def invalidate_weight_cache(self) -> None:
        """Forget everything derived from `weight`: the normalised rows a fused optimiser step left behind and the
        captured CUDA graph.  Needed only after writes that bypass autograd's version counter (`weight.data.copy_`,
        raw-pointer writes): ordinary in-place ops, optimiser steps and `load_state_dict` are noticed by themselves."""
        engine._W_CACHE.pop(self, None)
        engine.drop_plan(self)

    def extra_repr(self):
        return ""
