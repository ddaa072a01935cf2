"""Class-sharded (PartialFC-style) ArcFace head over one process per GPU.

Replaces the reference's only multi-GPU strategy, `nn.DataParallel(model)`
(/root/reference/nlp_classifier_train_daodian_v2_dist.py:82-85), which replicates the whole C x D head
weight to every GPU each step, gathers the B x C logits to cuda:0 and reduces a C x D gradient back.
Here rank r owns classes [lo_r, hi_r) (weight rows, their gradient and optimiser state never leave the
rank) and the per-step traffic is three small collectives over NCCL / NVLink:

  1. all-gather the local embeddings + labels          (B x D fp32 + B int64)
  2. all-gather the per-row softmax statistics         (R x B x 20 bytes: max, sum-exp, label logit, argmax)
  3. reduce-scatter the embedding gradient partials    (B x D fp32)

The compute between them is the same kernel sequence as the single-GPU head, run on the local class
shard for the whole (gathered) batch.  `kernels` is the object providing that sequence; it defaults to
`multimodalsimilar_b200.ops` (the CUDA library).  Tests inject a stand-in to exercise this
choreography with the gloo backend on CPU; the product has no such stand-in.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist
from torch import nn
from torch.nn import Parameter


def shard_range(num_classes: int, world_size: int, rank: int):
    """Contiguous class range [lo, hi) owned by `rank`: ceil(C / R) classes per rank, last one ragged."""
    per = (num_classes + world_size - 1) // world_size
    lo = min(num_classes, rank * per)
    hi = min(num_classes, lo + per)
    return lo, hi


def _all_gather_rows(t: torch.Tensor, group) -> torch.Tensor:
    world = dist.get_world_size(group)
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


def _reduce_scatter_rows(full: torch.Tensor, group) -> torch.Tensor:
    """Sum `full` [R * n, ...] over ranks and return this rank's n rows."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = full.shape[0] // world
    if dist.get_backend(group) == "gloo":  # gloo has no reduce-scatter; used by the CPU tests only
        buf = full.clone()
        dist.all_reduce(buf, group=group)
        return buf[rank * n:(rank + 1) * n].contiguous()
    out = torch.empty((n,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    dist.reduce_scatter_tensor(out, full.contiguous(), group=group)
    return out


def _pack_rows(rmax, rsum, z, rarg) -> torch.Tensor:
    """One byte buffer per rank so the statistics exchange is a single collective."""
    f = torch.stack([rmax, rsum, z]).contiguous().view(torch.uint8).reshape(-1)
    a = rarg.contiguous().view(torch.uint8).reshape(-1)
    return torch.cat([f, a]).unsqueeze(0)


def _unpack_rows(buf: torch.Tensor, B: int):
    R = buf.shape[0]
    f = buf[:, : 12 * B].contiguous().view(torch.float32).reshape(R, 3, B)
    a = buf[:, 12 * B:].contiguous().view(torch.int64).reshape(R, B)
    return f[:, 0].contiguous(), f[:, 1].contiguous(), f[:, 2].contiguous(), a.contiguous()


class ShardedArcFaceCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, w_shard, label_local, head):
        K, group = head.kernels, head.process_group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        b_loc = x_local.shape[0]
        B = world * b_loc
        x_all = _all_gather_rows(x_local, group)
        y_all = _all_gather_rows(label_local, group)
        xhat, inv_nx, xhat_t = K.normalize_cast(x_all, want_transpose=True)
        lm = K.label_margin(x_all, w_shard, inv_nx, None, y_all, head.class_lo, head.out_feature, float(head.s),
                            float(head.m), bool(head.easy_margin))
        what, inv_nw, rmax, rsum, rarg = K.forward_rows_fused(xhat, w_shard, lm.label_local, float(head.s),
                                                              head.class_lo)
        packed = _all_gather_rows(_pack_rows(rmax, rsum, lm.z_label, rarg), group)
        rows_max, rows_sum, rows_z, rows_arg = _unpack_rows(packed, B)
        lse, argmax, _z, omp, loss = K.finalize_rows(rows_max, rows_sum, rows_arg, rows_z, y_all)
        ctx.save_for_backward(x_local, inv_nx, xhat, xhat_t, what, inv_nw, lse, omp, lm.dphi, lm.label_local)
        ctx.head = head
        ctx.B = B
        argmax_local = argmax[rank * b_loc:(rank + 1) * b_loc].contiguous()
        ctx.mark_non_differentiable(argmax_local)
        return loss, argmax_local

    @staticmethod
    def backward(ctx, grad_loss, _grad_argmax):
        x_local, inv_nx, xhat, xhat_t, what, inv_nw, lse, omp, dphi, label_local = ctx.saved_tensors
        head = ctx.head
        K, group = head.kernels, head.process_group
        rank = dist.get_rank(group)
        b_loc = x_local.shape[0]
        g = grad_loss.to(torch.float32).contiguous()
        dxhat_part, dw = K.backward(xhat, xhat_t, what, inv_nw, lse, omp, dphi, label_local, float(head.s),
                                    1.0 / ctx.B, grad_loss_dev=g)
        dx = None
        if ctx.needs_input_grad[0]:
            dxhat_loc = _reduce_scatter_rows(dxhat_part, group)
            dx = K.normalize_bwd_x(x_local, inv_nx[rank * b_loc:(rank + 1) * b_loc].contiguous(), dxhat_loc)
        return dx, (dw if ctx.needs_input_grad[1] else None), None, None


class ShardedArcMarginProduct(nn.Module):
    """`ArcMarginProduct` with the class dimension sharded over `process_group`.

    Every rank passes its local batch rows (equal counts on all ranks) and their labels (global class
    ids).  `loss` is the mean cross-entropy over the GLOBAL batch (identical on all ranks, like the
    reference's single CrossEntropyLoss over the gathered logits); `dx` is its exact gradient for the
    local rows; `weight.grad` is the gradient of the local class shard.
    """

    def __init__(self, in_feature=128, out_feature=10575, s=64.0, m=0.40, easy_margin=False, *, in_features=None,
                 out_features=None, process_group=None, kernels=None):
        super().__init__()
        if in_features is not None:
            in_feature = in_features
        if out_features is not None:
            out_feature = out_features
        if kernels is None:
            from . import ops as kernels  # the CUDA library; raises later if it was not built
        self.kernels = kernels
        self.process_group = process_group if process_group is not None else dist.group.WORLD
        self.world_size = dist.get_world_size(self.process_group)
        self.rank = dist.get_rank(self.process_group)
        if out_feature < self.world_size:
            raise ValueError("need at least one class per rank")
        self.in_feature = in_feature
        self.out_feature = out_feature
        self.s = s
        self.m = m
        self.easy_margin = easy_margin
        self.class_lo, self.class_hi = shard_range(out_feature, self.world_size, self.rank)
        if self.class_hi <= self.class_lo:
            raise ValueError("rank %d owns no classes (C=%d, world=%d)" % (self.rank, out_feature, self.world_size))
        self.weight = Parameter(torch.empty(self.class_hi - self.class_lo, in_feature))
        bound = math.sqrt(6.0 / (out_feature + in_feature))  # xavier_uniform_ of the FULL matrix (arcface.py:25)
        nn.init.uniform_(self.weight, -bound, bound)
        self.cos_m = math.cos(m)
        self.sin_m = math.sin(m)
        self.th = math.cos(math.pi - m)
        self.mm = math.sin(math.pi - m) * m

    def update_m(self, delta):
        updated = self.m + delta
        if updated >= 1e-6 and updated <= 1.0:
            self.m = updated
            self.cos_m = math.cos(self.m)
            self.sin_m = math.sin(self.m)
            self.th = math.cos(math.pi - self.m)
            self.mm = math.sin(math.pi - self.m) * self.m

    def loss(self, x, label):
        x = x.to(torch.float32).contiguous()
        label = label.reshape(-1).to(device=x.device, dtype=torch.int64).contiguous()
        return ShardedArcFaceCE.apply(x, self.weight.contiguous(), label, self)

    def forward(self, x, label):
        from .head import FusedLogits

        return FusedLogits(self, x, label)

    def logits(self, x, label=None):
        raise NotImplementedError("the sharded head never materialises the B x C logits; use loss() / predict()")

    # ------------------------------------------------------------------ checkpoint compatibility
    @torch.no_grad()
    def gather_weight(self) -> torch.Tensor:
        """Full [C, D] fp32 weight in the reference's row order (state_dict key `weight`, arcface.py:24)."""
        per = (self.out_feature + self.world_size - 1) // self.world_size
        pad = torch.zeros((per, self.in_feature), dtype=self.weight.dtype, device=self.weight.device)
        pad[: self.weight.shape[0]] = self.weight
        full = _all_gather_rows(pad, self.process_group)
        return full[: self.out_feature].contiguous()

    @torch.no_grad()
    def load_full_weight(self, weight: torch.Tensor) -> None:
        """Take this rank's rows from a reference-layout [C, D] weight (e.g. a reference state_dict)."""
        if tuple(weight.shape) != (self.out_feature, self.in_feature):
            raise ValueError("expected weight of shape %s" % ((self.out_feature, self.in_feature),))
        self.weight.copy_(weight[self.class_lo:self.class_hi].to(self.weight.device, self.weight.dtype))
