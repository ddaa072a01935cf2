"""Class-sharded (PartialFC-style) ArcFace head over one process per GPU.

Replaces the reference's only multi-GPU strategy, `nn.DataParallel(model)`
(/root/reference/nlp_classifier_train_daodian_v2_dist.py:82-85), which replicates the whole C x D head
weight to every GPU each step, gathers the B x C logits to cuda:0 and reduces a C x D gradient back.
Here rank r owns classes [lo_r, hi_r) (weight rows, their gradient and optimiser state never leave the
rank) and the per-step traffic is three small collectives over NCCL / NVLink:

  1. all-gather the local embeddings + labels          (B x D fp32 + B int64, byte-packed: one collective)
  2. all-gather the per-row softmax statistics         (R x B x 20 bytes: max, sum-exp, label logit, argmax)
  3. reduce-scatter the embedding gradient partials    (B x D fp32)

The sequence itself (and its CUDA-graph replay, which is what keeps an 8-GPU step from being host-bound) lives
in `engine.py`; this module owns the shard bookkeeping and the reference-layout checkpoint helpers.

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

from . import engine


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


class ShardedArcMarginProduct(nn.Module):
    """`ArcMarginProduct` with the class dimension sharded over `process_group`.

    Every rank passes its local batch rows (equal counts on all ranks) and their labels (global class
    ids).  `loss` is the mean cross-entropy over the GLOBAL batch (identical on all ranks, like the
    reference's single CrossEntropyLoss over the gathered logits); `dx` is its exact gradient for the
    local rows; `weight.grad` is the gradient of the local class shard.
    """

    def __init__(self, in_feature=128, out_feature=10575, s=64.0, m=0.40, easy_margin=False, *, in_features=None,
                 out_features=None, process_group=None, kernels=None, use_cuda_graph=True, use_p2p=True,
                 precision="bf16", sample_rate=1.0, sample_seed=None, sparse_grad=False):
        super().__init__()
        if in_features is not None:
            in_feature = in_features
        if out_features is not None:
            out_feature = out_features
        if kernels is None:
            from . import ops as kernels  # the CUDA library; raises later if it was not built
        self.kernels = kernels
        self.use_cuda_graph = use_cuda_graph
        self.use_p2p = use_p2p      # exchanges through peer-mapped memory (p2p.py) instead of NCCL when available
        engine.precision_code(precision)
        self.precision = precision  # 'bf16' | 'bf16x3' (see ArcMarginProduct)
        # PartialFC-style class sampling of the LOCAL shard (see ArcMarginProduct): every rank draws its own negatives
        if not 0.0 < float(sample_rate) <= 1.0:
            raise ValueError("sample_rate must be in (0, 1], got %r" % (sample_rate,))
        self.sample_rate = float(sample_rate)
        self.sample_seed = None if sample_seed is None else int(sample_seed)   # offset by the rank below
        self.sparse_grad = bool(sparse_grad)
        self.process_group = process_group if process_group is not None else dist.group.WORLD
        self.world_size = dist.get_world_size(self.process_group)
        self.rank = dist.get_rank(self.process_group)
        if self.sample_seed is not None:
            self.sample_seed += 7919 * self.rank   # one seed for the job, a different stream per rank
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
        # ranks usually share one seed: drawing the shard from the default generator would start class i and class
        # per + i as duplicates, so every rank draws from its own stream derived from that seed
        gen = torch.Generator().manual_seed((torch.initial_seed() + 0x9E3779B1 * (self.rank + 1)) % (2 ** 63))
        with torch.no_grad():
            self.weight.uniform_(-bound, bound, generator=gen)
        self.cos_m = math.cos(m)
        self.sin_m = math.sin(m)
        self.th = math.cos(math.pi - m)
        self.mm = math.sin(math.pi - m) * m

    # torch.save(model) pickles the module's __dict__ (the reference checkpoints whole modules,
    # nlp_classifier_train.py:159): the kernel provider (a Python module) and the process group cannot be pickled
    # and are re-bound to the defaults on load
    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("kernels", None)
        state.pop("process_group", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        from . import ops as kernels

        self.kernels = kernels
        self.process_group = dist.group.WORLD if dist.is_initialized() else None

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
        cfg = engine.StepConfig(float(self.s), float(self.m), bool(self.easy_margin), self.class_lo, self.out_feature,
                                engine.precision_code(getattr(self, "precision", "bf16")))
        w = self.weight if self.weight.is_contiguous() else self.weight.contiguous()
        return engine.run_step(self, self.kernels, self.process_group, x, w, label, cfg, False)

    def forward(self, x, label):
        from .head import FusedLogits

        return FusedLogits(self, x, label)

    def logits(self, x, label=None):
        raise NotImplementedError("the sharded head never materialises the B x C logits; use loss() / predict()")

    @torch.no_grad()
    def predict_topk(self, x, k: int):
        """Global top-k classes by cosine for this rank's rows: every rank ranks the gathered batch against its class
        shard (fused top-k kernel), the per-rank lists are all-gathered and merged on the device.
        Returns (cosines fp32 [b_local, k], global class ids int64 [b_local, k])."""
        K, group = self.kernels, self.process_group
        x = x.to(torch.float32).contiguous()
        b_loc = x.shape[0]
        x_all = _all_gather_rows(x, group)
        if engine.precision_code(getattr(self, "precision", "bf16")):
            xhat, _, _ = K.normalize_cast3(x_all, 0)
            what, _, _ = K.normalize_cast3(self.weight.detach().contiguous(), 1)
        else:
            xhat, _, _ = K.normalize_cast(x_all)
            what, _, _ = K.normalize_cast(self.weight.detach().contiguous())
        v, i = K.cosine_topk(xhat, what, k, 1.0, self.class_lo)            # [B, k] over the local classes
        B = x_all.shape[0]
        vs = _all_gather_rows(v.unsqueeze(0), group)                       # [R, B, k]
        is_ = _all_gather_rows(i.unsqueeze(0), group)
        cand_v = vs.permute(1, 0, 2).reshape(B, -1).contiguous()
        cand_i = is_.permute(1, 0, 2).reshape(B, -1).contiguous()
        mv, mi = K.topk_merge(cand_v, cand_i, k)
        return (mv[self.rank * b_loc:(self.rank + 1) * b_loc].contiguous(),
                mi[self.rank * b_loc:(self.rank + 1) * b_loc].contiguous())

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
        self.invalidate_weight_cache()

    def last_sample_index(self):
        """Sorted LOCAL class ids (add class_lo for global ids) of the most recent sampled training step, or None."""
        return engine.last_sample(self)

    def invalidate_weight_cache(self) -> None:
        """See ArcMarginProduct.invalidate_weight_cache."""
        engine._W_CACHE.pop(self, None)
        engine.drop_plan(self)
