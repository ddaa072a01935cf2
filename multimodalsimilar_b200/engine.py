"""Step engine of the ArcFace head: the kernel sequence of one forward / backward, run either eagerly or as ONE
replayed CUDA graph.

One sequence serves both public modules: `ArcMarginProduct` (group = None, the whole class range on one GPU)
and `ShardedArcMarginProduct` (class shard per rank, three exchanges per step: all-gather of the packed local
embeddings + labels, all-gather of the packed per-row statistics, reduce-scatter of the embedding gradient;
reference counterpart: nn.DataParallel at nlp_classifier_train_daodian_v2_dist.py:85).  The exchanges run over
peer-mapped memory (p2p.py / csrc/p2p.cu) when the head enables it and the group supports it, else over NCCL.

Why graphs: at 8 GPUs one rank's kernels for the north-star shape take ~0.35 ms, less than the host needs to
issue ~60 small torch / ctypes calls, so the eager step is host-bound (0.84 ms measured).  `GraphedStep` captures
forward + backward (kernels and exchanges) once per signature -- every buffer, including `what`, the softmax
statistics and dW, lives in the graph's private pool -- and replays it: one launch per step (0.45 ms measured).
The captured kernels are the same C-ABI calls the eager path makes (`ops`), on the capture stream.

Also here: the weight cache a fused optimiser step leaves behind (`_W_CACHE`, optim.py), consumed by `run_step`.
"""
from __future__ import annotations

import warnings
import weakref
from dataclasses import dataclass
from typing import Any

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class StepConfig:
    s: float
    m: float
    easy_margin: bool
    class_lo: int     # first global class id of the local weight rows
    c_total: int      # classes over all ranks
    prec: int = 0     # PRECISIONS[...]: 0 = bf16 operands, 1 = bf16x3 (hi/lo pairs, three products per cosine)


PRECISIONS = {"bf16": 0, "bf16x3": 1}   # the `precision` keyword of the modules -> ARCFACE_B200_PREC_*


def precision_code(name) -> int:
    try:
        return PRECISIONS[name]
    except KeyError:
        raise ValueError("precision must be one of %s, got %r" % (sorted(PRECISIONS), name)) from None


@dataclass
class FwdState:
    """Everything the backward needs; no B x C tensor."""
    loss: torch.Tensor
    argmax_local: torch.Tensor
    bad_flag: torch.Tensor
    B: int
    inv_nx: torch.Tensor
    xhat: torch.Tensor
    xhat_t: torch.Tensor
    what: torch.Tensor
    inv_nw: torch.Tensor
    lse: torch.Tensor
    omp: torch.Tensor
    dphi: torch.Tensor
    label_local: torch.Tensor
    argmax_all: Any = None   # argmax of every row of the global batch (argmax_local is a view of it)
    sample: Any = None       # class sampling: int64 [S] sorted local class ids the step ran on (what / inv_nw are theirs)
    c_local: int = 0         # class rows of the full local weight (the shape of dW)


class LabelGuard:
    """Out-of-range labels without a host sync in the step.

    The reference raises from `one_hot.scatter_` (arcface.py:59) when a label is outside [0, C).  Here
    `label_margin` raises a device flag instead; reading it synchronously would stall every step (and is impossible
    inside a replayed CUDA graph), so the label kernel raises the flag directly in pinned host memory (device-visible
    under unified addressing: no copy node, no fill) and the host examines it at the START of the next call of the
    same head: a bad label raises IndexError one step late instead of training on silently.
    `validate_labels=True` keeps the synchronous check (eager launches, a device flag, one sync per step)."""

    def __init__(self):
        self.host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.event = torch.cuda.Event()
        self.armed = False

    def mark(self) -> None:
        self.event.record()
        self.armed = True

    def check(self, c_total: int) -> None:
        if self.armed and self.event.query():
            self.armed = False
            if int(self.host[0]) != 0:
                self.host.zero_()
                raise IndexError("ArcMarginProduct: a label of an earlier step was outside [0, %d) (detected one step "
                                 "late; construct the head with validate_labels=True to fail in the same step)" % c_total)


def _world(group) -> int:
    return 1 if group is None else dist.get_world_size(group)


def _rank(group) -> int:
    return 0 if group is None else dist.get_rank(group)


def _all_gather_bytes(buf: torch.Tensor, group) -> torch.Tensor:
    """buf uint8 [n] -> uint8 [R, n]."""
    out = torch.empty(_world(group) * buf.numel(), dtype=torch.uint8, device=buf.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out.view(_world(group), buf.numel())


def gather_batch(x_local: torch.Tensor, y_local: torch.Tensor, group, packed=None, peer=None):
    """All ranks' embeddings [B, D] fp32 and labels [B] int64 in rank order: ONE collective over the
    byte-packed (x | labels) of every rank.  `packed`: the caller already holds x_local / y_local as views of one
    uint8 buffer laid out that way (the graph's static input), so nothing has to be concatenated."""
    R = _world(group)
    if R == 1:
        return x_local, y_local
    b, D = x_local.shape
    xb = b * D * 4
    if packed is None:
        packed = torch.cat([x_local.reshape(-1).view(torch.uint8), y_local.view(torch.uint8)])
    if peer is not None and packed.numel() % 16 == 0 and xb % 16 == 0:
        # stores over NVLink + flags (csrc/p2p.cu); x rows and labels land as two contiguous rank-ordered regions
        xa, ya = peer.all_gather_split(0, packed, xb)
        # x is consumed (K1, label margin) before this rank raises its flag of the NEXT exchange, so no peer can be
        # overwriting it (csrc/p2p.cu); the labels are read again after that exchange (finalize), hence their copy
        return xa.view(torch.float32).view(R * b, D), ya.view(torch.int64).clone()
    allp = _all_gather_bytes(packed, group)
    x_all = allp[:, :xb].contiguous().view(torch.float32).reshape(R * b, D)
    y_all = allp[:, xb:].contiguous().view(torch.int64).reshape(R * b)
    return x_all, y_all


def exchange_rows(rmax, rsum, z_label, rarg, group):
    """Per-rank statistics [B] -> [R, B] each, in ONE collective (20 bytes per row per rank)."""
    R = _world(group)
    B = rmax.shape[0]
    if R == 1:
        return rmax.view(1, B), rsum.view(1, B), z_label.view(1, B), rarg.view(1, B)
    packed = torch.cat([rarg.view(torch.uint8), rmax.view(torch.uint8), rsum.view(torch.uint8),
                        z_label.view(torch.uint8)])
    allp = _all_gather_bytes(packed, group)
    a = allp[:, : 8 * B].contiguous().view(torch.int64).reshape(R, B)
    f = allp[:, 8 * B:].contiguous().view(torch.float32).reshape(R, 3, B)
    return f[:, 0].contiguous(), f[:, 1].contiguous(), f[:, 2].contiguous(), a


def reduce_scatter_rows(full: torch.Tensor, group) -> torch.Tensor:
    """Sum `full` [R * n, D] over ranks and return this rank's n rows."""
    R = _world(group)
    if R == 1:
        return full
    rank = _rank(group)
    n = full.shape[0] // R
    if dist.get_backend(group) == "gloo":  # gloo has no reduce-scatter; used by the CPU tests only
        buf = full.clone()
        dist.all_reduce(buf, group=group)
        return buf[rank * n:(rank + 1) * n].contiguous()
    out = torch.empty((n,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    dist.reduce_scatter_tensor(out, full.contiguous(), group=group)
    return out


def sample_classes(label_local: torch.Tensor, c_local: int, num_sample: int, generator=None) -> torch.Tensor:
    """Sorted int64 [S] class ids of this step's sub-matrix: every class that is a label of the batch plus uniformly
    drawn negatives.  The rule of PartialFC (insightface recognition/arcface_torch/partial_fc_v2.py `sample`: random
    scores, positives forced to the top, top-k, sort; restated in oracle/arcface_numpy.py:partial_fc_sample), with
    S = max(num_sample, min(B, c_local)) fixed on the host, so that the positives always fit and no device value has to
    be read back (PartialFC falls back to "positives only" when they outnumber num_sample).
    `label_local`: int [B], label - class_lo or -1 for labels of other ranks."""
    B = label_local.numel()
    S = min(c_local, max(int(num_sample), min(B, c_local)))
    dev = label_local.device
    perm = torch.rand(c_local + 1, device=dev, generator=generator)   # slot c_local swallows the rows without a label here
    idx = label_local.to(torch.int64)
    idx = torch.where(idx >= 0, idx, torch.full_like(idx, c_local))
    perm.scatter_(0, idx, 2.0)
    return torch.topk(perm[:c_local], S).indices.sort().values


def remap_labels(label_local: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """Position of every local label inside the sorted sample `index` (int32 [B]; -1 stays -1)."""
    lab = label_local.to(torch.int64)
    pos = torch.searchsorted(index, lab.clamp(min=0))
    return torch.where(lab >= 0, pos, torch.full_like(pos, -1)).to(torch.int32)


# Rows of the (gathered) batch one GEMM launch takes.  The CTA-pair kernels and the single-launch backward cover B <= 1024
# (include/arcface_b200.h); a larger global batch -- the PartialFC regime, per-rank batch x ranks -- runs K2 and K3 once
# per chunk of <= 1024 rows against the same normalised weights.  Rows are independent up to the mean of the loss and the
# sum over rows in dW, so the row kernels, the three exchanges and the softmax statistics stay one pass over the whole
# batch and only dW is accumulated over the chunks.
BATCH_CHUNK = 1024


def batch_chunks(B: int, prec: int = 0):
    """[(b0, b1)] row ranges of the GEMM launches: one range for B <= BATCH_CHUNK (and in the bf16x3 mode, whose
    transposed operand is laid out per launch), else equal chunks whose starts are multiples of 64."""
    if prec or B <= BATCH_CHUNK:
        return [(0, B)]
    n = -(-B // BATCH_CHUNK)
    size = -(-(-(-B // n)) // 64) * 64
    return [(b0, min(B, b0 + size)) for b0 in range(0, B, size)]


def _rows(K, xhat, w, label_local, cfg, w_cache, out=None, sample=None):
    """(what, inv_nw, max, sum, arg): K1 (w) fused into K2, or K2 alone on the rows a fused optimiser step left
    behind (`w_cache` = (what, inv_nw), optim.FusedHeadAdamW).  `sample`: K1 gathers the sampled rows, K2 runs on
    them (`label_local` already remapped), the argmax comes back as a global class id.  `out`: (max fp32 [B], sum fp32 [B],
    arg int64 [B]) to fill."""
    B = xhat.shape[0]
    chunks = batch_chunks(B, cfg.prec)
    class_lo = cfg.class_lo
    what = inv_nw = None
    if sample is not None:
        if cfg.prec:
            raise NotImplementedError("class sampling runs in precision='bf16' only")
        what, inv_nw = K.normalize_cast_gather(w, sample)
        class_lo = 0
    elif cfg.prec:   # bf16x3: K1 writes the three-part rows, K2 contracts over 3 D
        what, inv_nw, _ = K.normalize_cast3(w, 1)
    elif w_cache is not None:
        what, inv_nw = w_cache
    if len(chunks) > 1 and out is None:
        dev = xhat.device
        out = (torch.empty(B, dtype=torch.float32, device=dev), torch.empty(B, dtype=torch.float32, device=dev),
               torch.empty(B, dtype=torch.int64, device=dev))
    rmax = rsum = rarg = None
    for b0, b1 in chunks:
        whole = (b0, b1) == (0, B)
        kw = {} if out is None else {"out": out if whole else tuple(o[b0:b1] for o in out)}
        xs = xhat if whole else xhat[b0:b1]
        ls = label_local if (whole or label_local is None) else label_local[b0:b1]
        if what is None:   # first launch: the forward kernel normalises the weights itself and leaves them behind
            what, inv_nw, rmax, rsum, rarg = K.forward_rows_fused(xs, w, ls, cfg.s, class_lo, **kw)
        else:
            rmax, rsum, rarg = K.forward_rows(xs, what, ls, cfg.s, class_lo, **kw)
    if out is not None:
        rmax, rsum, rarg = out
    if sample is not None:
        rarg.copy_(sample[rarg] + cfg.class_lo)   # sampled position -> global class id (B values)
    return what, inv_nw, rmax, rsum, rarg


def forward_eager(K, group, x_local, w, y_local, cfg: StepConfig, packed_xy=None, w_cache=None, peer=None,
                  guard=None, num_sample=0, generator=None) -> FwdState:
    """K1 (x) -> label margin -> K1 (w) + K2 -> combine -> [exchange] -> finalize.  arcface.py:45-63 + the mean
    CrossEntropyLoss + argmax of the call sites, for the global batch against the local class rows."""
    R, rank = _world(group), _rank(group)
    b_loc = x_local.shape[0]
    x_all, y_all = gather_batch(x_local, y_local, group, packed_xy, peer)
    B = x_all.shape[0]
    gk = {"bad_flag_out": guard.host} if guard is not None else {}   # the label kernel raises the guard's host flag
    if cfg.prec:
        xhat, inv_nx, xhat_t = K.normalize_cast3(x_all, 0, want_transpose=True)
    else:
        xhat, inv_nx, xhat_t = K.normalize_cast(x_all, want_transpose=True)
    if R > 1 and B % 2 == 0 and hasattr(K, "finalize_rows_packed"):
        # the kernels fill one packed buffer, the exchange is one all-gather, the merge reads it in place
        buf, v_max, v_sum, v_z, v_arg = K.packed_stats(B, x_all.device)
        lm = K.label_margin(x_all, w, inv_nx, None, y_all, cfg.class_lo, cfg.c_total, cfg.s, cfg.m, cfg.easy_margin,
                            z_out=v_z, **gk)
        sample, label_k = _sample_for(lm.label_local, w.shape[0], num_sample, generator)
        what, inv_nw, _, _, _ = _rows(K, xhat, w, label_k, cfg, w_cache, out=(v_max, v_sum, v_arg), sample=sample)
        if peer is not None and buf.numel() % 16 == 0:
            allp = peer.all_gather_bytes(1, buf)
        else:
            allp = _all_gather_bytes(buf, group)
        lse, argmax, _z, omp, loss = K.finalize_rows_packed(allp, y_all)
    else:
        lm = K.label_margin(x_all, w, inv_nx, None, y_all, cfg.class_lo, cfg.c_total, cfg.s, cfg.m, cfg.easy_margin, **gk)
        sample, label_k = _sample_for(lm.label_local, w.shape[0], num_sample, generator)
        what, inv_nw, rmax, rsum, rarg = _rows(K, xhat, w, label_k, cfg, w_cache, sample=sample)
        rows_max, rows_sum, rows_z, rows_arg = exchange_rows(rmax, rsum, lm.z_label, rarg, group)
        lse, argmax, _z, omp, loss = K.finalize_rows(rows_max, rows_sum, rows_arg, rows_z, y_all)
    argmax_local = argmax if R == 1 else argmax[rank * b_loc:(rank + 1) * b_loc]
    st = FwdState(loss, argmax_local, lm.bad_flag, B, inv_nx, xhat, xhat_t, what, inv_nw, lse, omp, lm.dphi, label_k)
    st.argmax_all = argmax   # (with `loss`: one packed buffer, ops.packed_outputs)
    st.sample, st.c_local = sample, w.shape[0]
    return st


def _sample_for(label_local, c_local, num_sample, generator):
    """(sample index or None, the labels the GEMM kernels see)."""
    if not num_sample or num_sample >= c_local:
        return None, label_local
    sample = sample_classes(label_local, c_local, num_sample, generator)
    return sample, remap_labels(label_local, sample)


def backward_eager(K, group, x_local, st: FwdState, grad_loss, cfg: StepConfig, need_dx: bool = True, peer=None,
                   sparse_dw: bool = False):
    """K3 -> [reduce-scatter] -> normalise backward.  Returns (dx for the local rows or None, dW of the local
    class rows)."""
    R, rank = _world(group), _rank(group)
    b_loc = x_local.shape[0]
    g = grad_loss.to(torch.float32).contiguous()
    kw = {"prec": cfg.prec} if cfg.prec else {}
    chunks = batch_chunks(st.B, cfg.prec)
    if len(chunks) == 1:
        dxhat_part, dw = K.backward(st.xhat, st.xhat_t, st.what, st.inv_nw, st.lse, st.omp, st.dphi, st.label_local,
                                    cfg.s, 1.0 / st.B, grad_loss_dev=g, **kw)
    else:
        # one K3 launch per row chunk (batch_chunks): dX rows land in their slice, dW is summed over the chunks
        dxhat_part = torch.empty((st.B, st.xhat.shape[1]), dtype=torch.float32, device=st.xhat.device)
        dw = tmp = None
        for b0, b1 in chunks:
            if dw is not None and tmp is None:
                tmp = torch.empty_like(dw)
            _, dwi = K.backward(st.xhat[b0:b1], st.xhat_t[:, b0:b1], st.what, st.inv_nw, st.lse[b0:b1], st.omp[b0:b1],
                                st.dphi[b0:b1], st.label_local[b0:b1], cfg.s, 1.0 / st.B, grad_loss_dev=g,
                                dxhat_out=dxhat_part[b0:b1], dw_out=tmp)
            if dw is None:
                dw = dwi
            else:
                K.accumulate(dw, tmp)
    if st.sample is not None:
        # class sampling: the rows that were not drawn took no part in the softmax -- their gradient is exactly zero
        if sparse_dw:
            dw = torch.sparse_coo_tensor(st.sample.view(1, -1), dw, (st.c_local, dw.shape[1]), check_invariants=False,
                                         is_coalesced=True)   # the sample is sorted and free of duplicates
        else:
            dw = K.scatter_rows(dw, st.sample, torch.zeros((st.c_local, dw.shape[1]), dtype=dw.dtype, device=dw.device))
    dx = None
    if need_dx:
        inv_loc = st.inv_nx if R == 1 else st.inv_nx[rank * b_loc:(rank + 1) * b_loc].contiguous()
        parts = peer.scatter_rows(dxhat_part) if (peer is not None and R > 1) else None
        if parts is not None:
            # every rank stored its partial rows into the owner's buffer; the sum is fused into the normalise backward
            dx = K.normalize_bwd_x_sum(x_local, inv_loc, parts)
        else:
            dx = K.normalize_bwd_x(x_local, inv_loc, reduce_scatter_rows(dxhat_part, group))
    return dx, dw


# ----------------------------------------------------------------------------------------- CUDA graphs
class GraphedStep:
    """One signature of the step captured as ONE CUDA graph over static buffers.

    with_backward: the graph holds forward AND backward, the backward run ahead of time with an upstream
    gradient of 1.  The gradients of a scalar loss are linear in its upstream gradient, so `loss.backward()` only
    has to apply that factor (`ops.scale_grads`, which returns immediately when it is exactly 1 -- the
    reference's `loss.backward()` on the head's own loss) and hand the buffers to autograd: one graph launch and
    one near-empty kernel per training step.  Without gradients (torch.no_grad) the graph holds the forward only.
    """

    WARMUP = 2

    def __init__(self, K, group, w: torch.Tensor, b_loc: int, cfg: StepConfig, with_backward: bool, w_cache=None,
                 peer=None, guard=None):
        dev = w.device
        D = w.shape[1]
        self.K, self.group, self.cfg, self.peer, self.guard = K, group, cfg, peer, guard
        # static inputs as views of one buffer laid out (x | labels): the sharded gather sends it as is
        self.xy = torch.zeros(b_loc * D * 4 + b_loc * 8, dtype=torch.uint8, device=dev)
        self.x = self.xy[: b_loc * D * 4].view(torch.float32).view(b_loc, D)
        self.y = self.xy[b_loc * D * 4:].view(torch.int64)
        self.one = torch.ones((), dtype=torch.float32, device=dev)
        self.version = 0
        self.with_backward = with_backward
        self.y.fill_(cfg.class_lo)  # labels of the warm-up / capture runs must be valid class ids
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self.WARMUP):  # lazy initialisation (NCCL communicators, kernel attributes) outside capture
                st = forward_eager(K, group, self.x, w, self.y, cfg, self.xy, w_cache, peer, guard)
                if with_backward:
                    backward_eager(K, group, self.x, st, self.one, cfg, True, peer)
            del st
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        self.dx = self.dw = None
        with torch.cuda.graph(self.graph):
            self.st = forward_eager(K, group, self.x, w, self.y, cfg, self.xy, w_cache, peer, guard)
            if with_backward:
                self.dx, self.dw = backward_eager(K, group, self.x, self.st, self.one, cfg, True, peer)

    def outputs(self):
        """(loss, argmax of the local rows) copied out of the graph's static storage: one device copy."""
        st = self.st
        full = st.argmax_all if st.argmax_all is not None else st.argmax_local
        if hasattr(self.K, "clone_outputs"):
            arg, loss = self.K.clone_outputs(full, st.loss)
        else:
            arg, loss = full.clone(), st.loss.clone()
        if arg.numel() != st.argmax_local.numel():
            lo = st.argmax_local.storage_offset() - full.storage_offset()
            arg = arg[lo: lo + st.argmax_local.numel()]
        return loss, arg

    def run(self, x_local, y_local, param=None):
        if self.with_backward and param is not None and param.grad is not None and \
                param.grad.untyped_storage().data_ptr() == self.dw.untyped_storage().data_ptr():
            # the caller accumulates gradients and .grad still aliases the buffer this replay overwrites
            param.grad = param.grad.clone()
        if hasattr(self.K, "pack_xy") and x_local.is_contiguous() and y_local.is_contiguous():
            self.K.pack_xy(x_local, y_local, self.xy)   # both inputs into the static buffer: one launch
        else:
            self.x.copy_(x_local)
            self.y.copy_(y_local)
        if self.peer is not None:
            self.peer.check()   # an exchange of an earlier replay timed out waiting for a rank
        self.graph.replay()
        if self.guard is not None:
            self.guard.mark()
        if self.peer is not None:  # what the replay just issued on the device (p2p.PeerExchange.last_channel)
            self.peer.last_channel = 2 if self.with_backward else 1
        self.version += 1
        return self.version


class _GraphedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, label, plan: GraphedStep, param_ref):
        ctx.version = plan.run(x, label, param_ref() if param_ref is not None else None)
        ctx.plan = plan
        loss, argmax = plan.outputs()
        ctx.mark_non_differentiable(argmax)
        return loss, argmax

    @staticmethod
    def backward(ctx, grad_loss, _grad_argmax):
        plan = ctx.plan
        if ctx.version != plan.version:
            raise RuntimeError("ArcMarginProduct (CUDA-graph mode): backward() of a forward whose buffers were reused "
                               "by a later forward of the same head; set head.use_cuda_graph = False to keep several "
                               "forwards in flight")
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g = grad_loss.to(torch.float32).reshape(1).contiguous()
        if need_dx and hasattr(plan.K, "scale_copy"):
            # dx leaves the static buffer scaled, dW is scaled in place (no-op for a factor of 1): one launch
            dx = plan.K.scale_copy(plan.dx, plan.dw if need_dw else None, g)
        else:
            plan.K.scale_grads(plan.dx if need_dx else None, plan.dw if need_dw else None, g)
            dx = plan.dx.clone() if need_dx else None
        dw = plan.dw.detach() if need_dw else None  # alias: autograd adopts it without a copy
        ctx.version = -1  # the factor has been applied in place: a second backward would apply it twice
        return dx, dw, None, None, None


class ArcFaceCEFunction(torch.autograd.Function):
    """loss, argmax = ArcFaceCEFunction.apply(x, weight, label, kernels, group, cfg, validate_labels[, w_cache, peer,
    guard]) -- the autograd.Function both modules run (eagerly; `_GraphedCE` is its CUDA-graph twin).

    forward : K1 (x), label margin, K1 (weight) fused into K2, combine, [exchange], finalize
              (arcface.py:45-63 + the call sites' mean CrossEntropyLoss and argmax)
    backward: K3 (dC^T producer, dW GEMM, dX GEMM), [reduce-scatter], normalise backward for x
    Saved for backward: xhat / xhat^T / what (bf16), the inverse norms, lse, 1 - p_label, dphi, labels -- no B x C
    tensor.  `kernels` is `multimodalsimilar_b200.ops`, `group` None (one GPU) or the class-shard process group,
    `cfg` a StepConfig."""

    @staticmethod
    def forward(ctx, x, w, label, K, group, cfg, validate_labels, w_cache=None, peer=None, guard=None, sampling=None):
        num_sample, generator, sparse_dw, owner = (tuple(sampling) + (None,))[:4] if sampling is not None else (0, None, False, None)
        st = forward_eager(K, group, x, w, label, cfg, None, w_cache, peer, None if validate_labels else guard,
                           num_sample, generator)
        if owner is not None and owner() is not None and st.sample is not None:
            _LAST_SAMPLE[owner()] = st.sample
        if validate_labels:
            if int(st.bad_flag.item()) != 0:
                raise IndexError("ArcMarginProduct: a label is outside [0, %d)" % cfg.c_total)
        elif guard is not None:
            guard.mark()
        ctx.save_for_backward(x, st.inv_nx, st.xhat, st.xhat_t, st.what, st.inv_nw, st.lse, st.omp, st.dphi,
                              st.label_local)
        ctx.meta = (K, group, cfg, st.B, peer, st.sample, st.c_local, sparse_dw)
        ctx.mark_non_differentiable(st.argmax_local)
        return st.loss, st.argmax_local

    @staticmethod
    def backward(ctx, grad_loss, _grad_argmax):
        x, inv_nx, xhat, xhat_t, what, inv_nw, lse, omp, dphi, label_local = ctx.saved_tensors
        K, group, cfg, B, peer, sample, c_local, sparse_dw = ctx.meta
        st = FwdState(None, None, None, B, inv_nx, xhat, xhat_t, what, inv_nw, lse, omp, dphi, label_local)
        st.sample, st.c_local = sample, c_local
        if peer is not None and not ctx.needs_input_grad[0]:
            # the exchange is a rendezvous of all ranks: it cannot be skipped by one of them
            dx, dw = backward_eager(K, group, x, st, grad_loss, cfg, True, peer, sparse_dw)
            dx = None
        else:
            dx, dw = backward_eager(K, group, x, st, grad_loss, cfg, ctx.needs_input_grad[0], peer, sparse_dw)
        return dx, (dw if ctx.needs_input_grad[1] else None), None, None, None, None, None, None, None, None, None


_EagerCE = ArcFaceCEFunction


# per-head graph state lives outside the module's __dict__ so that torch.save(model) keeps working
_PLANS: "weakref.WeakKeyDictionary[Any, dict]" = weakref.WeakKeyDictionary()

# head -> (what bf16 [C, D], inv_nw fp32 [C], weight._version, weight.data_ptr()) written by optim.FusedHeadAdamW
_W_CACHE: "weakref.WeakKeyDictionary[Any, tuple]" = weakref.WeakKeyDictionary()

# head -> PeerExchange (csrc/p2p.cu) or False when symmetric memory is unavailable for its group
_PEERS: "weakref.WeakKeyDictionary[Any, Any]" = weakref.WeakKeyDictionary()

# head -> LabelGuard (asynchronous out-of-range-label check)
_GUARDS: "weakref.WeakKeyDictionary[Any, LabelGuard]" = weakref.WeakKeyDictionary()


def _peer_for(head, group, x):
    """The head's peer-memory exchange for this batch shape, created collectively on first use.  Any rank failing
    (no NVLink peer mapping, gloo group ...) sends every rank back to the NCCL collectives."""
    st = _PEERS.get(head)
    if st is False:
        return None
    b_loc, D = x.shape
    if st is not None and st.matches(b_loc, D):
        return st
    if dist.get_world_size(group) < 2 or dist.get_backend(group) != "nccl":
        _PEERS[head] = False
        return None
    peer, err = None, None
    try:
        from .p2p import PeerExchange

        peer = PeerExchange(group, x.device, b_loc, D)
    except Exception as e:  # noqa: BLE001
        err = e
    ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=x.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 0:
        warnings.warn("multimodalsimilar_b200: peer-memory exchange unavailable (%r); using NCCL collectives" % (err,))
        _PEERS[head] = False
        return None
    _PEERS[head] = peer
    return peer


ENGAGE_AFTER = 2  # eager calls with an unchanged signature before a graph is captured


# head -> torch.Generator of its class sampling (created from head.sample_seed on first use)
_SAMPLERS: "weakref.WeakKeyDictionary[Any, Any]" = weakref.WeakKeyDictionary()


def _sampling_for(head, c_local: int, device):
    """(num_sample, generator, sparse_grad) when the head samples classes in this call, else None.  Sampling applies in
    training mode only (`head.training`), like PartialFC; evaluation always sees every class."""
    rate = float(getattr(head, "sample_rate", 1.0))
    if rate >= 1.0 or not getattr(head, "training", True) or not torch.is_grad_enabled():
        return None
    if not rate > 0.0:
        raise ValueError("sample_rate must be in (0, 1], got %r" % rate)
    num = max(1, int(round(rate * c_local)))
    gen = _SAMPLERS.get(head)
    seed = getattr(head, "sample_seed", None)
    if gen is None and seed is not None:
        gen = _SAMPLERS[head] = torch.Generator(device=device)
        gen.manual_seed(int(seed))
    return num, gen, bool(getattr(head, "sparse_grad", False)), weakref.ref(head)


# head -> sorted int64 [S] local class ids its most recent training step ran on (PartialFC calls this `weight_index`)
_LAST_SAMPLE: "weakref.WeakKeyDictionary[Any, torch.Tensor]" = weakref.WeakKeyDictionary()


def last_sample(head):
    """The class sample of `head`'s most recent sampled step (sorted local class ids), or None."""
    return _LAST_SAMPLE.get(head)


def run_step(head, K, group, x, w, label, cfg: StepConfig, validate_labels: bool):
    """loss, argmax = one forward of `head` (autograd-connected).  Graph replay when `head.use_cuda_graph` and the
    signature has repeated; the eager kernel sequence otherwise."""
    # normalised rows left behind by a fused optimiser step: valid while the weight has not been touched since
    w_cache = None
    cache = _W_CACHE.get(head)
    if cache is not None and x.is_cuda and cache[2] == w._version and cache[3] == w.data_ptr() and not cfg.prec:
        w_cache = (cache[0], cache[1])
    peer = _peer_for(head, group, x) if (group is not None and x.is_cuda and getattr(head, "use_p2p", False)) else None
    guard = None
    if x.is_cuda and not validate_labels:
        guard = _GUARDS.get(head)
        if guard is None:
            guard = _GUARDS[head] = LabelGuard()
        guard.check(cfg.c_total)   # raises for a bad label of the previous step
    sampling = _sampling_for(head, w.shape[0], x.device)
    if sampling is not None:
        # a fresh sample every step (torch RNG + top-k): eager launches; the weight cache covers all rows, not the sample
        return _EagerCE.apply(x, w, label, K, group, cfg, validate_labels, None, peer, guard, sampling)
    use_graph = bool(getattr(head, "use_cuda_graph", False)) and x.is_cuda and not validate_labels
    if not use_graph:
        return _EagerCE.apply(x, w, label, K, group, cfg, validate_labels, w_cache, peer, guard)
    with_bwd = torch.is_grad_enabled() and (x.requires_grad or w.requires_grad)
    sig = (tuple(x.shape), x.device, w.data_ptr(), tuple(w.shape), cfg, with_bwd, id(group),
           w_cache[0].data_ptr() if w_cache is not None else 0, id(peer))
    state = _PLANS.setdefault(head, {"sig": None, "seen": 0, "plan": None, "failed": False})
    if state["failed"]:
        return _EagerCE.apply(x, w, label, K, group, cfg, validate_labels, w_cache, peer, guard)
    if state["sig"] != sig:
        state.update(sig=sig, seen=0, plan=None)
    if state["plan"] is None:
        state["seen"] += 1
        if state["seen"] <= ENGAGE_AFTER:
            return _EagerCE.apply(x, w, label, K, group, cfg, validate_labels, w_cache, peer, guard)
        try:
            state["plan"] = GraphedStep(K, group, w.detach(), x.shape[0], cfg, with_bwd, w_cache, peer, guard)
        except Exception as e:  # keep training: the eager sequence computes the same thing
            state["failed"] = True
            warnings.warn("multimodalsimilar_b200: CUDA-graph capture failed (%r); continuing with eager launches" % (e,))
            return _EagerCE.apply(x, w, label, K, group, cfg, validate_labels, w_cache, peer, guard)
    plan: GraphedStep = state["plan"]
    if not with_bwd:
        plan.run(x, label)
        return plan.outputs()
    param = getattr(head, "weight", None)
    return _GraphedCE.apply(x, w, label, plan, weakref.ref(param) if isinstance(param, torch.nn.Parameter) else None)


def drop_plan(head) -> None:
    """Forget the captured graphs of `head` (frees their memory pool)."""
    _PLANS.pop(head, None)
