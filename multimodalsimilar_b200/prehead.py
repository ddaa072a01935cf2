"""The steps around the head that the reference's model wrappers run per batch.

`two_stream_embed(img, text)`
    multimodal_classifier.py:50-56 -- `cat(F.normalize(img_emb), F.normalize(title_emb), dim=1)`, the embedding the
    two-stream model hands to its ArcFace head -- as one kernel forward and one kernel backward (the reference runs
    five launches forward and about ten in autograd, each a B x D round trip).

`MultiHeadArcFace(heads, weights)`
    nlp_classifier_multilabel.py:33-35 + nlp_classifier_train_daodian_v3_dist.py:164-166 -- several ArcFace heads on
    ONE embedding, trained on `10 * CE_1 + 5 * CE_2 + 1 * CE_3`.  The heads share K1 of the embedding (one normalise /
    cast / transpose instead of one per head), every head's kernels run back to back on the shared xhat, each head's
    backward writes its own dXhat partial and ONE normalise-backward kernel sums the partials in fixed order
    (`normalize_bwd_x_sum`, the kernel the class-sharded head uses for its reduce-scatter).  With `use_cuda_graph` the
    whole thing -- all heads, forward and backward -- replays as one CUDA graph: the heads of the reference model are
    small (38 / 590 / 10 205 classes), i.e. launch-bound.
"""
from __future__ import annotations

import torch
from torch import nn

from . import engine, ops


class TwoStreamConcat(torch.autograd.Function):
    """emb = cat(normalize(a), normalize(b), dim=1); fused forward / backward (csrc/rows.cu)."""

    @staticmethod
    def forward(ctx, a, b):
        a = a.to(torch.float32).contiguous()
        b = b.to(torch.float32).contiguous()
        emb, inv1, inv2 = ops.two_stream_concat(a, b)
        ctx.save_for_backward(emb, inv1, inv2)
        ctx.d1 = a.shape[1]
        return emb

    @staticmethod
    def backward(ctx, grad):
        emb, inv1, inv2 = ctx.saved_tensors
        da, db = ops.two_stream_concat_bwd(emb, inv1, inv2, grad.to(torch.float32).contiguous(), ctx.d1)
        return da, db


def two_stream_embed(img_embedding: torch.Tensor, title_embedding: torch.Tensor) -> torch.Tensor:
    """Drop-in for the last three lines of MultimodalClassifier.predict_emb (multimodal_classifier.py:53-55)."""
    if not img_embedding.is_cuda:
        raise RuntimeError("multimodalsimilar_b200.two_stream_embed runs on a B200 only; there is no CPU path")
    return TwoStreamConcat.apply(img_embedding, title_embedding)


# ----------------------------------------------------------------------------------------- several heads, one x
def _multi_forward(x, heads, ws, ys, cfgs):
    """Shared K1 (x), then per head: label margin, K1 (w) + K2, combine, finalize.  Returns the per-head states."""
    K = ops
    prec = cfgs[0].prec
    if prec:
        xhat, inv_nx, xhat_t = K.normalize_cast3(x, 0, want_transpose=True)
    else:
        xhat, inv_nx, xhat_t = K.normalize_cast(x, want_transpose=True)
    B = x.shape[0]
    states = []
    for w, y, cfg in zip(ws, ys, cfgs):
        lm = K.label_margin(x, w, inv_nx, None, y, 0, cfg.c_total, cfg.s, cfg.m, cfg.easy_margin)
        what, inv_nw, rmax, rsum, rarg = engine._rows(K, xhat, w, lm.label_local, cfg, None)
        lse, argmax, _z, omp, loss = K.finalize_rows(rmax.view(1, B), rsum.view(1, B), rarg.view(1, B),
                                                     lm.z_label.view(1, B), y)
        states.append(engine.FwdState(loss, argmax, lm.bad_flag, B, inv_nx, xhat, xhat_t, what, inv_nw, lse, omp,
                                      lm.dphi, lm.label_local))
    return states


def _multi_backward(x, states, cfgs, gdev, parts):
    """Per head K3 with its own upstream factor gdev[h] (DEVICE scalars) into parts[h]; one summed normalise backward."""
    K = ops
    dws = []
    for h, (st, cfg) in enumerate(zip(states, cfgs)):
        kw = {"prec": cfg.prec} if cfg.prec else {}
        _, dw = K.backward(st.xhat, st.xhat_t, st.what, st.inv_nw, st.lse, st.omp, st.dphi, st.label_local, cfg.s,
                           1.0 / st.B, grad_loss_dev=gdev[h:h + 1], dxhat_out=parts[h], **kw)
        dws.append(dw)
    dx = K.normalize_bwd_x_sum(x, states[0].inv_nx, parts)
    return dx, dws


class _MultiCE(torch.autograd.Function):
    """total, losses [H], argmax_0 .. argmax_{H-1} = f(x, w_0 .. w_{H-1}); labels / configs / weights ride in `pack`."""

    @staticmethod
    def forward(ctx, x, pack, *ws):
        ys, cfgs, weights = pack["labels"], pack["cfgs"], pack["weights_dev"]
        states = _multi_forward(x, pack["heads"], ws, ys, cfgs)
        losses = torch.stack([st.loss for st in states])
        total = (losses * weights).sum()
        ctx.save_for_backward(x, weights, *[t for st in states for t in (st.inv_nx, st.xhat, st.xhat_t, st.what, st.inv_nw,
                                                                         st.lse, st.omp, st.dphi, st.label_local)])
        ctx.cfgs = cfgs
        ctx.B = x.shape[0]
        argmaxes = tuple(st.argmax_local for st in states)
        ctx.mark_non_differentiable(*argmaxes)
        return (total, losses.detach()) + argmaxes

    @staticmethod
    def backward(ctx, g_total, _g_losses, *_g_arg):
        saved = ctx.saved_tensors
        x, weights = saved[0], saved[1]
        H = len(ctx.cfgs)
        states = []
        for h in range(H):
            t = saved[2 + 9 * h: 2 + 9 * (h + 1)]
            states.append(engine.FwdState(None, None, None, ctx.B, *t))
        gdev = (weights * g_total.to(torch.float32)).contiguous()
        parts = torch.empty((H,) + tuple(x.shape), dtype=torch.float32, device=x.device)
        dx, dws = _multi_backward(x, states, ctx.cfgs, gdev, parts)
        return (dx if ctx.needs_input_grad[0] else None, None) + tuple(dws)


class _MultiGraph:
    """Forward + backward of every head captured as ONE CUDA graph over static buffers (upstream gradient 1 for the
    weighted total; `loss.backward()` applies any other factor with `ops.scale_grads`, like engine.GraphedStep)."""

    def __init__(self, x_shape, heads, ws, cfgs, weights_dev, device):
        H = len(heads)
        self.x = torch.zeros(x_shape, dtype=torch.float32, device=device)
        self.ys = [torch.zeros(x_shape[0], dtype=torch.int64, device=device) for _ in range(H)]
        self.weights = weights_dev
        self.version = 0
        self.parts = torch.empty((H,) + tuple(x_shape), dtype=torch.float32, device=device)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))

        def run():
            states = _multi_forward(self.x, heads, ws, self.ys, cfgs)
            losses = torch.stack([st.loss for st in states])
            total = (losses * self.weights).sum()
            dx, dws = _multi_backward(self.x, states, cfgs, self.weights, self.parts)
            return states, losses, total, dx, dws

        with torch.cuda.stream(side):
            for _ in range(2):
                run()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.states, self.losses, self.total, self.dx, self.dws = run()

    def replay(self, x, ys, params):
        for p, dw in zip(params, self.dws):
            if p.grad is not None and p.grad.untyped_storage().data_ptr() == dw.untyped_storage().data_ptr():
                p.grad = p.grad.clone()   # the caller accumulates gradients: .grad still aliases the buffer overwritten now
        self.x.copy_(x)
        for dst, src in zip(self.ys, ys):
            dst.copy_(src)
        self.graph.replay()
        self.version += 1
        return self.version


class _MultiGraphedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pack, *ws):
        plan = pack["plan"]
        ctx.version = plan.replay(x, pack["labels"], [h.weight for h in pack["heads"]])
        ctx.plan = plan
        argmaxes = tuple(st.argmax_local.clone() for st in plan.states)
        ctx.mark_non_differentiable(*argmaxes)
        return (plan.total.clone(), plan.losses.clone()) + argmaxes

    @staticmethod
    def backward(ctx, g_total, _g_losses, *_g_arg):
        plan = ctx.plan
        if ctx.version != plan.version:
            raise RuntimeError("MultiHeadArcFace (CUDA-graph mode): backward() of a forward whose buffers were reused by a "
                               "later forward; set use_cuda_graph = False to keep several forwards in flight")
        g = g_total.to(torch.float32).reshape(1).contiguous()
        ops.scale_grads(plan.dx, None, g)
        for dw in plan.dws:
            ops.scale_grads(dw, None, g)
        ctx.version = -1
        return (plan.dx.clone() if ctx.needs_input_grad[0] else None, None) + tuple(dw.detach() for dw in plan.dws)


class MultiHeadArcFace(nn.Module):
    """Several `ArcMarginProduct` heads sharing one embedding (nlp_classifier_multilabel.py:15-17, 33-35).

    `loss(x, labels) -> (total, losses [H], argmaxes)` with total = sum_h weights[h] * CE_h -- the reference's
    `10 * loss_first + 5 * loss_second + 1 * loss_tag` (nlp_classifier_train_daodian_v3_dist.py:164-166).  The heads
    stay ordinary modules (their `weight` parameters, `update_m`, `forward_test`, checkpoints ... are untouched).
    """

    ENGAGE_AFTER = 2

    def __init__(self, heads, weights=None, use_cuda_graph=True):
        super().__init__()
        self.heads = nn.ModuleList(heads)
        self.loss_weights = [1.0] * len(heads) if weights is None else [float(w) for w in weights]
        if len(self.loss_weights) != len(heads):
            raise ValueError("one loss weight per head")
        if len({h.in_feature for h in heads}) != 1 or len({h.precision for h in heads}) != 1:
            raise ValueError("the heads must share the embedding width and the precision mode")
        self.use_cuda_graph = use_cuda_graph
        self._state = {"sig": None, "seen": 0, "plan": None}

    def __getstate__(self):   # the captured graph cannot be pickled (torch.save(model))
        state = dict(self.__dict__)
        state["_state"] = {"sig": None, "seen": 0, "plan": None}
        return state

    def loss(self, x, labels):
        if not x.is_cuda:
            raise RuntimeError("multimodalsimilar_b200.MultiHeadArcFace runs on a B200 only; there is no CPU path")
        x = x.to(torch.float32).contiguous()
        if x.shape[0] > ops.MAX_BATCH:
            # (the single heads run larger batches in row chunks, engine.batch_chunks; the shared-K1 multi-head step does
            # not: its backward is one launch group per head)
            raise ValueError("MultiHeadArcFace takes up to %d rows per call, got %d" % (ops.MAX_BATCH, x.shape[0]))
        ys = [y.reshape(-1).to(device=x.device, dtype=torch.int64).contiguous() for y in labels]
        heads = list(self.heads)
        ws = [h.weight if h.weight.is_contiguous() else h.weight.contiguous() for h in heads]
        cfgs = [engine.StepConfig(float(h.s), float(h.m), bool(h.easy_margin), 0, w.shape[0],
                                  engine.precision_code(h.precision)) for h, w in zip(heads, ws)]
        wdev = getattr(self, "_wdev", None)
        if wdev is None or wdev.device != x.device or wdev.tolist() != self.loss_weights:
            wdev = self._wdev = torch.tensor(self.loss_weights, dtype=torch.float32, device=x.device)
        pack = {"heads": heads, "labels": ys, "cfgs": cfgs, "weights_dev": wdev}
        st = self._state
        graph_ok = self.use_cuda_graph and torch.is_grad_enabled() and all(w.requires_grad for w in ws)
        if graph_ok:
            sig = (tuple(x.shape), x.device, tuple(w.data_ptr() for w in ws), tuple(cfgs), tuple(self.loss_weights))
            if st["sig"] != sig:
                st.update(sig=sig, seen=0, plan=None)
            if st["plan"] is None:
                st["seen"] += 1
                if st["seen"] > self.ENGAGE_AFTER:
                    st["plan"] = _MultiGraph(tuple(x.shape), heads, [w.detach() for w in ws], cfgs, wdev, x.device)
            if st["plan"] is not None:
                pack["plan"] = st["plan"]
                out = _MultiGraphedCE.apply(x, pack, *ws)
                return out[0], out[1], list(out[2:])
        out = _MultiCE.apply(x, pack, *ws)
        return out[0], out[1], list(out[2:])

    def forward(self, x, labels):
        return self.loss(x, labels)
