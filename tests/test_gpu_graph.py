"""CUDA-graph replay of the head step (multimodalsimilar_b200/engine.py) against the eager kernel sequence:
same kernels, so results must be bit-identical; plus the autograd corner cases of static buffers."""
import numpy as np
import pytest
import torch

from oracle import arcface_numpy as onp

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _head(w, s, m, graph):
    import multimodalsimilar_b200 as mm

    h = mm.ArcMarginProduct(w.shape[1], w.shape[0], s=s, m=m, use_cuda_graph=graph).to(dev())
    with torch.no_grad():
        h.weight.copy_(torch.from_numpy(w))
    return h


def _step(head, x, y, grad=1.0):
    xt = torch.from_numpy(x).to(dev()).requires_grad_(True)
    yt = torch.from_numpy(y).to(dev())
    head.weight.grad = None
    loss, pred = head.loss(xt, yt)
    (loss * grad).backward()
    return loss.detach().clone(), pred.clone(), xt.grad.clone(), head.weight.grad.clone()


def _same(a, b, what, grad=1.0):
    """Same kernels either way: loss / argmax / dW / dX bit-identical (the single-launch backward sums the class splits'
    dX tiles in a fixed order; the three-launch fallbacks use fp32 reduce-adds and are reproducible to rounding only).
    With an upstream gradient != 1 the eager path rounds dC = bf16(grad * ...) while the graph path runs the
    backward with 1 and scales the fp32 result: equal up to one bf16 rounding of dC."""
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), what
    if grad == 1.0:
        assert torch.equal(a[2], b[2]), what + ": dX"
        assert torch.equal(a[3], b[3]), what
    else:
        for u, v in ((a[2], b[2]), (a[3], b[3])):
            assert float((u - v).norm() / v.norm()) <= 6e-3, what


@pytest.mark.parametrize("B,D,C", [(64, 128, 3000), (96, 1024, 5000)])   # CTA-pair kernels / generic (D > 512) kernels
def test_graph_replay_equals_eager(B, D, C):
    from multimodalsimilar_b200 import engine

    s, m = 64.0, 0.4
    _, w, _ = onp.synthetic_inputs(B, D, C, seed=1, trained_like=False)
    eager, graph = _head(w, s, m, False), _head(w, s, m, True)
    for it in range(6):
        x, _, y = onp.synthetic_inputs(B, D, C, seed=10 + it, trained_like=False)
        g = 1.0 if it % 2 == 0 else 2.5
        a = _step(eager, x, y, g)
        b = _step(graph, x, y, g)
        _same(a, b, "iteration %d" % it, g)
    st = engine._PLANS[graph]
    assert st["plan"] is not None and not st["failed"], "the graph was never captured"
    assert engine._PLANS.get(eager) is None


@pytest.mark.parametrize("B,D,C", [(512, 512, 200000), (256, 1792, 30000), (1024, 512, 60000), (2304, 256, 20000)])
def test_step_is_bit_reproducible(B, D, C):
    """Run to run (and launch schedule to launch schedule): loss, argmax, dX and dW of the bf16 mode are bit-identical --
    partial statistics are merged in class order, q / dW have one writer per element, the dX tiles of the class splits
    are summed in split order (deterministic reductions)."""
    s, m = 64.0, 0.5
    x, w, y = onp.synthetic_inputs(B, D, C, seed=4, trained_like=True)
    head = _head(w, s, m, False)
    first = _step(head, x, y, 1.0)
    for it in range(4):
        if it == 2:   # perturb the timing of the roles: another kernel's leftovers in L2, a different clock state
            torch.empty(64 << 20, device=dev()).normal_()
        again = _step(head, x, y, 1.0)
        for u, v, name in zip(first, again, ("loss", "argmax", "dX", "dW")):
            assert torch.equal(u, v), "%s differs between runs (iteration %d)" % (name, it)


def test_graph_mode_gradient_accumulation_and_margin_update():
    B, D, C, s, m = 32, 64, 1000, 64.0, 0.3
    _, w, _ = onp.synthetic_inputs(B, D, C, seed=2, trained_like=False)
    eager, graph = _head(w, s, m, False), _head(w, s, m, True)
    batches = [onp.synthetic_inputs(B, D, C, seed=20 + i, trained_like=False) for i in range(5)]
    for h in (eager, graph):
        h.weight.grad = None
        for x, _, y in batches:          # no zero_grad between steps: .grad accumulates
            xt = torch.from_numpy(x).to(dev()).requires_grad_(True)
            loss, _ = h.loss(xt, torch.from_numpy(y).to(dev()))
            loss.backward()
    torch.testing.assert_close(graph.weight.grad, eager.weight.grad, rtol=1e-6, atol=1e-7)
    # a margin update changes the kernel arguments: the plan is re-captured, results keep matching
    for h in (eager, graph):
        h.update_m(0.04)
    for it in range(4):
        x, _, y = batches[it]
        a = _step(eager, x, y)
        b = _step(graph, x, y)
        _same(a, b, "after update_m, iteration %d" % it)


def test_graph_mode_rejects_stale_backward():
    B, D, C = 16, 64, 500
    x, w, y = onp.synthetic_inputs(B, D, C, seed=3, trained_like=False)
    h = _head(w, 64.0, 0.4, True)
    for _ in range(4):
        _step(h, x, y)
    xt = torch.from_numpy(x).to(dev()).requires_grad_(True)
    yt = torch.from_numpy(y).to(dev())
    l1, _ = h.loss(xt, yt)
    l2, _ = h.loss(xt, yt)
    l2.backward()
    with pytest.raises(RuntimeError, match="use_cuda_graph"):
        l1.backward()


def test_graph_mode_no_grad_and_pickle():
    import io

    B, D, C = 16, 64, 500
    x, w, y = onp.synthetic_inputs(B, D, C, seed=4, trained_like=False)
    h = _head(w, 64.0, 0.4, True)
    e = _head(w, 64.0, 0.4, False)
    xt, yt = torch.from_numpy(x).to(dev()), torch.from_numpy(y).to(dev())
    with torch.no_grad():
        ref = e.loss(xt, yt)
        for _ in range(5):
            got = h.loss(xt, yt)
    assert torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1])
    buf = io.BytesIO()
    torch.save(h, buf)      # the reference checkpoints whole modules (cv_classifier_train_daodian.py:298-306)
    buf.seek(0)
    h2 = torch.load(buf, weights_only=False)
    assert torch.equal(h2.weight, h.weight)


def test_graph_mode_two_heads_share_an_embedding():
    """Two heads on one embedding with the 10 / 5 loss weights of nlp_classifier_train_daodian_v3_dist.py:164-166:
    every head replays its own graph, upstream gradients != 1 go through scale_grads, dx accumulates over heads."""
    B, D, s = 32, 64, 64.0
    _, w1, _ = onp.synthetic_inputs(B, D, 700, seed=5, trained_like=False)
    _, w2, _ = onp.synthetic_inputs(B, D, 90, seed=6, trained_like=False)
    eager = (_head(w1, s, 0.4, False), _head(w2, s, 0.2, False))
    graph = (_head(w1, s, 0.4, True), _head(w2, s, 0.2, True))
    for it in range(5):
        x, _, y1 = onp.synthetic_inputs(B, D, 700, seed=40 + it, trained_like=False)
        y2 = (y1 % 90).astype(np.int64)
        res = []
        for heads in (eager, graph):
            xt = torch.from_numpy(x).to(dev()).requires_grad_(True)
            for h in heads:
                h.weight.grad = None
            l1, p1 = heads[0].loss(xt, torch.from_numpy(y1).to(dev()))
            l2, p2 = heads[1].loss(xt, torch.from_numpy(y2).to(dev()))
            (10.0 * l1 + 5.0 * l2).backward()
            res.append((l1.detach().clone(), l2.detach().clone(), p1.clone(), p2.clone(), xt.grad.clone(),
                        heads[0].weight.grad.clone(), heads[1].weight.grad.clone()))
        a, b = res
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
        for u, v in zip(a[4:], b[4:]):
            assert float((u - v).norm() / v.norm()) <= 6e-3, "iteration %d" % it


@pytest.mark.parametrize("graph", [False, True])
def test_bad_label_raises_one_step_late(graph):
    """arcface.py:59 (`scatter_`) raises on a label outside [0, C).  Without validate_labels the device flag is
    examined at the start of the NEXT call of the head -- eagerly and in CUDA-graph mode -- instead of never."""
    B, D, C = 32, 64, 500
    x, w, y = onp.synthetic_inputs(B, D, C, seed=3)
    head = _head(w, 64.0, 0.4, graph)
    for _ in range(4):      # graph mode: the capture happens on the third call
        _step(head, x, y)
    torch.cuda.synchronize()
    bad = y.copy()
    bad[5] = C + 7
    _step(head, x, bad)     # trains on silently ...
    torch.cuda.synchronize()
    with pytest.raises(IndexError):
        _step(head, x, y)   # ... and is reported here
    torch.cuda.synchronize()
    _step(head, x, y)       # the report is not sticky
    strict = _head(w, 64.0, 0.4, graph)
    strict.validate_labels = True
    with pytest.raises(IndexError):
        _step(strict, x, bad)


def test_autograd_function_is_the_modules_path():
    """`ArcFaceCEFunction` (north_star: "drop-in torch.nn.Module / autograd.Function") applied by hand equals the
    module that applies it through engine.run_step."""
    import multimodalsimilar_b200 as mm
    from multimodalsimilar_b200 import engine, ops

    B, D, C = 48, 128, 1000
    x, w, y = onp.synthetic_inputs(B, D, C, seed=4)
    head = _head(w, 64.0, 0.4, False)
    ref = _step(head, x, y)
    xt = torch.from_numpy(x).to(dev()).requires_grad_(True)
    wt = torch.from_numpy(w).to(dev()).requires_grad_(True)
    cfg = engine.StepConfig(64.0, 0.4, False, 0, C)
    assert mm.ArcFaceCEFunction is engine.ArcFaceCEFunction
    loss, pred = mm.ArcFaceCEFunction.apply(xt, wt, torch.from_numpy(y).to(dev()), ops, None, cfg, False)
    loss.backward()
    assert torch.equal(loss.detach(), ref[0]) and torch.equal(pred, ref[1])
    torch.testing.assert_close(xt.grad, ref[2], rtol=1e-4, atol=1e-7)
    assert torch.equal(wt.grad, ref[3])
