"""Fused head optimiser: `FusedHeadAdamW` against torch.optim.AdamW, the normalised rows it
emits against K1, and a training loop whose forwards run from those rows against the ordinary loop."""
import pytest
import torch

from oracle import arcface_numpy as onp

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("C,D", [(1000, 64), (777, 512), (300, 2816)])
def test_adamw_step_matches_torch(C, D):
    import multimodalsimilar_b200 as mm
    from multimodalsimilar_b200 import ops

    g = torch.Generator(device=dev()).manual_seed(0)
    w0 = torch.randn(C, D, device=dev(), generator=g) * 0.05
    a = torch.nn.Parameter(w0.clone())
    b = torch.nn.Parameter(w0.clone())
    ref = torch.optim.AdamW([a], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    fus = mm.FusedHeadAdamW([b], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    for it in range(5):
        grad = torch.randn(C, D, device=dev(), generator=g) * (10.0 ** (-it))
        a.grad = grad.clone()
        b.grad = grad.clone()
        ref.step()
        fus.step()
        torch.testing.assert_close(b.detach(), a.detach(), rtol=2e-6, atol=1e-8)
        torch.testing.assert_close(fus.state[b]["exp_avg"], ref.state[a]["exp_avg"], rtol=2e-6, atol=1e-12)
        torch.testing.assert_close(fus.state[b]["exp_avg_sq"], ref.state[a]["exp_avg_sq"], rtol=2e-6, atol=1e-20)
    assert float(fus.state[b]["step"]) == float(ref.state[a]["step"]) == 5.0
    # optimiser checkpoints interchange with torch.optim.AdamW
    ref.load_state_dict(fus.state_dict())


def test_emitted_rows_equal_k1():
    import multimodalsimilar_b200 as mm
    from multimodalsimilar_b200 import engine, ops

    D, C = 256, 5000
    head = mm.ArcMarginProduct(D, C).to(dev())
    opt = mm.FusedHeadAdamW.for_head(head, lr=1e-2)
    head.weight.grad = torch.randn_like(head.weight) * 0.01
    opt.step()
    what, inv_nw, version, ptr = engine._W_CACHE[head]
    rw, rinv, _ = ops.normalize_cast(head.weight.detach())
    # same definition as K1; the sum of squares is accumulated in a different lane order, so 1 / ||w|| may differ in
    # the last bit and flip the bf16 rounding of a few elements by one ulp
    torch.testing.assert_close(inv_nw, rinv, rtol=1e-6, atol=0)
    torch.testing.assert_close(what.float(), rw.float(), rtol=2.0 ** -7, atol=0)
    assert float((what != rw).float().mean()) <= 1e-3
    assert version == head.weight._version and ptr == head.weight.data_ptr()
    with torch.no_grad():
        head.weight.mul_(1.0)          # any other in-place change invalidates the rows
    assert engine._W_CACHE[head][2] != head.weight._version


@pytest.mark.parametrize("graph", [False, True])
def test_training_loop_with_fused_optimizer_matches_plain_loop(graph):
    """Same data, same hyper-parameters: head + torch.optim.AdamW (K1 on the weights every forward) against head +
    FusedHeadAdamW (forwards run the GEMM from the rows the optimiser emitted)."""
    import multimodalsimilar_b200 as mm

    B, D, C, s, m = 64, 128, 3000, 64.0, 0.4
    _, w, _ = onp.synthetic_inputs(B, D, C, seed=1, trained_like=False)
    heads = []
    for _ in range(2):
        h = mm.ArcMarginProduct(D, C, s=s, m=m, use_cuda_graph=graph).to(dev())
        with torch.no_grad():
            h.weight.copy_(torch.from_numpy(w))
        heads.append(h)
    plain, fused = heads
    opt_p = torch.optim.AdamW(plain.parameters(), lr=1e-3)
    opt_f = mm.FusedHeadAdamW.for_head(fused, lr=1e-3)
    for it in range(8):
        x, _, y = onp.synthetic_inputs(B, D, C, seed=50 + it, trained_like=False)
        losses = []
        for h, opt in ((plain, opt_p), (fused, opt_f)):
            xt = torch.from_numpy(x).to(dev())
            opt.zero_grad(set_to_none=True)
            loss, _ = h.loss(xt, torch.from_numpy(y).to(dev()))
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        assert abs(losses[0] - losses[1]) <= 1e-4 * max(1.0, abs(losses[0])), "iteration %d: %r" % (it, losses)
    rel = float((plain.weight - fused.weight).norm() / plain.weight.norm())
    assert rel <= 1e-4, rel
