"""GPU parity of PartialFC-style class sampling: the gathered K1 and the row scatter
against their definitions, and the sampled head (ArcMarginProduct(sample_rate=...)) against the reference head
evaluated on the sampled rows (oracle.sampled_head) for the very sample the step drew."""
import numpy as np
import pytest
import torch

import multimodalsimilar_b200 as mm
from multimodalsimilar_b200 import ops
from oracle import arcface_numpy as onp
from tests.test_gpu_parity import check_grad, loss_tol

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("C,D,S", [(1000, 512, 257), (333, 1024, 333), (5000, 64, 1), (700, 2816, 90), (400, 3200, 33)])
def test_normalize_cast_gather_is_k1_of_the_gathered_rows(C, D, S):
    g = torch.Generator(device=dev()).manual_seed(C + D)
    w = torch.randn(C, D, device=dev(), generator=g)
    w[C // 2] = 0.0                                                  # a zero row stays finite
    index = torch.randperm(C, device=dev(), generator=g)[:S].sort().values
    if S > 1:
        index[0] = C // 2
        index = index.unique()
    got, inv = ops.normalize_cast_gather(w, index)
    ref, rinv, _ = ops.normalize_cast(w[index].contiguous())
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16)) and torch.equal(inv, rinv)


def test_scatter_rows():
    g = torch.Generator(device=dev()).manual_seed(1)
    src = torch.randn(300, 512, device=dev(), generator=g)
    index = torch.randperm(5000, device=dev(), generator=g)[:300].sort().values
    dst = ops.scatter_rows(src, index, torch.zeros(5000, 512, device=dev()))
    assert torch.equal(dst, torch.zeros(5000, 512, device=dev()).index_copy_(0, index, src))
    with pytest.raises(ValueError):
        ops.scatter_rows(src, index[:10], dst)


CASES = [
    # B, D, C, s, m, easy, rate, trained_like
    (64, 128, 3000, 64.0, 0.5, False, 0.1, True),
    (64, 128, 3000, 64.0, 0.5, False, 0.1, False),
    (256, 512, 20000, 64.0, 0.4, False, 0.25, True),
    (48, 1024, 5000, 64.0, 0.2, True, 0.05, True),
    (200, 64, 150, 30.0, 0.5, False, 0.2, False),      # more rows than sampled classes: S = min(B, C)
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("sparse", [False, True])
def test_sampled_head_matches_reference_on_its_sample(case, sparse):
    B, D, C, s, m, easy, rate, trained = case
    x, w, y = onp.synthetic_inputs(B, D, C, seed=B + C, trained_like=trained)
    head = mm.ArcMarginProduct(D, C, s=s, m=m, easy_margin=easy, sample_rate=rate, sample_seed=5, sparse_grad=sparse).to(dev())
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(w))
    xt = torch.from_numpy(x).to(dev()).requires_grad_(True)
    yt = torch.from_numpy(y).to(dev())
    for step in range(2):       # the second step draws a different sample
        xt.grad = None
        head.weight.grad = None
        preds = head(xt, yt)
        loss = torch.nn.CrossEntropyLoss()(preds, yt)
        (loss * 3.0).backward()
        pred = torch.argmax(preds, dim=-1)
        index = head.last_sample_index().cpu().numpy()
        S = min(C, max(int(round(rate * C)), min(B, C)))
        assert index.size == S and np.all(np.diff(index) > 0) and set(y.tolist()).issubset(set(index.tolist()))
        if step == 0:
            first = index
        else:
            assert S == C or not np.array_equal(index, first)
        rloss, rarg, rdx, rdw = onp.sampled_head(x, w, y, index, s, m, easy, grad_loss=3.0)
        assert abs(float(loss.detach()) - rloss) <= loss_tol(s, D, rloss)
        zs = onp.forward_logits(x, w[index], np.searchsorted(index, y), s, m, easy, dtype=np.float64)
        top2 = np.sort(zs, axis=1)[:, -2:]
        sep = (top2[:, 1] - top2[:, 0]) > 0.2 * (s / 30.0)
        np.testing.assert_array_equal(pred.cpu().numpy()[sep], rarg[sep])
        gdx = xt.grad.cpu().numpy()
        gw = head.weight.grad
        assert gw.is_sparse == sparse
        gdw = (gw.to_dense() if sparse else gw).cpu().numpy()
        check_grad(gdx / 3.0, rdx / 3.0, s, D, "dx")     # the suite's gradient gate, per unit of upstream gradient
        check_grad(gdw / 3.0, rdw / 3.0, s, D, "dw")
        outside = np.setdiff1d(np.arange(C), index)
        assert np.all(gdw[outside] == 0.0)


def test_eval_mode_and_full_rate_use_every_class():
    B, D, C = 32, 128, 2000
    x, w, y = onp.synthetic_inputs(B, D, C, seed=1, trained_like=True)
    xt, yt = torch.from_numpy(x).to(dev()), torch.from_numpy(y).to(dev())
    full = mm.ArcMarginProduct(D, C).to(dev())
    samp = mm.ArcMarginProduct(D, C, sample_rate=0.1).to(dev())
    with torch.no_grad():
        full.weight.copy_(torch.from_numpy(w))
        samp.weight.copy_(torch.from_numpy(w))
    samp.eval()
    l0, p0 = full.loss(xt, yt)
    l1, p1 = samp.loss(xt, yt)
    assert float(l0) == float(l1) and torch.equal(p0, p1) and samp.last_sample_index() is None
