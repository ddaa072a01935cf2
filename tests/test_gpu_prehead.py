"""SURVEY section 8f row N4 on the GPU: the fused two-stream normalise-concat in front of the multimodal head
(multimodal_classifier.py:50-56) and several heads sharing one embedding (nlp_classifier_multilabel.py:33-35 with the
10 / 5 / 1 loss weights of nlp_classifier_train_daodian_v3_dist.py:164-166), each against plain fp32 PyTorch / the
separate single-head modules."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import arcface_numpy as onp

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("B,D1,D2", [(64, 1792, 1024), (37, 512, 768), (5, 4, 8)])
def test_two_stream_embed_matches_reference_ops(B, D1, D2):
    import multimodalsimilar_b200 as mm

    g = torch.Generator().manual_seed(B + D1)
    img = (torch.randn(B, D1, generator=g) * 3).to(dev()).requires_grad_(True)
    txt = (torch.randn(B, D2, generator=g) * 0.2).to(dev()).requires_grad_(True)
    with torch.no_grad():
        img[B // 2] = 0.0   # zero-norm stream -> zeros (eps clamp), finite gradients
    up = torch.randn(B, D1 + D2, generator=g).to(dev())
    emb = mm.two_stream_embed(img, txt)
    (emb * up).sum().backward()
    i2 = img.detach().clone().requires_grad_(True)
    t2 = txt.detach().clone().requires_grad_(True)
    ref = torch.cat((F.normalize(i2, p=2, dim=1), F.normalize(t2, p=2, dim=1)), 1)   # multimodal_classifier.py:53-55
    (ref * up).sum().backward()
    torch.testing.assert_close(emb, ref, rtol=2e-6, atol=1e-7)
    keep = torch.arange(B, device=dev()) != B // 2
    torch.testing.assert_close(img.grad[keep], i2.grad[keep], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(txt.grad, t2.grad, rtol=1e-4, atol=1e-5)
    assert bool(torch.isfinite(img.grad).all())


def test_two_stream_head_end_to_end():
    """The multimodal model's head path: two_stream_embed -> ArcMarginProduct(m=0.5) -> CE, against the fp32 oracle."""
    import multimodalsimilar_b200 as mm
    from oracle import arcface_torch_chunked as och

    B, D1, D2, C = 96, 256, 128, 2000
    g = torch.Generator().manual_seed(5)
    img = torch.randn(B, D1, generator=g).to(dev()).requires_grad_(True)
    txt = torch.randn(B, D2, generator=g).to(dev()).requires_grad_(True)
    y = torch.randint(0, C, (B,), generator=g).to(dev())
    head = mm.ArcMarginProduct(in_feature=D1 + D2, out_feature=C, m=0.5).to(dev())
    loss, pred = head.loss(mm.two_stream_embed(img, txt), y)
    loss.backward()
    i2 = img.detach().clone().requires_grad_(True)
    t2 = txt.detach().clone().requires_grad_(True)
    emb = torch.cat((F.normalize(i2), F.normalize(t2)), 1)
    r = och.head_step_chunked(emb.detach(), head.weight.detach(), y, 64.0, 0.5, False)
    emb.backward(r["dx"])
    assert abs(float(loss) - float(r["loss"])) <= 1e-3 * float(r["loss"])
    assert float((img.grad - i2.grad).norm() / i2.grad.norm()) <= 3e-2
    assert float((txt.grad - t2.grad).norm() / t2.grad.norm()) <= 3e-2


@pytest.mark.parametrize("graph", [False, True])
def test_multi_head_equals_separate_heads(graph):
    import multimodalsimilar_b200 as mm

    B, D = 128, 768
    Cs, ms, weights = (38, 590, 10205), (0.4, 0.2, 0.1), (10.0, 5.0, 1.0)
    g = torch.Generator().manual_seed(1)
    heads = [mm.ArcMarginProduct(D, c, m=m, use_cuda_graph=False).to(dev()) for c, m in zip(Cs, ms)]
    multi = mm.MultiHeadArcFace([mm.ArcMarginProduct(D, c, m=m).to(dev()) for c, m in zip(Cs, ms)], weights,
                                use_cuda_graph=graph)
    for a, b in zip(heads, multi.heads):
        with torch.no_grad():
            b.weight.copy_(a.weight)
    for it in range(5):   # graph mode: two eager calls, the capture, replays
        x = torch.randn(B, D, generator=g).to(dev())
        ys = [torch.randint(0, c, (B,), generator=g).to(dev()) for c in Cs]
        up = 1.0 if it % 2 == 0 else 0.5
        xa = x.clone().requires_grad_(True)
        total_ref = 0
        preds_ref = []
        for h, y, wgt in zip(heads, ys, weights):
            h.weight.grad = None
            l, p = h.loss(xa, y)
            total_ref = total_ref + wgt * l            # nlp_classifier_train_daodian_v3_dist.py:164-166
            preds_ref.append(p)
        (total_ref * up).backward()
        xb = x.clone().requires_grad_(True)
        for h in multi.heads:
            h.weight.grad = None
        total, losses, preds = multi.loss(xb, ys)
        (total * up).backward()
        assert abs(float(total) - float(total_ref)) <= 1e-5 * abs(float(total_ref))
        for p, q in zip(preds, preds_ref):
            assert torch.equal(p, q)
        assert float((xb.grad - xa.grad).norm() / xa.grad.norm()) <= 6e-3   # bf16 rounding of dC after a different scale
        for h, m in zip(heads, multi.heads):
            assert float((m.weight.grad - h.weight.grad).norm() / h.weight.grad.norm()) <= 6e-3
    if graph:
        assert multi._state["plan"] is not None
