"""Reference checkpoint compatibility against fixtures minted from the UNMODIFIED
reference by tests/golden/make_ref_checkpoints.py: a whole-module pickle of `arcface.ArcMarginProduct`, a
DataParallel-style state_dict and a multi-label state_dict."""
import math
import os
import sys

import pytest
import torch

import multimodalsimilar_b200 as mm
from multimodalsimilar_b200 import checkpoint

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture()
def shim():
    saved = sys.modules.pop("arcface", None)
    mod = checkpoint.install_reference_shim()
    yield mod
    sys.modules.pop("arcface", None)
    if saved is not None:
        sys.modules["arcface"] = saved


def test_reference_module_pickle_loads_as_b200_head(shim):
    h = torch.load(os.path.join(GOLDEN, "ref_head_module.pt"), weights_only=False)
    assert type(h) is mm.ArcMarginProduct
    # attributes exactly as the reference left them (arcface.py:20-33 after update_m(+0.04))
    assert (h.in_feature, h.out_feature, h.s, h.easy_margin) == (16, 32, 30.0, False)
    assert abs(h.m - 0.54) < 1e-12
    assert abs(h.cos_m - math.cos(0.54)) < 1e-12 and abs(h.mm - math.sin(math.pi - 0.54) * 0.54) < 1e-12
    assert tuple(h.weight.shape) == (32, 16) and h.weight.requires_grad
    # attributes the reference never had resolve to the class defaults, and the methods work
    assert h.validate_labels is False and h.use_cuda_graph is True
    h.update_m(0.04)
    assert abs(h.m - 0.58) < 1e-12
    assert list(h.state_dict().keys()) == ["weight"]


def test_state_dict_head_extraction_and_load():
    sd = torch.load(os.path.join(GOLDEN, "ref_model_state.pt"))
    heads = checkpoint.head_weights(sd)
    assert list(heads) == ["classifier"] and tuple(heads["classifier"].shape) == (32, 16)
    h = mm.ArcMarginProduct(16, 32)
    mm.load_reference_head(h, sd)
    assert torch.equal(h.weight.detach(), sd["module.classifier.weight"])
    out = mm.reference_state_dict(h)
    assert list(out) == ["classifier.weight"] and torch.equal(out["classifier.weight"], sd["module.classifier.weight"])
    with pytest.raises(ValueError):
        mm.load_reference_head(mm.ArcMarginProduct(16, 40), sd)


def test_multilabel_state_dict_needs_a_name():
    sd = torch.load(os.path.join(GOLDEN, "ref_multilabel_state.pt"))
    assert sorted(checkpoint.head_weights(sd)) == ["classifier1", "classifier2", "classifier3"]
    h2 = mm.ArcMarginProduct(16, 16)
    with pytest.raises(KeyError):
        mm.load_reference_head(h2, sd)
    mm.load_reference_head(h2, sd, name="classifier2")
    assert torch.equal(h2.weight.detach(), sd["classifier2.weight"])


def test_load_from_module_and_tensor(shim):
    ref = torch.load(os.path.join(GOLDEN, "ref_head_module.pt"), weights_only=False)
    a, b = mm.ArcMarginProduct(16, 32), mm.ArcMarginProduct(16, 32)
    mm.load_reference_head(a, ref)
    mm.load_reference_head(b, ref.weight.detach())
    assert torch.equal(a.weight, ref.weight) and torch.equal(b.weight, ref.weight)


@pytest.mark.gpu
def test_unpickled_reference_head_trains_on_gpu(shim):
    import numpy as np
    from oracle import arcface_numpy as onp

    dev = torch.device("cuda:0")
    h = torch.load(os.path.join(GOLDEN, "ref_head_module.pt"), weights_only=False).to(dev)
    rng = np.random.RandomState(3)
    x = rng.standard_normal((8, 16)).astype(np.float32)
    y = rng.randint(0, 32, size=(8,)).astype(np.int64)
    xt = torch.from_numpy(x).to(dev).requires_grad_(True)
    preds = h(xt, torch.from_numpy(y).to(dev))
    loss = torch.nn.CrossEntropyLoss()(preds, torch.from_numpy(y).to(dev))
    loss.backward()
    w = h.weight.detach().cpu().numpy()
    z = onp.forward_logits(x, w, y, h.s, h.m, False, dtype=np.float64)
    from tests.test_gpu_parity import check_grad, loss_tol   # the suite's bf16 tolerance model

    ref_loss = onp.cross_entropy(z, y)
    assert abs(float(loss.detach()) - ref_loss) <= loss_tol(h.s, 16, ref_loss)
    dx, dw = onp.backward(x, w, y, h.s, h.m, False, dtype=np.float64)
    check_grad(xt.grad.cpu().numpy(), dx, h.s, 16, "dx")
    check_grad(h.weight.grad.cpu().numpy(), dw, h.s, 16, "dw")
