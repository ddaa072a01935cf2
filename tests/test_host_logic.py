"""Host-side mirror of the reference interface (arcface.py:17-67): constructor, attributes, margin
schedule, state_dict / pickle round trips, shard partition.  CPU only -- no kernel is launched."""
import io
import math
import pickle

import numpy as np
import pytest
import torch
from torch import nn

import multimodalsimilar_b200 as mm
from multimodalsimilar_b200 import ops
from multimodalsimilar_b200.head import FusedLogits


def test_constructor_matches_reference_signature():
    h = mm.ArcMarginProduct(16, 32)
    assert (h.in_feature, h.out_feature, h.s, h.m, h.easy_margin) == (16, 32, 64.0, 0.40, False)
    assert tuple(h.weight.shape) == (32, 16) and h.weight.dtype == torch.float32
    assert [n for n, _ in h.named_parameters()] == ["weight"]
    assert list(h.state_dict().keys()) == ["weight"]
    assert repr(h) == "ArcMarginProduct()"
    # keyword spellings: singular (multimodal_classifier.py:22) and plural (BASELINE.json)
    a = mm.ArcMarginProduct(in_feature=24, out_feature=7, m=0.5)
    b = mm.ArcMarginProduct(in_features=24, out_features=7, m=0.5)
    assert (a.in_feature, a.out_feature) == (b.in_feature, b.out_feature) == (24, 7)
    # xavier_uniform_ bound (arcface.py:25)
    bound = math.sqrt(6.0 / (32 + 16))
    assert float(h.weight.abs().max()) <= bound + 1e-7


def test_margin_constants_and_update_m(golden):
    h = mm.ArcMarginProduct(16, 32, s=30.0, m=0.5)
    h.update_m(0.04)
    u = golden["update_m"]
    assert [h.m, h.cos_m, h.sin_m, h.th, h.mm] == list(u[:5])
    h.update_m(2.0)
    assert h.m == u[5]
    h2 = mm.ArcMarginProduct(16, 32, s=30.0, m=0.5)
    h2.update_m(-0.6)
    assert h2.m == u[6] == 0.5
    h3 = mm.ArcMarginProduct(16, 32)
    assert [h3.cos_m, h3.sin_m, h3.th, h3.mm] == list(golden["const_m04"])


def test_pickle_and_state_dict_round_trip():
    h = mm.ArcMarginProduct(16, 32, s=30.0, m=0.5, easy_margin=True)
    buf = io.BytesIO()
    torch.save(h, buf)  # whole-module pickle, as nlp_classifier_train.py:159 does
    buf.seek(0)
    h2 = torch.load(buf, weights_only=False)
    assert torch.equal(h.weight, h2.weight) and h2.easy_margin and h2.m == 0.5
    h3 = mm.ArcMarginProduct(16, 32)
    h3.load_state_dict(h.state_dict())
    assert torch.equal(h3.weight, h.weight)
    assert pickle.loads(pickle.dumps(h)).s == 30.0


def test_optimizer_sees_the_head_parameters():
    h = mm.ArcMarginProduct(16, 32)
    opt = torch.optim.AdamW(h.parameters(), lr=1e-2)  # nlp_classifier_train.py:94
    assert sum(p.numel() for g in opt.param_groups for p in g["params"]) == 32 * 16


def test_cpu_inputs_are_refused_loudly():
    h = mm.ArcMarginProduct(16, 32)
    x = torch.randn(4, 16)
    y = torch.randint(0, 32, (4,))
    with pytest.raises(RuntimeError, match="no CPU path"):
        h.loss(x, y)
    with pytest.raises(RuntimeError, match="no CPU path"):
        nn.CrossEntropyLoss()(h(x, y), y)


def test_forward_returns_lazy_handle_with_logit_shape():
    h = mm.ArcMarginProduct(16, 32)
    p = h(torch.randn(4, 16), torch.randint(0, 32, (4,)))
    assert isinstance(p, FusedLogits) and tuple(p.shape) == (4, 32) and p.size(1) == 32 and p.dim() == 2


def test_shard_range_partitions_the_classes():
    for C, R in [(10, 2), (1000000, 8), (10205, 8), (37, 4), (8, 8)]:
        spans = [mm.shard_range(C, R, r) for r in range(R)]
        assert spans[0][0] == 0 and spans[-1][1] == C
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and a <= b and c <= d
        assert sum(b - a for a, b in spans) == C


def test_margin_constants_helper():
    for m in (0.1, 0.2, 0.4, 0.5, 0.54):
        cm, sm_, th, mmv = ops.margin_constants(m)
        assert (cm, sm_, th, mmv) == (math.cos(m), math.sin(m), math.cos(math.pi - m), math.sin(math.pi - m) * m)
        assert np.isclose(th, -cm) and np.isclose(mmv, m * sm_)
