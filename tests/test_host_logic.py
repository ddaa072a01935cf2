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


def test_fused_optimizer_rejects_what_it_cannot_update():
    """FusedHeadAdamW is for fp32 [C, D] CUDA head weights; anything else is an error, not a silent fallback."""
    import multimodalsimilar_b200 as mm

    p = torch.nn.Parameter(torch.zeros(8, 16))
    opt = mm.FusedHeadAdamW([p], lr=1e-2)
    assert opt.param_groups[0]["lr"] == 1e-2 and opt.param_groups[0]["betas"] == (0.9, 0.999)
    opt.step()                       # no gradient yet: nothing to do
    p.grad = torch.zeros_like(p)
    with pytest.raises(RuntimeError, match="CUDA head weights"):
        opt.step()
    with pytest.raises(ValueError):
        mm.FusedHeadAdamW([p], lr=-1.0)
    # LR schedulers drive it like any torch optimiser
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda step: 0.5)
    assert abs(opt.param_groups[0]["lr"] - 5e-3) < 1e-12 and sched is not None


def test_cosine_index_argument_checks():
    import multimodalsimilar_b200 as mm

    assert mm.CosineIndex(100, device="cpu").dp == 104   # any width: rows are zero-padded to a multiple of 8
    with pytest.raises(ValueError):
        mm.CosineIndex(0)
    with pytest.raises(ValueError):
        mm.CosineIndex(16, precision="fp64")
    index = mm.CosineIndex(16, device="cpu")
    assert index.ntotal == 0
    with pytest.raises(ValueError):
        index.search(torch.zeros(2, 16), 0)
    with pytest.raises(RuntimeError, match="empty"):
        index.search(torch.zeros(2, 16), 5)
    with pytest.raises(RuntimeError, match="empty"):
        index.search(torch.zeros(2, 16), 500)   # k > 128 is served (materialising path), not rejected
    with pytest.raises(RuntimeError, match="CUDA"):
        index.add(torch.zeros(4, 16))   # there is no CPU path


def test_stand_in_shard_merge_equals_global_statistics():
    """The per-shard statistics + finalize merge (the contract of include/arcface_b200.h, restated by the test-only
    stand-in) reproduce the oracle's global loss / argmax for any contiguous class split."""
    import numpy as np

    from oracle import arcface_numpy as onp
    from tests import _cpu_kernels as K

    B, D, C, s, m = 6, 16, 41, 30.0, 0.5
    x, w, y = onp.synthetic_inputs(B, D, C, seed=3, trained_like=False)
    z = onp.forward_logits(x, w, y, s, m, False, dtype=np.float64)
    for R in (1, 2, 3, 5):
        bounds = [round(i * C / R) for i in range(R + 1)]
        xt, yt = torch.from_numpy(x), torch.from_numpy(y)
        xhat, inv_nx, _ = K.normalize_cast(xt)
        rows = []
        for r in range(R):
            lo, hi = bounds[r], bounds[r + 1]
            ws = torch.from_numpy(w[lo:hi])
            lm = K.label_margin(xt, ws, inv_nx, None, yt, lo, C, s, m, False)
            what, inv_nw, rmax, rsum, rarg = K.forward_rows_fused(xhat, ws, lm.label_local, s, lo)
            rows.append((rmax, rsum, lm.z_label, rarg))
        lse, arg, zl, omp, loss = K.finalize_rows(torch.stack([t[0] for t in rows]), torch.stack([t[1] for t in rows]),
                                                  torch.stack([t[3] for t in rows]), torch.stack([t[2] for t in rows]), yt)
        assert abs(float(loss) - onp.cross_entropy(z, y)) <= 1e-5
        np.testing.assert_array_equal(arg.numpy(), onp.argmax(z))


def test_bench_weights_do_not_depend_on_the_rank_count():
    """bench.py's synthetic weight matrix is a function of (seed, class id) only: any rank's shard is a slice of the
    matrix a single GPU generates, which is what lets the driver compare losses and parity across N = 1, 2, 4, 8."""
    import bench
    import multimodalsimilar_b200 as mm

    C, D = 3 * bench.WEIGHT_BLOCK + 1234, 16
    cpu = torch.device("cpu")
    full = bench.class_weights(torch, C, D, 0, C, cpu)
    assert full.shape == (C, D)
    assert float(full.abs().max()) <= (6.0 / (C + D)) ** 0.5 * (1 + 1e-6)
    for world in (2, 3, 8):
        for rank in range(world):
            lo, hi = mm.shard_range(C, world, rank)
            assert torch.equal(bench.class_weights(torch, C, D, lo, hi, cpu), full[lo:hi])
    # ragged slices that start and end inside a block
    assert torch.equal(bench.class_weights(torch, C, D, 70000, 140001, cpu), full[70000:140001])


def test_bench_config_table_matches_baseline():
    """The configurations bench.py measures are BASELINE.json's."""
    import bench

    assert bench.CONFIGS["ns"] == dict(bench.CONFIGS["ns"], B=512, D=512, C=1000000, s=64.0, m=0.5)
    assert (bench.CONFIGS["c1"]["B"], bench.CONFIGS["c1"]["D"], bench.CONFIGS["c1"]["C"], bench.CONFIGS["c1"]["s"]) == (64, 512, 1000, 30.0)
    assert (bench.CONFIGS["c2"]["B"], bench.CONFIGS["c2"]["D"], bench.CONFIGS["c2"]["C"]) == (256, 1792, 100000)
    assert (bench.CONFIGS["c3"]["B"], bench.CONFIGS["c3"]["D"], bench.CONFIGS["c3"]["C"]) == (512, 1024, 1000000)
    assert (bench.CONFIGS["c4"]["B"], bench.CONFIGS["c4"]["D"], bench.CONFIGS["c4"]["C"]) == (512, 2816, 1000000)
    assert [bench.CONFIGS[k]["C"] for k in ("c5_100k", "c5_1m", "c5_10m")] == [100000, 1000000, 10000000]
    assert all(bench.CONFIGS[k]["B"] == 1024 and bench.CONFIGS[k]["D"] == 512 for k in ("c5_100k", "c5_1m", "c5_10m"))
    assert set(bench.EXTRA_CONFIGS) <= set(bench.CONFIGS)
