"""PartialFC-style class sampling: the sampling rule against its numpy restatement, the
label remap, and the whole sampled step (single process and two gloo ranks, kernels replaced by the test-only CPU
stand-in) against the reference head evaluated on the sampled rows.  The GPU kernels are covered by
tests/test_gpu_sampling.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodalsimilar_b200 import engine
from oracle import arcface_numpy as onp
from tests import _cpu_kernels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _index_like_engine(label_local, c_local, num, seed):
    """What engine.sample_classes draws from a generator seeded with `seed`, recomputed through the oracle's rule."""
    g = torch.Generator().manual_seed(seed)
    scores = torch.rand(c_local + 1, generator=g)[:c_local].numpy()
    return onp.partial_fc_sample(label_local, c_local, num, scores)


@pytest.mark.parametrize("B,C,num", [(8, 100, 20), (16, 40, 4), (5, 7, 3), (4, 1000, 1000), (6, 50, 1)])
def test_sample_rule_matches_oracle(B, C, num):
    rs = np.random.RandomState(B * 131 + C)
    lab = rs.randint(-1, C, size=B)          # -1: the label lives on another rank
    idx = engine.sample_classes(torch.from_numpy(lab), C, num, torch.Generator().manual_seed(99))
    want = _index_like_engine(lab, C, num, 99)
    np.testing.assert_array_equal(idx.numpy(), want)
    S = min(C, max(num, min(B, C)))
    assert idx.numel() == S and idx.dtype == torch.int64
    assert np.all(np.diff(idx.numpy()) > 0)                      # sorted, no duplicates
    assert set(lab[lab >= 0]).issubset(set(idx.numpy().tolist()))  # every positive is in the sample
    remap = engine.remap_labels(torch.from_numpy(lab).int(), idx).numpy()
    assert np.all(remap[lab < 0] == -1)
    np.testing.assert_array_equal(idx.numpy()[remap[lab >= 0]], lab[lab >= 0])


def test_sample_without_local_positives_and_with_tiny_shards():
    # a rank none of whose classes is a label of the batch still samples (pure negatives)
    lab = torch.full((6,), -1)
    idx = engine.sample_classes(lab, 50, 10, torch.Generator().manual_seed(1))
    assert idx.numel() == 10 and np.all(np.diff(idx.numpy()) > 0) and int(idx.min()) >= 0 and int(idx.max()) < 50
    assert np.all(engine.remap_labels(lab.int(), idx).numpy() == -1)
    # a shard smaller than the batch: every class is taken
    lab = torch.tensor([0, 2, 2, 1, -1, 0, 1, 2])
    idx = engine.sample_classes(lab, 3, 1, torch.Generator().manual_seed(1))
    assert idx.tolist() == [0, 1, 2]
    # one class
    assert engine.sample_classes(torch.tensor([0, 0]), 1, 1).tolist() == [0]


def test_sample_draws_change_and_negatives_are_uniform():
    lab = torch.tensor([3, 3, 9])
    g = torch.Generator().manual_seed(5)
    a = engine.sample_classes(lab, 1000, 100, g)
    b = engine.sample_classes(lab, 1000, 100, g)
    assert not torch.equal(a, b)
    hits = torch.zeros(1000)
    for _ in range(200):
        hits[engine.sample_classes(lab, 1000, 100, g)] += 1
    assert hits[3] == 200 and hits[9] == 200
    neg = torch.cat([hits[:3], hits[4:9], hits[10:]])
    assert abs(float(neg.mean()) - 200 * 98 / 998) < 1.0 and float(neg.max()) < 60


def test_oracle_sampled_head_with_every_class_is_the_full_head():
    x, w, y = onp.synthetic_inputs(8, 16, 32, seed=0)
    loss, arg, dx, dw = onp.sampled_head(x, w, y, np.arange(32), 30.0, 0.5)
    z = onp.forward_logits(x, w, y, 30.0, 0.5, dtype=np.float64)
    assert abs(loss - onp.cross_entropy(z, y)) < 1e-12
    np.testing.assert_array_equal(arg, onp.argmax(z))
    fdx, fdw = onp.backward(x, w, y, 30.0, 0.5)
    np.testing.assert_allclose(dx, fdx, atol=1e-12)
    np.testing.assert_allclose(dw, fdw, atol=1e-12)


@pytest.mark.parametrize("sparse", [False, True])
def test_sampled_step_matches_reference_on_the_sampled_rows(sparse):
    B, D, C, s, m = 8, 16, 60, 64.0, 0.4
    x, w, y = onp.synthetic_inputs(B, D, C, seed=3, trained_like=True)
    cfg = engine.StepConfig(s, m, False, 0, C)
    xt = torch.from_numpy(x).clone().requires_grad_(True)
    wt = torch.from_numpy(w).clone().requires_grad_(True)
    gen = torch.Generator().manual_seed(11)
    loss, pred = engine.ArcFaceCEFunction.apply(xt, wt, torch.from_numpy(y), _cpu_kernels, None, cfg, False, None, None,
                                                None, (15, gen, sparse))
    (loss * 2.0).backward()
    index = _index_like_engine(y, C, 15, 11)
    assert index.size == 15
    rloss, rarg, rdx, rdw = onp.sampled_head(x, w, y, index, s, m, False, grad_loss=2.0)
    assert abs(float(loss) - rloss) <= 1e-5 * max(1.0, abs(rloss))
    np.testing.assert_array_equal(pred.numpy(), rarg)
    np.testing.assert_allclose(xt.grad.numpy(), rdx, atol=1e-5 * max(1.0, np.abs(rdx).max()))
    gw = wt.grad
    assert gw.is_sparse == sparse
    gw = gw.to_dense() if sparse else gw
    np.testing.assert_allclose(gw.numpy(), rdw, atol=1e-5 * max(1.0, np.abs(rdw).max()))
    outside = np.setdiff1d(np.arange(C), index)
    assert np.all(gw.numpy()[outside] == 0.0)


def test_module_keywords_and_eval_mode():
    import multimodalsimilar_b200 as mm

    h = mm.ArcMarginProduct(16, 100, sample_rate=0.25, sample_seed=1)
    assert h.sample_rate == 0.25 and h.sparse_grad is False
    assert engine._sampling_for(h, 100, torch.device("cpu"))[0] == 25
    h.eval()
    assert engine._sampling_for(h, 100, torch.device("cpu")) is None      # evaluation sees every class
    h.train()
    with torch.no_grad():
        assert engine._sampling_for(h, 100, torch.device("cpu")) is None
    assert engine._sampling_for(mm.ArcMarginProduct(16, 100), 100, torch.device("cpu")) is None
    with pytest.raises(ValueError):
        mm.ArcMarginProduct(16, 100, sample_rate=0.0)
    with pytest.raises(ValueError):
        mm.ArcMarginProduct(16, 100, sample_rate=1.5)


# ------------------------------------------------------------------------------------------- two ranks (gloo)
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multimodalsimilar_b200 import ShardedArcMarginProduct

        B, D, C, s, m = 8, 16, 90, 64.0, 0.4
        x, w, y = onp.synthetic_inputs(B, D, C, seed=5, trained_like=True)
        head = ShardedArcMarginProduct(D, C, s=s, m=m, kernels=_cpu_kernels, sample_rate=0.2, sample_seed=77)
        head.load_full_weight(torch.from_numpy(w))
        b_loc = B // world
        xl = torch.from_numpy(x[rank * b_loc:(rank + 1) * b_loc]).clone().requires_grad_(True)
        yl = torch.from_numpy(y[rank * b_loc:(rank + 1) * b_loc])
        loss, pred = head.loss(xl, yl)
        loss.backward()
        c_local = head.class_hi - head.class_lo
        lab_loc = np.where((y >= head.class_lo) & (y < head.class_hi), y - head.class_lo, -1)
        index = _index_like_engine(lab_loc, c_local, int(round(0.2 * c_local)), 77 + 7919 * rank) + head.class_lo
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), loss=loss.item(), pred=pred.numpy(), dx=xl.grad.numpy(),
                 dw=head.weight.grad.numpy(), lo=head.class_lo, hi=head.class_hi, index=index)
    finally:
        dist.destroy_process_group()


def test_two_rank_sampled_head_matches_reference_on_the_union_of_samples(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    B, D, C, s, m = 8, 16, 90, 64.0, 0.4
    x, w, y = onp.synthetic_inputs(B, D, C, seed=5, trained_like=True)
    got = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    index = np.concatenate([g["index"] for g in got])        # rank order = ascending class ranges
    assert index.size == 2 * max(9, 8) and np.all(np.diff(index) > 0)
    loss, arg, dx, dw = onp.sampled_head(x, w, y, index, s, m)
    b_loc = B // world
    for r, g in enumerate(got):
        assert abs(float(g["loss"]) - loss) <= 1e-5 * max(1.0, abs(loss))
        np.testing.assert_array_equal(g["pred"], arg[r * b_loc:(r + 1) * b_loc])
        np.testing.assert_allclose(g["dx"], dx[r * b_loc:(r + 1) * b_loc], atol=1e-5 * max(1.0, np.abs(dx).max()))
        np.testing.assert_allclose(g["dw"], dw[int(g["lo"]):int(g["hi"])], atol=1e-5 * max(1.0, np.abs(dw).max()))
