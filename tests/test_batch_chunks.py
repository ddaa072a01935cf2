"""Global batches above one GEMM launch (engine.BATCH_CHUNK rows): K2 / K3 run once per row chunk, the statistics,
exchanges and the loss stay one pass, dW is summed over the chunks.  Checked with the test-only CPU stand-in kernels
(chunk size lowered so that small batches exercise the path) against the oracle; the GPU kernels are covered by
tests/test_gpu_parity.py::test_large_batch_runs_in_row_chunks."""
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


def test_chunk_ranges():
    assert engine.batch_chunks(512) == [(0, 512)]
    assert engine.batch_chunks(1024) == [(0, 1024)]
    assert engine.batch_chunks(1025) == [(0, 576), (576, 1025)]
    assert engine.batch_chunks(2048) == [(0, 1024), (1024, 2048)]
    assert engine.batch_chunks(4096) == [(0, 1024), (1024, 2048), (2048, 3072), (3072, 4096)]
    assert engine.batch_chunks(3000) == [(0, 1024), (1024, 2048), (2048, 3000)]
    assert engine.batch_chunks(4096, prec=1) == [(0, 4096)]          # the bf16x3 operands are laid out per launch
    for B in (1030, 2500, 5000, 8192):
        ch = engine.batch_chunks(B)
        assert ch[0][0] == 0 and ch[-1][1] == B and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
        assert all(b0 % 64 == 0 and 0 < b1 - b0 <= engine.BATCH_CHUNK for b0, b1 in ch)


@pytest.mark.parametrize("B,sampled", [(200, False), (130, False), (200, True), (67, False), (193, True)])
def test_chunked_step_matches_oracle(monkeypatch, B, sampled):
    monkeypatch.setattr(engine, "BATCH_CHUNK", 64)
    assert len(engine.batch_chunks(B)) >= 2
    D, C, s, m = 16, 90, 64.0, 0.4
    x, w, y = onp.synthetic_inputs(B, D, C, seed=B, trained_like=True)
    cfg = engine.StepConfig(s, m, False, 0, C)
    xt = torch.from_numpy(x).clone().requires_grad_(True)
    wt = torch.from_numpy(w).clone().requires_grad_(True)
    sampling = (30, torch.Generator().manual_seed(3), False) if sampled else None
    loss, pred = engine.ArcFaceCEFunction.apply(xt, wt, torch.from_numpy(y), _cpu_kernels, None, cfg, False, None, None,
                                                None, sampling)
    (loss * 0.5).backward()
    if sampled:
        g = torch.Generator().manual_seed(3)
        scores = torch.rand(C + 1, generator=g)[:C].numpy()
        index = onp.partial_fc_sample(y, C, 30, scores)
        rloss, rarg, rdx, rdw = onp.sampled_head(x, w, y, index, s, m, False, grad_loss=0.5)
    else:
        z = onp.forward_logits(x, w, y, s, m, False, dtype=np.float64)
        rloss, rarg = onp.cross_entropy(z, y), onp.argmax(z)
        rdx, rdw = onp.backward(x, w, y, s, m, False, grad_loss=0.5)
    assert abs(float(loss.detach()) - rloss) <= 1e-5 * max(1.0, abs(rloss))
    np.testing.assert_array_equal(pred.numpy(), rarg)
    np.testing.assert_allclose(xt.grad.numpy(), rdx, atol=1e-5 * max(1.0, np.abs(rdx).max()))
    np.testing.assert_allclose(wt.grad.numpy(), rdw, atol=1e-5 * max(1.0, np.abs(rdw).max()))


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
        from multimodalsimilar_b200 import engine as eng

        eng.BATCH_CHUNK = 64     # (spawned process: its own module state)
        B, D, C, s, m = 160, 16, 75, 64.0, 0.4
        x, w, y = onp.synthetic_inputs(B, D, C, seed=9, trained_like=True)
        head = ShardedArcMarginProduct(D, C, s=s, m=m, kernels=_cpu_kernels)
        head.load_full_weight(torch.from_numpy(w))
        b_loc = B // world
        xl = torch.from_numpy(x[rank * b_loc:(rank + 1) * b_loc]).clone().requires_grad_(True)
        yl = torch.from_numpy(y[rank * b_loc:(rank + 1) * b_loc])
        loss, pred = head.loss(xl, yl)
        loss.backward()
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), loss=loss.item(), pred=pred.numpy(), dx=xl.grad.numpy(),
                 dw=head.weight.grad.numpy(), lo=head.class_lo, hi=head.class_hi)
    finally:
        dist.destroy_process_group()


def test_two_rank_chunked_global_batch(tmp_path):
    """The PartialFC regime in miniature: the gathered batch (2 x 80 rows) exceeds one launch and the chunk boundaries
    (64, 128) do not coincide with the rank boundary (80)."""
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    B, D, C, s, m = 160, 16, 75, 64.0, 0.4
    x, w, y = onp.synthetic_inputs(B, D, C, seed=9, trained_like=True)
    z = onp.forward_logits(x, w, y, s, m, False, dtype=np.float64)
    loss = onp.cross_entropy(z, y)
    dx, dw = onp.backward(x, w, y, s, m, False)
    b_loc = B // world
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert abs(float(got["loss"]) - loss) <= 1e-5 * max(1.0, abs(loss))
        np.testing.assert_array_equal(got["pred"], onp.argmax(z)[r * b_loc:(r + 1) * b_loc])
        np.testing.assert_allclose(got["dx"], dx[r * b_loc:(r + 1) * b_loc], atol=1e-5 * max(1.0, np.abs(dx).max()))
        np.testing.assert_allclose(got["dw"], dw[int(got["lo"]):int(got["hi"])], atol=1e-5 * max(1.0, np.abs(dw).max()))
