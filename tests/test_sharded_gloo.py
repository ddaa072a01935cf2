"""World-size-2 run of the class-sharded head on CPU (gloo): the collective choreography of
multimodalsimilar_b200/sharded.py (all-gather embeddings, one statistics exchange, reduce-scatter of the
embedding gradient) against the single-process oracle.  The kernel sequence is a test-only CPU stand-in
(tests/_cpu_kernels.py); the GPU tests cover the real kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.manual_seed(1234)   # the usual same-seed-on-all-ranks setup
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multimodalsimilar_b200 import ShardedArcMarginProduct
        from oracle import arcface_numpy as onp
        from tests import _cpu_kernels

        B, D, C, s, m, easy, trained, grad = case
        x, w, y = onp.synthetic_inputs(B, D, C, seed=7, trained_like=trained)
        head = ShardedArcMarginProduct(D, C, s=s, m=m, easy_margin=easy, kernels=_cpu_kernels)
        # fresh shards of ranks that share a seed must not start as copies of each other
        peers = [torch.empty_like(head.weight) for _ in range(world)] if (C % world == 0) else None
        if peers is not None:
            dist.all_gather(peers, head.weight.detach())
            assert not torch.equal(peers[0], peers[1])
        head.load_full_weight(torch.from_numpy(w))
        assert torch.equal(head.gather_weight(), torch.from_numpy(w))
        # torch.save(model) must keep working (the reference checkpoints whole modules, nlp_classifier_train.py:159)
        import io

        buf = io.BytesIO()
        torch.save(head, buf)
        buf.seek(0)
        again = torch.load(buf, weights_only=False)
        assert torch.equal(again.weight, head.weight) and again.class_lo == head.class_lo
        b_loc = B // world
        xl = torch.from_numpy(x[rank * b_loc:(rank + 1) * b_loc]).clone().requires_grad_(True)
        yl = torch.from_numpy(y[rank * b_loc:(rank + 1) * b_loc])
        loss, pred = head.loss(xl, yl)
        (loss * grad).backward()
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), loss=loss.item(), pred=pred.numpy(), dx=xl.grad.numpy(),
                 dw=head.weight.grad.numpy(), lo=head.class_lo, hi=head.class_hi)
    finally:
        dist.destroy_process_group()


CASES = [
    (8, 16, 37, 30.0, 0.5, False, False, 1.0),   # ragged class split (19 + 18)
    (8, 16, 32, 64.0, 0.4, False, True, 1.0),    # trained-like: phi branch, separated argmax
    (6, 24, 50, 64.0, 0.2, True, False, 10.0),   # easy margin, upstream grad != 1
]


@pytest.mark.parametrize("case", CASES)
def test_two_rank_sharded_head_matches_oracle(case, tmp_path):
    from oracle import arcface_numpy as onp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    B, D, C, s, m, easy, trained, grad = case
    x, w, y = onp.synthetic_inputs(B, D, C, seed=7, trained_like=trained)
    z = onp.forward_logits(x, w, y, s, m, easy, dtype=np.float64)
    loss = onp.cross_entropy(z, y)
    dx, dw = onp.backward(x, w, y, s, m, easy, grad_loss=grad, dtype=np.float64)
    b_loc = B // world
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert abs(float(got["loss"]) - loss) <= 1e-5 * max(1.0, abs(loss))
        np.testing.assert_array_equal(got["pred"], onp.argmax(z)[r * b_loc:(r + 1) * b_loc])
        np.testing.assert_allclose(got["dx"], dx[r * b_loc:(r + 1) * b_loc], rtol=0, atol=1e-5 * max(1.0, np.abs(dx).max()))
        lo, hi = int(got["lo"]), int(got["hi"])
        assert (lo, hi) == ((0, (C + 1) // 2) if r == 0 else ((C + 1) // 2, C))
        np.testing.assert_allclose(got["dw"], dw[lo:hi], rtol=0, atol=1e-5 * max(1.0, np.abs(dw).max()))
