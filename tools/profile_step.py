"""Minimal driver for ncu: a few fwd+bwd steps of the head at the bench workload, nothing else.

    python tools/profile_step.py [--steps 3] [--B 512 --D 512 --C 1000000]

61 kernel launches per step at the north-star shape (K1 x2, label margin, K2, combine, finalize,
18 x (dC^T producer, dW GEMM, dX GEMM), normalise backward).
"""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--B", type=int, default=512)
    ap.add_argument("--D", type=int, default=512)
    ap.add_argument("--C", type=int, default=1000000)
    args = ap.parse_args()
    import torch

    import multimodalsimilar_b200 as mm

    dev = torch.device("cuda:0")
    B, D, C = args.B, args.D, args.C
    head = mm.ArcMarginProduct(D, 8, s=64.0, m=0.5)
    head.out_feature = C
    bound = math.sqrt(6.0 / (C + D))
    g = torch.Generator(device=dev).manual_seed(0)
    head.weight = torch.nn.Parameter(torch.empty(C, D, device=dev).uniform_(-bound, bound, generator=g))
    x = torch.randn(B, D, device=dev, generator=g).requires_grad_(True)
    y = torch.randint(0, C, (B,), device=dev, generator=g)
    for _ in range(args.steps):
        x.grad = None
        head.weight.grad = None
        loss, pred = head.loss(x, y)
        loss.backward()
    torch.cuda.synchronize()
    print("loss %.6f" % float(loss.detach()))


if __name__ == "__main__":
    main()
