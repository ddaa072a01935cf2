"""Whole head training step INCLUDING the optimiser (what a training loop pays; bench.py's metric excludes the
optimiser): head + torch.optim.AdamW against head + FusedHeadAdamW, north-star shape, CUDA events."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import multimodalsimilar_b200 as mm

dev = torch.device("cuda:0")
B, D, C = 512, 512, int(os.environ.get("C", 1000000))
g = torch.Generator(device=dev).manual_seed(0)
bound = math.sqrt(6.0 / (C + D))
x = torch.randn(B, D, device=dev, generator=g)
y = torch.randint(0, C, (B,), device=dev, generator=g)


def build(kind):
    h = mm.ArcMarginProduct(D, 8, s=64.0, m=0.5)
    h.out_feature = C
    h.weight = torch.nn.Parameter(torch.empty(C, D, device=dev).uniform_(-bound, bound, generator=g))
    if kind == "torch":
        opt = torch.optim.AdamW(h.parameters(), lr=1e-3)
    elif kind == "torch_fused":
        opt = torch.optim.AdamW(h.parameters(), lr=1e-3, fused=True)
    else:
        opt = mm.FusedHeadAdamW.for_head(h, lr=1e-3)
    return h, opt


for kind in ("torch", "torch_fused", "b200_fused"):
    h, opt = build(kind)

    def step():
        opt.zero_grad(set_to_none=True)
        loss, _ = h.loss(x, y)
        loss.backward()
        opt.step()

    for _ in range(8):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("%-12s fwd + bwd + optimiser step: %.3f ms" % (kind, e0.elapsed_time(e1) / n), flush=True)
    del h, opt
    torch.cuda.empty_cache()
