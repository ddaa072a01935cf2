"""Eager vs CUDA-graph step time of the dense head at the bench shape, alternating in one process."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import multimodalsimilar_b200 as mm
from multimodalsimilar_b200 import engine

dev = torch.device("cuda:0")
B, D, C = 512, 512, int(os.environ.get("C", 1000000))
g = torch.Generator(device=dev).manual_seed(0)
bound = math.sqrt(6.0 / (C + D))
w = torch.empty(C, D, device=dev).uniform_(-bound, bound, generator=g)
heads = {}
for name, graph in (("eager", False), ("graph", True)):
    h = mm.ArcMarginProduct(D, 8, s=64.0, m=0.5, use_cuda_graph=graph)
    h.out_feature = C
    h.weight = torch.nn.Parameter(w)   # shared storage: the probe never updates it
    heads[name] = h
x = torch.randn(B, D, device=dev, generator=g).requires_grad_(True)
y = torch.randint(0, C, (B,), device=dev, generator=g)


def run(h, n):
    for _ in range(n):
        x.grad = None
        h.weight.grad = None
        loss, pred = h.loss(x, y)
        loss.backward()


for h in heads.values():
    run(h, 6)
torch.cuda.synchronize()
plan = engine._PLANS[heads["graph"]]["plan"]
print("dw adopted without copy:", heads["graph"].weight.grad.data_ptr() == plan.dw.data_ptr())
for rep in range(3):
    for name, h in heads.items():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        run(h, 30)
        e1.record()
        torch.cuda.synchronize()
        print(rep, name, "ms/step %.4f" % (e0.elapsed_time(e1) / 30))
