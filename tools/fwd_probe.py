"""Times the fused forward (K1(w)+K2) alone with CUDA events; ARCFACE_B200_FWD_DEBUG selects measurement modes."""
import os
import math, os, sys
os.environ.setdefault("ARCFACE_B200_DIAG", "1")   # the ARCFACE_B200_* knobs exist in the diagnostic library only
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodalsimilar_b200 import ops

dev = torch.device("cuda:0")
B, D, C = 512, 512, 1000000
g = torch.Generator(device=dev).manual_seed(0)
bound = math.sqrt(6.0 / (C + D))
w = torch.empty(C, D, device=dev).uniform_(-bound, bound, generator=g)
x = torch.randn(B, D, device=dev, generator=g)
y = torch.randint(0, C, (B,), device=dev, generator=g)
xhat, inv_nx, _ = ops.normalize_cast(x)
lm = ops.label_margin(x, w, inv_nx, None, y, 0, C, 64.0, 0.5, False)
for it in range(3):
    ops.forward_rows_fused(xhat, w, lm.label_local, 64.0, 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 10
for it in range(n):
    ops.forward_rows_fused(xhat, w, lm.label_local, 64.0, 0)
e1.record()
torch.cuda.synchronize()
print("mode", os.environ.get("ARCFACE_B200_FWD_DEBUG", "0"), os.environ.get("ARCFACE_B200_FWD_IMPL", "fused"), "ms", e0.elapsed_time(e1) / n)
