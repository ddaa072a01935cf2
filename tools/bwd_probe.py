"""Times the backward (K3) alone with CUDA events under several ARCFACE_B200_BWD_* settings."""
import os
import math, os, sys
os.environ.setdefault("ARCFACE_B200_DIAG", "1")   # the ARCFACE_B200_* knobs exist in the diagnostic library only
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodalsimilar_b200 import ops

dev = torch.device("cuda:0")
B, D, C = 512, 512, int(os.environ.get("PROBE_C", "1000000"))
g = torch.Generator(device=dev).manual_seed(0)
bound = math.sqrt(6.0 / (C + D))
w = torch.empty(C, D, device=dev).uniform_(-bound, bound, generator=g)
x = torch.randn(B, D, device=dev, generator=g)
y = torch.randint(0, C, (B,), device=dev, generator=g)
xhat, inv_nx, xhat_t = ops.normalize_cast(x, want_transpose=True)
lm = ops.label_margin(x, w, inv_nx, None, y, 0, C, 64.0, 0.5, False)
what, inv_nw, rmax, rsum, rarg = ops.forward_rows_fused(xhat, w, lm.label_local, 64.0, 0)
lse, arg, zl, omp, loss = ops.finalize_rows(rmax.view(1, B), rsum.view(1, B), rarg.view(1, B), lm.z_label.view(1, B), y)
dw = torch.empty_like(w)

def run(tag, env):
    for k in ("ARCFACE_B200_BWD_IMPL", "ARCFACE_B200_BWD_SPLIT", "ARCFACE_B200_BWD_RING"):
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        for _ in range(2):
            ops.backward(xhat, xhat_t, what, inv_nw, lse, omp, lm.dphi, lm.label_local, 64.0, 1.0 / B, dw_out=dw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 5
        for _ in range(n):
            ops.backward(xhat, xhat_t, what, inv_nw, lse, omp, lm.dphi, lm.label_local, 64.0, 1.0 / B, dw_out=dw)
        e1.record()
        torch.cuda.synchronize()
        print("%-28s %.4f ms" % (tag, e0.elapsed_time(e1) / n), flush=True)
    except Exception as e:
        print(tag, "FAILED", e, flush=True)

configs = [("split", {"ARCFACE_B200_BWD_IMPL": "split"}), ("fused default", {})]
for s in sys.argv[1:]:
    parts = s.split(":")
    env = {"ARCFACE_B200_BWD_SPLIT": parts[0]}
    if len(parts) > 1:
        env["ARCFACE_B200_BWD_RING"] = parts[1]
    configs.append(("fused " + s, env))
for tag, env in configs:
    run(tag, env)
