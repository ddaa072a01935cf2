"""Board power / SM clock (NVML) while one kernel of the step runs back to back for ~1.5 s: tells whether a kernel
is limited by the 1000 W cap (time = energy / power) or by stalls.

    python tools/power_probe.py            # fused forward, helpers only, split GEMM, backward, cuBLAS bf16 GEMM
"""
import os
import math, os, sys, threading, time
os.environ.setdefault("ARCFACE_B200_DIAG", "1")   # the ARCFACE_B200_* knobs exist in the diagnostic library only
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pynvml
import torch
from multimodalsimilar_b200 import ops

dev = torch.device("cuda:0")
B, D, C = 512, 512, 1000000
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

pynvml.nvmlInit()
vis = os.environ.get("CUDA_VISIBLE_DEVICES")
h = pynvml.nvmlDeviceGetHandleByIndex(int(vis.split(",")[0]) if vis else 0)


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.stop = threading.Event()
        self.p, self.c = [], []

    def run(self):
        while not self.stop.is_set():
            self.p.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            self.c.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            time.sleep(0.01)


def probe(name, fn, env=None, seconds=1.5):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    n = max(10, int(seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3)))
    s = Sampler(); s.start()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    s.stop.set(); s.join()
    k = len(s.p) // 3   # skip the ramp
    p, c = s.p[k:], s.c[k:]
    print("%-34s %8.4f ms/iter   power avg %6.1f W max %6.1f W   sm clock avg %6.0f MHz min %5d" % (
        name, e0.elapsed_time(e1) / n, sum(p) / len(p), max(p), sum(c) / len(c), min(c)), flush=True)
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    time.sleep(0.5)


if os.environ.get("PROBE_ONLY") == "fwd":   # A/B of forward variants: two passes over the fused forward only
    for _ in range(2):
        probe("forward fused (K1w + K2)", lambda: ops.forward_rows_fused(xhat, w, lm.label_local, 64.0, 0))
    sys.exit(0)
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
probe("cuBLAS bf16 8192^3", lambda: torch.matmul(a, b))
probe("forward fused (K1w + K2)", lambda: ops.forward_rows_fused(xhat, w, lm.label_local, 64.0, 0))
probe("forward helpers only (debug 2)", lambda: ops.forward_rows_fused(xhat, w, lm.label_local, 64.0, 0),
      {"ARCFACE_B200_FWD_DEBUG": "2"})
probe("K2 GEMM only (what given)", lambda: ops.forward_rows(xhat, what, lm.label_local, 64.0, 0))
probe("K1 standalone (weights)", lambda: ops.normalize_cast(w))
probe("backward (single launch)", lambda: ops.backward(xhat, xhat_t, what, inv_nw, lse, omp, lm.dphi, lm.label_local, 64.0,
                                                     1.0 / B, dw_out=dw))
