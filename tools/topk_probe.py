"""Fused cosine top-k (K4) timing against the dense alternative (materialise B x C cosines with the logits kernel,
then torch.topk), CUDA events.   python tools/topk_probe.py [B,D,C,k ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodalsimilar_b200 import ops

dev = torch.device("cuda:0")
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    peak = 1400.0
specs = sys.argv[1:] or ["512,512,1000000,100", "512,1024,1000000,100", "512,1792,500000,13", "2048,512,1000000,100"]
for spec in specs:
    B, D, C, k = (int(v) for v in spec.split(","))
    g = torch.Generator(device=dev).manual_seed(0)
    xhat, _, _ = ops.normalize_cast(torch.randn(B, D, device=dev, generator=g))
    what, _, _ = ops.normalize_cast(torch.randn(C, D, device=dev, generator=g))

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    t_fused = timeit(lambda: ops.cosine_topk(xhat, what, k))
    t_dense = timeit(lambda: torch.topk(ops.logits(xhat, what, None, None, 1.0), k, dim=1), n=3)
    flops = 2.0 * B * D * C
    print("B=%d D=%d C=%d k=%d: fused %.3f ms (%.0f TFLOP/s algorithmic = %.0f %% of sustained bf16 peak; executes 2 GEMM "
          "passes) | logits kernel + torch.topk %.3f ms (%.1fx)" % (B, D, C, k, t_fused, flops / t_fused / 1e9,
          100 * flops / t_fused / 1e9 / peak, t_dense, t_dense / t_fused), flush=True)
    del xhat, what
    torch.cuda.empty_cache()
