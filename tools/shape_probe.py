"""fwd+bwd step time of the dense head at arbitrary shapes (the BASELINE.json configs that are parity cases, not
bench lines), with the roofline of SURVEY.md section 8d beside it.

    python tools/shape_probe.py 256,1792,100000 512,1024,125000 512,2816,100000
"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import multimodalsimilar_b200 as mm

dev = torch.device("cuda:0")
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
for spec in sys.argv[1:]:
    B, D, C = (int(v) for v in spec.split(","))
    g = torch.Generator(device=dev).manual_seed(0)
    bound = math.sqrt(6.0 / (C + D))
    h = mm.ArcMarginProduct(D, 8, s=64.0, m=0.4)
    h.out_feature = C
    h.weight = torch.nn.Parameter(torch.empty(C, D, device=dev).uniform_(-bound, bound, generator=g))
    x = torch.randn(B, D, device=dev, generator=g).requires_grad_(True)
    y = torch.randint(0, C, (B,), device=dev, generator=g)

    def step():
        x.grad = None
        h.weight.grad = None
        loss, _ = h.loss(x, y)
        loss.backward()

    for _ in range(6):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    t_tensor = 6.0 * B * D * C / (peaks["bf16_tflops_sustained"] * 1e12) * 1e3
    t_hbm = (8.0 * C * D + 8.0 * B * D + 24.0 * B) / (peaks["hbm_gbs"] * 1e9) * 1e3
    print("B=%d D=%d C=%d: %.3f ms/step, %.0f samples/s; roofline tensor %.3f ms / hbm %.3f ms -> %.1f %% of the %s bound"
          % (B, D, C, ms, B / ms * 1e3, t_tensor, t_hbm, 100 * max(t_tensor, t_hbm) / ms,
             "tensor" if t_tensor >= t_hbm else "hbm"), flush=True)
    del h, x, y
    torch.cuda.empty_cache()
