"""Forward / backward kernel times per shape and per kernel-selection knob (diagnostic library: ARCFACE_B200_DIAG=1).

    python tools/stage_probe.py 512,1024,125000 "256,1792,100000@11,49,12;9,49,16" [--splits "24,32,16;30,28,16"] [--prof]

For every shape: the fused forward (K1(w)+K2) with the CTA-pair kernels and with the one-CTA streaming kernel, and the
backward as the single launch (default role split, then every --splits entry), as three CTA-pair launches, and as the
generic streaming kernels.  CUDA events, 3 warm-up + 10 timed launches each.
"""
import math
import os
import sys

os.environ["ARCFACE_B200_DIAG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from multimodalsimilar_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
splits = []
prof = "--prof" in sys.argv
for i, a in enumerate(sys.argv):
    if a == "--splits":
        splits = sys.argv[i + 1].split(";")
        args.remove(sys.argv[i + 1])


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def setenv(**kw):
    for k in ("ARCFACE_B200_FWD_IMPL", "ARCFACE_B200_BWD_IMPL", "ARCFACE_B200_BWD_SPLIT", "ARCFACE_B200_BWD_PROF",
              "ARCFACE_B200_FWD_PROF", "ARCFACE_B200_BWD_DX"):
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ["ARCFACE_B200_" + k] = v


for spec in args:
    shape_splits = list(splits)
    if "@" in spec:   # B,D,C@a,b,c;a,b,c : role splits tried for this shape only
        spec, sp = spec.split("@")
        shape_splits += sp.split(";")
    B, D, C = (int(v) for v in spec.split(","))
    g = torch.Generator(device=dev).manual_seed(0)
    bound = math.sqrt(6.0 / (C + D))
    w = torch.empty(C, D, device=dev).uniform_(-bound, bound, generator=g)
    x = torch.randn(B, D, device=dev, generator=g)
    y = torch.randint(0, C, (B,), device=dev, generator=g)
    xhat, inv_nx, xhat_t = ops.normalize_cast(x, want_transpose=True)
    lm = ops.label_margin(x, w, inv_nx, None, y, 0, C, 64.0, 0.5, False)
    out = {}
    for name, env in (("pairs", {}), ("generic", {"FWD_IMPL": "generic"})):
        setenv(**env)
        out["fwd " + name] = timed(lambda: ops.forward_rows_fused(xhat, w, lm.label_local, 64.0, 0))
    setenv()
    what, inv_nw, rmax, rsum, rarg = ops.forward_rows_fused(xhat, w, lm.label_local, 64.0, 0)
    out["k1(w) alone"] = timed(lambda: ops.normalize_cast(w))
    out["k2 alone (pairs)"] = timed(lambda: ops.forward_rows(xhat, what, lm.label_local, 64.0, 0))
    lse, arg, zl, omp, loss = ops.finalize_rows(rmax.view(1, B), rsum.view(1, B), rarg.view(1, B), lm.z_label.view(1, B), y)
    dw = torch.empty_like(w)

    def bwd():
        ops.backward(xhat, xhat_t, what, inv_nw, lse, omp, lm.dphi, lm.label_local, 64.0, 1.0 / B, dw_out=dw)

    variants = [("fused default", {}), ("fused, dX on single CTAs", {"BWD_DX": "cta"})] + \
               [("fused " + sp, {"BWD_SPLIT": sp}) for sp in shape_splits] + \
               [("3 pair launches", {"BWD_IMPL": "split"}), ("generic", {"BWD_IMPL": "generic"})]
    for name, env in variants:
        setenv(**env)
        out["bwd " + name] = timed(bwd)
        if prof and name.startswith("fused"):
            os.environ["ARCFACE_B200_BWD_PROF"] = "1"
            print("== %s %s" % (spec, name), file=sys.stderr, flush=True)
            bwd()
            torch.cuda.synchronize()
    setenv()
    t_tensor = 2.0 * B * D * C / 1653.2e12 * 1e3
    print("B=%d D=%d C=%d (one GEMM at the burst bf16 peak: %.3f ms; W fp32: %.3f ms of HBM)" %
          (B, D, C, t_tensor, C * D * 4 / 6530.3e9 * 1e3))
    for k, v in out.items():
        print("   %-28s %.3f ms" % (k, v), flush=True)
    del w, dw, what
    torch.cuda.empty_cache()
