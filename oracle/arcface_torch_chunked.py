"""Class-chunked fp32 restatement of the reference head -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.

The dense port (oracle/arcface_torch_cpu.py) follows /root/reference/arcface.py:45-63 + nn.CrossEntropyLoss +
loss.backward() op by op, which needs ~12 B x C fp32 temporaries (2 GB each at B=512, C=1M; 41 GB each at
BASELINE config 5's C=10M).  This module evaluates the SAME formulas streaming over class chunks, so that the
full-size configurations can be checked on the GPU box (`bench.py`'s in-run `parity` object, the full-size
`-m gpu` tests) without the B x C matrix:

    pass 1, per class chunk:  cos = normalize(x) @ normalize(w_chunk)^T          arcface.py:47
                              label column: phi / th / mm branch                 arcface.py:49-55
                              z = s * blend                                      arcface.py:58-61
                              running row max / sum-exp / first argmax / top-2   CrossEntropyLoss, torch.argmax
    pass 2, per class chunk:  p = exp(z - lse); dz = (p - onehot) * grad / B     CrossEntropyLoss backward
                              dcos = s * dz (label column: * d blend / d cos)    autograd through :49-61
                              dxhat += dcos @ what_chunk; dwhat = dcos^T @ xhat  F.linear backward
                              dw = (dwhat - what (what . dwhat)) / max(||w||, eps)  F.normalize backward

Everything is fp32 torch (TF32 off), on whatever device the inputs live on.  tests/test_oracle.py pins it
against the golden vectors minted from the unmodified reference (tests/golden/make_golden.py) and against the
dense port.  Only tests/ and bench.py import it, and only as the checker / the reported baseline.
"""
from __future__ import annotations

import math

import torch

EPS = 1e-12  # F.normalize default (arcface.py:47)


def _margin(t: torch.Tensor, m: float, easy_margin: bool):
    """arcface.py:49-55 on the label cosines `t` -> (blended value u, d u / d t)."""
    cos_m, sin_m = math.cos(m), math.sin(m)
    th, mm = math.cos(math.pi - m), math.sin(math.pi - m) * m
    sine = torch.sqrt(1.0 - t * t)                    # :49 (unclamped, like the reference)
    phi = t * cos_m - sine * sin_m                    # :50
    dphi = cos_m + t * sin_m / sine
    if easy_margin:                                   # :52-53
        take = t > 0
        u = torch.where(take, phi, t)
    else:                                             # :54-55
        take = (t - th) > 0
        u = torch.where(take, phi, t - mm)
    du = torch.where(take, dphi, torch.ones_like(t))
    return u, du


@torch.no_grad()
def head_step_chunked(x, w, label, s, m, easy_margin=False, grad_loss=1.0, chunk=65536, dw_range=None,
                      want_grads=True):
    """One fwd+bwd of the reference head + mean cross-entropy, streamed over class chunks.

    Returns a dict: loss (0-d), argmax int64 [B], top2_gap fp32 [B] (largest minus second-largest logit),
    z_label [B], lse [B], and -- when want_grads -- dx [B, D] and dw [hi - lo, D] for the class rows
    dw_range = (lo, hi) (default: all)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        x = x.to(torch.float32)
        w = w.to(torch.float32)
        label = label.reshape(-1).to(torch.int64)
        B, D = x.shape
        C = w.shape[0]
        dev = x.device
        nx = x.norm(dim=1, keepdim=True).clamp_min(EPS)
        xh = x / nx
        rows = torch.arange(B, device=dev)

        def chunk_logits(c0, c1):
            wc = w[c0:c1]
            nw = wc.norm(dim=1, keepdim=True).clamp_min(EPS)
            wh = wc / nw
            cos = xh @ wh.t()
            z = cos * s
            inside = (label >= c0) & (label < c1)
            r = rows[inside]
            col = label[inside] - c0
            t = cos[r, col]
            u, du = _margin(t, m, easy_margin)
            z[r, col] = u * s
            return z, cos, wh, nw, r, col, du

        run_max = torch.full((B,), -float("inf"), device=dev)
        run_2nd = torch.full((B,), -float("inf"), device=dev)
        run_sum = torch.zeros(B, device=dev)
        run_arg = torch.zeros(B, dtype=torch.int64, device=dev)
        z_label = torch.zeros(B, device=dev)
        for c0 in range(0, C, chunk):
            c1 = min(C, c0 + chunk)
            z, _, _, _, r, col, _ = chunk_logits(c0, c1)
            z_label[r] = z[r, col]
            k = min(2, c1 - c0)
            top, idx = torch.topk(z, k, dim=1)
            cmax = top[:, 0]
            carg = torch.argmax(z, dim=1) + c0        # first maximum inside the chunk, like torch.argmax
            c2nd = top[:, 1] if k == 2 else torch.full_like(cmax, -float("inf"))
            new_max = torch.maximum(run_max, cmax)
            run_2nd = torch.maximum(torch.minimum(run_max, cmax), torch.maximum(run_2nd, c2nd))
            run_sum = run_sum * torch.exp(run_max - new_max) + torch.exp(z - new_max[:, None]).sum(dim=1)
            run_arg = torch.where(cmax > run_max, carg, run_arg)   # strict: an earlier equal maximum wins
            run_max = new_max
        lse = run_max + torch.log(run_sum)
        loss = (lse - z_label).mean()
        out = {"loss": loss, "argmax": run_arg, "top2_gap": run_max - run_2nd, "z_label": z_label, "lse": lse}
        if not want_grads:
            return out

        lo, hi = (0, C) if dw_range is None else dw_range
        dxh = torch.zeros(B, D, device=dev)
        dw = torch.empty(hi - lo, D, device=dev)
        g = float(grad_loss) / B
        for c0 in range(0, C, chunk):
            c1 = min(C, c0 + chunk)
            z, cos, wh, nw, r, col, du = chunk_logits(c0, c1)
            dz = torch.exp(z - lse[:, None]) * g
            dz[r, col] -= g
            dcos = dz * s
            dcos[r, col] *= du
            dxh += dcos @ wh
            a, b = max(c0, lo), min(c1, hi)
            if a < b:
                sl = slice(a - c0, b - c0)
                dwh = dcos[:, sl].t() @ xh
                whs = wh[sl]
                dw[a - lo:b - lo] = (dwh - whs * (whs * dwh).sum(dim=1, keepdim=True)) / nw[sl]
        out["dx"] = (dxh - xh * (xh * dxh).sum(dim=1, keepdim=True)) / nx
        out["dw"] = dw
        return out
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
