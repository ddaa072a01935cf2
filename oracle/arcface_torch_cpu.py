"""Timed CPU baseline for bench.py -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's hot path is "PyTorch eager fp32 on the host" (arcface.py:45-63 + nn.CrossEntropyLoss +
loss.backward()).  /root/reference does not exist on the GPU box, so bench.py's `cpu_baseline` leg and
`--impl reference` arm time this port instead (kind = "port"): the same dense op sequence -- every
element of the B x C cosine matrix goes through the margin formula and the one-hot blend, autograd does
the backward -- on all host threads torch will use.  tests/test_oracle.py checks it against the golden
vectors generated from the real reference, so it computes what the reference computes.

Only bench.py and tests/ import this module.
"""
from __future__ import annotations

import math
import time

import torch
import torch.nn.functional as F


def head_step(x: torch.Tensor, w: torch.Tensor, label: torch.Tensor, s: float, m: float, easy_margin: bool = False):
    """One fwd+bwd of the dense fp32 head + mean cross-entropy.  Returns (loss, argmax, dx, dw)."""
    x = x.detach().clone().requires_grad_(True)
    w = w.detach().clone().requires_grad_(True)
    c = F.linear(F.normalize(x), F.normalize(w))          # arcface.py:47
    sn = (1.0 - c.pow(2)).sqrt()                            # :49
    ph = c * math.cos(m) - sn * math.sin(m)                 # :50
    if easy_margin:                                         # :52-55
        ph = torch.where(c > 0, ph, c)
    else:
        ph = torch.where(c - math.cos(math.pi - m) > 0, ph, c - math.sin(math.pi - m) * m)
    hot = torch.zeros_like(c).scatter_(1, label.view(-1, 1), 1)   # :58-59
    z = (hot * ph + (1.0 - hot) * c) * s                    # :60-61
    loss = F.cross_entropy(z, label.view(-1))
    pred = torch.argmax(z, dim=-1)
    loss.backward()
    return loss.detach(), pred, x.grad, w.grad


def time_head_step(B, D, C, s, m, budget_s=20.0, max_steps=5, seed=0):
    """Times head_step on a bounded sample of the (B, D, C) workload: the class count is cut to what one
    step finishes in ~budget_s/3 of host time (cost is linear in C), then scaled back.  Returns a dict."""
    threads = torch.get_num_threads()
    g = torch.Generator().manual_seed(seed)
    # calibrate on a small slice, then pick the sample size
    c_probe = int(min(C, 4096))
    xs = torch.randn(B, D, generator=g)
    ws = torch.randn(c_probe, D, generator=g) * 0.05
    ys = torch.randint(0, c_probe, (B,), generator=g)
    head_step(xs, ws, ys, s, m)
    t0 = time.perf_counter()
    head_step(xs, ws, ys, s, m)
    per_class = (time.perf_counter() - t0) / c_probe
    c_sample = int(min(C, max(c_probe, (budget_s / 3.0) / max(per_class, 1e-9))))
    ws = torch.randn(c_sample, D, generator=g) * 0.05
    ys = torch.randint(0, c_sample, (B,), generator=g)
    head_step(xs, ws, ys, s, m)  # warm-up at the sample size
    best = float("inf")
    spent = 0.0
    steps = 0
    while steps < max_steps and (spent < budget_s or steps < 2):
        t0 = time.perf_counter()
        head_step(xs, ws, ys, s, m)
        dt = time.perf_counter() - t0
        best = min(best, dt)
        spent += dt
        steps += 1
    t_full = best * (C / c_sample)  # linear in C
    return {
        "value": B / t_full,
        "unit": "samples/s",
        "cores": threads,
        "kind": "port",
        "sample": "B=%d D=%d C=%d of %d classes, fp32 torch eager fwd+bwd, best of %d, scaled linearly in C"
        % (B, D, c_sample, C, steps),
        "ms_per_step_sample": best * 1e3,
        "c_sample": c_sample,
    }
