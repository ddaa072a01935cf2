"""TEST-ONLY stand-in for `multimodalsimilar_b200.ops`: the kernel-level contract of
include/arcface_b200.h restated with float64 torch on the CPU (using the same formulas as the oracle),
so the multi-rank choreography in multimodalsimilar_b200/sharded.py can run under gloo without a GPU.
The product never imports this module."""
import math
from types import SimpleNamespace

import torch


def normalize_cast(src, want_transpose=False):
    v = src.double()
    inv = 1.0 / v.norm(dim=1).clamp_min(1e-12)
    vh = v * inv[:, None]
    return vh, inv, (vh.t().contiguous() if want_transpose else None)


def label_margin(x, w, inv_nx, inv_nw, label, class_offset, c_total, s, m, easy_margin):
    B = x.shape[0]
    C = w.shape[0]
    loc = label - class_offset
    own = (loc >= 0) & (loc < C)
    lc = loc.clamp(0, C - 1)
    t = (x.double() * w.double()[lc]).sum(1) * inv_nx * inv_nw[lc]
    sine = (1.0 - t * t).clamp_min(0).sqrt()
    phi = t * math.cos(m) - sine * math.sin(m)
    take = (t > 0) if easy_margin else ((t - math.cos(math.pi - m)) > 0)
    u = torch.where(take, phi, t if easy_margin else t - math.sin(math.pi - m) * m)
    dphi = torch.where(take, math.cos(m) + t * math.sin(m) / sine.clamp_min(1e-6), torch.ones_like(t))
    z = torch.zeros(B, dtype=torch.float64)
    return SimpleNamespace(
        t_label=torch.where(own, t, z).float(),
        z_label=torch.where(own, u * s, z).float(),
        dphi=torch.where(own, dphi, z).float(),
        label_local=torch.where(own, loc, torch.full_like(loc, -1)).int(),
        bad_flag=((label < 0) | (label >= c_total)).any().int().reshape(1),
    )


def _local_logits(xhat, what, z_label, label_local, s):
    z = (xhat @ what.t()) * s
    if z_label is not None:
        rows = torch.nonzero(label_local >= 0).flatten()
        z[rows, label_local[rows].long()] = z_label[rows].double()
    return z


def forward_rows(xhat, what, z_label, label_local, s, class_offset=0):
    z = _local_logits(xhat, what, z_label, label_local, s)
    rmax, rarg = z.max(dim=1)
    # torch.max returns an arbitrary index on ties for some backends; take the first maximum explicitly
    rarg = (z == rmax[:, None]).int().argmax(dim=1)
    rsum = (z - rmax[:, None]).exp().sum(1)
    return rmax.float(), rsum.float(), (rarg + class_offset).long()


def finalize_rows(rows_max, rows_sum, rows_arg, rows_z):
    M, r = rows_max.double().max(dim=0)
    r = (rows_max.double() == M[None]).int().argmax(dim=0)  # lowest rank (= lowest class range) on ties
    S = (rows_sum.double() * (rows_max.double() - M[None]).exp()).sum(0)
    lse = M + S.log()
    z = rows_z.double().sum(0)
    arg = rows_arg.gather(0, r[None]).squeeze(0)
    return lse.float(), arg, z.float(), (lse - z).mean().float()


def backward(xhat, xhat_t, what, inv_nw, lse, z_label, dphi, label_local, s, grad_scale, grad_loss_dev=None,
             dw_out=None):
    g = grad_scale * (float(grad_loss_dev) if grad_loss_dev is not None else 1.0)
    cos = xhat @ what.t()
    z = cos * s
    rows = torch.nonzero(label_local >= 0).flatten()
    cols = label_local[rows].long()
    z[rows, cols] = z_label[rows].double()
    p = (z - lse.double()[:, None]).exp()
    dc = p * (s * g)
    dc[rows, cols] = (p[rows, cols] - 1.0) * (s * g) * dphi[rows].double()
    dxhat = dc @ what
    dwh = dc.t() @ xhat
    q = (dc * cos).sum(0)
    dw = (dwh - q[:, None] * what) * inv_nw[:, None]
    return dxhat.float(), dw.float()


def normalize_bwd_x(x, inv_nx, dxhat):
    xh = x.double() * inv_nx[:, None]
    g = dxhat.double()
    return ((g - xh * (xh * g).sum(1, keepdim=True)) * inv_nx[:, None]).float()
