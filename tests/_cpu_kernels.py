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
    if inv_nw is None:
        inv_nw = 1.0 / w.double().norm(dim=1).clamp_min(1e-12)
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


def _fill(out, vals):
    if out is None:
        return vals
    for o, v in zip(out, vals):
        o.copy_(v)
    return tuple(out)


def forward_rows(xhat, what, label_local, s, class_offset=0, out=None):
    """Statistics over every local column except the row's label column (merged later by finalize_rows)."""
    z = (xhat @ what.t()) * s
    if label_local is not None:
        rows = torch.nonzero(label_local >= 0).flatten()
        z[rows, label_local[rows].long()] = float("-inf")
    rmax = z.max(dim=1).values
    rarg = (z == rmax[:, None]).int().argmax(dim=1)  # first maximum
    rsum = torch.where(torch.isinf(rmax), torch.zeros_like(rmax), (z - rmax[:, None]).exp().sum(1))
    return _fill(out, (rmax.float(), rsum.float(), (rarg + class_offset).long()))


def forward_rows_fused(xhat, weight, label_local, s, class_offset=0, out=None):
    what, inv_nw, _ = normalize_cast(weight)
    return (what, inv_nw) + forward_rows(xhat, what, label_local, s, class_offset, out)


def finalize_rows(rows_max, rows_sum, rows_arg, rows_z, label):
    rm = rows_max.double()
    Mx = rm.max(dim=0).values
    r = (rm == Mx[None]).int().argmax(dim=0)  # lowest rank (= lowest class range) on ties
    scale = torch.where(torch.isinf(rm), torch.zeros_like(rm), (rm - Mx[None]).exp())
    Sx = (rows_sum.double() * scale).sum(0)
    Ax = rows_arg.gather(0, r[None]).squeeze(0)
    Z = rows_z.double().sum(0)
    M = torch.maximum(Mx, Z)
    ex = torch.where(torch.isinf(Mx), torch.zeros_like(Mx), Sx * (Mx - M).exp())
    ey = (Z - M).exp()
    S = ex + ey
    lse = M + S.log()
    omp = ex / S
    arg = torch.where((Z > Mx) | ((Z == Mx) & (label < Ax)), label, Ax)
    return lse.float(), arg, Z.float(), omp.float(), (lse - Z).mean().float()


def accumulate(dst, src):
    dst += src
    return dst


def backward(xhat, xhat_t, what, inv_nw, lse, one_minus_p, dphi, label_local, s, grad_scale, grad_loss_dev=None,
             dw_out=None, dxhat_out=None):
    g = grad_scale * (float(grad_loss_dev) if grad_loss_dev is not None else 1.0)
    cos = xhat @ what.t()
    p = (cos * s - lse.double()[:, None]).exp()
    dc = p * (s * g)
    rows = torch.nonzero(label_local >= 0).flatten()
    cols = label_local[rows].long()
    dc[rows, cols] = -one_minus_p[rows].double() * (s * g) * dphi[rows].double()
    dxhat = dc @ what
    dwh = dc.t() @ xhat
    q = (dc * cos).sum(0)
    dw = (dwh - q[:, None] * what) * inv_nw[:, None]
    dxhat, dw = dxhat.float(), dw.float()
    if dxhat_out is not None:
        dxhat = dxhat_out.copy_(dxhat)
    if dw_out is not None:
        dw = dw_out.copy_(dw)
    return dxhat, dw


def normalize_bwd_x(x, inv_nx, dxhat):
    xh = x.double() * inv_nx[:, None]
    g = dxhat.double()
    return ((g - xh * (xh * g).sum(1, keepdim=True)) * inv_nx[:, None]).float()


def normalize_cast_gather(src, index):
    vh, inv, _ = normalize_cast(src[index])
    return vh, inv


def scatter_rows(src, index, dst):
    dst[index] = src
    return dst
