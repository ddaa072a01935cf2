"""Tensor-level wrappers over the C ABI (include/arcface_b200.h).

PyTorch is used for device memory, streams and autograd plumbing only: every function here takes CUDA
tensors, passes `data_ptr()` + sizes + the current stream to libarcface_b200.so, and returns tensors that
the library filled.  Nothing here computes on the host or through ATen; if the library is missing the
import of `_lib` raises.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass

import torch

from . import _lib

MAX_BATCH = _lib.MAX_BATCH


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: multimodalsimilar_b200 has no CPU path" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    return t


def margin_constants(m: float):
    """(cos_m, sin_m, th, mm) as the reference stores them (arcface.py:28-33)."""
    return math.cos(m), math.sin(m), math.cos(math.pi - m), math.sin(math.pi - m) * m


def device_ok() -> None:
    _lib.call("arcface_b200_device_ok")


def round_up(v: int, k: int) -> int:
    return (v + k - 1) // k * k


def normalize_cast(src: torch.Tensor, want_transpose: bool = False):
    """K1.  src [R, D] fp32 -> (bf16 [R, D], inv_norm fp32 [R], optional bf16 transpose [D, ld_t])."""
    _req(src, torch.float32, "src")
    R, D = src.shape
    dst = torch.empty((R, D), dtype=torch.bfloat16, device=src.device)
    inv = torch.empty((R,), dtype=torch.float32, device=src.device)
    dst_t = None
    ld_t = 0
    if want_transpose:
        ld_t = round_up(R, 64)
        dst_t = torch.empty((D, ld_t), dtype=torch.bfloat16, device=src.device)
    _lib.call("arcface_b200_normalize_cast", _ptr(src), R, D, _ptr(dst), _ptr(inv), _ptr(dst_t), ld_t, _stream())
    return dst, inv, dst_t


PREC = {"bf16": 0, "bf16x3": 1}   # ARCFACE_B200_PREC_* (include/arcface_b200.h)


def normalize_cast3(src: torch.Tensor, order: int, want_transpose: bool = False):
    """K1 of the bf16x3 mode.  src [R, D] fp32 -> (bf16 [R, 3 D] laid out [hi|hi|lo] (order 0, embeddings) or
    [hi|lo|hi] (order 1, class weights), inv_norm fp32 [R], optional transposed operand of the dW GEMM: bf16
    [D, 3 ld_t] = [hi^T | lo^T | hi^T] with ld_t = R rounded up to 64 and zero padding)."""
    _req(src, torch.float32, "src")
    R, D = src.shape
    dst = torch.empty((R, 3 * D), dtype=torch.bfloat16, device=src.device)
    inv = torch.empty((R,), dtype=torch.float32, device=src.device)
    dst_t = None
    ld_t = 0
    if want_transpose:
        ld_t = round_up(R, 64)
        dst_t = torch.zeros((D, 3 * ld_t), dtype=torch.bfloat16, device=src.device)   # the padding columns are contracted
    _lib.call("arcface_b200_normalize_cast3", _ptr(src), R, D, int(order), _ptr(dst), _ptr(inv), _ptr(dst_t), ld_t,
              _stream())
    return dst, inv, dst_t


def normalize_cast_gather(src: torch.Tensor, index: torch.Tensor):
    """K1 over sampled rows.  src [R, D] fp32, index int64 [S] -> (bf16 [S, D] rows normalise(src[index]), inv_norm [S])."""
    _req(src, torch.float32, "src")
    _req(index, torch.int64, "index")
    R, D = src.shape
    S = index.numel()
    dst = torch.empty((S, D), dtype=torch.bfloat16, device=src.device)
    inv = torch.empty((S,), dtype=torch.float32, device=src.device)
    _lib.call("arcface_b200_normalize_cast_gather", _ptr(src), R, _ptr(index), S, D, _ptr(dst), _ptr(inv), _stream())
    return dst, inv


def scatter_rows(src: torch.Tensor, index: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst[index[r]] = src[r] (fp32 rows): the sampled classes' gradient into the full-size dW."""
    _req(src, torch.float32, "src")
    _req(index, torch.int64, "index")
    _req(dst, torch.float32, "dst")
    if src.shape[1] != dst.shape[1] or index.numel() != src.shape[0]:
        raise ValueError("scatter_rows: %s rows by %s indices into %s" % (tuple(src.shape), tuple(index.shape), tuple(dst.shape)))
    _lib.call("arcface_b200_scatter_rows", _ptr(src), _ptr(index), src.shape[0], src.shape[1], _ptr(dst), dst.shape[0],
              _stream())
    return dst


@dataclass
class LabelMargin:
    t_label: torch.Tensor      # fp32 [B] exact label cosine (0 where the label is on another rank)
    z_label: torch.Tensor      # fp32 [B] s * margin(t)
    dphi: torch.Tensor         # fp32 [B] d margin / d t
    label_local: torch.Tensor  # int32 [B] label - class_offset, or -1
    bad_flag: torch.Tensor     # int32 [1] set when a label is outside [0, C_total)


def label_margin(x, w, inv_nx, inv_nw, label, class_offset, c_total, s, m, easy_margin, z_out=None,
                 bad_flag_out=None) -> LabelMargin:
    """bad_flag_out: optional int32 [1] tensor the kernel raises on an out-of-range label instead of a fresh device
    flag -- may be PINNED HOST memory (device-visible under unified addressing): nothing resets it but its owner."""
    _req(x, torch.float32, "x")
    _req(w, torch.float32, "weight")
    _req(label, torch.int64, "label")
    B, D = x.shape
    dev = x.device
    out = LabelMargin(
        torch.empty(B, dtype=torch.float32, device=dev),
        z_out if z_out is not None else torch.empty(B, dtype=torch.float32, device=dev),
        torch.empty(B, dtype=torch.float32, device=dev),
        torch.empty(B, dtype=torch.int32, device=dev),
        bad_flag_out if bad_flag_out is not None else torch.zeros(1, dtype=torch.int32, device=dev),
    )
    cos_m, sin_m, th, mm = margin_constants(m)
    _lib.call("arcface_b200_label_margin", _ptr(x), _ptr(w), _ptr(inv_nx), _ptr(inv_nw), _ptr(label), B, D,
              w.shape[0], class_offset, c_total, s, cos_m, sin_m, th, mm, int(bool(easy_margin)),
              _ptr(out.t_label), _ptr(out.z_label), _ptr(out.dphi), _ptr(out.label_local), _ptr(out.bad_flag), _stream())
    return out


def forward_parts(B: int, D: int, c_local: int) -> int:
    n = ctypes.c_int32(0)
    _lib.call("arcface_b200_forward_parts", B, D, c_local, ctypes.byref(n))
    return n.value


def forward_rows(xhat, what, label_local, s: float, class_offset: int = 0, out=None):
    """K2 + per-shard combine over every column except the row's label column (label_local, or None for
    the eval path).  Returns (row_max fp32 [B], row_sum fp32 [B], row_arg int64 [B])."""
    _req(xhat, torch.bfloat16, "xhat")
    _req(what, torch.bfloat16, "what")
    B, D = xhat.shape
    C = what.shape[0]
    dev = xhat.device
    n_parts = forward_parts(B, D, C)
    pmax = torch.empty((n_parts, B), dtype=torch.float32, device=dev)
    psum = torch.empty((n_parts, B), dtype=torch.float32, device=dev)
    parg = torch.empty((n_parts, B), dtype=torch.int32, device=dev)
    _lib.call("arcface_b200_forward_stats", _ptr(xhat), _ptr(what), _ptr(label_local), B, D, C, s,
              _ptr(pmax), _ptr(psum), _ptr(parg), n_parts, _stream())
    if out is not None:
        rmax, rsum, rarg = out
    else:
        rmax = torch.empty(B, dtype=torch.float32, device=dev)
        rsum = torch.empty(B, dtype=torch.float32, device=dev)
        rarg = torch.empty(B, dtype=torch.int64, device=dev)
    _lib.call("arcface_b200_combine_partials", _ptr(pmax), _ptr(psum), _ptr(parg), n_parts, B, class_offset,
              _ptr(rmax), _ptr(rsum), _ptr(rarg), _stream())
    return rmax, rsum, rarg


def forward_rows_fused(xhat, weight, label_local, s: float, class_offset: int = 0, out=None):
    """K1 (class weights) + K2 + per-shard combine in one GEMM launch: the forward kernel's helper warps
    normalise and cast `weight` (fp32 [C, D]) while its tcgen05 pipeline consumes the rows already
    published.  Returns (what bf16 [C, D], inv_nw fp32 [C], row_max, row_sum, row_arg) -- the first two are
    bit-identical to `normalize_cast(weight)`."""
    _req(xhat, torch.bfloat16, "xhat")
    _req(weight, torch.float32, "weight")
    B, D = xhat.shape
    C = weight.shape[0]
    dev = xhat.device
    what = torch.empty((C, D), dtype=torch.bfloat16, device=dev)
    inv_nw = torch.empty((C,), dtype=torch.float32, device=dev)
    n_parts = forward_parts(B, D, C)
    pmax = torch.empty((n_parts, B), dtype=torch.float32, device=dev)
    psum = torch.empty((n_parts, B), dtype=torch.float32, device=dev)
    parg = torch.empty((n_parts, B), dtype=torch.int32, device=dev)
    nws = ctypes.c_size_t(0)
    _lib.call("arcface_b200_forward_fused_workspace_bytes", B, D, C, ctypes.byref(nws))
    ws = torch.empty(max(16, nws.value), dtype=torch.uint8, device=dev)
    _lib.call("arcface_b200_forward_stats_fused", _ptr(xhat), _ptr(weight), _ptr(label_local), B, D, C, s,
              _ptr(what), _ptr(inv_nw), _ptr(pmax), _ptr(psum), _ptr(parg), n_parts, _ptr(ws), ws.numel(), _stream())
    if out is not None:  # (row_max fp32 [B], row_sum fp32 [B], row_arg int64 [B]) to fill, e.g. views of a packed buffer
        rmax, rsum, rarg = out
    else:
        rmax = torch.empty(B, dtype=torch.float32, device=dev)
        rsum = torch.empty(B, dtype=torch.float32, device=dev)
        rarg = torch.empty(B, dtype=torch.int64, device=dev)
    _lib.call("arcface_b200_combine_partials", _ptr(pmax), _ptr(psum), _ptr(parg), n_parts, B, class_offset,
              _ptr(rmax), _ptr(rsum), _ptr(rarg), _stream())
    return what, inv_nw, rmax, rsum, rarg


def finalize_rows(rows_max, rows_sum, rows_arg, rows_z, label):
    """Merge [R, B] per-rank rows of the non-label columns with the label logits ->
    (lse [B], argmax int64 [B], z_label [B], one_minus_p [B], loss [])."""
    R, B = rows_max.shape
    dev = rows_max.device
    lse = torch.empty(B, dtype=torch.float32, device=dev)
    arg, loss = packed_outputs(B, dev)
    z = torch.empty(B, dtype=torch.float32, device=dev)
    omp = torch.empty(B, dtype=torch.float32, device=dev)
    _lib.call("arcface_b200_finalize_rows", _ptr(_req(rows_max, torch.float32, "rows_max")),
              _ptr(_req(rows_sum, torch.float32, "rows_sum")), _ptr(_req(rows_arg, torch.int64, "rows_arg")),
              _ptr(_req(rows_z, torch.float32, "rows_z")), _ptr(_req(label, torch.int64, "label")), R, B, _ptr(lse),
              _ptr(arg), _ptr(z), _ptr(omp), _ptr(loss), _stream())
    return lse, arg, z, omp, loss


def packed_outputs(B: int, device):
    """(argmax int64 [B], loss fp32 []) as views of ONE buffer, [argmax | loss], so that a caller that must copy the
    step's results out of static storage (graph replay) needs a single copy: `clone_outputs(argmax, loss)`."""
    buf = torch.empty(8 * B + 16, dtype=torch.uint8, device=device)
    arg = buf[: 8 * B].view(torch.int64)
    arg._pack = buf   # Python-side note for clone_outputs; views made from `arg` do not carry it
    return arg, buf[8 * B: 8 * B + 4].view(torch.float32).reshape(())


def clone_outputs(arg: torch.Tensor, loss: torch.Tensor):
    """(argmax, loss) copied out of the buffer behind `packed_outputs` with one device copy."""
    buf = getattr(arg, "_pack", None)
    if buf is None:
        return arg.clone(), loss.clone()
    B = arg.numel()
    raw = buf.clone()
    return raw[: 8 * B].view(torch.int64), raw[8 * B: 8 * B + 4].view(torch.float32).reshape(())


STATS_BYTES_PER_ROW = 20  # packed per-rank statistics: [arg int64 x B | max fp32 x B | sum fp32 x B | z_label fp32 x B]


def packed_stats(B: int, device):
    """One rank's statistics buffer (uint8 [20 B]) and the typed views the kernels fill: (buf, max, sum, z, arg)."""
    buf = torch.empty(STATS_BYTES_PER_ROW * B, dtype=torch.uint8, device=device)
    arg = buf[: 8 * B].view(torch.int64)
    f = buf[8 * B:].view(torch.float32)
    return buf, f[:B], f[B:2 * B], f[2 * B:], arg


def finalize_rows_packed(allp, label):
    """`finalize_rows` over the all-gathered packed statistics (uint8 [R, 20 B], B even), read in place."""
    R, nbytes = allp.shape
    B = nbytes // STATS_BYTES_PER_ROW
    if B % 2 != 0:
        raise ValueError("packed statistics need an even batch size")
    # `allp` may be a [R, n] view of a wider receive buffer (p2p.PeerExchange sizes its slots for the largest batch
    # seen, a later smaller batch leaves rows `slot` bytes apart): the kernel takes the rank stride, only the bytes
    # of one rank have to be contiguous
    if allp.dtype != torch.uint8 or not allp.is_cuda or allp.stride(1) != 1:
        raise RuntimeError("packed statistics must be a CUDA uint8 [R, 20 B] tensor with contiguous rows")
    stride = allp.stride(0) if R > 1 else nbytes
    if stride % 8 != 0 or stride < nbytes:
        raise RuntimeError("packed statistics: rank stride %d must be a multiple of 8 and >= %d" % (stride, nbytes))
    dev = allp.device
    base = allp.data_ptr()
    lse = torch.empty(B, dtype=torch.float32, device=dev)
    arg, loss = packed_outputs(B, dev)
    z = torch.empty(B, dtype=torch.float32, device=dev)
    omp = torch.empty(B, dtype=torch.float32, device=dev)
    _lib.call("arcface_b200_finalize_rows_strided", ctypes.c_void_p(base + 8 * B), ctypes.c_void_p(base + 12 * B),
              ctypes.c_void_p(base), ctypes.c_void_p(base + 16 * B), _ptr(_req(label, torch.int64, "label")), R, B,
              stride // 4, stride // 8, _ptr(lse), _ptr(arg), _ptr(z), _ptr(omp), _ptr(loss), _stream())
    return lse, arg, z, omp, loss


def logits(xhat, what, z_label, label_local, scale: float) -> torch.Tensor:
    """Materialise scale * cos (label column overridden when z_label is given): fp32 [B, C]."""
    _req(xhat, torch.bfloat16, "xhat")
    _req(what, torch.bfloat16, "what")
    B, D = xhat.shape
    C = what.shape[0]
    out = torch.empty((B, C), dtype=torch.float32, device=xhat.device)
    for lo in range(0, B, MAX_BATCH):   # one launch takes MAX_BATCH rows
        hi = min(B, lo + MAX_BATCH)
        _lib.call("arcface_b200_logits", _ptr(xhat[lo:hi]), _ptr(what), _ptr(None if z_label is None else z_label[lo:hi]),
                  _ptr(None if label_local is None else label_local[lo:hi]), hi - lo, D, C, scale, _ptr(out[lo:hi]), C,
                  _stream())
    return out


def cosine_topk(xhat, what, k: int, scale: float = 1.0, class_offset: int = 0):
    """K4.  (values fp32 [B, k], indices int64 [B, k]) of the k largest scale * cos per row, descending."""
    _req(xhat, torch.bfloat16, "xhat")
    _req(what, torch.bfloat16, "what")
    B, D = xhat.shape
    if B > MAX_BATCH:   # one launch takes MAX_BATCH query rows
        parts = [cosine_topk(xhat[lo:lo + MAX_BATCH], what, k, scale, class_offset) for lo in range(0, B, MAX_BATCH)]
        return torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
    C = what.shape[0]
    dev = xhat.device
    n = ctypes.c_size_t(0)
    _lib.call("arcface_b200_topk_workspace_bytes", B, D, C, k, ctypes.byref(n))
    ws = torch.empty(max(16, n.value), dtype=torch.uint8, device=dev)
    val = torch.empty((B, k), dtype=torch.float32, device=dev)
    idx = torch.empty((B, k), dtype=torch.int64, device=dev)
    _lib.call("arcface_b200_cosine_topk", _ptr(xhat), _ptr(what), B, D, C, k, scale, class_offset, _ptr(val), _ptr(idx),
              _ptr(ws), ws.numel(), _stream())
    return val, idx


def topk_merge(val, idx, k: int):
    """Merge [B, n] (value, global index) candidates into the k best per row."""
    _req(val, torch.float32, "val")
    _req(idx, torch.int64, "idx")
    B, n = val.shape
    out_v = torch.empty((B, k), dtype=torch.float32, device=val.device)
    out_i = torch.empty((B, k), dtype=torch.int64, device=val.device)
    _lib.call("arcface_b200_topk_merge", _ptr(val), _ptr(idx), B, n, k, _ptr(out_v), _ptr(out_i), _stream())
    return out_v, out_i


def backward_workspace_bytes(B: int, D: int, c_local: int) -> int:
    n = ctypes.c_size_t(0)
    _lib.call("arcface_b200_backward_workspace_bytes", B, D, c_local, ctypes.byref(n))
    return n.value


def backward_plan(B: int, D: int, c_local: int):
    """(classes per scratch chunk, number of chunks) of the backward walk."""
    cc = ctypes.c_int64(0)
    n = ctypes.c_int32(0)
    _lib.call("arcface_b200_backward_plan", B, D, c_local, ctypes.byref(cc), ctypes.byref(n))
    return cc.value, n.value


def backward_launches(B: int, D: int, c_local: int) -> int:
    """Kernels one `backward` call launches (1 = single-launch backward, else 3 per scratch chunk)."""
    n = ctypes.c_int32(0)
    _lib.call("arcface_b200_backward_launches", B, D, c_local, ctypes.byref(n))
    return n.value


def backward(xhat, xhat_t, what, inv_nw, lse, one_minus_p, dphi, label_local, s: float, grad_scale: float,
             grad_loss_dev=None, dw_out=None, prec: int = 0, dxhat_out=None):
    """K3.  Returns (dxhat fp32 [B, D] partial over this shard's classes, dW fp32 [C_local, D]).
    prec = PREC['bf16x3']: xhat / what are the 3 D wide rows of `normalize_cast3`."""
    B, D = xhat.shape
    if prec:
        D //= 3
    C = what.shape[0]
    dev = xhat.device
    dxhat = dxhat_out if dxhat_out is not None else torch.empty((B, D), dtype=torch.float32, device=dev)
    _req(dxhat, torch.float32, "dxhat")
    dw = dw_out if dw_out is not None else torch.empty((C, D), dtype=torch.float32, device=dev)
    _req(dw, torch.float32, "dw")
    nbytes = backward_workspace_bytes(B, D, C)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if grad_loss_dev is not None:
        grad_loss_dev = _req(grad_loss_dev.reshape(1), torch.float32, "grad_loss")
    # xhat_t may be a column slice [D, b0:b1] of the full transpose (batch chunks): its leading dimension is the stride
    ld_t = xhat_t.stride(0) // (3 if prec else 1)
    _lib.call("arcface_b200_backward_prec", _ptr(xhat), _ptr(xhat_t), ld_t, _ptr(what), _ptr(inv_nw),
              _ptr(lse), _ptr(one_minus_p), _ptr(dphi), _ptr(label_local), B, D, C, s, grad_scale, _ptr(grad_loss_dev),
              _ptr(dxhat), _ptr(dw), _ptr(ws), nbytes, int(prec), _stream())
    return dxhat, dw


def accumulate(dst: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    """dst += src (fp32, same shape): dW of a further batch chunk onto the first one's."""
    _req(dst, torch.float32, "dst")
    _req(src, torch.float32, "src")
    if dst.shape != src.shape:
        raise ValueError("accumulate: %s += %s" % (tuple(dst.shape), tuple(src.shape)))
    _lib.call("arcface_b200_accumulate", _ptr(dst), _ptr(src), dst.numel(), _stream())
    return dst


def two_stream_concat(a, b):
    """cat(F.normalize(a), F.normalize(b), dim=1) in one kernel -> (emb fp32 [B, D1 + D2], inv1 [B], inv2 [B])."""
    _req(a, torch.float32, "a")
    _req(b, torch.float32, "b")
    B, D1 = a.shape
    D2 = b.shape[1]
    if b.shape[0] != B:
        raise ValueError("both streams need the same batch size")
    out = torch.empty((B, D1 + D2), dtype=torch.float32, device=a.device)
    inv1 = torch.empty(B, dtype=torch.float32, device=a.device)
    inv2 = torch.empty(B, dtype=torch.float32, device=a.device)
    _lib.call("arcface_b200_two_stream_concat", _ptr(a), _ptr(b), B, D1, D2, _ptr(out), _ptr(inv1), _ptr(inv2), _stream())
    return out, inv1, inv2


def two_stream_concat_bwd(emb, inv1, inv2, grad, D1: int):
    """Backward of `two_stream_concat`: (d a [B, D1], d b [B, D2])."""
    _req(emb, torch.float32, "emb")
    _req(grad, torch.float32, "grad")
    B, D = emb.shape
    D2 = D - D1
    da = torch.empty((B, D1), dtype=torch.float32, device=emb.device)
    db = torch.empty((B, D2), dtype=torch.float32, device=emb.device)
    _lib.call("arcface_b200_two_stream_concat_bwd", _ptr(emb), _ptr(inv1), _ptr(inv2), _ptr(grad), B, D1, D2, _ptr(da),
              _ptr(db), _stream())
    return da, db


def normalize_bwd_x(x, inv_nx, dxhat) -> torch.Tensor:
    _req(x, torch.float32, "x")
    _req(dxhat, torch.float32, "dxhat")
    B, D = x.shape
    dx = torch.empty_like(x)
    _lib.call("arcface_b200_normalize_bwd_x", _ptr(x), _ptr(inv_nx), _ptr(dxhat), B, D, _ptr(dx), _stream())
    return dx


def normalize_bwd_x_sum(x, inv_nx, parts) -> torch.Tensor:
    """normalize_bwd_x over the sum of `parts` fp32 [R, B, D] (fixed order)."""
    _req(x, torch.float32, "x")
    if parts.dtype != torch.float32 or parts.stride(-1) != 1 or parts.stride(1) != parts.shape[2]:
        raise RuntimeError("parts must be fp32 [R, B, D] with contiguous rows")
    B, D = x.shape
    dx = torch.empty_like(x)
    _lib.call("arcface_b200_normalize_bwd_x_sum", _ptr(x), _ptr(inv_nx), _ptr(parts), parts.shape[0], parts.stride(0), B, D,
              _ptr(dx), _stream())
    return dx


def scale_grads(a, b, scale_dev) -> None:
    """In-place a *= scale, b *= scale (fp32 tensors or None; scale = DEVICE scalar); free when scale == 1."""
    na = 0 if a is None else _req(a, torch.float32, "a").numel()
    nb = 0 if b is None else _req(b, torch.float32, "b").numel()
    _lib.call("arcface_b200_scale_grads", _ptr(a), na, _ptr(b), nb, _ptr(_req(scale_dev, torch.float32, "scale")),
              _stream())


def scale_copy(src, b, scale_dev) -> torch.Tensor:
    """Returns a fresh tensor src * scale (DEVICE scalar) and scales `b` (fp32 tensor or None) in place unless the
    factor is exactly 1: one launch."""
    _req(src, torch.float32, "src")
    dst = torch.empty_like(src)
    nb = 0 if b is None else _req(b, torch.float32, "b").numel()
    _lib.call("arcface_b200_scale_copy", _ptr(src), _ptr(dst), src.numel(), _ptr(b), nb,
              _ptr(_req(scale_dev, torch.float32, "scale")), _stream())
    return dst


def pack_xy(x, y, dst) -> None:
    """dst uint8 [b * D * 4 + b * 8] <- (x fp32 [b, D] | y int64 [b]) in one launch."""
    _req(x, torch.float32, "x")
    _req(y, torch.int64, "y")
    b, D = x.shape
    _lib.call("arcface_b200_pack_xy", _ptr(x), _ptr(y), b, D, _ptr(dst), _stream())


def adamw_normalize(w, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, what=None, inv_nw=None):
    """One AdamW step in place on (w, exp_avg, exp_avg_sq) [rows, D] fp32; optionally emits the next forward's
    normalised bf16 rows `what` [rows, D] and `inv_nw` [rows] in the same pass."""
    for t, name in ((w, "w"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _req(t, torch.float32, name)
    rows, D = w.shape
    if what is not None:
        _req(what, torch.bfloat16, "what")
        _req(inv_nw, torch.float32, "inv_nw")
    _lib.call("arcface_b200_adamw_normalize", _ptr(w), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), rows, D, lr, beta1,
              beta2, eps, weight_decay, step, _ptr(what), _ptr(inv_nw), _stream())


def step_workspace_bytes(B: int, D: int, C: int) -> int:
    n = ctypes.c_size_t(0)
    _lib.call("arcface_b200_step_workspace_bytes", B, D, C, ctypes.byref(n))
    return n.value


def step_host(x_host, label_host, w_dev, s, m, easy_margin, grad_loss, loss_host, argmax_host, dx_host, dw_dev, ws):
    """The one-call host-buffer step (pinned HOST x / labels in, HOST loss / argmax / dx out)."""
    for t, name in ((x_host, "x_host"), (label_host, "label_host"), (loss_host, "loss_host"),
                    (argmax_host, "argmax_host"), (dx_host, "dx_host")):
        if t.is_cuda or not t.is_contiguous():
            raise RuntimeError("%s must be a contiguous host tensor" % name)
    B, D = x_host.shape
    C = w_dev.shape[0]
    _lib.call("arcface_b200_step_host", _ptr(x_host), _ptr(label_host), _ptr(_req(w_dev, torch.float32, "w")), B, D, C,
              s, m, int(bool(easy_margin)), grad_loss, _ptr(loss_host), _ptr(argmax_host), _ptr(dx_host),
              _ptr(_req(dw_dev, torch.float32, "dw")), _ptr(ws), ws.numel(), _stream())
