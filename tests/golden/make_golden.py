"""Mint known-answer vectors for the ArcFace head from the UNMODIFIED reference.

Run in the build container only (it imports /root/reference/arcface.py, which does not exist on the
GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/arcface_golden.npz.  Every case stores its inputs (numpy RandomState, so they are
reproducible bit for bit) and the outputs of
    ArcMarginProduct(...).forward(x, label) -> nn.CrossEntropyLoss() -> torch.argmax -> loss.backward()
executed by torch (CPU, fp32) on the reference module.  Cases follow SURVEY.md section 8c.
"""
import math
import os
import sys

import numpy as np
import torch
from torch import nn

sys.path.insert(0, "/root/reference")
from arcface import ArcMarginProduct  # noqa: E402  (the reference, imported verbatim)

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.arcface_numpy import synthetic_inputs  # noqa: E402  (seeded input recipe only)


def run_reference(x, w, label, s, m, easy, grad_loss=1.0, label_shape=None):
    head = ArcMarginProduct(in_feature=x.shape[1], out_feature=w.shape[0], s=s, m=m, easy_margin=easy)
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(w))
    xt = torch.from_numpy(x).clone().requires_grad_(True)
    lt = torch.from_numpy(label)
    if label_shape is not None:
        lt = lt.view(*label_shape)
    logits = head(xt, lt)
    loss = nn.CrossEntropyLoss()(logits, lt.view(-1))
    pred = torch.argmax(logits, dim=-1)
    (loss * grad_loss).backward()
    cosines = head.forward_test(xt.detach())
    return {
        "logits": logits.detach().numpy(),
        "loss": np.float32(loss.item()),
        "argmax": pred.numpy(),
        "dx": xt.grad.numpy(),
        "dw": head.weight.grad.numpy(),
        "cos": cosines.detach().numpy(),
    }


def xavier(rng, C, D):
    b = math.sqrt(6.0 / (C + D))
    return rng.uniform(-b, b, size=(C, D)).astype(np.float32)


def main():
    out = {}
    meta = []

    def add(name, x, w, label, s, m, easy, grad_loss=1.0, label_shape=None, keep_dw_rows=None):
        r = run_reference(x, w, label, s, m, easy, grad_loss, label_shape)
        out[name + "/x"] = x
        out[name + "/label"] = label
        out[name + "/hp"] = np.array([s, m, float(easy), grad_loss], dtype=np.float64)
        if keep_dw_rows is None:
            out[name + "/w"] = w
            for k, v in r.items():
                out[name + "/" + k] = v
        else:  # large case: inputs are regenerated from the seed; store sampled outputs only
            del out[name + "/x"]
            out[name + "/loss"] = r["loss"]
            out[name + "/argmax"] = r["argmax"]
            out[name + "/dx"] = r["dx"]
            out[name + "/dw_rows"] = np.asarray(keep_dw_rows, dtype=np.int64)
            out[name + "/dw"] = r["dw"][keep_dw_rows]
            out[name + "/logits_rows"] = r["logits"][:, keep_dw_rows]
        meta.append(name)

    rng = np.random.RandomState(0)
    # (1) base case, all defaults of the path: B=8, D=16, C=32, s=30, m=0.5
    B, D, C = 8, 16, 32
    x = rng.standard_normal((B, D)).astype(np.float32)
    w = xavier(rng, C, D)
    y = rng.randint(0, C, size=(B,)).astype(np.int64)
    add("base", x, w, y, 30.0, 0.5, False)
    # (2) easy_margin=True
    add("easy", x, w, y, 30.0, 0.5, True)
    # (3) row with cos_label < th  -> fallback s * (cos - mm): x = -w_y + noise
    x3 = x.copy()
    x3[0] = -w[y[0]] + 0.01 * rng.standard_normal(D).astype(np.float32)
    add("fallback", x3, w, y, 30.0, 0.5, False)
    # (4) easy-margin row with cos <= 0 -> s * cos
    add("easy_neg", x3, w, y, 30.0, 0.5, True)
    # (6) zero-norm embedding row
    x6 = x.copy()
    x6[1] = 0.0
    add("zero_row", x6, w, y, 30.0, 0.5, False)
    # (7) argmax tie: two identical class rows that win -> lowest index
    w7 = w.copy()
    x7 = x.copy()
    w7[5] = w7[20]
    x7[2] = 4.0 * w7[20] + 0.05 * rng.standard_normal(D).astype(np.float32)  # not collinear: cos^2 < 1
    y7 = y.copy()
    y7[2] = 3
    add("tie", x7, w7, y7, 30.0, 0.5, False)
    # (8) label given as (B, 1)
    add("label_col", x, w, y, 30.0, 0.5, False, label_shape=(B, 1))
    # (10) trained-like high-cosine batch, reference defaults s=64, m=0.4
    wh = w[y] / np.linalg.norm(w[y], axis=1, keepdims=True)
    n = rng.standard_normal((B, D)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    x10 = (3.0 * (wh + 0.5 * n)).astype(np.float32)
    add("trained", x10, w, y, 64.0, 0.4, False)
    # (12) upstream grad != 1 (multi-head weighting, nlp_classifier_train_daodian_v3_dist.py:164-166)
    add("grad10", x, w, y, 64.0, 0.2, False, grad_loss=10.0)
    # ragged shapes: B not a multiple of anything, C not a multiple of the tile
    Br, Dr, Cr = 5, 24, 37
    xr = rng.standard_normal((Br, Dr)).astype(np.float32)
    wr = xavier(rng, Cr, Dr)
    yr = rng.randint(0, Cr, size=(Br,)).astype(np.int64)
    add("ragged", xr, wr, yr, 64.0, 0.4, False)
    # (9) BASELINE config 1: B=64, D=512, C=1000, s=30, m=0.5
    # inputs come from the oracle's seeded recipe so the 2 MB weight matrix need not be stored
    x1, w1, y1 = synthetic_inputs(64, 512, 1000, seed=1)
    rows = np.unique(np.concatenate([y1[:16], np.arange(0, 1000, 97)]))
    add("c1", x1, w1, y1, 30.0, 0.5, False, keep_dw_rows=rows)

    # (5) update_m schedule (arcface.py:35-42)
    h = ArcMarginProduct(16, 32, s=30.0, m=0.5)
    h.update_m(0.04)
    upd = [h.m, h.cos_m, h.sin_m, h.th, h.mm]
    h.update_m(2.0)  # rejected
    upd += [h.m]
    h2 = ArcMarginProduct(16, 32, s=30.0, m=0.5)
    h2.update_m(-0.6)  # rejected (below 1e-6)
    upd += [h2.m]
    out["update_m"] = np.array(upd, dtype=np.float64)
    h3 = ArcMarginProduct(16, 32, s=64.0, m=0.4)
    out["const_m04"] = np.array([h3.cos_m, h3.sin_m, h3.th, h3.mm], dtype=np.float64)

    out["cases"] = np.array(meta)
    path = os.path.join(HERE, "arcface_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(meta), "cases")
    for name in meta:
        print("  %-10s loss=%.6f argmax[:4]=%s" % (name, float(out[name + "/loss"]), out[name + "/argmax"][:4]))


if __name__ == "__main__":
    torch.manual_seed(0)
    main()
