"""CPU oracle for the ArcFace head hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference algorithm, function by function:

    /root/reference/arcface.py:28-33   margin constants          -> margin_constants
    /root/reference/arcface.py:35-42   update_m                  -> update_m
    /root/reference/arcface.py:45-63   ArcMarginProduct.forward  -> forward_logits
    /root/reference/arcface.py:65-67   forward_test              -> forward_test
    nn.CrossEntropyLoss() / torch.argmax / loss.backward() at the call sites
    (e.g. /root/reference/nlp_classifier_train.py:100,120-123)   -> cross_entropy, argmax, backward

The arithmetic of the reference lives in PyTorch (unpinned in the reference; this project pins the
container's torch 2.11.0, CPU, fp32).  The reference ships no golden vectors for this path
(SURVEY.md section 8c), so the oracle is pinned against outputs of the reference itself, generated in
the build container by tests/golden/make_golden.py (which imports /root/reference/arcface.py unmodified)
and committed under tests/golden/.  tests/test_oracle.py checks every function here against them.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker or the timed CPU baseline.  The product (multimodalsimilar_b200)
never imports it and has no CPU path.
"""
from __future__ import annotations

import math

import numpy as np

EPS = 1e-12  # torch.nn.functional.normalize default eps (arcface.py:47)


def margin_constants(m: float):
    """cos_m, sin_m, th, mm exactly as arcface.py:28-33 (Python floats)."""
    cos_m = math.cos(m)
    sin_m = math.sin(m)
    th = math.cos(math.pi - m)
    mm = math.sin(math.pi - m) * m
    return cos_m, sin_m, th, mm


def update_m(m: float, delta: float) -> float:
    """arcface.py:35-42: the update is accepted only while 1e-6 <= m + delta <= 1.0."""
    updated = m + delta
    if updated >= 1e-6 and updated <= 1.0:
        return updated
    return m


def normalize(v: np.ndarray) -> np.ndarray:
    """F.normalize(v) along dim 1: v / max(||v||_2, eps)  (arcface.py:47)."""
    n = np.sqrt(np.sum(v * v, axis=1, keepdims=True, dtype=v.dtype))
    return v / np.maximum(n, v.dtype.type(EPS))


def cosines(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """F.linear(F.normalize(x), F.normalize(weight))  (arcface.py:47, :66)."""
    return normalize(x) @ normalize(w).T


def forward_test(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """arcface.py:65-67: bare cosines, no margin, no scale."""
    return cosines(x, w)


def _phi(cosine: np.ndarray, m: float, easy_margin: bool):
    """arcface.py:49-55 on an array of cosines.  Returns (phi_after_where, took_phi_branch, sine)."""
    dt = cosine.dtype.type
    cos_m, sin_m, th, mm = margin_constants(m)
    with np.errstate(invalid="ignore"):
        sine = np.sqrt(dt(1.0) - cosine * cosine)  # unclamped, as in the reference (:49)
    phi = cosine * dt(cos_m) - sine * dt(sin_m)
    if easy_margin:
        take = cosine > 0
        out = np.where(take, phi, cosine)
    else:
        take = (cosine - dt(th)) > 0
        out = np.where(take, phi, cosine - dt(mm))
    return out, take, sine


def forward_logits(x, w, label, s=64.0, m=0.40, easy_margin=False, dtype=np.float32) -> np.ndarray:
    """ArcMarginProduct.forward (arcface.py:45-63): s * (one_hot * phi + (1 - one_hot) * cosine)."""
    x = np.asarray(x, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    label = np.asarray(label).reshape(-1).astype(np.int64)
    cosine = cosines(x, w)
    phi, _, _ = _phi(cosine, m, easy_margin)
    one_hot = np.zeros_like(cosine)
    one_hot[np.arange(cosine.shape[0]), label] = 1
    out = one_hot * phi + (dtype(1.0) - one_hot) * cosine  # 0 * NaN hazard kept on purpose (:60)
    return out * dtype(s)


def log_softmax(z: np.ndarray) -> np.ndarray:
    zmax = np.max(z, axis=1, keepdims=True)
    e = np.exp(z - zmax)
    return z - zmax - np.log(np.sum(e, axis=1, keepdims=True, dtype=z.dtype))


def cross_entropy(logits: np.ndarray, label) -> float:
    """nn.CrossEntropyLoss() defaults: mean over the batch of -log_softmax(logits)[b, y_b]."""
    label = np.asarray(label).reshape(-1).astype(np.int64)
    ls = log_softmax(logits)
    return float(-np.mean(ls[np.arange(logits.shape[0]), label], dtype=logits.dtype))


def argmax(logits: np.ndarray) -> np.ndarray:
    """torch.argmax(preds, dim=-1): index of the first maximal element of each row."""
    return np.argmax(logits, axis=1).astype(np.int64)


def loss_and_argmax(x, w, label, s=64.0, m=0.40, easy_margin=False, dtype=np.float32):
    z = forward_logits(x, w, label, s, m, easy_margin, dtype)
    return cross_entropy(z, label), argmax(z)


def row_stats(logits: np.ndarray):
    """Per-row max and log-sum-exp: the statistics the fused kernel saves instead of the logits."""
    zmax = np.max(logits, axis=1)
    lse = zmax + np.log(np.sum(np.exp(logits - zmax[:, None]), axis=1, dtype=logits.dtype))
    return zmax, lse


def backward(x, w, label, s=64.0, m=0.40, easy_margin=False, grad_loss=1.0, dtype=np.float64):
    """Analytic loss.backward() through CrossEntropyLoss and arcface.py:45-63.

    g      = grad_loss / B * (softmax(z) - one_hot)
    dcos   = s * g                                        off the label column, and on it when the
                                                          torch.where at :53/:55 took the non-phi branch
    dcos_y = s * g_y * (cos_m + cos * sin_m / sine)       on the phi branch (d/dcos of :50 through :49)
    dxhat = dcos @ what, dwhat = dcos^T @ xhat, then the backward of F.normalize for both operands:
    dv = (dvhat - vhat * <vhat, dvhat>) / max(||v||, eps).
    Returns (dx, dw).
    """
    x = np.asarray(x, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    label = np.asarray(label).reshape(-1).astype(np.int64)
    B = x.shape[0]
    rows = np.arange(B)
    cos_m, sin_m, _, _ = margin_constants(m)
    nx = np.maximum(np.sqrt(np.sum(x * x, axis=1, keepdims=True)), EPS)
    nw = np.maximum(np.sqrt(np.sum(w * w, axis=1, keepdims=True)), EPS)
    xh = x / nx
    wh = w / nw
    cosine = xh @ wh.T
    z = forward_logits(x, w, label, s, m, easy_margin, dtype)
    p = np.exp(log_softmax(z))
    g = p
    g[rows, label] = 0.0
    g[rows, label] = -np.sum(g, axis=1)  # p_y - 1 = -(sum of the other probabilities): no cancellation
    g *= grad_loss / B
    dcos = g * s
    t = cosine[rows, label]
    _, take, sine = _phi(t, m, easy_margin)
    with np.errstate(divide="ignore", invalid="ignore"):
        dphi = np.where(take, cos_m + t * sin_m / sine, 1.0)
    dcos[rows, label] *= dphi
    dxh = dcos @ wh
    dwh = dcos.T @ xh
    dx = (dxh - xh * np.sum(xh * dxh, axis=1, keepdims=True)) / nx
    dw = (dwh - wh * np.sum(wh * dwh, axis=1, keepdims=True)) / nw
    return dx, dw


def synthetic_inputs(B, D, C, seed=0, trained_like=False, dtype=np.float32):
    """Deterministic inputs shared by the parity tests and bench.py (numpy RandomState: identical bits
    on every box).  W is xavier-uniform (bound sqrt(6 / (C + D)), as nn.init.xavier_uniform_ at
    arcface.py:25); `trained_like` puts every row near its class centre (cos_label ~ 0.9) so the
    phi branch and the argmax are well separated."""
    rng = np.random.RandomState(seed)
    bound = math.sqrt(6.0 / (C + D))
    w = rng.uniform(-bound, bound, size=(C, D)).astype(dtype)
    label = rng.randint(0, C, size=(B,)).astype(np.int64)
    if trained_like:
        wh = w[label] / np.linalg.norm(w[label], axis=1, keepdims=True)
        n = rng.standard_normal((B, D)).astype(dtype)
        n /= np.linalg.norm(n, axis=1, keepdims=True)
        x = (3.0 * (wh + 0.5 * n)).astype(dtype)
    else:
        x = rng.standard_normal((B, D)).astype(dtype)
    return x, w, label


def cosine_topk(x, w, k):
    """Exact top-k by cosine, descending, ties by ascending index: (values [B, k], indices int64 [B, k]).

    Restates top-k over ArcMarginProduct.forward_test (/root/reference/arcface.py:65-67) and the product's retrieval
    call faiss.normalize_L2 + IndexFlat(d, METRIC_INNER_PRODUCT).search(x, k) (/root/reference/daodian_infer.py:
    225-230, 295-302).  faiss is a third-party dependency absent from the container (unpinned in the reference); its
    documented semantics for a flat inner-product index are an exhaustive scan returning the k highest scores in
    decreasing order, padded with (-inf, -1) when the index holds fewer than k rows.  tests/test_oracle.py pins this
    function against the golden `cos` matrices produced by the reference's own forward_test."""
    cos = forward_test(np.asarray(x, dtype=np.float64), np.asarray(w, dtype=np.float64))
    B, C = cos.shape
    order = np.argsort(-cos, axis=1, kind="stable")[:, :k]
    vals = np.take_along_axis(cos, order, axis=1)
    if C < k:
        vals = np.concatenate([vals, np.full((B, k - C), -np.inf, dtype=vals.dtype)], axis=1)
        order = np.concatenate([order, np.full((B, k - C), -1, dtype=order.dtype)], axis=1)
    return vals, order.astype(np.int64)



# ---------------------------------------------------------------------------------------------- class sampling
# SURVEY.md section 8f row N4.  The reference always trains every class; the sampling rule is PartialFC's
# (insightface, recognition/arcface_torch/partial_fc_v2.py `PartialFC_V2.sample`, not vendored in /root/reference --
# "parity unpinned" for the RULE; the head evaluated on the sampled rows is arcface.py itself and is pinned by the
# goldens like everything above):
#     positive = unique(labels of this shard);  perm = rand(num_local);  perm[positive] = 2.0
#     index = sort(topk(perm, num_sample).indices);  labels = searchsorted(index, labels)
#     sub_weight = weight[index]  -> normalised, multiplied with the embeddings, margin + softmax over these rows only
def partial_fc_sample(label_local, c_local: int, num_sample: int, scores: np.ndarray) -> np.ndarray:
    """Sorted class ids of the sample given the random scores `scores` [c_local] in [0, 1): the labels' classes
    (score forced to 2) and the best-scoring negatives, S = max(num_sample, min(B, c_local)) rows (the product fixes S
    on the host; PartialFC's fallback for num_sample < #positives is 'positives only').  label_local: -1 = not here."""
    label_local = np.asarray(label_local).reshape(-1)
    S = min(c_local, max(int(num_sample), min(label_local.size, c_local)))
    perm = np.array(scores, dtype=np.float64, copy=True)
    perm[label_local[label_local >= 0]] = 2.0
    # top-S by score; ties (only among the 2.0s, all of which are taken) do not matter
    index = np.argsort(-perm, kind="stable")[:S]
    return np.sort(index).astype(np.int64)


def sampled_head(x, w, label, index, s=64.0, m=0.40, easy_margin=False, grad_loss=1.0, dtype=np.float64):
    """The reference head (arcface.py:45-63 + CrossEntropyLoss + argmax + backward) on the sampled rows w[index] with
    the labels remapped into the sample (every label must be in `index`).  Returns (loss, argmax as ORIGINAL class
    ids, dx, dw of the full weight: zero rows outside the sample)."""
    index = np.asarray(index, dtype=np.int64)
    label = np.asarray(label).reshape(-1)
    pos = np.searchsorted(index, label)
    assert np.array_equal(index[pos], label), "a label is missing from the sample"
    ws = w[index]
    z = forward_logits(x, ws, pos, s, m, easy_margin, dtype=dtype)
    dx, dws = backward(x, ws, pos, s, m, easy_margin, grad_loss, dtype=dtype)
    dw = np.zeros(w.shape, dtype=dws.dtype)
    dw[index] = dws
    return cross_entropy(z, pos), index[argmax(z)], dx, dw
