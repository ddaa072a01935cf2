"""K4 -- fused cosine top-k through the C ABI against the oracle and against a dense
fp32 evaluation of the same bf16 operands; faiss-style CosineIndex; edge cases (ties, tails, C < k, k = 1 / 128)."""
import numpy as np
import pytest
import torch

from oracle import arcface_numpy as onp

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _dense_topk(xhat, what, k):
    """top-k of the cosines the kernel's operands define (bf16 values, fp32 accumulate), on the GPU."""
    cos = xhat.float() @ what.float().t()
    kk = min(k, cos.shape[1])
    v, i = torch.topk(cos, kk, dim=1)
    return cos, v, i


def _check(xhat, what, k, vals, idx, atol=2e-5):
    cos, rv, ri = _dense_topk(xhat, what, k)
    kk = rv.shape[1]
    assert vals.shape == (xhat.shape[0], k) and idx.dtype == torch.int64
    # values: exact top-k of the same operands up to fp32 summation order
    torch.testing.assert_close(vals[:, :kk], rv, rtol=0, atol=atol)
    # descending
    assert bool((vals[:, 1:kk] <= vals[:, :kk - 1]).all())
    # indices point at those values, are valid and distinct per row
    got = torch.gather(cos, 1, idx[:, :kk])
    torch.testing.assert_close(got, vals[:, :kk], rtol=0, atol=atol)
    srt = torch.sort(idx[:, :kk], dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all()) if kk > 1 else True
    # where the reference ranking is separated by more than the noise, the indices agree exactly
    if kk > 1:
        gap_ok = torch.ones_like(ri, dtype=torch.bool)
        gap_ok[:, 1:] &= (rv[:, :-1] - rv[:, 1:]) > 4 * atol
        gap_ok[:, :-1] &= (rv[:, :-1] - rv[:, 1:]) > 4 * atol
        # the k-th entry also needs a gap to the (k+1)-th
        if kk < cos.shape[1]:
            nxt = torch.topk(cos, kk + 1, dim=1).values[:, kk]
            gap_ok[:, -1] &= (rv[:, -1] - nxt) > 4 * atol
        assert bool((idx[:, :kk][gap_ok] == ri[gap_ok]).all())
    if k > kk:
        assert bool(torch.isneginf(vals[:, kk:]).all()) and bool((idx[:, kk:] == -1).all())


@pytest.mark.parametrize("B,D,C,k", [
    (64, 128, 3000, 13),      # ragged tail: 3000 = 11 * 256 + 184
    (64, 128, 3000, 1),
    (64, 128, 3000, 128),
    (200, 64, 257, 26),       # B not a multiple of 128, one column past a tile
    (8, 16, 32, 100),         # C < k: padded with (-inf, -1)
    (130, 1024, 40000, 100),  # D > 512 (RoBERTa-large width)
    (512, 512, 300000, 100),  # bench-like
])
def test_topk_matches_dense(B, D, C, k):
    from multimodalsimilar_b200 import ops

    x, w, _ = onp.synthetic_inputs(B, D, C, seed=9, trained_like=(C >= 3000))
    xhat, _, _ = ops.normalize_cast(torch.from_numpy(x).to(dev()))
    what, _, _ = ops.normalize_cast(torch.from_numpy(w).to(dev()))
    vals, idx = ops.cosine_topk(xhat, what, k)
    _check(xhat, what, k, vals, idx)


def test_topk_against_oracle_small():
    """fp64 oracle on fp32 inputs: bf16 operand rounding bounds the value error (|d cos| <~ 2^-8 / sqrt(D) * few)."""
    import multimodalsimilar_b200 as mm

    B, D, C, k = 32, 256, 2000, 10
    x, w, _ = onp.synthetic_inputs(B, D, C, seed=2, trained_like=True)
    head = mm.ArcMarginProduct(D, C).to(dev())
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(w))
    vals, idx = head.predict_topk(torch.from_numpy(x).to(dev()), k)
    rv, ri = onp.cosine_topk(x, w, k)
    np.testing.assert_allclose(vals.cpu().numpy(), rv, rtol=0, atol=4e-3)
    # trained-like rows: the best class is the label's, separated by far more than bf16 noise
    np.testing.assert_array_equal(idx[:, 0].cpu().numpy(), ri[:, 0])
    cos = onp.forward_test(x.astype(np.float64), w.astype(np.float64))
    np.testing.assert_allclose(np.take_along_axis(cos, idx.cpu().numpy(), axis=1), rv, rtol=0, atol=8e-3)


def test_topk_ties_prefer_lower_index():
    from multimodalsimilar_b200 import ops

    B, D, C, k = 16, 64, 1000, 8
    x, w, _ = onp.synthetic_inputs(B, D, C, seed=3, trained_like=False)
    w[500:520] = w[7]            # twenty duplicates of one catalogue row
    x[:] = w[7] + 0.01 * x       # every query close to it: the top of every list is a 21-way tie
    xhat, _, _ = ops.normalize_cast(torch.from_numpy(x).to(dev()))
    what, _, _ = ops.normalize_cast(torch.from_numpy(w).to(dev()))
    vals, idx = ops.cosine_topk(xhat, what, k)
    assert bool((vals[:, :1] == vals).all())                     # all k winners are the tied value
    expect = torch.tensor([7] + list(range(500, 500 + k - 1)), device=dev())
    assert bool((idx == expect[None]).all())                     # lowest indices of the tie, ascending


def test_cosine_index_is_a_flat_ip_index():
    import multimodalsimilar_b200 as mm

    n, d, k = 5000, 128, 13
    rng = np.random.RandomState(0)
    arr = rng.standard_normal((n, d)).astype(np.float32) * 3.0   # un-normalised, like the embeddings before normalize_L2
    index = mm.CosineIndex(d)
    index.add(arr[:3000])
    index.add(arr[3000:])
    assert index.ntotal == n
    D_, I_ = index.search(arr[:2500], k)                         # > MAX_BATCH queries: chunked launches
    assert D_.shape == (2500, k) and I_.shape == (2500, k)
    assert bool((I_[:, 0] == torch.arange(2500, device=I_.device)).all())     # self-match first
    assert float((D_[:, 0] - 1.0).abs().max()) <= 1e-2                        # cos(x, x) = 1 up to bf16 rounding
    rv, ri = onp.cosine_topk(arr[:64], arr, k)
    np.testing.assert_allclose(D_[:64].cpu().numpy(), rv, rtol=0, atol=1e-2)
    # k beyond the fused kernel's lists (daodian_infer.py:230 searches with k = len(slice)): materialising path
    D2, I2 = index.search(arr[:8], 200)
    rv2, _ = onp.cosine_topk(arr[:8], arr, 200)
    np.testing.assert_allclose(D2.cpu().numpy(), rv2, rtol=0, atol=1e-2)
    assert bool((I2[:, 0] == torch.arange(8, device=I2.device)).all())


def test_cosine_index_any_width_and_high_precision():
    """The reference's fastText retrieval builds IndexFlat(100) (daodian_infer.py:227): widths that are not multiples of
    8 are zero-padded; precision='bf16x3' brings the scores within 1e-5 of fp32 cosines (they are compared against
    tuned thresholds downstream)."""
    import multimodalsimilar_b200 as mm

    n, d, k = 3000, 100, 26
    rng = np.random.RandomState(1)
    arr = rng.standard_normal((n, d)).astype(np.float32)
    rv, ri = onp.cosine_topk(arr[:50], arr, k)
    for prec, tol in (("bf16", 1e-2), ("bf16x3", 2e-5)):
        index = mm.CosineIndex(d, precision=prec)
        index.add(arr)
        D_, I_ = index.search(arr[:50], k)
        np.testing.assert_allclose(D_.cpu().numpy(), rv, rtol=0, atol=tol)
        # the returned ids are the top-k up to near-ties: the fp64 cosines at those ids are the fp64 top-k values
        an = arr.astype(np.float64) / np.linalg.norm(arr.astype(np.float64), axis=1, keepdims=True)
        cos = an[:50] @ an.T
        np.testing.assert_allclose(np.take_along_axis(cos, I_.cpu().numpy(), axis=1), rv, rtol=0, atol=tol)


def test_topk_merge_of_shard_lists():
    """The sharded path's merge kernel: per-shard top-k lists -> global top-k == top-k over the whole catalogue."""
    from multimodalsimilar_b200 import ops

    B, D, C, k, R = 48, 128, 9000, 20, 3
    x, w, _ = onp.synthetic_inputs(B, D, C, seed=5, trained_like=False)
    xhat, _, _ = ops.normalize_cast(torch.from_numpy(x).to(dev()))
    what, _, _ = ops.normalize_cast(torch.from_numpy(w).to(dev()))
    per = C // R
    vs, is_ = [], []
    for r in range(R):
        v, i = ops.cosine_topk(xhat, what[r * per:(r + 1) * per].contiguous(), k, 1.0, r * per)
        vs.append(v)
        is_.append(i)
    mv, mi = ops.topk_merge(torch.cat(vs, dim=1).contiguous(), torch.cat(is_, dim=1).contiguous(), k)
    gv, gi = ops.cosine_topk(xhat, what, k)
    assert torch.equal(mv, gv) and torch.equal(mi, gi)
