"""The oracle (oracle/arcface_numpy.py) and the timed CPU port (oracle/arcface_torch_cpu.py) against the
golden vectors minted from the unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import arcface_numpy as onp
from oracle import arcface_torch_cpu as otc

SMALL = ["base", "easy", "fallback", "easy_neg", "zero_row", "tie", "label_col", "trained", "grad10", "ragged"]


def _case(golden, name):
    s, m, easy, grad = golden[name + "/hp"]
    if name == "c1":
        x, w, y = onp.synthetic_inputs(64, 512, 1000, seed=1)
    else:
        x, w, y = golden[name + "/x"], golden[name + "/w"], golden[name + "/label"]
    return x, w, y, float(s), float(m), bool(easy), float(grad)


@pytest.mark.parametrize("name", SMALL)
def test_forward_matches_reference(golden, name):
    x, w, y, s, m, easy, _ = _case(golden, name)
    z = onp.forward_logits(x, w, y, s, m, easy, dtype=np.float32)
    np.testing.assert_allclose(z, golden[name + "/logits"], rtol=0, atol=2e-5 * s)
    assert abs(onp.cross_entropy(z, y) - float(golden[name + "/loss"])) <= 1e-5 * max(1.0, float(golden[name + "/loss"]))
    np.testing.assert_array_equal(onp.argmax(golden[name + "/logits"]), golden[name + "/argmax"])
    np.testing.assert_allclose(onp.forward_test(x, w), golden[name + "/cos"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", SMALL)
def test_backward_matches_reference(golden, name):
    x, w, y, s, m, easy, grad = _case(golden, name)
    dx, dw = onp.backward(x, w, y, s, m, easy, grad_loss=grad, dtype=np.float64)
    gx, gw = golden[name + "/dx"], golden[name + "/dw"]
    if name == "zero_row":  # dx of the zero-norm row is dXhat / 1e-12 in the reference: compare relatively
        np.testing.assert_allclose(dx, gx, rtol=2e-3, atol=1e-5 * max(1.0, np.abs(gx).max()) * 1e-3)
    else:
        np.testing.assert_allclose(dx, gx, rtol=0, atol=5e-5 * max(1.0, np.abs(gx).max()))
    np.testing.assert_allclose(dw, gw, rtol=0, atol=5e-5 * max(1.0, np.abs(gw).max()))


def test_branches_are_exercised(golden):
    # fallback: row 0 has cos_label < th -> s * (cos - mm); easy_neg: row 0 has cos <= 0 -> s * cos
    x, w, y, s, m, easy, _ = _case(golden, "fallback")
    t = onp.cosines(x, w)[0, y[0]]
    _, _, th, mm = onp.margin_constants(m)
    assert t < th
    assert abs(golden["fallback/logits"][0, y[0]] - s * (t - mm)) < 1e-3
    assert abs(golden["easy_neg/logits"][0, y[0]] - s * t) < 1e-3
    # zero-norm row: logits 0 off the label, -s * sin(m) on it
    zr = golden["zero_row/logits"][1]
    yz = golden["zero_row/label"][1]
    assert abs(zr[yz] + 30.0 * np.sin(0.5)) < 1e-4 and np.abs(np.delete(zr, yz)).max() == 0.0
    # tie: two identical winning columns -> the lower index
    assert golden["tie/argmax"][2] == 5


def test_c1_config(golden):
    x, w, y, s, m, easy, grad = _case(golden, "c1")
    loss, am = onp.loss_and_argmax(x, w, y, s, m, easy)
    assert abs(loss - float(golden["c1/loss"])) <= 1e-4 * float(golden["c1/loss"])
    np.testing.assert_array_equal(am, golden["c1/argmax"])
    rows = golden["c1/dw_rows"]
    z = onp.forward_logits(x, w, y, s, m, easy)
    np.testing.assert_allclose(z[:, rows], golden["c1/logits_rows"], rtol=0, atol=2e-5 * s)
    dx, dw = onp.backward(x, w, y, s, m, easy, grad)
    np.testing.assert_allclose(dx, golden["c1/dx"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(dw[rows], golden["c1/dw"], rtol=0, atol=2e-6)


def test_margin_schedule(golden):
    u = golden["update_m"]
    m = onp.update_m(0.5, 0.04)
    assert m == u[0]
    np.testing.assert_allclose(onp.margin_constants(m), u[1:5], rtol=0, atol=0)
    assert onp.update_m(m, 2.0) == u[5] == m          # rejected: above 1.0
    assert onp.update_m(0.5, -0.6) == u[6] == 0.5     # rejected: below 1e-6
    np.testing.assert_allclose(onp.margin_constants(0.4), golden["const_m04"], rtol=0, atol=0)


@pytest.mark.parametrize("name", ["base", "easy", "fallback", "trained", "grad10", "ragged"])
def test_torch_cpu_port_matches_reference(golden, name):
    x, w, y, s, m, easy, grad = _case(golden, name)
    loss, pred, dx, dw = otc.head_step(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(y), s, m, easy)
    assert abs(float(loss) - float(golden[name + "/loss"])) <= 1e-6 * max(1.0, float(golden[name + "/loss"]))
    np.testing.assert_array_equal(pred.numpy(), golden[name + "/argmax"])
    np.testing.assert_allclose(dx.numpy() * grad, golden[name + "/dx"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dw.numpy() * grad, golden[name + "/dw"], rtol=1e-5, atol=1e-6)


def test_row_stats_consistency():
    x, w, y = onp.synthetic_inputs(16, 32, 100, seed=3)
    z = onp.forward_logits(x, w, y, 64.0, 0.4, False, dtype=np.float64)
    zmax, lse = onp.row_stats(z)
    assert abs(np.mean(lse - z[np.arange(16), y]) - onp.cross_entropy(z, y)) < 1e-9
    assert np.all(zmax == z.max(axis=1))


@pytest.mark.parametrize("name", ["base", "tie", "trained", "ragged"])
def test_topk_oracle_against_reference_cosines(golden, name):
    """oracle.cosine_topk against top-k of the golden `cos` matrix (the reference's forward_test output)."""
    x, w, _, _, _, _, _ = _case(golden, name)
    cos = golden[name + "/cos"].astype(np.float64)
    for k in (1, 5, cos.shape[1], cos.shape[1] + 3):
        vals, idx = onp.cosine_topk(x, w, k)
        kk = min(k, cos.shape[1])
        ref_vals = -np.sort(-cos, axis=1)[:, :kk]
        np.testing.assert_allclose(vals[:, :kk], ref_vals, rtol=0, atol=2e-6)
        # indices: the reference's own cosines at the returned positions are the top-k values
        np.testing.assert_allclose(np.take_along_axis(cos, idx[:, :kk], axis=1), ref_vals, rtol=0, atol=2e-6)
        if k > kk:
            assert np.all(np.isneginf(vals[:, kk:])) and np.all(idx[:, kk:] == -1)


@pytest.mark.parametrize("easy", [False, True])
def test_oracle_backward_is_the_gradient_of_the_oracle_loss(easy):
    """Central finite differences of the fp64 oracle loss against oracle.backward (independent of the goldens)."""
    B, D, C, s, m = 4, 8, 11, 16.0, 0.35
    x, w, y = onp.synthetic_inputs(B, D, C, seed=13, trained_like=False)
    x, w = x.astype(np.float64), w.astype(np.float64)

    def loss(xx, ww):
        return onp.cross_entropy(onp.forward_logits(xx, ww, y, s, m, easy, dtype=np.float64), y)

    dx, dw = onp.backward(x, w, y, s, m, easy, dtype=np.float64)
    h = 1e-6
    rng = np.random.RandomState(0)
    for _ in range(12):
        b, d = rng.randint(B), rng.randint(D)
        xp, xm = x.copy(), x.copy()
        xp[b, d] += h
        xm[b, d] -= h
        assert abs((loss(xp, w) - loss(xm, w)) / (2 * h) - dx[b, d]) <= 1e-6 * max(1.0, abs(dx[b, d]))
        c, d = rng.randint(C), rng.randint(D)
        wp, wm = w.copy(), w.copy()
        wp[c, d] += h
        wm[c, d] -= h
        assert abs((loss(x, wp) - loss(x, wm)) / (2 * h) - dw[c, d]) <= 1e-6 * max(1.0, abs(dw[c, d]))


@pytest.mark.parametrize("name", ["base", "easy", "fallback", "easy_neg", "tie", "label_col", "trained", "grad10", "ragged"])
@pytest.mark.parametrize("chunk", [7, 1 << 16])
def test_chunked_torch_oracle_matches_reference(golden, name, chunk):
    """oracle/arcface_torch_chunked.py (the full-size checker of bench.py / the -m gpu tests) against the goldens
    minted from the unmodified reference, with the classes cut into ragged chunks and in one piece."""
    from oracle import arcface_torch_chunked as och

    x, w, y, s, m, easy, grad = _case(golden, name)
    C = w.shape[0]
    lo, hi = C // 3, C - 1
    r = och.head_step_chunked(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(y), s, m, easy,
                              grad_loss=grad, chunk=chunk, dw_range=(lo, hi))
    gl = float(golden[name + "/loss"])
    assert abs(float(r["loss"]) - gl) <= 2e-6 * max(1.0, abs(gl))
    np.testing.assert_array_equal(r["argmax"].numpy(), golden[name + "/argmax"])
    z = golden[name + "/logits"]
    top = -np.sort(-z, axis=1)
    np.testing.assert_allclose(r["top2_gap"].numpy(), top[:, 0] - top[:, 1], rtol=0, atol=3e-5 * s)
    np.testing.assert_allclose(r["z_label"].numpy(), z[np.arange(len(y)), y.reshape(-1)], rtol=0, atol=2e-5 * s)
    gx, gw = golden[name + "/dx"], golden[name + "/dw"]
    np.testing.assert_allclose(r["dx"].numpy(), gx, rtol=1e-4, atol=5e-6 * max(1.0, np.abs(gx).max()))
    np.testing.assert_allclose(r["dw"].numpy(), gw[lo:hi], rtol=1e-4, atol=5e-6 * max(1.0, np.abs(gw).max()))


def test_chunked_torch_oracle_matches_dense_port():
    from oracle import arcface_torch_chunked as och

    x, w, y = onp.synthetic_inputs(48, 96, 1500, seed=5, trained_like=True)
    xt, wt, yt = torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(y)
    loss, pred, dx, dw = otc.head_step(xt, wt, yt, 64.0, 0.4, False)
    r = och.head_step_chunked(xt, wt, yt, 64.0, 0.4, False, chunk=256)
    assert abs(float(r["loss"]) - float(loss)) <= 1e-5 * max(1.0, float(loss))
    assert torch.equal(r["argmax"], pred)
    torch.testing.assert_close(r["dx"], dx, rtol=1e-4, atol=2e-6 * max(1.0, float(dx.abs().max())))
    torch.testing.assert_close(r["dw"], dw, rtol=1e-4, atol=2e-6 * max(1.0, float(dw.abs().max())))
