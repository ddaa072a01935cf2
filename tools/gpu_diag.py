"""Stage-by-stage GPU diagnostics: each check runs in its own subprocess (a device fault in one stage
must not hide the others) with a timeout, and prints the error statistics of that stage.

    python tools/gpu_diag.py            # all stages
    python tools/gpu_diag.py k1 logits  # selected stages
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["k1", "logits", "stats", "bwd_dw", "bwd_dx", "module"]


def stage_k1():
    import torch
    from multimodalsimilar_b200 import ops
    src = torch.randn(1000, 512, device="cuda") * 0.3
    dst, inv, dst_t = ops.normalize_cast(src, want_transpose=True)
    ref = torch.nn.functional.normalize(src)
    print("k1: max|dst-ref| %.3e  inv rel err %.3e  transpose ok %s" % (
        float((dst.float() - ref).abs().max()), float((inv * src.norm(dim=1) - 1).abs().max()),
        bool(torch.equal(dst_t[:, :1000], dst.t()))))


def stage_logits():
    import torch
    from multimodalsimilar_b200 import ops
    for (B, D, C) in [(128, 64, 256), (64, 512, 1000), (5, 24, 37), (512, 512, 4099)]:
        g = torch.Generator().manual_seed(B + C)
        xh = torch.randn(B, D, generator=g).cuda().bfloat16()
        wh = torch.randn(C, D, generator=g).cuda().bfloat16()
        out = ops.logits(xh, wh, None, None, 1.0)
        torch.cuda.synchronize()
        ref = (xh.double() @ wh.double().t()).float()
        err = (out - ref).abs()
        print("logits B=%d D=%d C=%d: max err %.3e (ref max %.2f)  bad elems %d" % (
            B, D, C, float(err.max()), float(ref.abs().max()), int((err > 1e-2).sum())))
        if float(err.max()) > 1e-2:
            bad = torch.nonzero(err > 1e-2)
            print("   first bad (row, col):", bad[:8].tolist(), " rows", bad[:, 0].unique()[:16].tolist(),
                  " cols%64", (bad[:, 1] % 64).unique()[:16].tolist())


def _setup(B, D, C, s=64.0, m=0.4, trained=True):
    import numpy as np
    import torch
    from multimodalsimilar_b200 import ops
    from oracle import arcface_numpy as onp
    x, w, y = onp.synthetic_inputs(B, D, C, seed=3, trained_like=trained)
    xt, wt, yt = (torch.from_numpy(a).cuda() for a in (x, w, y))
    xhat, inv_nx, xhat_t = ops.normalize_cast(xt, want_transpose=True)
    what, inv_nw, _ = ops.normalize_cast(wt)
    lm = ops.label_margin(xt, wt, inv_nx, inv_nw, yt, 0, C, s, m, False)
    return dict(x=x, w=w, y=y, xt=xt, wt=wt, yt=yt, xhat=xhat, inv_nx=inv_nx, xhat_t=xhat_t, what=what,
                inv_nw=inv_nw, lm=lm, s=s, m=m, B=B, D=D, C=C, np=np, onp=onp, ops=ops, torch=torch)


def stage_stats():
    c = _setup(200, 64, 3000)
    torch, ops, onp, np = c["torch"], c["ops"], c["onp"], c["np"]
    B = c["B"]
    rmax, rsum, rarg = ops.forward_rows(c["xhat"], c["what"], c["lm"].label_local, c["s"], 0)
    lse, arg, zl, omp, loss = ops.finalize_rows(rmax.view(1, B), rsum.view(1, B), rarg.view(1, B),
                                                c["lm"].z_label.view(1, B), c["yt"])
    torch.cuda.synchronize()
    z = onp.forward_logits(c["x"], c["w"], c["y"], c["s"], c["m"], False, dtype=np.float64)
    zmax, rlse = onp.row_stats(z)
    print("stats: loss %.6f (oracle %.6f)  max|lse| err %.3e  max|rowmax| err %.3e  argmax mismatches %d  t_label err %.2e" % (
        float(loss), onp.cross_entropy(z, c["y"]), float(np.abs(lse.cpu().numpy() - rlse).max()),
        float(np.abs(rmax.cpu().numpy() - zmax).max()), int((arg.cpu().numpy() != onp.argmax(z)).sum()),
        float(np.abs(c["lm"].t_label.cpu().numpy() - onp.cosines(c["x"].astype(np.float64), c["w"].astype(np.float64))[np.arange(B), c["y"]]).max())))


def _bwd(c):
    torch, ops = c["torch"], c["ops"]
    B = c["B"]
    rmax, rsum, rarg = ops.forward_rows(c["xhat"], c["what"], c["lm"].label_local, c["s"], 0)
    lse, arg, zl, omp, loss = ops.finalize_rows(rmax.view(1, B), rsum.view(1, B), rarg.view(1, B),
                                                c["lm"].z_label.view(1, B), c["yt"])
    dxhat, dw = ops.backward(c["xhat"], c["xhat_t"], c["what"], c["inv_nw"], lse, omp, c["lm"].dphi, c["lm"].label_local,
                             c["s"], 1.0 / B)
    dx = ops.normalize_bwd_x(c["xt"], c["inv_nx"], dxhat)
    torch.cuda.synchronize()
    return dx, dw


def stage_bwd_dw():
    for shape in [(200, 64, 3000), (64, 512, 1000), (300, 256, 2049)]:
        c = _setup(*shape)
        np, onp = c["np"], c["onp"]
        dx, dw = _bwd(c)
        rdx, rdw = onp.backward(c["x"], c["w"], c["y"], c["s"], c["m"], False, dtype=np.float64)
        gdw = dw.cpu().numpy()
        print("bwd_dw %s: max|ddw| %.3e  rel fro %.3e  (|dw| max %.3e)" % (
            shape, np.abs(gdw - rdw).max(), np.linalg.norm(gdw - rdw) / np.linalg.norm(rdw), np.abs(rdw).max()))


def stage_bwd_dx():
    for shape in [(200, 64, 3000), (64, 512, 1000), (300, 256, 2049)]:
        c = _setup(*shape)
        np, onp = c["np"], c["onp"]
        dx, dw = _bwd(c)
        rdx, rdw = onp.backward(c["x"], c["w"], c["y"], c["s"], c["m"], False, dtype=np.float64)
        gdx = dx.cpu().numpy()
        print("bwd_dx %s: max|ddx| %.3e  rel fro %.3e  (|dx| max %.3e)" % (
            shape, np.abs(gdx - rdx).max(), np.linalg.norm(gdx - rdx) / np.linalg.norm(rdx), np.abs(rdx).max()))


def stage_full():
    """Large shapes against the fp32 torch restatement on the GPU; prints where the errors sit."""
    import numpy as np
    import torch
    import multimodalsimilar_b200 as mm
    from oracle import arcface_numpy as onp
    from tests.test_gpu_parity import torch_oracle
    for (B, D, C, s, m) in [(256, 1792, 100000, 64.0, 0.2), (128, 64, 600000, 64.0, 0.4), (512, 512, 200000, 64.0, 0.5)]:
        for trained in (False, True):
            x, w, y = onp.synthetic_inputs(B, D, C, seed=5, trained_like=trained)
            head = mm.ArcMarginProduct(D, 8, s=s, m=m)
            head.out_feature = C
            head.weight = torch.nn.Parameter(torch.from_numpy(w).cuda())
            xt = torch.from_numpy(x).cuda().requires_grad_(True)
            yt = torch.from_numpy(y).cuda()
            loss, pred = head.loss(xt, yt)
            loss.backward()
            torch.cuda.synchronize()
            rloss, _rp, z, rdx, rdw = torch_oracle(xt.detach(), head.weight.detach(), yt, s, m, False)
            dw, dx = head.weight.grad, xt.grad
            ew = (dw - rdw).abs()
            ex = (dx - rdx).abs()
            print("full B=%d D=%d C=%d trained=%s: loss %.6f ref %.6f | dx max err %.3e rel fro %.3e (max |dx| %.3e) | "
                  "dw max err %.3e rel fro %.3e (max |dw| %.3e) nan %d" % (
                      B, D, C, trained, float(loss), float(rloss), float(ex.max()), float((dx - rdx).norm() / rdx.norm()),
                      float(rdx.abs().max()), float(ew.max()), float((dw - rdw).norm() / rdw.norm()),
                      float(rdw.abs().max()), int(torch.isnan(dw).sum())))
            thr = max(1e-3 * float(rdw.abs().max()), 10 * float(ew.median()))
            bad = torch.nonzero(ew > max(thr, 1e-9))
            if bad.numel():
                rows = bad[:, 0].unique()
                print("   dw: %d bad elems in %d rows; first rows %s ; rows %% 128: %s ; cols min/max %d/%d ; label rows? %s" % (
                    bad.shape[0], rows.numel(), rows[:12].tolist(), (rows % 128).unique()[:20].tolist(),
                    int(bad[:, 1].min()), int(bad[:, 1].max()),
                    bool(torch.isin(rows, yt).all())))
                r0 = int(rows[0])
                print("   row %d: got %s ref %s" % (r0, dw[r0, :4].tolist(), rdw[r0, :4].tolist()))
            del head, z, rdw, rdx, dw, dx
            torch.cuda.empty_cache()


def stage_module():
    import __graft_entry__ as g
    g.smoke()


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        globals()["stage_" + sys.argv[2]]()
        return
    todo = sys.argv[1:] or STAGES
    failed = []
    for st in todo:
        print("=== %s" % st, flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", st], timeout=420)
            if r.returncode != 0:
                failed.append(st)
                print("!!! stage %s exited with %d" % (st, r.returncode), flush=True)
        except subprocess.TimeoutExpired:
            failed.append(st)
            print("!!! stage %s timed out" % st, flush=True)
    print("diag done; failed stages:", failed)


if __name__ == "__main__":
    main()
