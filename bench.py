#!/usr/bin/env python
"""bench.py -- ArcFace head fwd+bwd samples/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload: the north-star shape the metric is quoted on -- B=512, D=512, C=1,000,000, s=64, m=0.5, synthetic
embeddings / labels / xavier-uniform weights.  One step = K1 (normalise + cast of x) + label margin + K1 of the
class weights fused into K2 (cosine GEMM with margin / softmax / argmax epilogue) + K3 (dC^T, dW, dX GEMMs)
+ normalise backward; the optimiser is excluded.  At N > 1 the head is class-sharded
over the ranks (fixed global problem: strong scaling) with three NCCL collectives per step.

`value`   : device-timed (CUDA events, max over ranks), inputs resident in HBM.
`e2e`     : the same step through the host-buffer C-ABI call (N=1) / the public module API (N>1), with
            the pinned-host -> device copy of x / labels and the device -> host read of loss / argmax / dx
            inside the timed region.
`roofline`: for the stage with the largest share of the step, timed live with CUDA events.
`cpu_baseline` (N=1, rank 0): the reference's dense fp32 PyTorch path (oracle/arcface_torch_cpu.py port)
            on the host cores, on a bounded class sample scaled linearly in C.
--impl reference: only that CPU path, as its own JSON line.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = {"B": 512, "D": 512, "C": 1000000, "s": 64.0, "m": 0.5}
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` capture of this workload on one
# B200 (profiles/r1_v9_fwd_bwd_full_raw.csv); only meaningful for the single-GPU north-star shape.
NCU_TRAFFIC_BYTES = {"fwd": 2.543024e9 + 0.999243e9, "k3": 1.399593e9 + 2.114164e9}
METRIC = "arcface_head_fwd_bwd_samples_per_sec_1M_classes"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    except Exception:
        # fallback stated in /opt/skills/guides/B200_PROFILING.md
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.sm = []
        self.power = []
        self.mask = 0
        self.max_mhz = None
        self.error = None

    def run(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except Exception:
                    idx = self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag.is_set():
                self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                try:
                    self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                except Exception:
                    pass
                time.sleep(0.02)
        except Exception as e:  # pragma: no cover
            self.error = repr(e)

    def result(self):
        self.stop_flag.set()
        self.join(timeout=2.0)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "error": self.error or "no samples"}
        s = sorted(self.sm)
        out = {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "samples": len(s),
               "reasons": [n for b, n in self.REASONS.items() if self.mask & b]}
        if self.power:
            out["power_w_max"] = max(self.power)
        return out


def run_reference(args):
    """The reference's own CPU implementation of the path (port, see oracle/arcface_torch_cpu.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import arcface_torch_cpu as otc

    w = WORKLOAD
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    total_budget = args.ref_budget
    per_step = total_budget / max(1, args.steps + args.warmup)
    g = torch.Generator().manual_seed(0)
    c_probe = 2048
    xs = torch.randn(w["B"], w["D"], generator=g)
    ws = torch.randn(c_probe, w["D"], generator=g) * 0.05
    ys = torch.randint(0, c_probe, (w["B"],), generator=g)
    otc.head_step(xs, ws, ys, w["s"], w["m"])
    t0 = time.perf_counter()
    otc.head_step(xs, ws, ys, w["s"], w["m"])
    per_class = (time.perf_counter() - t0) / c_probe
    c_sample = int(min(w["C"], max(1024, per_step / max(per_class, 1e-9))))
    ws = torch.randn(c_sample, w["D"], generator=g) * 0.05
    ys = torch.randint(0, c_sample, (w["B"],), generator=g)
    for _ in range(args.warmup):
        otc.head_step(xs, ws, ys, w["s"], w["m"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        otc.head_step(xs, ws, ys, w["s"], w["m"])
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    t_full = dt * (w["C"] / c_sample)
    value = w["B"] / t_full
    sample = ("B=%d D=%d, %d of %d classes per step, fp32 torch eager fwd+bwd on %d host threads, "
              "time scaled linearly in C" % (w["B"], w["D"], c_sample, w["C"], threads))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_full * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    w = WORKLOAD
    return {"workload": "ArcFace head fwd+bwd, north-star shape B=%d D=%d C=%d s=%g m=%g (BASELINE.json metric shape)"
            % (w["B"], w["D"], w["C"], w["s"], w["m"]),
            "global_batch": w["B"], "embedding_dim": w["D"], "classes": w["C"],
            "parallelism": "single GPU" if n_gpus == 1 else "class-sharded x%d (PartialFC-style), NCCL" % n_gpus,
            "l2": "inputs exceed L2: fp32 class weights %.2f GB per step, no flush needed"
            % (w["C"] * w["D"] * 4 / 1e9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="host seconds the reference arm may spend")
    ap.add_argument("--classes", type=int, default=None, help="override C (debug only; invalidates the metric)")
    args = ap.parse_args()
    if args.classes:
        WORKLOAD["C"] = args.classes
    if args.impl == "reference":
        run_reference(args)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist

    import multimodalsimilar_b200 as mm
    from multimodalsimilar_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run for N > 1)" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    w = WORKLOAD
    B, D, C, s, m = w["B"], w["D"], w["C"], w["s"], w["m"]
    peaks = load_peaks()
    gen = torch.Generator(device="cpu").manual_seed(0)
    x_host = torch.randn(B, D, generator=gen).pin_memory()
    y_host = torch.randint(0, C, (B,), generator=gen).pin_memory()
    bound = math.sqrt(6.0 / (C + D))

    if world == 1:
        head = mm.ArcMarginProduct(D, 8, s=s, m=m)
        head.out_feature = C
        c_lo, c_hi = 0, C
    else:
        # ARCFACE_B200_P2P=0: exchanges through NCCL instead of peer-mapped memory (A/B measurements)
        head = mm.ShardedArcMarginProduct(D, world, s=s, m=m, use_p2p=os.environ.get("ARCFACE_B200_P2P", "1") != "0")
        head.out_feature = C
        head.class_lo, head.class_hi = mm.shard_range(C, world, rank)
        c_lo, c_hi = head.class_lo, head.class_hi
    gdev = torch.Generator(device=dev).manual_seed(1234 + rank)
    head.weight = torch.nn.Parameter(torch.empty(c_hi - c_lo, D, device=dev).uniform_(-bound, bound, generator=gdev))
    head = head.to(dev)
    b_loc = B // world
    x_dev = x_host[rank * b_loc:(rank + 1) * b_loc].to(dev).requires_grad_(True)
    y_dev = y_host[rank * b_loc:(rank + 1) * b_loc].to(dev)

    def step():
        x_dev.grad = None
        head.weight.grad = None
        loss, pred = head.loss(x_dev, y_dev)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        loss = step()
    ev1.record()
    barrier()
    clocks = sampler.result()
    ms_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    value = B / (ms_step * 1e-3)
    loss_value = float(loss.detach())

    # ---- kernel launches per step (our kernels only; memset / NCCL not counted)
    _, n_chunks = ops.backward_plan(B, D, c_hi - c_lo)
    k3_launches = ops.backward_launches(B, D, c_hi - c_lo)
    # K1(x), label, K1(w)+K2, combine, finalize, K3, bwd-x, scale_grads -- replayed from one CUDA graph per step
    launches_per_step = 1 + 1 + 1 + 1 + 1 + k3_launches + 1 + 1
    if world > 1 and getattr(head, "use_p2p", False):
        launches_per_step += 3  # the three peer-memory exchange kernels (csrc/p2p.cu) that replace the NCCL collectives
    gpu_launches = launches_per_step * args.steps

    # ---- end-to-end with host buffers
    h2d = b_loc * D * 4 + b_loc * 8
    d2h = 4 + b_loc * 8 + b_loc * D * 4
    e2e_steps = max(5, min(args.steps, 50))
    if world == 1:
        loss_h = torch.zeros(1).pin_memory()
        arg_h = torch.zeros(B, dtype=torch.int64).pin_memory()
        dx_h = torch.zeros(B, D).pin_memory()
        dw = torch.empty(C, D, device=dev)
        ws = torch.empty(ops.step_workspace_bytes(B, D, C), dtype=torch.uint8, device=dev)
        wdet = head.weight.detach()
        head.weight.grad = None
        torch.cuda.empty_cache()

        def e2e_step():
            ops.step_host(x_host, y_host, wdet, s, m, False, 1.0, loss_h, arg_h, dx_h, dw, ws)
    else:
        xl_host = x_host[rank * b_loc:(rank + 1) * b_loc].contiguous().pin_memory()
        yl_host = y_host[rank * b_loc:(rank + 1) * b_loc].contiguous().pin_memory()
        dx_h = torch.zeros(b_loc, D).pin_memory()
        arg_h = torch.zeros(b_loc, dtype=torch.int64).pin_memory()

        def e2e_step():
            xd = xl_host.to(dev, non_blocking=True).requires_grad_(True)
            yd = yl_host.to(dev, non_blocking=True)
            head.weight.grad = None
            l, p = head.loss(xd, yd)
            l.backward()
            dx_h.copy_(xd.grad, non_blocking=True)
            arg_h.copy_(p, non_blocking=True)
            return float(l)  # device -> host read of the loss; also orders the copies above

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
    barrier()
    e2e = {"value": B / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d * world,
           "d2h_bytes_per_step": d2h * world, "ms_per_step": e2e_ms,
           "api": "arcface_b200_step_host (C ABI, pinned host buffers)" if world == 1 else
                  "ShardedArcMarginProduct.loss + backward with pinned host copies"}

    # ---- stage breakdown + roofline (rank 0 view; live CUDA events, same inputs)
    c_loc = c_hi - c_lo
    stages = stage_times(torch, ops, head, x_host, y_host, dev, s, m, c_lo, C, max(5, min(args.steps, 30)))
    flops_alg = 6.0 * B * D * c_loc
    bytes_alg = 8.0 * c_loc * D + 8.0 * B * D + 24.0 * B
    p_tensor = peaks["bf16_tflops_sustained"]
    t_tensor = flops_alg / (p_tensor * 1e12)
    t_hbm = bytes_alg / (peaks["hbm_gbs"] * 1e9)
    t_roof = max(t_tensor, t_hbm)
    comp_ms = stages["k1_x"] + stages["label"] + stages["fwd"] + stages["k3"] + stages["bwd_x"]
    roofline_step = {"bound": "tensor" if t_tensor >= t_hbm else "hbm",
                     "achieved": flops_alg / (ms_step * 1e-3) / 1e12, "peak": p_tensor, "unit": "TFLOP/s",
                     "frac": t_roof / (ms_step * 1e-3), "t_roof_ms": t_roof * 1e3,
                     "algorithmic_flops": flops_alg, "algorithmic_bytes": bytes_alg,
                     "executed_flops": 8.0 * B * D * c_loc, "peak_source": peaks["source"] + " (sustained bf16)"}
    fwd_hbm_s = (c_loc * D * 6.0 + 4.0 * c_loc) / (peaks["hbm_gbs"] * 1e9)   # fp32 W in, bf16 What + 1/||w|| out
    fwd_tensor_s = 2.0 * B * D * c_loc / (p_tensor * 1e12)
    fwd_cand = (("hbm", (c_loc * D * 6.0 + 4.0 * c_loc) / 1e9, "GB/s", peaks["hbm_gbs"]) if fwd_hbm_s >= fwd_tensor_s
                else ("tensor", 2.0 * B * D * c_loc / 1e12, "TFLOP/s", p_tensor))
    cand = {
        "fwd": fwd_cand + ("K1(w)+K2 forward: in-kernel weight normalise/cast + cosine GEMM + softmax epilogue",),
        "k3": ("tensor", 4.0 * B * D * c_loc / 1e12, "TFLOP/s", p_tensor,
               "K3 backward (dC^T producer + dW GEMM + dX GEMM, %s)"
               % ("one persistent launch, dC^T through an L2 ring" if k3_launches == 1 else "%d chunks" % n_chunks)),
    }
    top = max(cand, key=lambda k: stages[k])
    bnd, work, unit, peak, name = cand[top]
    achieved = work / (stages[top] * 1e-3)
    roofline = {"kernel": name, "bound": bnd, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                "traffic": NCU_TRAFFIC_BYTES[top] if (world == 1 and not args.classes) else None,
                "traffic_source": "profiles/r1_v9_fwd_bwd_full_raw.csv (ncu --set full, per launch)",
                "ms_per_launch_group": stages[top], "share_of_step": stages[top] / comp_ms,
                "peak_source": peaks["source"]}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import arcface_torch_cpu as otc

        torch.set_num_threads(os.cpu_count() or 1)
        cpu_baseline = otc.time_head_step(B, D, C, s, m, budget_s=20.0)
        cpu_baseline = {k: cpu_baseline[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline,
            "roofline_step": roofline_step, "stages_ms": stages, "loss": loss_value,
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if world > 1:
        # Captured graphs hold NCCL kernels: release them before the communicator goes away, and leave without
        # running destructors (destroy_process_group() with live graphs was seen to hang after the line was printed).
        from multimodalsimilar_b200 import engine

        engine.drop_plan(head)
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def stage_times(torch, ops, head, x_host, y_host, dev, s, m, c_lo, c_total, iters):
    """Average milliseconds of each stage of one rank's step (whole batch, local class shard), timed with
    CUDA events on the launching stream, after warm-up.  The collectives of the sharded head are not part
    of this breakdown."""
    x = x_host.to(dev)
    y = y_host.to(dev)
    w = head.weight.detach()
    B, D = x.shape
    names = ["k1_x", "label", "fwd", "k3", "bwd_x"]
    acc = {n: 0.0 for n in names}
    dw = torch.empty_like(w)
    for it in range(iters + 2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        ev[0].record()
        xhat, inv_nx, xhat_t = ops.normalize_cast(x, want_transpose=True)
        ev[1].record()
        lm = ops.label_margin(x, w, inv_nx, None, y, c_lo, c_total, s, m, False)
        ev[2].record()
        what, inv_nw, rmax, rsum, rarg = ops.forward_rows_fused(xhat, w, lm.label_local, s, c_lo)
        lse, arg, zl, omp, loss = ops.finalize_rows(rmax.view(1, B), rsum.view(1, B), rarg.view(1, B),
                                                    lm.z_label.view(1, B), y)
        ev[3].record()
        dxhat, _ = ops.backward(xhat, xhat_t, what, inv_nw, lse, omp, lm.dphi, lm.label_local, s, 1.0 / B, dw_out=dw)
        ev[4].record()
        ops.normalize_bwd_x(x, inv_nx, dxhat)
        ev[5].record()
        torch.cuda.synchronize()
        if it >= 2:
            for i, n in enumerate(names):
                acc[n] += ev[i].elapsed_time(ev[i + 1])
    return {n: acc[n] / iters for n in names}


if __name__ == "__main__":
    main()
