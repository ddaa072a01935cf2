#!/usr/bin/env python
"""bench.py -- ArcFace head fwd+bwd samples/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config ns|c1|c2|c3|c4|c5_100k|c5_1m|c5_10m]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload: the north-star shape the metric is quoted on -- B=512, D=512, C=1,000,000, s=64, m=0.5,
synthetic embeddings / labels / xavier-uniform weights.  One step = K1 (normalise + cast of x) + label margin +
K1 of the class weights fused into K2 (cosine GEMM with margin / softmax / argmax epilogue) + K3 (dC^T, dW, dX
GEMMs) + normalise backward; the optimiser is excluded.  At N > 1 the head is class-sharded
over the ranks (fixed global problem: strong scaling) with three exchanges per step (peer memory or NCCL: see
`exchange`).  The synthetic weight matrix is a function of (seed, class id) only, so every N sees the same problem.

`value`      : device-timed (CUDA events, max over ranks), inputs resident in HBM; mean over the K steps
               (`ms_per_step`), with the per-step median and best beside it.
`sustained`  : the same step looped for >= 1 s (the power-capped regime), with the clocks seen there.
`e2e`        : the same step through the host-buffer C-ABI call (N=1) / the public module API (N>1), with
               the pinned-host -> device copy of x / labels and the device -> host read of loss / argmax / dx
               inside the timed region.
`parity`     : computed IN THIS RUN on every rank: loss / argmax / dx / dW-shard of the timed head against the fp32
               restatement of the reference evaluated on the same GPU on identical inputs
               (oracle/arcface_torch_chunked.py, the checker), max over ranks, with the north-star gates.
`roofline`   : for the stage with the largest share of the step, timed live with CUDA events (N=1).
`roofline_step`: t_roof / t_step for the whole step against BOTH measured bf16 peaks (burst and sustained);
               `frac` is the one matching the clocks sampled during the timed region (`regime`).
`configs`    : the other BASELINE.json configurations (c2..c5), each timed and parity-checked the same way
               (short runs; not the headline).  --config X makes X the headline instead and skips the others.
`gpu_reference`: the reference's dense fp32 eager op sequence (oracle/arcface_torch_cpu.py port) timed on the same
               B200 (N=1, rank 0): the GPU kernel chain to beat, reported, not part of the product.
`cpu_baseline` (N=1, rank 0): the same port on the host cores, on a bounded class sample scaled linearly in C.
--impl reference: only that CPU path, as its own JSON line.
"""
import argparse
import gc
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json `configs` (c1..c5 + the north-star shape the metric is quoted on)
CONFIGS = {
    "ns": {"B": 512, "D": 512, "C": 1000000, "s": 64.0, "m": 0.5,
           "what": "north-star shape (BASELINE.json metric shape)"},
    "c1": {"B": 64, "D": 512, "C": 1000, "s": 30.0, "m": 0.5, "what": "BASELINE config 1 (the reference's CPU case)"},
    "c2": {"B": 256, "D": 1792, "C": 100000, "s": 64.0, "m": 0.2,
           "what": "BASELINE config 2: EfficientNet-B4 embedding head"},
    "c3": {"B": 512, "D": 1024, "C": 1000000, "s": 64.0, "m": 0.4,
           "what": "BASELINE config 3: RoBERTa-wwm-ext-large embedding head"},
    "c4": {"B": 512, "D": 2816, "C": 1000000, "s": 64.0, "m": 0.5,
           "what": "BASELINE config 4: two-stream multimodal concat embedding"},
    "c5_100k": {"B": 1024, "D": 512, "C": 100000, "s": 64.0, "m": 0.5, "what": "BASELINE config 5: class sweep, C=100k"},
    "c5_1m": {"B": 1024, "D": 512, "C": 1000000, "s": 64.0, "m": 0.5, "what": "BASELINE config 5: class sweep, C=1M"},
    "c5_10m": {"B": 1024, "D": 512, "C": 10000000, "s": 64.0, "m": 0.5, "what": "BASELINE config 5: class sweep, C=10M"},
}
CONFIGS["ns_bf16x3"] = dict(CONFIGS["ns"], precision="bf16x3",
                            what="north-star shape in the high-precision mode (precision='bf16x3': the north star's "
                                 "'1e-4 under a TF32 mode' gates)")
CONFIGS["c5_10m_pfc10"] = dict(CONFIGS["c5_10m"], sample_rate=0.1, sparse_grad=True,
                               what="BASELINE config 5 at C=10M with PartialFC-style class sampling (sample_rate=0.1: "
                                    "the batch's label classes + uniform negatives, 1M rows per step; sparse dW)")
# "the realistic PartialFC regime (per-rank B = 512 => global B = 512 R), separately, labelled": weak scaling
# in the batch, run at N > 1 only (at N = 1 it is the north-star shape).  B is resolved in resolve_config().
CONFIGS["ns_weak"] = dict(CONFIGS["ns"], B_per_rank=512, scaling="weak",
                          what="PartialFC regime: north-star head with a per-rank batch of 512 rows (global batch 512 x N; "
                               "above 1024 rows the GEMM kernels run once per row chunk)")
EXTRA_CONFIGS = ["c2", "c3", "c4", "c5_100k", "c5_1m", "c5_10m", "c5_10m_pfc10", "ns_bf16x3", "ns_weak"]


def resolve_config(name, world):
    """CONFIGS[name] with the world-size dependent entries filled in."""
    c = dict(CONFIGS[name])
    if "B_per_rank" in c:
        c["B"] = c["B_per_rank"] * world
    return c
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` capture of the north-star
# workload on one B200; only meaningful for the single-GPU north-star shape.
NCU_TRAFFIC = {"fwd": 2.346882e9 + 0.999903e9, "k3": 1.200321e9 + 2.148333e9,
               "source": "profiles/r2b_fwd_bwd_full_raw.csv (ncu --set full, per launch)"}
METRIC = "arcface_head_fwd_bwd_samples_per_sec_1M_classes"
WEIGHT_SEED = 1234
WEIGHT_BLOCK = 65536

# north_star gates
LOSS_RTOL = 1e-3
GRAD_ATOL = 2e-2
LOGIT_ATOL = 2e-2


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
                time.sleep(0.003)
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
            tail = self.power[len(self.power) // 2:]   # NVML's power reading lags the load by ~100 ms: the settled half
            out["power_w_mean_settled"] = sum(tail) / len(tail)
        return out


def workload_config(name, cfg, n_gpus, exchange=None):
    par = "single GPU" if n_gpus == 1 else "class-sharded x%d (PartialFC-style), exchanges over %s" % (
        n_gpus, {"p2p": "peer-mapped memory (csrc/p2p.cu)", "nccl": "NCCL"}.get(exchange, "peer memory or NCCL"))
    return {"workload": "ArcFace head fwd+bwd, %s: B=%d D=%d C=%d s=%g m=%g" % (cfg["what"], cfg["B"], cfg["D"], cfg["C"],
                                                                                cfg["s"], cfg["m"]),
            "name": name, "global_batch": cfg["B"], "embedding_dim": cfg["D"], "classes": cfg["C"], "parallelism": par,
            "l2": "inputs exceed L2: fp32 class weights %.2f GB per step over all ranks, no flush needed"
            % (cfg["C"] * cfg["D"] * 4 / 1e9) if cfg["C"] * cfg["D"] * 4 / max(1, n_gpus) > 126e6 else
            "weights of one rank fit in L2 (%.0f MB): warm-L2 number" % (cfg["C"] * cfg["D"] * 4 / max(1, n_gpus) / 1e6)}


def run_reference(args):
    """The reference's own CPU implementation of the path (port, see oracle/arcface_torch_cpu.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import arcface_torch_cpu as otc

    name = args.config or "ns"
    w = CONFIGS[name]
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
        "config": workload_config(name, w, args.gpus),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def class_weights(torch, C, D, lo, hi, dev):
    """Rows [lo, hi) of the synthetic [C, D] weight matrix (xavier-uniform bound of the FULL matrix, arcface.py:25).
    The matrix is defined block by block -- WEIGHT_BLOCK classes per block, each from its own Philox stream seeded
    by the block index -- so a row depends on (seed, class id) only, not on how many ranks share the classes."""
    bound = math.sqrt(6.0 / (C + D))
    out = torch.empty(hi - lo, D, device=dev)
    g = torch.Generator(device=dev)
    b0 = lo // WEIGHT_BLOCK * WEIGHT_BLOCK
    while b0 < hi:
        b1 = min(C, b0 + WEIGHT_BLOCK)
        g.manual_seed(WEIGHT_SEED + b0 // WEIGHT_BLOCK)
        blk = torch.empty(b1 - b0, D, device=dev).uniform_(-bound, bound, generator=g)
        a, b = max(b0, lo), min(b1, hi)
        out[a - lo:b - lo] = blk[a - b0:b - b0]
        b0 = b1
    return out


class Job:
    """One configuration on this process group: head, deterministic weights, synthetic batch, timing, parity."""

    def __init__(self, torch, dist, mm, name, cfg, world, rank, dev):
        self.torch, self.dist, self.mm = torch, dist, mm
        self.name, self.cfg, self.world, self.rank, self.dev = name, cfg, world, rank, dev
        B, D, C, s, m = cfg["B"], cfg["D"], cfg["C"], cfg["s"], cfg["m"]
        gen = torch.Generator(device="cpu").manual_seed(0)
        self.x_host = torch.randn(B, D, generator=gen).pin_memory()
        self.y_host = torch.randint(0, C, (B,), generator=gen).pin_memory()
        prec = cfg.get("precision", "bf16")
        samp = {"sample_rate": cfg.get("sample_rate", 1.0), "sparse_grad": cfg.get("sparse_grad", False), "sample_seed": 4321}
        self.sampled = samp["sample_rate"] < 1.0
        if world == 1:
            head = mm.ArcMarginProduct(D, 8, s=s, m=m, precision=prec, **samp)
            head.out_feature = C
            self.c_lo, self.c_hi = 0, C
        else:
            # ARCFACE_B200_P2P=0: exchanges through NCCL instead of peer-mapped memory (A/B measurements)
            head = mm.ShardedArcMarginProduct(D, world, s=s, m=m, precision=prec,
                                              use_p2p=os.environ.get("ARCFACE_B200_P2P", "1") != "0", **samp)
            head.out_feature = C
            head.class_lo, head.class_hi = mm.shard_range(C, world, rank)
            self.c_lo, self.c_hi = head.class_lo, head.class_hi
        head.weight = torch.nn.Parameter(class_weights(torch, C, D, self.c_lo, self.c_hi, dev))
        self.head = head.to(dev)
        self.b_loc = B // world
        lo = rank * self.b_loc
        self.x_dev = self.x_host[lo:lo + self.b_loc].to(dev).requires_grad_(True)
        self.y_dev = self.y_host[lo:lo + self.b_loc].to(dev)
        self.pred = None

    # -- the step
    def step(self):
        self.x_dev.grad = None
        self.head.weight.grad = None
        loss, self.pred = self.head.loss(self.x_dev, self.y_dev)
        loss.backward()
        return loss

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, steps, warmup, sample_clocks=True):
        """`warmup` untimed steps, then `steps` steps between barriers; device time by CUDA events, max over ranks."""
        torch = self.torch
        for _ in range(warmup):
            self.step()
        self.barrier()
        sampler = None
        if sample_clocks and self.rank == 0:   # the line is rank 0's; NVML polling from every rank only adds host jitter
            sampler = ClockSampler(self.dev.index or 0)
            sampler.start()
            t_wait = time.perf_counter()   # NVML initialisation takes tens of ms: short timed regions would see no sample
            while not sampler.sm and sampler.error is None and time.perf_counter() - t_wait < 2.0:
                time.sleep(0.002)
            sampler.sm.clear()
            sampler.power.clear()
            sampler.mask = 0
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        # a generational GC pause in ONE rank stalls every rank at the next exchange: collect now, not inside the region
        gc.collect()
        gc.disable()
        try:
            self.barrier()
            evs[0].record()
            loss = None
            for i in range(steps):
                loss = self.step()
                evs[i + 1].record()
            self.barrier()
        finally:
            gc.enable()
        clocks = sampler.result() if sampler is not None else ({"sm_mhz": None, "note": "sampled on rank 0 only"} if sample_clocks else None)
        raw = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        per = sorted(raw)
        out = {"ms_per_step": self.max_over_ranks(evs[0].elapsed_time(evs[steps]) / steps),
               # the first step after the barrier + synchronize starts on an idle device: the host's launch work, hidden
               # behind the previous step's kernels everywhere else, is exposed there
               "ms_first_step": self.max_over_ranks(raw[0]),
               "ms_per_step_median": self.max_over_ranks(per[len(per) // 2]),
               "ms_per_step_best": self.max_over_ranks(per[0]), "ms_per_step_worst": self.max_over_ranks(per[-1]),
               "loss": float(loss.detach())}
        if clocks is not None:
            out["clocks"] = clocks
        return out

    def exchange(self):
        if self.world == 1:
            return None
        from multimodalsimilar_b200 import engine

        return "p2p" if engine._PEERS.get(self.head) else "nccl"

    # -- roofline of the whole step
    def roofline_step(self, ms_step, peaks, clocks):
        cfg = self.cfg
        B, D = cfg["B"], cfg["D"]
        c_loc = self.c_hi - self.c_lo
        if self.sampled:   # the step's work is the sampled sub-matrix (engine.sample_classes: the size is fixed on the host)
            c_loc = min(c_loc, max(int(round(cfg["sample_rate"] * c_loc)), min(B, c_loc)))
        flops_alg = 6.0 * B * D * c_loc
        bytes_alg = 8.0 * c_loc * D + 8.0 * B * D + 24.0 * B
        t_hbm = bytes_alg / (peaks["hbm_gbs"] * 1e9)
        t_b = max(flops_alg / (peaks["bf16_tflops"] * 1e12), t_hbm)
        t_s = max(flops_alg / (peaks["bf16_tflops_sustained"] * 1e12), t_hbm)
        bound = "tensor" if flops_alg / (peaks["bf16_tflops_sustained"] * 1e12) >= t_hbm else "hbm"
        regime = "burst"
        if clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz"):
            regime = "burst" if clocks["sm_mhz"] >= 0.9 * clocks["sm_max_mhz"] else "sustained"
        t = ms_step * 1e-3
        fb, fs = t_b / t, t_s / t
        if bound == "tensor":
            ach, peak, unit = flops_alg / t / 1e12, peaks["bf16_tflops" if regime == "burst" else "bf16_tflops_sustained"], "TFLOP/s"
        else:
            ach, peak, unit = bytes_alg / t / 1e9, peaks["hbm_gbs"], "GB/s"
        return {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "regime": regime,
                "frac": fb if regime == "burst" else fs, "frac_burst": fb, "frac_sustained": fs,
                "t_roof_burst_ms": t_b * 1e3, "t_roof_sustained_ms": t_s * 1e3, "algorithmic_flops": flops_alg,
                "algorithmic_bytes": bytes_alg, "executed_flops": 8.0 * B * D * c_loc, "peak_source": peaks["source"],
                "note": "per rank: the local class shard (1/%d of the classes) for the global batch" % self.world}

    # -- parity against the fp32 restatement of the reference, in this run
    def parity(self):
        torch = self.torch
        from oracle import arcface_torch_chunked as och  # the checker

        cfg = self.cfg
        B, D, C, s, m = cfg["B"], cfg["D"], cfg["C"], cfg["s"], cfg["m"]
        if self.sampled:
            return self.parity_sampled()
        loss = self.step()
        torch.cuda.synchronize()
        got_loss = float(loss.detach())
        got_pred = self.pred.clone()
        got_dx = self.x_dev.grad.clone()
        got_dw = self.head.weight.grad   # alias of the step's buffer: no step runs until the comparison is done
        w_full = self.head.weight.detach() if self.world == 1 else class_weights(torch, C, D, 0, C, self.dev)
        ref = och.head_step_chunked(self.x_host.to(self.dev), w_full, self.y_host.to(self.dev), s, m, False,
                                    chunk=32768 if B > 512 else 65536, dw_range=(self.c_lo, self.c_hi))
        del w_full
        lo = self.rank * self.b_loc
        rows = slice(lo, lo + self.b_loc)
        ref_loss = float(ref["loss"])
        # bf16 noise floor of a logit: rows whose reference winner leads by less cannot be
        # expected to keep their argmax under ANY bf16 evaluation
        x3 = cfg.get("precision", "bf16") == "bf16x3"
        noise = 1e-3 if x3 else LOGIT_ATOL * (s / 30.0) * max(1.0, math.sqrt(512.0 / D))
        sep = ref["top2_gap"][rows] > 2.0 * noise
        mism_sep = int((got_pred[sep] != ref["argmax"][rows][sep]).sum())
        match_all = float((got_pred == ref["argmax"][rows]).float().mean())
        ddx = got_dx - ref["dx"][rows]
        ddw = got_dw - ref["dw"]
        vals = {
            "loss_rel": abs(got_loss - ref_loss) / max(1.0, abs(ref_loss)),
            "dx_max_abs": float(ddx.abs().max()), "dx_rel_fro": float(ddx.norm() / ref["dx"][rows].norm().clamp_min(1e-30)),
            "dw_max_abs": float(ddw.abs().max()), "dw_rel_fro": float(ddw.norm() / ref["dw"].norm().clamp_min(1e-30)),
            "dx_max_rel": float(ddx.abs().max() / ref["dx"][rows].abs().max().clamp_min(1e-30)),
            "dw_max_rel": float(ddw.abs().max() / ref["dw"].abs().max().clamp_min(1e-30)),
            "argmax_mismatch_separated_rows": float(mism_sep), "argmax_mismatch_all_rows_frac": 1.0 - match_all,
        }
        n_sep = int(sep.sum())
        del ref, ddx, ddw
        if self.world > 1:
            t = torch.tensor(list(vals.values()) + [-float(n_sep)], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            for k, v in zip(list(vals.keys()), t.tolist()):
                vals[k] = v
            n_sep = int(-t[-1].item())   # min over ranks
        gates = {"loss_rel<=1e-3": vals["loss_rel"] <= LOSS_RTOL,
                 "argmax_exact_on_separated_rows": vals["argmax_mismatch_separated_rows"] == 0,
                 "dx_max_abs<=2e-2": vals["dx_max_abs"] <= GRAD_ATOL, "dw_max_abs<=2e-2": vals["dw_max_abs"] <= GRAD_ATOL}
        if x3:   # the high-precision mode: gradients within 1e-4 (absolute, and relative to their largest element)
            gates.update({"loss_rel<=2e-5": vals["loss_rel"] <= 2e-5,
                          "dx_max_abs<=1e-4": vals["dx_max_abs"] <= 1e-4, "dw_max_abs<=1e-4": vals["dw_max_abs"] <= 1e-4,
                          "dx_max_rel<=1e-4": vals["dx_max_rel"] <= 1e-4, "dw_max_rel<=1e-4": vals["dw_max_rel"] <= 1e-4})
        out = {"vs": "fp32 restatement of arcface.py:45-63 + CrossEntropyLoss + backward on the same GPU, identical "
                     "inputs (oracle/arcface_torch_chunked.py); max over ranks",
               "precision": cfg.get("precision", "bf16"), "argmax_gap_floor": 2.0 * noise,
               "loss": got_loss, "loss_ref": ref_loss, "separated_rows_per_rank_min": n_sep}
        out.update(vals)
        out["argmax_mismatch_separated_rows"] = int(out["argmax_mismatch_separated_rows"])
        out["gates"] = gates
        out["ok"] = all(gates.values())
        torch.cuda.empty_cache()
        return out

    def parity_sampled(self):
        """Class sampling: the step against the fp32 restatement evaluated on the rows the step sampled (one GPU)."""
        torch = self.torch
        if self.world > 1:
            return None   # every rank draws its own sample; the 2-rank case is covered by tests/test_sampling.py
        from oracle import arcface_torch_chunked as och  # the checker

        cfg = self.cfg
        B, s, m = cfg["B"], cfg["s"], cfg["m"]
        loss = self.step()
        torch.cuda.synchronize()
        index = self.head.last_sample_index()
        y_all = self.y_host.to(self.dev)
        pos = torch.searchsorted(index, y_all)
        labels_in = bool((index[pos] == y_all).all())
        ref = och.head_step_chunked(self.x_host.to(self.dev), self.head.weight.detach()[index], pos, s, m, False,
                                    chunk=32768, dw_range=(0, index.numel()))
        g = self.head.weight.grad
        got_dw = g.coalesce().values() if g.is_sparse else g[index]
        ddx = self.x_dev.grad - ref["dx"]
        ddw = got_dw - ref["dw"]
        noise = LOGIT_ATOL * (s / 30.0) * max(1.0, math.sqrt(512.0 / cfg["D"]))
        sep = ref["top2_gap"] > 2.0 * noise
        got_arg = torch.searchsorted(index, self.pred)
        vals = {"loss_rel": abs(float(loss.detach()) - float(ref["loss"])) / max(1.0, abs(float(ref["loss"]))),
                "dx_max_abs": float(ddx.abs().max()), "dx_rel_fro": float(ddx.norm() / ref["dx"].norm().clamp_min(1e-30)),
                "dw_max_abs": float(ddw.abs().max()), "dw_rel_fro": float(ddw.norm() / ref["dw"].norm().clamp_min(1e-30)),
                "argmax_mismatch_separated_rows": int((got_arg[sep] != ref["argmax"][sep]).sum())}
        gates = {"every_label_in_sample": labels_in, "loss_rel<=1e-3": vals["loss_rel"] <= LOSS_RTOL,
                 "argmax_exact_on_separated_rows": vals["argmax_mismatch_separated_rows"] == 0,
                 "dx_max_abs<=2e-2": vals["dx_max_abs"] <= GRAD_ATOL, "dw_max_abs<=2e-2": vals["dw_max_abs"] <= GRAD_ATOL}
        out = {"vs": "fp32 restatement of arcface.py:45-63 + CrossEntropyLoss + backward on the rows the step sampled "
                     "(oracle/arcface_torch_chunked.py on weight[index])", "sampled_classes": int(index.numel()),
               "loss": float(loss.detach()), "loss_ref": float(ref["loss"]), "separated_rows": int(sep.sum())}
        out.update(vals)
        out["gates"] = gates
        out["ok"] = all(gates.values())
        del ref, ddx, ddw
        torch.cuda.empty_cache()
        return out

    def close(self):
        from multimodalsimilar_b200 import engine

        engine.drop_plan(self.head)
        self.torch.cuda.synchronize()
        del self.head
        self.torch.cuda.empty_cache()


def e2e_measure(job, ops, steps):
    """The step through host buffers: pinned x / labels in, loss / argmax / dx out, copies inside the timed region."""
    torch = job.torch
    cfg = job.cfg
    B, D, C, s, m = cfg["B"], cfg["D"], cfg["C"], cfg["s"], cfg["m"]
    world, rank, dev, b_loc, head = job.world, job.rank, job.dev, job.b_loc, job.head
    h2d = b_loc * D * 4 + b_loc * 8
    d2h = 4 + b_loc * 8 + b_loc * D * 4
    if world == 1:
        loss_h = torch.zeros(1).pin_memory()
        arg_h = torch.zeros(B, dtype=torch.int64).pin_memory()
        dx_h = torch.zeros(B, D).pin_memory()
        head.weight.grad = None
        from multimodalsimilar_b200 import engine

        engine.drop_plan(head)   # the graph's private pool holds a dW of its own
        torch.cuda.empty_cache()
        dw = torch.empty(C, D, device=dev)
        ws = torch.empty(ops.step_workspace_bytes(B, D, C), dtype=torch.uint8, device=dev)
        wdet = head.weight.detach()

        def e2e_step():
            ops.step_host(job.x_host, job.y_host, wdet, s, m, False, 1.0, loss_h, arg_h, dx_h, dw, ws)
    else:
        xl_host = job.x_host[rank * b_loc:(rank + 1) * b_loc].contiguous().pin_memory()
        yl_host = job.y_host[rank * b_loc:(rank + 1) * b_loc].contiguous().pin_memory()
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
    job.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = job.max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)
    job.barrier()
    return {"value": B / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d * world,
            "d2h_bytes_per_step": d2h * world, "ms_per_step": e2e_ms,
            "api": "arcface_b200_step_host (C ABI, pinned host buffers)" if world == 1 else
                   "ShardedArcMarginProduct.loss + backward with pinned host copies"}


def gpu_reference(job):
    """The reference's dense fp32 eager op chain (the port bench.py's reference arm times on the CPU) on this B200."""
    torch = job.torch
    from oracle import arcface_torch_cpu as otc

    cfg = job.cfg
    B, D, C, s, m = cfg["B"], cfg["D"], cfg["C"], cfg["s"], cfg["m"]
    dev = job.dev
    torch.backends.cuda.matmul.allow_tf32 = False  # the reference never enables TF32
    x, y = job.x_host.to(dev), job.y_host.to(dev)
    w = job.head.weight.detach()
    c_used = C
    while True:
        try:
            ws, ys = w[:c_used], y.clamp_max(c_used - 1)
            otc.head_step(x, ws, ys, s, m)
            torch.cuda.synchronize()
            best = float("inf")
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                otc.head_step(x, ws, ys, s, m)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            break
        except torch.OutOfMemoryError:
            torch.cuda.empty_cache()
            c_used //= 2
            if c_used < 1024:
                return {"unavailable": "out of memory"}
    torch.cuda.empty_cache()
    ms = best * (C / c_used)
    return {"value": B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "kind": "port",
            "what": "dense fp32 PyTorch eager op sequence of arcface.py:45-63 + CrossEntropyLoss + backward "
                    "(oracle/arcface_torch_cpu.py) on the same B200, TF32 off, best of 3",
            "classes_timed": c_used, "extrapolated": c_used != C}


def stage_times(torch, ops, head, x_host, y_host, dev, s, m, c_lo, c_total, iters):
    """Average milliseconds of each stage of one rank's step (whole batch, local class shard).  The step's kernel
    sequence is captured as one CUDA graph PER STAGE (the product replays the whole step as one graph; eager launches of
    the big kernels carry ~0.1 ms of host-side launch work that a replay does not have) and the stage graphs are replayed
    in step order `iters` times back to back, a CUDA event between them on the launching stream, one synchronize at the
    end: every kernel runs next to the kernels it runs next to in the real step, in the same clock / power state.
    N = 1 only."""
    x = x_host.to(dev)
    y = y_host.to(dev)
    w = head.weight.detach()
    B, D = x.shape
    names = ["k1_x", "label", "fwd", "k3", "bwd_x"]
    dw = torch.empty_like(w)
    st = {}

    def k1_x():
        st["xhat"], st["inv_nx"], st["xhat_t"] = ops.normalize_cast(x, want_transpose=True)

    def label():
        st["lm"] = ops.label_margin(x, w, st["inv_nx"], None, y, c_lo, c_total, s, m, False)

    def fwd():
        lm = st["lm"]
        st["what"], st["inv_nw"], rmax, rsum, rarg = ops.forward_rows_fused(st["xhat"], w, lm.label_local, s, c_lo)
        st["lse"], _, _, st["omp"], _ = ops.finalize_rows(rmax.view(1, B), rsum.view(1, B), rarg.view(1, B),
                                                          lm.z_label.view(1, B), y)

    def k3():
        lm = st["lm"]
        st["dxhat"], _ = ops.backward(st["xhat"], st["xhat_t"], st["what"], st["inv_nw"], st["lse"], st["omp"], lm.dphi,
                                      lm.label_local, s, 1.0 / B, dw_out=dw)

    def bwd_x():
        ops.normalize_bwd_x(x, st["inv_nx"], st["dxhat"])

    fns = [k1_x, label, fwd, k3, bwd_x]
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            for fn in fns:
                fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    graphs = []
    pool = None
    for fn in fns:   # one pool: a later stage reads what an earlier one left in its static buffers
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):
            fn()
        pool = g.pool()
        graphs.append(g)
    marks = []
    for it in range(iters + 2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        ev[0].record()
        for i, g in enumerate(graphs):
            g.replay()
            ev[i + 1].record()
        if it >= 2:
            marks.append(ev)
    torch.cuda.synchronize()
    acc = {n: 0.0 for n in names}
    for ev in marks:
        for i, n in enumerate(names):
            acc[n] += ev[i].elapsed_time(ev[i + 1])
    del graphs, st, dw
    return {n: acc[n] / len(marks) for n in names}


def dominant_kernel_roofline(job, ops, peaks, stages, regime, traffic_ok, ms_step=None):
    cfg = job.cfg
    B, D = cfg["B"], cfg["D"]
    c_loc = job.c_hi - job.c_lo
    p_tensor = peaks["bf16_tflops" if regime == "burst" else "bf16_tflops_sustained"]
    _, n_chunks = ops.backward_plan(B, D, c_loc)
    k3_launches = ops.backward_launches(B, D, c_loc)
    # share of the (graph-replayed) step; the small stages are host-bound when launched eagerly, so their sum is not used
    comp_ms = ms_step if ms_step else sum(stages.values())
    fwd_bytes = c_loc * D * 6.0 + 4.0 * c_loc   # fp32 W in, bf16 What + 1/||w|| out
    fwd_hbm_s = fwd_bytes / (peaks["hbm_gbs"] * 1e9)
    fwd_tensor_s = 2.0 * B * D * c_loc / (p_tensor * 1e12)
    fwd_cand = (("hbm", fwd_bytes / 1e9, "GB/s", peaks["hbm_gbs"]) if fwd_hbm_s >= fwd_tensor_s
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
    return {"kernel": name, "bound": bnd, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
            "peak_regime": regime, "traffic": NCU_TRAFFIC[top] if traffic_ok else None,
            "traffic_source": NCU_TRAFFIC["source"], "ms_per_launch_group": stages[top],
            "share_of_step": stages[top] / comp_ms, "peak_source": peaks["source"]}


def batch_chunks_of(cfg):
    from multimodalsimilar_b200 import engine

    return engine.batch_chunks(cfg["B"], 1 if cfg.get("precision", "bf16") == "bf16x3" else 0)


def k3_launches(job, ops):
    """Backward kernels per step: per row chunk of the batch, 1 (single-launch backward) or 3 per scratch chunk; for a
    sampled head over the sampled class count."""
    cfg = job.cfg
    c_loc = job.c_hi - job.c_lo
    if job.sampled:
        c_loc = min(c_loc, max(int(round(cfg["sample_rate"] * c_loc)), min(cfg["B"], c_loc)))
    return sum(ops.backward_launches(b1 - b0, cfg["D"], c_loc) for b0, b1 in batch_chunks_of(cfg))


def launches_per_step(job, ops):
    """Our kernels per step (memset / copy / NCCL nodes not counted): pack_xy, then -- replayed from one CUDA graph --
    K1(x), label, K1(w)+K2, combine, finalize, K3, bwd-x, then scale_copy; plus the three peer-memory exchange kernels
    when used."""
    cfg = job.cfg
    n = 1 + 1 + 1 + 1 + 1 + 1 + k3_launches(job, ops) + 1 + 1
    chunks = len(batch_chunks_of(cfg))
    n += 3 * (chunks - 1)   # a batch above one GEMM launch: K2 + combine and the dW accumulate once more per further chunk
    if cfg["D"] > 512:
        n += 1   # K1 of the class weights is its own launch when the in-kernel normaliser does not cover D
    if job.exchange() == "p2p":
        n += 3
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="headline configuration (default: ns, "
                    "followed by short runs of the other BASELINE configs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="host seconds the reference arm may spend")
    ap.add_argument("--classes", type=int, default=None, help="override C (debug only; invalidates the metric)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.classes:
            CONFIGS[args.config or "ns"]["C"] = args.classes
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

    head_name = args.config or "ns"
    cfg = resolve_config(head_name, world)
    if args.classes:
        cfg["C"] = args.classes
    peaks = load_peaks()

    job = Job(torch, dist, mm, head_name, cfg, world, rank, dev)
    B, D, C = cfg["B"], cfg["D"], cfg["C"]
    t = job.timed(args.steps, args.warmup)
    ms_step = t["ms_per_step"]
    value = B / (ms_step * 1e-3)
    clocks = t["clocks"]
    roof_step = job.roofline_step(ms_step, peaks, clocks)
    exchange = job.exchange()

    sustained = None
    if not args.no_sustained:
        n_sus = int(min(4000, max(args.steps, math.ceil(1000.0 / ms_step))))
        ts = job.timed(n_sus, 0)
        sustained = {"steps": n_sus, "ms_per_step": ts["ms_per_step"], "ms_per_step_median": ts["ms_per_step_median"],
                     "value": B / (ts["ms_per_step"] * 1e-3), "clocks": ts["clocks"],
                     "frac_sustained": roof_step["t_roof_sustained_ms"] / ts["ms_per_step"],
                     "note": "the same step looped back to back for >= 1 s right after the timed region"}
        pw = (ts["clocks"] or {}).get("power_w_mean_settled")
        if pw and world == 1 and head_name == "ns" and not args.classes:
            # Board energy of one step, and what the step's executed work costs at the energy per flop / per DRAM byte of
            # the two pure kernels measured at the same power cap (profiles/README.md, tools/power_probe.py: cuBLAS bf16
            # 8192^3 0.69 pJ/flop, K1 streaming 0.13 nJ/byte): under the cap, time = energy / power.
            exe_flops, dram = roof_step["executed_flops"], NCU_TRAFFIC["fwd"] + NCU_TRAFFIC["k3"]
            sustained["energy"] = {
                "power_w_mean": pw, "j_per_step": pw * ts["ms_per_step"] * 1e-3,
                "model_j_per_step": exe_flops * 0.69e-12 + dram * 0.13e-9,
                "model_j_algorithmic_floor": roof_step["algorithmic_flops"] * 0.69e-12 + roof_step["algorithmic_bytes"] * 0.13e-9,
                "note": "model = executed flops x 0.69 pJ + ncu DRAM bytes x 0.13 nJ; floor = the same for 6 B D C flops "
                        "and 8 C D bytes (no recompute, no bf16 copy of the weights)"}

    parity = None if args.no_parity else job.parity()
    gpu_launches = launches_per_step(job, ops) * args.steps

    e2e = e2e_measure(job, ops, max(5, min(args.steps, 50)))

    stages = roofline = gpu_ref = cpu_baseline = None
    if world == 1:
        if not job.sampled:   # (the per-stage breakdown and the dense GPU reference run on every class)
            stages = stage_times(torch, ops, job.head, job.x_host, job.y_host, dev, cfg["s"], cfg["m"], 0, C,
                                 max(5, min(args.steps, 30)))
            roofline = dominant_kernel_roofline(job, ops, peaks, stages, roof_step["regime"],
                                                head_name == "ns" and not args.classes, ms_step)
            gpu_ref = gpu_reference(job)
    job.close()
    del job

    extra = None
    if args.config is None and not args.no_extra and not args.classes:
        extra = {}
        for name in EXTRA_CONFIGS:
            c = resolve_config(name, world)
            if "B_per_rank" in c and world == 1:
                continue   # identical to the north-star shape on one GPU
            try:
                j = Job(torch, dist, mm, name, c, world, rank, dev)
                tt = j.timed(10, 6, sample_clocks=True)
                r = {"B": c["B"], "D": c["D"], "C": c["C"], "s": c["s"], "m": c["m"], "ms_per_step": tt["ms_per_step"],
                     "ms_per_step_median": tt["ms_per_step_median"], "ms_per_step_best": tt["ms_per_step_best"],
                     "value": c["B"] / (tt["ms_per_step"] * 1e-3), "unit": "samples/s", "steps": 10, "warmup": 6,
                     "clocks": tt["clocks"], "roofline_step": j.roofline_step(tt["ms_per_step"], peaks, tt["clocks"]),
                     "exchange": j.exchange(), "k3_launches": k3_launches(j, ops),
                     "sample_rate": c.get("sample_rate", 1.0), "scaling": c.get("scaling", "strong"),
                     "precision": c.get("precision", "bf16"), "workload": c["what"]}
                if not args.no_parity:
                    r["parity"] = j.parity()
                extra[name] = r
                j.close()
                del j
            except Exception as e:  # noqa: BLE001 -- one configuration failing must not lose the headline line
                extra[name] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()

    if world == 1 and not args.no_cpu_baseline:
        from oracle import arcface_torch_cpu as otc

        torch.set_num_threads(os.cpu_count() or 1)
        cpu_baseline = otc.time_head_step(B, D, C, cfg["s"], cfg["m"], budget_s=20.0)
        cpu_baseline = {k: cpu_baseline[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "ms_per_step_median": t["ms_per_step_median"],
            "ms_per_step_best": t["ms_per_step_best"], "ms_per_step_worst": t["ms_per_step_worst"],
            "ms_first_step": t["ms_first_step"],
            "higher_is_better": True, "scaling": cfg.get("scaling", "strong"),
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(head_name, cfg, world, exchange),
            "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "roofline_step": roof_step, "loss": t["loss"],
        }
        if exchange is not None:
            line["exchange"] = exchange
        if roofline is not None:
            line["roofline"] = roofline
        else:
            # N > 1: no per-kernel breakdown is taken (an eager, host-bound one would not describe the replayed graph);
            # the step-level figure of the rank's shard stands in
            line["roofline"] = {k: roof_step[k] for k in ("bound", "achieved", "peak", "unit", "frac")}
            line["roofline"]["traffic"] = None
            line["roofline"]["kernel"] = "whole step of one rank (class shard)"
        if stages is not None:
            line["stages_ms"] = stages
        if sustained is not None:
            line["sustained"] = sustained
        if parity is not None:
            line["parity"] = parity
        if gpu_ref is not None:
            line["gpu_reference"] = gpu_ref
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        if extra is not None:
            line["configs"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        # Captured graphs hold NCCL kernels: leave without running destructors (destroy_process_group() with live
        # graphs was seen to hang after the line was printed).
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
