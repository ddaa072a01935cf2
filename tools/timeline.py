"""Kernel timeline of one rank's step (CUDA-graph replay) from torch.profiler / CUPTI: per kernel, average
device duration and the idle gap before it.  Single process: C is one rank's class shard.

    python tools/timeline.py [--C 125000] [--B 512] [--D 512] [--steps 5]
    torchrun --nproc-per-node 2 tools/timeline.py --world     (adds the NCCL kernels; rank 0 prints)
"""
import argparse, collections, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import multimodalsimilar_b200 as mm

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=512)
ap.add_argument("--D", type=int, default=512)
ap.add_argument("--C", type=int, default=125000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--world", action="store_true")
ap.add_argument("--eager", action="store_true")
args = ap.parse_args()
rank, world = 0, 1
if args.world:
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
dev = torch.device("cuda", torch.cuda.current_device())
B, D, C = args.B, args.D, args.C
g = torch.Generator(device=dev).manual_seed(rank)
bound = math.sqrt(6.0 / (C * world + D))
if world == 1:
    h = mm.ArcMarginProduct(D, 8, s=64.0, m=0.5, use_cuda_graph=not args.eager)
    h.out_feature = C
else:
    h = mm.ShardedArcMarginProduct(D, world, s=64.0, m=0.5, use_cuda_graph=not args.eager)
    h.out_feature = C * world
    h.class_lo, h.class_hi = rank * C, (rank + 1) * C
h.weight = torch.nn.Parameter(torch.empty(C, D, device=dev).uniform_(-bound, bound, generator=g))
b_loc = B // world
x = torch.randn(b_loc, D, device=dev, generator=g).requires_grad_(True)
y = torch.randint(0, C * world, (b_loc,), device=dev, generator=g)


def step():
    x.grad = None
    h.weight.grad = None
    loss, _ = h.loss(x, y)
    loss.backward()


for _ in range(6):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    t_first, t_last = ev[0].time_range.start, ev[-1].time_range.end
    dur, gap, cnt = collections.OrderedDict(), collections.Counter(), collections.Counter()
    prev_end = None
    for e in ev:
        name = e.name[:70]
        dur[name] = dur.get(name, 0.0) + (e.time_range.end - e.time_range.start)
        cnt[name] += 1
        if prev_end is not None:
            gap[name] += max(0.0, e.time_range.start - prev_end)
        prev_end = max(prev_end or 0, e.time_range.end)
    n = args.steps
    print("span per step %.1f us   (B=%d D=%d C_local=%d world=%d %s)" % ((t_last - t_first) / n, B, D, C, world,
                                                                         "eager" if args.eager else "graph"))
    tot_d = tot_g = 0.0
    for name in dur:
        print("%-72s x%-3d dur %8.1f us  gap-before %7.1f us" % (name, cnt[name] // n, dur[name] / n, gap[name] / n))
        tot_d += dur[name] / n
        tot_g += gap[name] / n
    print("sum of kernel time %.1f us, sum of gaps %.1f us" % (tot_d, tot_g))
if args.world:
    from multimodalsimilar_b200 import engine
    engine.drop_plan(h)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)
