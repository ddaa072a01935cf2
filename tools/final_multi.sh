#!/bin/bash
# Round-end multi-GPU bench lines.  usage: tools/final_multi.sh N [full]
cd "$(dirname "$0")/.."
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 5 "${@:3}" 2> gpurun_out/bench_n${N}_$2.err | grep '^{' > gpurun_out/bench_n${N}_$2.json; echo "n=$N $2 rc=$?"; }
if [ "$2" = "full" ]; then run 29531 full; else run 29531 ns --no-extra; run 29532 weak --config ns_weak --no-sustained; fi
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_n${N}_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, {k: (round(d[k], 4) if isinstance(d[k], float) else d[k]) for k in ("ms_per_step", "ms_per_step_median", "ms_per_step_best", "value", "exchange", "scaling") if k in d},
          "B", d["config"]["global_batch"], "frac_burst", round(d["roofline_step"]["frac_burst"], 3), "parity", (d.get("parity") or {}).get("ok"))
    for n, c in (d.get("configs") or {}).items():
        print("   ", n, c.get("error") or (round(c["ms_per_step"], 3), round(c["roofline_step"]["frac_burst"], 3), (c.get("parity") or {}).get("ok")))
PY
