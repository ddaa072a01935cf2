cd $GRAFT_REPO_ROOT
export ARCFACE_B200_DIAG=1
for maxd in 512 1024; do
for dbg in 0 1 2; do
echo "== NORM_MAXD=$maxd FWD_DEBUG=$dbg"
ARCFACE_B200_FWD_NORM_MAXD=$maxd ARCFACE_B200_FWD_DEBUG=$dbg ARCFACE_B200_FWD_PROF=1 python - <<'PY' 2>&1 | grep -v Warn
import math, os, sys, torch
sys.path.insert(0, os.environ["GRAFT_REPO_ROOT"])
from multimodalsimilar_b200 import ops
dev = torch.device("cuda:0")
B, D, C = 512, 1024, 125000
g = torch.Generator(device=dev).manual_seed(0)
w = torch.empty(C, D, device=dev).uniform_(-0.01, 0.01, generator=g)
x = torch.randn(B, D, device=dev, generator=g)
y = torch.randint(0, C, (B,), device=dev, generator=g)
xhat, inv_nx, _ = ops.normalize_cast(x)
lm = ops.label_margin(x, w, inv_nx, None, y, 0, C, 64.0, 0.5, False)
os.environ.pop("ARCFACE_B200_FWD_PROF", None)
for _ in range(3): ops.forward_rows_fused(xhat, w, lm.label_local, 64.0, 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.forward_rows_fused(xhat, w, lm.label_local, 64.0, 0)
e1.record(); torch.cuda.synchronize()
print("ms", e0.elapsed_time(e1) / 10)
os.environ["ARCFACE_B200_FWD_PROF"] = "1"
ops.forward_rows_fused(xhat, w, lm.label_local, 64.0, 0)
torch.cuda.synchronize()
PY
done; done
