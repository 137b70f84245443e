"""Per-launch table of ag_gemm_fwd inside one trunk forward (CUDA events around each launch)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from alphagenome_b200 import ops
from alphagenome_b200.init_weights import apply_synth_weights
from alphagenome_b200.model import AlphaGenome

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
m = AlphaGenome().cuda().eval()
for _mod in m.modules():
    if type(_mod).__name__ == 'BatchRMSNorm':
        _mod.update_running_var = False
apply_synth_weights(m, 0)
seq = torch.from_numpy(np.random.default_rng(42).integers(0, 4, size=(1, L)).astype(np.int64)).cuda()
org = torch.zeros(1, dtype=torch.long).cuda()
for _ in range(2):
    m(seq, org)
torch.cuda.synchronize()
shapes = []
orig = ops.gemm


def wrapped(A, W, **kw):
    shapes.append((tuple(A.shape), tuple(W.shape), kw.get("w_batched", False), [k for k in ("res", "out", "actA", "actB", "pool", "actP") if kw.get(k) is not None]))
    return orig(A, W, **kw)


import alphagenome_b200.trunk as T

T.ops.gemm = wrapped
ops.PROFILE_GEMM = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
m(seq, org)
e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1)
rows = []
for (a, b, fl, _nb), (sa, sw, wb, outs) in zip(ops.PROFILE_GEMM, shapes):
    ms = a.elapsed_time(b)
    rows.append(dict(A=sa, W=sw, batched=wb, outs=outs, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1), gflop=round(fl / 1e9, 2)))
gsum = sum(r["ms"] for r in rows)
print(f"step {total:.2f} ms, gemm launches {len(rows)} sum {gsum:.2f} ms")
for r in rows:
    print(json.dumps(r))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"gemm_table_{L}.json").write_text(json.dumps(dict(step_ms=total, rows=rows)))
