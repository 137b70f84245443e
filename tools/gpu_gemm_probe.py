"""Three representative ag_gemm_fwd launches with their real epilogues (for `ncu --set full -k regex:gemm_conv`)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from alphagenome_b200 import ops
from alphagenome_b200.ops import ACT_GELU1702, Act

dev = "cuda"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
C = 768
f32 = dict(device=dev, dtype=torch.float32)
bf = dict(device=dev, dtype=torch.bfloat16)
s, t = torch.rand(C, **f32) + 0.5, torch.randn(C, **f32)
bias = torch.randn(C, **f32)


ONCE = os.environ.get("PROBE_ONCE") == "1"   # one launch per case (for an ncu --set full capture)
CASES = [c for c in os.environ.get("PROBE_CASES", "").split(",") if c]   # subset of case names (default: all)


def timeit(fn, name, flops):
    if CASES and name not in CASES:
        return
    if ONCE:
        fn()
        torch.cuda.synchronize()
        print(f"{name}: launched once", flush=True)
        return
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):            # min over 4 repeats of 8 launches: the first repeats also absorb the clock ramp
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 8)
    print(f"{name}: {best:.3f} ms  {flops / best / 1e9:.0f} TFLOP/s", flush=True)


# (1) embed: K=64 -> out f32 + actA bf16
X = torch.randn(1, M, 64, **f32).to(torch.bfloat16)
W0 = (torch.randn(1, C, 64, **f32) * 0.1).to(torch.bfloat16)
x0 = torch.empty(1, M, C, **f32)
a0 = torch.empty(1, M, C, **bf)
timeit(lambda: ops.gemm(X, W0, pre_shift=bias, out=x0, actA=Act(a0, s, t, ACT_GELU1702)), "embed_k64", 2.0 * M * C * 64)
# (2) conv5 768->768 with res + actA + pool + actP (DNAEmbed.pointwise / DownresBlock.conv_out epilogue)
W5 = (torch.randn(5, C, C, **f32) * 0.02).to(torch.bfloat16)
skipact = torch.empty(1, M, C, **bf)
xp = torch.empty(1, M // 2, C, **f32)
ap = torch.empty(1, M // 2, C, **bf)
timeit(lambda: ops.gemm(a0, W5, pre_shift=bias, res=x0, actA=Act(skipact, s, t, ACT_GELU1702), pool=xp, actP=Act(ap, s, t, ACT_GELU1702)),
       "conv5_pool", 2.0 * M * C * C * 5)
# (3) w1 768->768 with upsampled residual + out + actA (UpresBlock.unet_conv epilogue)
W1 = (torch.randn(1, C, C, **f32) * 0.05).to(torch.bfloat16)
low = torch.randn(1, M // 2, C, **f32)
y = torch.empty(1, M, C, **f32)
ay = torch.empty(1, M, C, **bf)
timeit(lambda: ops.gemm(skipact, W1, pre_shift=bias, res=low, res_shift=1, out=y, actA=Act(ay, s, t, ACT_GELU1702)), "w1_upres", 2.0 * M * C * C)
# (4) conv5 768->768 res + out + actA (UpresBlock.conv_out)
timeit(lambda: ops.gemm(ay, W5, pre_shift=bias, res=y, out=x0, actA=Act(a0, s, t, ACT_GELU1702)), "conv5_out", 2.0 * M * C * C * 5)
# (5) outembed_1bp: 768 -> 1536, residual at row>>7, fp32 output with accurate GELU1702 (OutputEmbedding)
W2 = (torch.randn(1, 2 * C, C, **f32) * 0.05).to(torch.bfloat16)
skp = torch.randn(1, M // 128, 2 * C, **f32)
s2, t2 = torch.rand(2 * C, **f32) + 0.5, torch.randn(1, 2 * C, **f32)
e1 = torch.empty(1, M, 2 * C, **f32)
b2 = torch.randn(2 * C, **f32)
timeit(lambda: ops.gemm(a0, W2, pre_shift=b2, res=skp, res_shift=7, actA=Act(e1, s2, t2, ACT_GELU1702)), "outembed_1bp", 2.0 * M * C * 2 * C)
# (6) pair linear K=128 -> 256, bf16 out (PairwiseRowAttention.to_qk)
pa = torch.randn(1, M, 128, **f32).to(torch.bfloat16)
Wp = (torch.randn(1, 256, 128, **f32) * 0.1).to(torch.bfloat16)
qk = torch.empty(1, M, 256, **bf)
timeit(lambda: ops.gemm(pa, Wp, actA=Act(qk)), "pair_k128", 2.0 * M * 128 * 256)
