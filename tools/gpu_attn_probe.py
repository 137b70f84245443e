"""Two representative ag_attention_fwd launches (for `ncu --set full -k regex:attention_kernel`) with CUDA-event timing:
mode 0 = tower attention at 1 Mbp (n = 8192 tokens, 8 heads, MQA, soft-cap + pair bias), mode 1 = pair row attention
(P = 512 rows x 512 keys).  Usage: python tools/gpu_attn_probe.py [n_tokens]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from alphagenome_b200 import ops

dev = "cuda"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
H, d, dv = 8, 128, 192
bf = dict(device=dev, dtype=torch.bfloat16)
f32 = dict(device=dev, dtype=torch.float32)
torch.manual_seed(0)


def timeit(fn, name, flops, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name}: {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s", flush=True)


# mode 0
Q = torch.randn(1, n, H * d, **f32).to(torch.bfloat16)
K = torch.randn(1, 1, n, d, **f32).to(torch.bfloat16)
Vt = torch.randn(1, 1, dv, n, **f32).to(torch.bfloat16)
P = n // 16
bias = torch.randn(1, H, P, P, **f32)
O = torch.empty(1, n, H * dv, **bf)
timeit(lambda: ops.attention(Q.view(1, n, H, d).permute(0, 2, 1, 3), K, Vt, O.view(1, n, H, dv).permute(0, 2, 1, 3), mode=0,
                             scale=d ** -0.5, bias=bias, softcap=5.0), f"attn_mode0 n={n}", 1.0 * H * n * n * (2 * d + 2 * dv))
# a sequence-parallel rank's shape at 8 GPUs: 1/8 of the queries against all keys (64 tiles: the key split matters)
nq = n // 8
Qs, Os = Q[:, :nq].contiguous(), torch.empty(1, nq, H * dv, **bf)
bias_s = bias[:, :, : nq // 16].contiguous()
timeit(lambda: ops.attention(Qs.view(1, nq, H, d).permute(0, 2, 1, 3), K, Vt, Os.view(1, nq, H, dv).permute(0, 2, 1, 3), mode=0,
                             scale=d ** -0.5, bias=bias_s, softcap=5.0), f"attn_mode0 n_q={nq} n_k={n}", 1.0 * H * nq * n * (2 * d + 2 * dv))
print("max |shard - full| =", (Os.float() - O[:, :nq].float()).abs().max().item(), flush=True)
# mode 1: rows = P, keys = P, d = dv = 128
dp = 128
qk = torch.randn(1, P, P, 2 * dp, **f32).to(torch.bfloat16)
vt = torch.randn(1, P, dp, P, **f32).to(torch.bfloat16)
pair = torch.randn(1, P, P, dp, **f32)
bv = torch.randn(dp, **f32)
timeit(lambda: ops.attention(qk[..., :dp], qk[..., dp:], vt, pair, mode=1, scale=dp ** -0.5, out_bias=bv, res=pair),
       f"attn_mode1 P={P}", 1.0 * P * P * P * (2 * dp + 2 * dp))
