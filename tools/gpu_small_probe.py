"""CUDA-event timings of the HBM-bound pair kernels at the 1 Mbp shapes (P = 512): achieved GB/s against the
algorithmic bytes of each (read-once + write-once)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from alphagenome_b200 import ops

dev = "cuda"
P, H, D = 512, 32, 128
f32 = dict(device=dev, dtype=torch.float32)
bf = dict(device=dev, dtype=torch.bfloat16)
torch.manual_seed(0)


def timeit(fn, name, nbytes, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name}: {ms * 1e3:.1f} us  {nbytes / ms / 1e6:.0f} GB/s (algorithmic {nbytes / 1e6:.0f} MB)", flush=True)


pair = torch.randn(1, P, P, D, **f32)
s2, t2 = torch.rand(2, D, **f32) + 0.5, torch.randn(2, D, **f32)
W2 = torch.randn(2, 8, D, **f32) * 0.1
bias2 = torch.empty(2, 1, 8, P, P, **f32)
timeit(lambda: ops.pair_bias(pair, s2, t2, W2, bias2), "pair_bias x2 layers", pair.numel() * 4 + bias2.numel() * 4)
bias1 = torch.empty(1, 8, P, P, **f32)
timeit(lambda: ops.pair_bias(pair, s2[0], t2[0], W2[0], bias1), "pair_bias x1 layer", pair.numel() * 4 + bias1.numel() * 4)

Np, N2 = P, 2 * P
qk = torch.randn(1, H, P, Np, **f32)
relq = torch.randn(1, H, P, N2, **f32)
relk = torch.randn(1, H, P, N2, **f32)
Wp, bp = torch.randn(D, H, **f32) * 0.2, torch.randn(D, **f32) * 0.1
outer = torch.randn(1, P, 2 * D, **f32)
out = torch.empty(1, P, P, D, **f32)
pn = torch.empty(1, P, P, D, **bf)
nw, nb = torch.rand(D, **f32) + 0.5, torch.randn(D, **f32)
nbytes = (qk.numel() + relq.numel() + relk.numel() + 2 * pair.numel()) * 4 + pn.numel() * 2
timeit(lambda: ops.pair_combine(qk, relq, relk, Wp, bp, outer, pair, out, P=P, norm_w=nw, norm_b=nb, norm_out=pn), "pair_combine + norm", nbytes)
timeit(lambda: ops.rmsnorm(pair, nw, nb, pn.view(1, P * P, D)), "rmsnorm pair", pair.numel() * 4 + pn.numel() * 2)
x = torch.randn(P * P, D, **f32).to(torch.bfloat16).view(P, P, D)
xt = torch.empty(P, D, P, **bf)
timeit(lambda: ops.transpose_bf16(x, xt), "transpose pair V", 2 * x.numel() * 2)
emb = torch.randn(1, D, **f32)
timeit(lambda: ops.pair_out_embed(pair, nw, nb, emb, out), "pair_out_embed", 2 * pair.numel() * 4)
