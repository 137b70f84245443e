"""Row attention (ag_attention_fwd mode 1 with the fused RMSNorm, as the pair block calls it) alone at P x P cells:
CUDA-event time per launch.  AGB_ROWATTN=0 selects the one-tile-per-CTA two-pass kernel.

    python tools/gpu_rowattn_probe.py [P] [launches]
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from alphagenome_b200 import ops

P = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
D = 128
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
qk = (torch.randn(1, P, P, 2 * D, device=dev, generator=g) * 1.5).bfloat16()
vt = torch.randn(P, D, P, device=dev, generator=g).bfloat16()
pair = torch.randn(1, P, P, D, device=dev, generator=g)
pn = torch.empty(1, P, P, D, device=dev, dtype=torch.bfloat16)
bv = torch.randn(D, device=dev, generator=g) * 0.1
nw = torch.ones(D, device=dev)
nb = torch.zeros(D, device=dev)
big = torch.empty(256 << 20, device=dev, dtype=torch.uint8)   # L2 flush between launches


def launch():
    ops.attention(qk[..., :D], qk[..., D:], vt.view(1, P, D, P), pair, mode=1, scale=D ** -0.5, out_bias=bv, res=pair,
                  norm_w=nw, norm_b=nb, norm_out=pn)


for _ in range(3):
    launch()
torch.cuda.synchronize()
ts = []
for _ in range(n):
    big.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
print(f"ROWATTN P={P} median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f} us  ({n} launches, L2 flushed)")
