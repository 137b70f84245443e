"""Three eager trunk forwards at 1 Mbp (two warm-ups + one), for `ncu -k regex:<kernel> --launch-skip <2*per_forward + i>
--launch-count 1` captures of a particular launch.  Launch order of one forward (B=1, no heads), by kernel family:

  gemm_conv_kernel (119): 0 embed K=64 | 1 pointwise w5 | 2..13 downs | 14..94 tower (13 per pair layer, 4 otherwise) |
                          95..115 ups (conv, unet_conv w1, conv_out per block) | 116 outembed_128bp | 117 skip_proj | 118 outembed_1bp
  attention_kernel  (23): per layer [pair row attention (mode 1) if the layer has a pair block], full waves, split tail
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from alphagenome_b200.init_weights import apply_synth_weights
from alphagenome_b200.model import AlphaGenome, set_update_running_var

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1_048_576
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m = AlphaGenome().cuda().eval()
set_update_running_var(m, False)
apply_synth_weights(m, 0)
seq = torch.from_numpy(np.random.default_rng(42).integers(0, 4, size=(1, L)).astype(np.int64)).cuda()
org = torch.zeros(1, dtype=torch.long).cuda()
for _ in range(n):
    m(seq, org)
torch.cuda.synchronize()
print("done")
