"""GPU bring-up diagnostic for the whole trunk: per-stage error vs the CPU oracle on one golden case."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from alphagenome_b200 import _lib
from alphagenome_b200.init_weights import apply_synth_weights
from alphagenome_b200.model import AlphaGenome
from alphagenome_b200.trunk import engine_for
from oracle import trunk_oracle as O
from tests.golden_util import CASES, make_inputs

case = CASES[int(sys.argv[1]) if len(sys.argv) > 1 else 1]
seq, org = make_inputs(case)
m = AlphaGenome().cuda().eval()
for _mod in m.modules():
    if type(_mod).__name__ == 'BatchRMSNorm':
        _mod.update_running_var = False
apply_synth_weights(m, seed=case["weight_seed"])
sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
ref = {}
t0 = time.time()
out_ref = O.alphagenome_embeds(sd, seq, org, collect=lambda k, v: ref.__setitem__(k, None if v is None else v.clone()))
print("oracle secs", round(time.time() - t0, 2), flush=True)
ref.update(out_ref)
mine = {}
eng = engine_for(m.transformer_unet, m)
eng.collect = lambda k, v: (torch.cuda.synchronize(), mine.__setitem__(k, v.detach().float().cpu().clone()), print("  stage", k, "ok", flush=True))
try:
    emb, tu = m.get_embeds(seq.cuda(), org.cuda(), return_unet_transformer_outputs=True)
    torch.cuda.synchronize()
except Exception as e:  # noqa: BLE001
    print("FAILED", repr(e)[:600], "watchdog", [hex(x) for x in _lib.watchdog_status(0)], flush=True)
    raise
mine.update(dict(embeds_1bp=emb.embeds_1bp, embeds_128bp=emb.embeds_128bp, embeds_pair=emb.embeds_pair,
                 unet_output=tu.unet_output, single_repr=tu.single_repr, pairwise_repr=tu.pairwise_repr))
for k, v in mine.items():
    if k in ref and ref[k] is not None:
        v = v.float().cpu()
        e = ((v - ref[k]).abs().max() / ref[k].abs().max()).item()
        print(f"{k:28s} rel_err {e:.3e}   absmax {ref[k].abs().max().item():.3e}", flush=True)
print("launches", _lib.launch_count())
