"""Golden fixture for BatchRMSNorm with UPDATING statistics (the reference's default: update_running_var=True, eval
included, AG:324-361): run the UNMODIFIED reference twice on CPU fp32 WITHOUT freezing the norms and store the second
forward's embeddings plus all 81 running_var buffers after each forward.

    python tools/make_golden_update.py          # -> tests/golden/update_B2_L4096.npz
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "shims"))
sys.path.append("/root/reference")

from alphagenome_b200.init_weights import apply_synth_weights  # noqa: E402
from tests.golden_util import UPDATE_CASE, make_update_inputs, sample_rows  # noqa: E402


def main():
    from alphagenome_pytorch.alphagenome import AlphaGenome

    torch.set_num_threads(8)
    model = AlphaGenome().eval()             # update_running_var stays at its default (True)
    apply_synth_weights(model, seed=UPDATE_CASE["weight_seed"])
    rec = {}
    for step, (seq, org) in enumerate(make_update_inputs(UPDATE_CASE)):
        with torch.no_grad():
            emb = model(seq, org)
        for k, v in model.state_dict().items():
            if k.endswith("running_var"):
                rec[f"rv{step}/{k}"] = v.numpy().astype(np.float32).copy()
    for k, v in zip(("embeds_1bp", "embeds_128bp", "embeds_pair"), emb):
        idx = sample_rows(k, v.shape)
        rec[k] = (v if idx is None else v[:, idx]).numpy().astype(np.float32)
        rec[k + "__absmax"] = np.float32(v.abs().max().item())
    out = ROOT / "tests" / "golden" / f"{UPDATE_CASE['name']}.npz"
    np.savez_compressed(out, **rec)
    print(out, len([k for k in rec if k.startswith("rv1/")]), "running_var buffers")


if __name__ == "__main__":
    main()
