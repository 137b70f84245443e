"""Generate tests/golden/heads_*.npz: the UNMODIFIED reference with the publication heads of both organisms
(`add_heads(**publication_heads_config[...])`, AG:57-74, 1653-1682), CPU fp32, synthetic weights — see
tools/make_golden.py for how the reference is imported and why fixtures (not the reference) travel to the GPU box.

    python tools/make_golden_heads.py

Stores the inputs (tokens, organism_index, splice donor / acceptor indices) and, per organism and head, a
deterministic row-subsample of the output (dim 1 for the 1 bp heads; everything for the small ones).
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
from tests.golden_util import HEAD_CASES, make_head_inputs, sample_rows  # noqa: E402


def main():
    from alphagenome_pytorch.alphagenome import AlphaGenome, publication_heads_config, set_update_running_var

    torch.set_num_threads(8)
    out_dir = ROOT / "tests" / "golden"
    model = AlphaGenome()
    for org in ("human", "mouse"):
        model.add_heads(**publication_heads_config[org])
    model.eval()
    set_update_running_var(model, False)
    sd = model.state_dict()
    (out_dir / "state_dict_keys_heads.txt").write_text(
        "\n".join(f"{k} {tuple(sd[k].shape)}" for k in sorted(sd) if k.startswith("heads.")) + "\n")
    for case in HEAD_CASES:
        apply_synth_weights(model, seed=case["weight_seed"])
        seq, org, donor, acceptor = make_head_inputs(case)
        with torch.no_grad():
            out = model(seq, org, splice_donor_idx=donor, splice_acceptor_idx=acceptor)
        rec = dict(seq=seq.numpy().astype(np.int8), organism_index=org.numpy(), splice_donor_idx=donor.numpy(),
                   splice_acceptor_idx=acceptor.numpy())
        for organism, heads in out.items():
            for name, v in heads.items():
                key = f"{organism}/{name}"
                idx = sample_rows("head", v.shape) if name in ("1bp_tracks", "splice_logits", "splice_usage") else None
                rec[key] = (v if idx is None else v[:, idx]).numpy().astype(np.float32)
                rec[key + "__absmax"] = np.float32(v.abs().max().item())
                print(key, tuple(v.shape), "->", rec[key].shape)
        np.savez_compressed(out_dir / f"{case['name']}.npz", **rec)


if __name__ == "__main__":
    main()
