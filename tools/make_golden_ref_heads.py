"""Generate tests/golden/refheads_*.npz: the UNMODIFIED reference with the JAX-aligned heads of both organisms
(`add_reference_heads`, AG:1684-1731), CPU fp32, synthetic weights (see tools/make_golden.py).

    python tools/make_golden_ref_heads.py
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
from tests.golden_util import REF_HEAD_CASES, make_ref_head_inputs, sample_rows  # noqa: E402


def flatten(organism, name, v, rec):
    """head outputs are tensors, dicts of tensors (GenomeTracksHead, SpliceSitesJunctionHead) or None"""
    if v is None:
        return
    if isinstance(v, dict):
        for k, t in v.items():
            flatten(organism, f"{name}.{k}", t, rec)
        return
    key = f"{organism}/{name}"
    t = v.float() if v.dtype == torch.bool else v
    idx = sample_rows("head", t.shape) if t.dim() == 3 and t.shape[1] > 96 else None
    rec[key] = (t if idx is None else t[:, idx]).numpy().astype(np.float32)
    rec[key + "__absmax"] = np.float32(max(t.abs().max().item(), 1e-30))
    print(key, tuple(v.shape), "->", rec[key].shape)


def main():
    from alphagenome_pytorch.alphagenome import AlphaGenome, set_update_running_var

    torch.set_num_threads(8)
    out_dir = ROOT / "tests" / "golden"
    model = AlphaGenome()
    for org in ("human", "mouse"):
        model.add_reference_heads(org)
    model.eval()
    set_update_running_var(model, False)
    sd = model.state_dict()
    (out_dir / "state_dict_keys_ref_heads.txt").write_text(
        "\n".join(f"{k} {tuple(sd[k].shape)}" for k in sorted(sd) if k.startswith("heads.")) + "\n")
    for case in REF_HEAD_CASES:
        apply_synth_weights(model, seed=case["weight_seed"])
        seq, org, pos = make_ref_head_inputs(case)
        with torch.no_grad():
            out = model(seq, org, splice_site_positions=pos)
        rec = dict(seq=seq.numpy().astype(np.int8), organism_index=org.numpy(), splice_site_positions=pos.numpy())
        for organism, heads in out.items():
            for name, v in heads.items():
                flatten(organism, name, v, rec)
        np.savez_compressed(out_dir / f"{case['name']}.npz", **rec)


if __name__ == "__main__":
    main()
