"""Golden fixture for the DNAModel callers: the UNMODIFIED reference wrapper
(alphagenome_pytorch/dna_model.py: create, predict with the two-pass splice flow, ism) over the reference model with the
JAX-aligned heads of both organisms, CPU fp32, synthetic weights, frozen norms.

    python tools/make_golden_dna_model.py        # -> tests/golden/dna_model_L2048.npz
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
from tests.golden_util import DNA_MODEL_CASE, make_dna_model_inputs  # noqa: E402


def main():
    from alphagenome_pytorch import dna_model as ref_dm
    from alphagenome_pytorch.alphagenome import set_update_running_var

    torch.set_num_threads(8)
    c = DNA_MODEL_CASE
    dm = ref_dm.create(add_reference_heads=True, settings=ref_dm.ModelSettings(num_splice_sites=c["num_splice_sites"], splice_site_threshold=c["threshold"]))
    dm.eval()
    set_update_running_var(dm.model, False)
    apply_synth_weights(dm.model, seed=c["weight_seed"])
    seq, ism_positions = make_dna_model_inputs(c)
    rec = dict(seq=np.frombuffer(seq.encode(), dtype=np.uint8))
    pred = dm.predict(seq, organism="mouse")
    junc = pred["splice_sites_junction"]
    rec["splice_site_positions"] = junc["splice_site_positions"].numpy()
    rec["junction_predictions"] = junc["predictions"].numpy().astype(np.float32)
    rec["junction_mask"] = junc["splice_junction_mask"].numpy()
    rec["dnase_1bp"] = pred["dnase"]["scaled_predictions_1bp"][:, ::16].numpy().astype(np.float32)
    rec["chip_tf_128bp"] = pred["chip_tf"]["scaled_predictions_128bp"].numpy().astype(np.float32)
    rec["splice_logits"] = pred["splice_sites_classification"].numpy().astype(np.float32)
    for key in ("dnase", "chip_tf"):
        out = dm.ism(seq, output_key=key, organism="mouse", positions=ism_positions, batch_size=c["ism_batch"], return_predictions=True)
        rec[f"ism_{key}_positions"] = out["positions"]
        rec[f"ism_{key}_alt_bases"] = np.array([ord(b) for b in out["alt_bases"]], dtype=np.uint8)
        rec[f"ism_{key}_delta"] = out["delta"].astype(np.float32)
        rec[f"ism_{key}_predictions"] = out["predictions"].astype(np.float32)
    path = ROOT / "tests" / "golden" / f"{c['name']}.npz"
    np.savez_compressed(path, **rec)
    print(path, {k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    main()
