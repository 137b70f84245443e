"""Generate tests/golden/*.npz by running the UNMODIFIED reference (lucidrains/alphagenome, /root/reference)
on CPU fp32 in the build container (it cannot travel to the GPU box, the fixtures can).

    python tools/make_golden.py

The reference is imported through oracle/shims (einx / torch_einops_utils / PoPE_pytorch are not installed and
there is no network; the shims only cover named-axis broadcasting, all arithmetic is the reference's own torch
code).  Weights come from alphagenome_b200.init_weights (numpy PCG64 keyed by parameter name), loaded into the
reference model with load_state_dict(strict=True); norms are frozen with set_update_running_var(model, False)
as the reference's own tests do (tests/test_selective_inference.py:17).

Each fixture stores the inputs and a deterministic row-subsample of the six trunk outputs (see
`sample_rows`), so a full-size model is pinned with a few MB of data.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "shims"))
sys.path.append("/root/reference")  # after the repo so `tests` resolves to ours

from alphagenome_b200.init_weights import apply_synth_weights  # noqa: E402
from tests.golden_util import CASES, make_inputs, sample_rows  # noqa: E402


def main():
    from alphagenome_pytorch.alphagenome import AlphaGenome, set_update_running_var

    torch.set_num_threads(8)
    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(exist_ok=True)
    model = AlphaGenome()
    model.eval()
    set_update_running_var(model, False)
    keys = sorted(model.state_dict().keys())
    (out_dir / "state_dict_keys.txt").write_text(
        "\n".join(f"{k} {tuple(model.state_dict()[k].shape)}" for k in keys) + "\n")
    for case in CASES:
        apply_synth_weights(model, seed=case["weight_seed"])
        seq, org = make_inputs(case)
        with torch.no_grad():
            emb, tu = model.get_embeds(seq, org, return_unet_transformer_outputs=True)
        full = dict(embeds_1bp=emb.embeds_1bp, embeds_128bp=emb.embeds_128bp, embeds_pair=emb.embeds_pair,
                    unet_output=tu.unet_output, single_repr=tu.single_repr, pairwise_repr=tu.pairwise_repr)
        rec = dict(seq=seq.numpy().astype(np.int8), organism_index=org.numpy().astype(np.int64))
        for k, v in full.items():
            idx = sample_rows(k, v.shape)
            rec[k] = (v if idx is None else v[:, idx]).numpy().astype(np.float32)
            rec[k + "__absmax"] = np.float32(v.abs().max().item())
        np.savez(out_dir / f"{case['name']}.npz", **rec)
        print(case["name"], {k: rec[k].shape for k in full})


if __name__ == "__main__":
    main()
