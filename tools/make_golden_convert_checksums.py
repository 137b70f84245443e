"""tests/golden/convert_checksums.json: per-entry (sum, abs-sum, shape) of the state_dict the REFERENCE converter
(alphagenome_pytorch/convert/convert_checkpoint.py:108-233) produces from the synthetic JAX-layout checkpoint of
tests/convert_util.py — the travel-safe pin for alphagenome_b200.convert.

    python tools/make_golden_convert_checksums.py
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

from make_golden_convert import load_reference_convert  # noqa: E402

from alphagenome_b200.model import AlphaGenome  # noqa: E402
from tests.convert_util import synth_checkpoint  # noqa: E402


def main():
    m = AlphaGenome()
    for organism in ("human", "mouse"):
        m.add_reference_heads(organism)
    params, state = synth_checkpoint(m, seed=0)
    sd = load_reference_convert()["convert_checkpoint"].convert_checkpoint(params, state, verbose=False)
    out = {k: [float(v.double().sum()), float(v.double().abs().sum()), list(v.shape)] for k, v in sorted(sd.items())}
    path = ROOT / "tests" / "golden" / "convert_checksums.json"
    path.write_text(json.dumps(out, indent=0) + "\n")
    print(path, len(out), "entries")


if __name__ == "__main__":
    main()
