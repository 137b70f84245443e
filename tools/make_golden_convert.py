"""Dump the reference's JAX -> PyTorch conversion table (alphagenome_pytorch/convert/mapping_spec.py:130-984) as data:
tests/golden/convert_mapping.json, one row per state_dict entry.  The table is what `alphagenome_b200.convert` must
reproduce; the dump lets the CPU tests check that without the reference tree (which does not travel).

    python tools/make_golden_convert.py
"""
import importlib.util
import json
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/alphagenome_pytorch/convert")


def load_reference_convert():
    """mapping_spec / shape_transforms / convert_checkpoint without importing the package __init__ (which pulls in the
    model and its uninstalled dependencies)."""
    pkg = types.ModuleType("ag_ref_convert")
    pkg.__path__ = [str(REF)]
    sys.modules["ag_ref_convert"] = pkg
    mods = {}
    for name in ("shape_transforms", "mapping_spec", "convert_checkpoint"):
        spec = importlib.util.spec_from_file_location(f"ag_ref_convert.{name}", REF / f"{name}.py")
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"ag_ref_convert.{name}"] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods


def main():
    ms = load_reference_convert()["mapping_spec"]
    rows = []
    for kind, fn in (("param", ms.build_trunk_mapping), ("state", ms.build_state_mapping)):
        for m in fn():
            rows.append(dict(kind=kind, torch_key=m.torch_key, jax_keys=list(m.jax_keys), transform=m.transform.name,
                             extra=dict(m.extra_jax_keys), optional=bool(m.optional), select_index=m.select_index))
    out = ROOT / "tests" / "golden" / "convert_mapping.json"
    out.write_text(json.dumps(rows, indent=0) + "\n")
    print(out, len(rows), "rows")


if __name__ == "__main__":
    main()
