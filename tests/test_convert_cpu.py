"""CPU: the checkpoint conversion contract (reference alphagenome_pytorch/convert/).

  * the generated rule table equals the reference's mapping table entry by entry (golden dump of mapping_spec.py);
  * the layout rules on the reference's own known-answer vectors (tests/test_conversion_shape_transforms.py);
  * a synthetic JAX-layout checkpoint converts into a state_dict that load_state_dict(strict=True) accepts on the host
    mirror with the checkpoint's heads — and, when the reference tree is present (build container), equals the
    reference converter's output tensor by tensor; everywhere, its per-entry checksums equal the committed ones that
    the reference converter produced (tests/golden/convert_checksums.json, tools/make_golden_convert_checksums.py).
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from alphagenome_b200 import convert as CV
from alphagenome_b200.model import AlphaGenome
from tests.convert_util import synth_checkpoint

GOLD = Path(__file__).resolve().parent / "golden"
OPS = {"IDENTITY": "copy", "TRANSPOSE_LINEAR": "linear", "TRANSPOSE_CONV1D": "conv", "TRANSPOSE_LINEAR_TO_CONV1D": "pointwise",
       "CONCAT_QKV": "concat", "CONCAT_QK": "concat", "STACK_QK_BIAS": "stack", "BAKE_STANDARDIZED_CONV": "standardized",
       "SQUEEZE_TO_1D": "squeeze", "UNSQUEEZE_SCALAR": "scalar"}


@pytest.fixture(scope="module")
def model():
    m = AlphaGenome()
    for organism in ("human", "mouse"):
        m.add_reference_heads(organism)
    return m


def test_rule_table_equals_reference_mapping():
    rows = json.loads((GOLD / "convert_mapping.json").read_text())
    ref = {r["torch_key"]: (tuple(r["jax_keys"]), OPS[r["transform"]], tuple(sorted(r["extra"].items())), r["select_index"], r["kind"] == "state")
           for r in rows}
    mine = {r.torch_key: (r.src, r.op, tuple(sorted(r.extra)), r.index, r.state) for r in CV.build_rules()}
    assert len(rows) == 626 and len(ref) == 626
    assert set(mine) == set(ref)
    assert [k for k in ref if ref[k] != mine[k]] == []


def test_layout_rules_known_answers():
    """The vectors of the reference's tests/test_conversion_shape_transforms.py."""
    lin = CV.apply_rule(CV.Rule("x.weight", ("w",), "linear"), [torch.tensor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])], {})
    assert torch.equal(lin, torch.tensor([[1.0, 4.0], [2.0, 5.0], [3.0, 6.0]]))
    w = torch.arange(5 * 2 * 3, dtype=torch.float32).reshape(5, 2, 3)
    wt = CV.apply_rule(CV.Rule("x.weight", ("w",), "conv"), [w], {})
    assert wt.shape == (3, 2, 5) and wt[0, 0, 0] == w[0, 0, 0] and wt[2, 1, 4] == w[4, 1, 2]
    q, k, v = (torch.tensor([[a], [b]]) for a, b in ((1.0, 2.0), (3.0, 4.0), (5.0, 6.0)))
    cat = CV.apply_rule(CV.Rule("x.weight", ("q", "k", "v"), "concat"), [q, k, v], {})
    assert torch.equal(cat, torch.tensor([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]]))
    # standardized conv: same seed and shapes as the reference test (:49-71), expected value restated from the
    # published StandardizedConv1D formula
    torch.manual_seed(0)
    width, cin, cout = 5, 4, 7
    w = torch.randn(width, cin, cout)
    scale = torch.randn(1, 1, cout)
    bias = torch.randn(cout)
    bw, bb = CV.apply_rule(CV.Rule("x.weight", ("w",), "standardized", extra=(("scale", "s"), ("bias", "b"))), [w], dict(scale=scale, bias=bias))
    c = w - w.mean(dim=(0, 1), keepdim=True)
    var = c.var(dim=(0, 1), keepdim=True, unbiased=False)
    want = (c * scale * torch.rsqrt(torch.maximum(width * cin * var, torch.tensor(1e-4)))).permute(2, 1, 0).contiguous()
    assert bw.shape == (cout, cin, width) and torch.equal(bw, want) and torch.equal(bb, bias)
    assert torch.equal(CV.apply_rule(CV.Rule("x", ("a",), "scalar"), [torch.tensor(2.0)], {}), torch.tensor([2.0]))
    assert CV.apply_rule(CV.Rule("x", ("a",), "squeeze"), [torch.ones(1, 1, 9)], {}).shape == (9,)
    with pytest.raises(ValueError):
        CV.apply_rule(CV.Rule("x.weight", ("w",), "linear"), [torch.ones(3)], {})


def test_synthetic_checkpoint_loads_strict(model):
    params, state = synth_checkpoint(model, seed=0)
    report = {}
    sd = CV.convert_checkpoint(params, state, report=report)
    assert report["unused"] == []
    own = model.state_dict()
    constants = {k for k in own if k.endswith("inv_freq") or k.endswith("rel_pos_features.dummy")}
    assert set(own) - set(sd) == constants                       # only data-independent buffers are not in a checkpoint
    missing, unexpected = CV.load_converted(model, params, state, strict=True)
    assert not missing and not unexpected
    got = model.state_dict()
    w = torch.from_numpy(params["alphagenome/transformer_tower/mha_block_3/linear_embedding/w"])
    assert torch.equal(got["transformer_unet.transformer.layers.3.0.block.to_out.weight"], w.t())
    rv = torch.from_numpy(state["alphagenome/sequence_decoder/up_res_block_2/conv_out/rms_batch_norm/var_ema"])
    assert torch.equal(got["transformer_unet.ups.2.conv_out.net.0.running_var"], rv.squeeze())
    hw = torch.from_numpy(params["alphagenome/head/cage/resolution_128/multi_organism_linear/w"])
    assert torch.equal(got["heads.mouse.cage.resolutions.resolution_128.to_pred.weight"], hw[1].t())
    # nested Haiku trees are accepted as well
    nested = {}
    for k, v in params.items():
        mod, _, leaf = k.rpartition("/")
        nested.setdefault(mod, {})[leaf] = v
    sd2 = CV.convert_checkpoint(nested, state)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)
    # a missing array is an error, as in the reference converter
    broken = dict(params)
    broken.pop("alphagenome/output_pair/embed/embeddings")
    with pytest.raises(KeyError):
        CV.convert_checkpoint(broken, state)
    trunk_only = {k: v for k, v in params.items() if "/head/" not in k}
    with pytest.raises(KeyError):
        CV.convert_checkpoint(trunk_only, state)
    assert not any(k.startswith("heads.") for k in CV.convert_checkpoint(trunk_only, state, require_heads=False))


def _checksums(sd):
    return {k: [float(v.double().sum()), float(v.double().abs().sum()), list(v.shape)] for k, v in sorted(sd.items())}


def test_converted_checkpoint_equals_reference_converter(model):
    params, state = synth_checkpoint(model, seed=0)
    sd = CV.convert_checkpoint(params, state)
    want = json.loads((GOLD / "convert_checksums.json").read_text())
    mine = _checksums(sd)
    assert set(mine) == set(want)
    for k in want:
        assert mine[k][2] == want[k][2], k
        np.testing.assert_allclose(mine[k][:2], want[k][:2], rtol=1e-6, atol=1e-6, err_msg=k)
    ref_dir = Path("/root/reference/alphagenome_pytorch/convert")
    if not ref_dir.exists():
        return
    import sys

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    from make_golden_convert import load_reference_convert

    ref_sd = load_reference_convert()["convert_checkpoint"].convert_checkpoint(params, state, verbose=False)
    assert set(ref_sd) == set(sd)
    for k in sd:
        assert torch.equal(ref_sd[k], sd[k]), k
