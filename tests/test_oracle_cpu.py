"""CPU: the oracle restatement (oracle/trunk_oracle.py) against golden outputs of the unmodified reference
(tests/golden/*.npz, produced by tools/make_golden.py).  This is what 'pins' the oracle."""
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from alphagenome_b200.init_weights import synth_state_dict
from oracle import trunk_oracle as O
from tests.golden_util import CASES, make_inputs, sample_rows

GOLD = Path(__file__).resolve().parent / "golden"
OUTPUTS = ["embeds_1bp", "embeds_128bp", "embeds_pair", "unet_output", "single_repr", "pairwise_repr"]


def reference_shapes():
    shapes = {}
    for line in (GOLD / "state_dict_keys.txt").read_text().splitlines():
        m = re.match(r"(\S+) \((.*)\)", line)
        dims = tuple(int(x) for x in m.group(2).replace(",", " ").split())
        shapes[m.group(1)] = dims
    return shapes


def synth_sd(seed):
    shapes = {k: v for k, v in reference_shapes().items() if not k.endswith("rel_pos_features.dummy")}
    sd = synth_state_dict(shapes, seed)
    sd["transformer_unet.transformer.rotary_emb.inv_freq"] = O.rotary_inv_freq(128)
    return sd


def test_reference_key_contract():
    shapes = reference_shapes()
    assert len(shapes) == 559  # head-less AlphaGenome state_dict
    assert shapes["transformer_unet.transformer.layers.0.2.to_qk.weight"] == (8192, 1536)
    assert shapes["transformer_unet.ups.6.conv_out.net.2.weight"] == (768, 768, 5)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_golden(case):
    torch.set_num_threads(max(1, torch.get_num_threads()))
    g = np.load(GOLD / f"{case['name']}.npz")
    seq, org = make_inputs(case)
    assert np.array_equal(g["seq"].astype(np.int64), seq.numpy())
    out = O.alphagenome_embeds(synth_sd(case["weight_seed"]), seq, org)
    for name in OUTPUTS:
        got = out[name]
        idx = sample_rows(name, got.shape)
        got = (got if idx is None else got[:, idx]).numpy()
        want = g[name]
        scale = float(g[name + "__absmax"])
        err = np.abs(got - want).max() / scale
        # fp32 summation-order noise only (measured ~1e-6); 2e-5 still fails on any semantic difference
        assert err < 2e-5, (name, err)


def test_seed42_input_is_reference_fixture():
    # tests/regression_data/input_sequence.npy of the reference == default_rng(42).integers(0, 4, (1, 2048))
    seq, org = make_inputs(CASES[0])
    assert seq.shape == (1, 2048) and int(org[0]) == 0
    assert seq[0, :8].tolist() == np.random.default_rng(42).integers(0, 4, size=(1, 2048))[0, :8].tolist()


# ----------------------------------------------------------------------------------------------- prediction heads
from tests.golden_util import HEAD_CASES, make_head_inputs  # noqa: E402

HEAD_NAMES = ["1bp_tracks", "128bp_tracks", "contact_head", "splice_logits", "splice_usage", "splice_juncs"]


def head_shapes():
    shapes = {}
    for line in (GOLD / "state_dict_keys_heads.txt").read_text().splitlines():
        m = re.match(r"(\S+) \((.*)\)", line)
        shapes[m.group(1)] = tuple(int(x) for x in m.group(2).replace(",", " ").split())
    return shapes


def synth_sd_with_heads(seed):
    sd = synth_sd(seed)
    sd.update(synth_state_dict(head_shapes(), seed))
    return sd


@pytest.mark.parametrize("case", HEAD_CASES, ids=[c["name"] for c in HEAD_CASES])
def test_oracle_heads_match_reference_golden(case):
    """The publication heads of both organisms (add_heads(**publication_heads_config[...])) on every sample."""
    g = np.load(GOLD / f"{case['name']}.npz")
    seq, org, donor, acceptor = make_head_inputs(case)
    assert np.array_equal(g["seq"].astype(np.int64), seq.numpy()) and np.array_equal(g["splice_donor_idx"], donor.numpy())
    sd = synth_sd_with_heads(case["weight_seed"])
    emb = O.alphagenome_embeds(sd, seq, org)
    out = O.alphagenome_heads(sd, emb, org, ("human", "mouse"), donor, acceptor)
    for organism in ("human", "mouse"):
        for name in HEAD_NAMES:
            got = out[organism][name]
            if name in ("1bp_tracks", "splice_logits", "splice_usage"):
                got = got[:, sample_rows("head", got.shape)]
            want = g[f"{organism}/{name}"]
            assert got.shape == want.shape, (organism, name, got.shape, want.shape)
            err = np.abs(got.numpy() - want).max() / float(g[f"{organism}/{name}__absmax"])
            assert err < 5e-5, (organism, name, err)


# ----------------------------------------------------------------------------------------------- JAX-aligned heads
from tests.golden_util import REF_HEAD_CASES, make_ref_head_inputs  # noqa: E402


def flatten_heads(out):
    flat = {}
    for organism, heads in out.items():
        for name, v in heads.items():
            if isinstance(v, dict):
                for k, t in v.items():
                    flat[f"{organism}/{name}.{k}"] = t
            elif v is not None:
                flat[f"{organism}/{name}"] = v
    return flat


@pytest.mark.parametrize("case", REF_HEAD_CASES, ids=[c["name"] for c in REF_HEAD_CASES])
def test_oracle_reference_heads_match_reference_golden(case):
    """add_reference_heads (the heads of the official checkpoint) for both organisms."""
    g = np.load(GOLD / f"{case['name']}.npz")
    seq, org, pos = make_ref_head_inputs(case)
    assert np.array_equal(g["splice_site_positions"], pos.numpy())
    shapes = {}
    for line in (GOLD / "state_dict_keys_ref_heads.txt").read_text().splitlines():
        m = re.match(r"(\S+) \((.*)\)", line)
        shapes[m.group(1)] = tuple(int(x) for x in m.group(2).replace(",", " ").split())
    sd = synth_sd(case["weight_seed"])
    sd.update(synth_state_dict(shapes, case["weight_seed"]))
    emb = O.alphagenome_embeds(sd, seq, org)
    flat = flatten_heads(O.alphagenome_reference_heads(sd, emb, org, ("human", "mouse"), pos))
    keys = [k for k in g.files if "/" in k and not k.endswith("__absmax")]
    assert sorted(flat) == sorted(keys)
    for k in keys:
        got = flat[k].float()
        if got.dim() == 3 and got.shape[1] > 96:
            got = got[:, sample_rows("head", got.shape)]
        want = g[k]
        assert tuple(got.shape) == want.shape, (k, tuple(got.shape), want.shape)
        err = np.abs(got.numpy() - want).max() / float(g[k + "__absmax"])
        assert err < 5e-5, (k, err)


def test_oracle_updating_batch_rmsnorm_matches_reference_golden():
    """BatchRMSNorm with updating statistics (the reference's default, AG:324-361): two consecutive forwards of the
    unmodified reference (tools/make_golden_update.py) — all 81 running_var buffers after each forward and the second
    forward's embeddings."""
    from tests.golden_util import UPDATE_CASE, make_update_inputs

    g = np.load(GOLD / f"{UPDATE_CASE['name']}.npz")
    sd = synth_sd(UPDATE_CASE["weight_seed"])
    sd[O.UPDATE_KEY] = True
    rv_keys = [k for k in sd if isinstance(k, str) and k.endswith("running_var")]
    assert len(rv_keys) == 81
    for step, (seq, org) in enumerate(make_update_inputs(UPDATE_CASE)):
        before = {k: sd[k].clone() for k in rv_keys}
        out = O.alphagenome_embeds(sd, seq, org)
        for k in rv_keys:
            want = g[f"rv{step}/{k}"]
            assert not torch.equal(sd[k], before[k]), k                      # every norm on the path was updated
            np.testing.assert_allclose(sd[k].numpy(), want, rtol=2e-4, atol=1e-6, err_msg=f"step {step} {k}")
    for name in ("embeds_1bp", "embeds_128bp", "embeds_pair"):
        got = out[name]
        idx = sample_rows(name, got.shape)
        got = (got if idx is None else got[:, idx]).numpy()
        err = np.abs(got - g[name]).max() / float(g[name + "__absmax"])
        assert err < 5e-5, (name, err)
