"""-m gpu: the publication prediction heads of both organisms on the B200 path against golden outputs
of the unmodified reference (tests/golden/heads_*.npz, tools/make_golden_heads.py).
north_star tolerance for head outputs: Pearson >= 0.999; max|d| / max|ref| is reported and bounded at 3e-2
(a head stacks one more bf16 contraction on embeddings that are themselves within 1e-2)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from alphagenome_b200.init_weights import apply_synth_weights
from alphagenome_b200.model import AlphaGenome, publication_heads_config, set_update_running_var
from tests.golden_util import HEAD_CASES, make_head_inputs, sample_rows

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
HEAD_NAMES = ["1bp_tracks", "128bp_tracks", "contact_head", "splice_logits", "splice_usage", "splice_juncs"]

_M = {}


def model_with_heads(seed):
    if "m" not in _M:
        m = AlphaGenome()
        for org in ("human", "mouse"):
            m.add_heads(**publication_heads_config[org])
        _M["m"] = m.cuda().eval()
        set_update_running_var(_M["m"], False)
    apply_synth_weights(_M["m"], seed=seed)
    return _M["m"]


def pearson(a, b):
    a = a.astype(np.float64).ravel() - a.mean()
    b = b.astype(np.float64).ravel() - b.mean()
    return float((a * b).sum() / np.sqrt((a * a).sum() * (b * b).sum()))


@pytest.mark.parametrize("case", HEAD_CASES, ids=[c["name"] for c in HEAD_CASES])
def test_heads_match_reference_golden(case):
    g = np.load(GOLD / f"{case['name']}.npz")
    seq, org, donor, acceptor = make_head_inputs(case)
    model = model_with_heads(case["weight_seed"])
    out = model(seq.cuda(), org.cuda(), splice_donor_idx=donor.cuda(), splice_acceptor_idx=acceptor.cuda())
    torch.cuda.synchronize()
    assert set(out) == {"human", "mouse"}
    report = {}
    for organism in ("human", "mouse"):
        assert list(out[organism]) == HEAD_NAMES
        for name in HEAD_NAMES:
            got = out[organism][name].float().cpu()
            if name in ("1bp_tracks", "splice_logits", "splice_usage"):
                got = got[:, sample_rows("head", got.shape)]
            want = g[f"{organism}/{name}"]
            assert tuple(got.shape) == want.shape, (organism, name, tuple(got.shape), want.shape)
            err = np.abs(got.numpy() - want).max() / float(g[f"{organism}/{name}__absmax"])
            r = pearson(got.numpy(), want)
            report[f"{organism}/{name}"] = (round(err, 5), round(r, 6))
            assert r >= 0.999, (organism, name, r, err)
            assert err <= 3e-2, (organism, name, err)
    print(report)


def test_forward_without_head_kwargs_matches_reference_behaviour():
    """splice_juncs needs splice_donor_idx / splice_acceptor_idx: like the reference (AG:1838), a missing head input is
    a KeyError; return_embeds still short-circuits the heads (AG:1802)."""
    model = model_with_heads(0)
    seq = torch.randint(0, 4, (1, 2048), device="cuda")
    emb = model(seq, 0, return_embeds=True)
    assert tuple(emb.embeds_1bp.shape) == (1, 2048, 1536)
    with pytest.raises(KeyError):
        model(seq, 0)


def test_custom_head_is_called_like_the_reference():
    """add_head with a user module and explicit input names (AG:1621-1651): the module is simply called on the embeddings."""

    class MeanHead(torch.nn.Module):
        def forward(self, x):
            return x.mean(dim=1)

    m = AlphaGenome().cuda().eval()
    set_update_running_var(m, False)
    apply_synth_weights(m, seed=0)
    m.add_head("human", "mean128", MeanHead(), "embeds_128bp")
    seq = torch.randint(0, 4, (1, 2048), device="cuda")
    out = m(seq, 0)
    emb = m(seq, 0, return_embeds=True)
    assert torch.allclose(out["human"]["mean128"], emb.embeds_128bp.mean(dim=1))


def test_inference_matches_forward_and_filters_heads():
    """Mirrors the reference's tests/test_selective_inference.py:42-146: inference() == forward(), head filtering,
    track masks, embedding reuse."""
    case = HEAD_CASES[0]
    seq, org, donor, acceptor = make_head_inputs(case)
    model = model_with_heads(case["weight_seed"])
    kw = dict(splice_donor_idx=donor.cuda(), splice_acceptor_idx=acceptor.cuda())
    full = model(seq.cuda(), org.cuda(), **kw)
    inf = model.inference(seq.cuda(), org.cuda(), **kw)
    for organism in full:
        for name in full[organism]:
            assert torch.equal(full[organism][name], inf[organism][name]), (organism, name)
    # head filtering: only the requested heads run; organisms without any requested head are dropped
    sel = model.inference(seq.cuda(), org.cuda(), requested_heads={"128bp_tracks", "contact_head"})
    assert set(sel) == {"human", "mouse"} and set(sel["human"]) == {"128bp_tracks", "contact_head"}
    assert torch.equal(sel["mouse"]["contact_head"], full["mouse"]["contact_head"])
    # track masks slice the last dimension
    mask = torch.tensor([True, False, True, True, False], device="cuda")      # splice_logits has 5 tracks for both organisms
    masked = model.inference(seq.cuda(), org.cuda(), requested_heads={"splice_logits"}, track_masks={"splice_logits": mask})
    for organism in ("human", "mouse"):
        assert torch.equal(masked[organism]["splice_logits"], full[organism]["splice_logits"][..., mask])
    # embedding reuse: heads on embeddings handed back by the caller (bf16 operands made by rounding the fp32 tensors)
    emb = model.get_embeds(seq.cuda(), org.cuda())
    emb = type(emb)(*[t.clone() for t in emb])
    reused, emb2 = model.inference(seq.cuda(), org.cuda(), embeds=emb, return_embeds=True, requested_heads={"1bp_tracks", "contact_head"})
    assert emb2 is emb
    for organism in ("human", "mouse"):
        for name in ("1bp_tracks", "contact_head"):
            a, b = reused[organism][name].float(), full[organism][name].float()
            assert ((a - b).abs().max() / b.abs().max()).item() < 5e-3, (organism, name)


# ----------------------------------------------------------------------------------------------- JAX-aligned heads
from tests.golden_util import REF_HEAD_CASES, make_ref_head_inputs  # noqa: E402


@pytest.mark.parametrize("case", REF_HEAD_CASES, ids=[c["name"] for c in REF_HEAD_CASES])
def test_reference_heads_match_reference_golden(case):
    """add_reference_heads (the heads of the official checkpoint, AG:1684-1731) for both organisms."""
    g = np.load(GOLD / f"{case['name']}.npz")
    seq, org, pos = make_ref_head_inputs(case)
    m = AlphaGenome()
    for o in ("human", "mouse"):
        m.add_reference_heads(o)
    m = m.cuda().eval()
    set_update_running_var(m, False)
    apply_synth_weights(m, seed=case["weight_seed"])
    out = m(seq.cuda(), org.cuda(), splice_site_positions=pos.cuda())
    torch.cuda.synchronize()
    flat = {}
    for organism, heads in out.items():
        for name, v in heads.items():
            if isinstance(v, dict):
                for k, t in v.items():
                    flat[f"{organism}/{name}.{k}"] = t
            elif v is not None:
                flat[f"{organism}/{name}"] = v
    keys = [k for k in g.files if "/" in k and not k.endswith("__absmax")]
    assert sorted(flat) == sorted(keys)
    worst = (0.0, 1.0)
    for k in keys:
        got = flat[k].float().cpu()
        if got.dim() == 3 and got.shape[1] > 96:
            got = got[:, sample_rows("head", got.shape)]
        want = g[k]
        assert tuple(got.shape) == want.shape, (k, tuple(got.shape), want.shape)
        if k.endswith("splice_site_positions") or k.endswith("splice_junction_mask"):
            assert np.array_equal(got.numpy(), want), k
            continue
        err = np.abs(got.numpy() - want).max() / float(g[k + "__absmax"])
        r = pearson(got.numpy(), want)
        worst = (max(worst[0], err), min(worst[1], r))
        assert r >= 0.999 and err <= 3e-2, (k, err, r)
    print("reference heads: worst max-rel-err %.4f, min Pearson %.6f" % worst)
    # without splice_site_positions the junction head returns None (AG:1506-1507)
    out2 = m.inference(seq.cuda(), org.cuda(), requested_heads={"splice_sites_junction", "contact_maps"})
    assert out2["human"]["splice_sites_junction"] is None and tuple(out2["mouse"]["contact_maps"].shape) == (2, 2, 2, 28)
