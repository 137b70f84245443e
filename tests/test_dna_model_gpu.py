"""-m gpu: the DNAModel callers on the B200 model against
  (1) a golden of the UNMODIFIED reference wrapper (alphagenome_pytorch/dna_model.py create / predict / ism over the
      reference model, tools/make_golden_dna_model.py), and
  (2) the reference's own schedule restated with plain model.inference calls (two full passes in predict, full outputs
      copied to the host in ism): identical results, about half the kernel launches."""
import enum
from pathlib import Path

import numpy as np
import pytest
import torch

from alphagenome_b200 import _lib
from alphagenome_b200 import dna_model as DM
from alphagenome_b200.init_weights import apply_synth_weights
from tests.golden_util import DNA_MODEL_CASE, make_dna_model_inputs

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def pearson(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    a, b = a - a.mean(), b - b.mean()
    return float((a * b).sum() / max(np.sqrt((a * a).sum() * (b * b).sum()), 1e-300))


@pytest.fixture(scope="module")
def dm():
    c = DNA_MODEL_CASE
    m = DM.create(add_reference_heads=True, device="cuda", cuda_graphs=False,
                  settings=DM.ModelSettings(num_splice_sites=c["num_splice_sites"], splice_site_threshold=c["threshold"]))
    m.eval()
    apply_synth_weights(m.model, seed=c["weight_seed"])
    return m


def test_predict_two_pass_matches_reference_wrapper_golden(dm):
    g = np.load(GOLD / f"{DNA_MODEL_CASE['name']}.npz")
    seq, _ = make_dna_model_inputs(DNA_MODEL_CASE)
    assert seq.encode() == g["seq"].tobytes()
    pred = dm.predict(seq, organism="mouse")
    assert set(pred) >= {"rna_seq", "dnase", "chip_tf", "contact_maps", "splice_sites_classification", "splice_sites_usage", "splice_sites_junction"}
    assert pearson(pred["splice_sites_classification"].cpu().numpy(), g["splice_logits"]) >= 0.999
    assert pearson(pred["dnase"]["scaled_predictions_1bp"][:, ::16].cpu().numpy(), g["dnase_1bp"]) >= 0.999
    assert pearson(pred["chip_tf"]["scaled_predictions_128bp"].cpu().numpy(), g["chip_tf_128bp"]) >= 0.999
    # inferred splice sites: top-k of softmax probabilities computed from bf16-operand logits — the sets agree up to
    # near-ties at the k-th place
    got_pos = pred["splice_sites_junction"]["splice_site_positions"].cpu().numpy()
    want_pos = g["splice_site_positions"]
    assert got_pos.shape == want_pos.shape
    overlap = np.mean([len(set(got_pos[0, c]) & set(want_pos[0, c])) / want_pos.shape[2] for c in range(4)])
    assert overlap >= 0.8, overlap
    # with the reference's positions the junction head agrees with the reference's junction predictions
    fixed = dm.predict(seq, organism="mouse", splice_site_positions=torch.from_numpy(want_pos).cuda())
    j = fixed["splice_sites_junction"]
    assert np.array_equal(j["splice_junction_mask"].cpu().numpy(), g["junction_mask"])
    assert pearson(j["predictions"].cpu().numpy(), g["junction_predictions"]) >= 0.999


def test_predict_reuses_the_trunk(dm):
    """Same outputs as the reference's schedule (two full inference passes, dna_model.py:431-470), one trunk forward."""
    seq, _ = make_dna_model_inputs(DNA_MODEL_CASE)
    tokens = DM.as_token_tensor(seq).cuda()
    org = torch.ones(1, dtype=torch.long, device="cuda")
    l0 = _lib.launch_count()
    first = dm.model.inference(tokens, org)
    positions = dm._infer_splice_site_positions(first["mouse"]["splice_sites_classification"])
    naive = dm.model.inference(tokens, org, splice_site_positions=positions)
    l_naive = _lib.launch_count() - l0
    n0 = dm.trunk_forwards
    l0 = _lib.launch_count()
    mine = dm.predict(seq, organism="mouse")
    l_mine = _lib.launch_count() - l0
    assert dm.trunk_forwards - n0 == 1
    assert l_mine <= 0.62 * l_naive, (l_mine, l_naive)

    def same(a, b, path):
        if torch.is_tensor(a):
            assert torch.equal(a, b), path
        elif isinstance(a, dict):
            assert set(a) == set(b), path
            for k in a:
                same(a[k], b[k], path + "/" + str(k))
        else:
            assert a is b or a == b, path

    same(mine, naive["mouse"], "")
    with pytest.raises(ValueError):      # like the reference: the inferred-sites flow needs ONE organism's classification head
        dm.predict(seq, organism="mouse", return_all_organisms=True)
    # requested_outputs: the classification head is pulled in for the first pass (dna_model.py:410-418)
    sel = dm.predict(seq, organism="mouse", requested_outputs={"splice_sites_junction", "cage"})
    assert set(sel) == {"cage", "splice_sites_classification", "splice_sites_junction"}
    same(sel["splice_sites_junction"], naive["mouse"]["splice_sites_junction"], "sel")
    # no junctions wanted: one pass, junction head reports None like the reference's first pass
    one = dm.predict(seq, organism="mouse", include_splice_junctions=False)
    assert one["splice_sites_junction"] is None
    with pytest.raises(ValueError):
        dm.predict(seq)          # two organisms' heads, junction inference needs to know which one (dna_model.py:441-446)
    emb = dm.embeddings([seq, seq], organism="human")
    assert tuple(emb.embeds_1bp.shape) == (2, 2048, 1536)


@pytest.mark.parametrize("key", ["dnase", "chip_tf"])
def test_ism_matches_reference_wrapper_golden_and_naive_schedule(dm, key):
    g = np.load(GOLD / f"{DNA_MODEL_CASE['name']}.npz")
    seq, positions = make_dna_model_inputs(DNA_MODEL_CASE)
    out = dm.ism(seq, output_key=key, organism="mouse", positions=positions, batch_size=DNA_MODEL_CASE["ism_batch"], return_predictions=True)
    assert np.array_equal(out["positions"], g[f"ism_{key}_positions"])
    assert "".join(out["alt_bases"]).encode() == g[f"ism_{key}_alt_bases"].tobytes()
    assert out["delta"].shape == g[f"ism_{key}_delta"].shape
    assert pearson(out["predictions"], g[f"ism_{key}_predictions"]) >= 0.999
    # delta = difference of two predictions that each carry bf16-operand rounding noise: the noise does not cancel, the
    # signal (a one-base change) is small, so the bar is lower than for the predictions themselves (measured 0.977-0.984)
    assert pearson(out["delta"], g[f"ism_{key}_delta"]) >= 0.95
    # the reference's schedule restated: every mutant batch's full output goes to the host and is indexed there
    tokens = DM.as_token_tensor(seq).cuda()
    org = torch.ones(1, dtype=torch.long, device="cuda")
    head = {"dnase": "scaled_predictions_1bp", "chip_tf": "scaled_predictions_128bp"}[key]
    res = 1 if key == "dnase" else 128
    ref = dm.model.inference(tokens, org, requested_heads={key})["mouse"][key][head].cpu()
    muts = list(zip(out["positions"].tolist(), out["alt_bases"].tolist()))
    deltas = []
    for s in range(0, len(muts), DNA_MODEL_CASE["ism_batch"]):
        chunk = muts[s:s + DNA_MODEL_CASE["ism_batch"]]
        bt = tokens.repeat(len(chunk), 1)
        for r, (p, b) in enumerate(chunk):
            bt[r, p] = "ACGT".index(b)
        o = dm.model.inference(bt, org.repeat(len(chunk)), requested_heads={key})["mouse"][key][head].cpu()
        rows = [p // res for p, _ in chunk]
        deltas.append((o[range(len(chunk)), rows, :] - ref[0, rows, :]).numpy())
    assert np.array_equal(out["delta"], np.concatenate(deltas, axis=0))
    # graphs: same numbers when every batch shape replays its own captured graph
    dm.model.enable_cuda_graphs(True)
    try:
        again = dm.ism(seq, output_key=key, organism="mouse", positions=positions, batch_size=DNA_MODEL_CASE["ism_batch"])
    finally:
        dm.model.enable_cuda_graphs(False)
    assert np.array_equal(again["delta"], out["delta"])
    empty = dm.ism(seq, output_key=key, organism="mouse", positions=[])
    assert empty["delta"].shape == (0, out["delta"].shape[1])


class _Out(enum.Enum):
    DNASE = 1
    RNA_SEQ = 2


class _FakeScorer:
    """Records what AlphaGenome.score_variant hands to the scorer protocol (AG:2311-2346)."""

    def get_masks_and_metadata(self, interval, variant, *, settings, track_metadata):
        self.track_metadata = track_metadata
        return "masks", "mask_metadata"

    def score_variant(self, ref, alt, *, masks, settings, variant, interval):
        self.ref, self.alt = ref, alt
        assert masks == "masks"
        return {"score": (alt[settings.requested_output] - ref[settings.requested_output]).abs().amax(dim=1)}

    def finalize_variant(self, scores, *, track_metadata, mask_metadata, settings):
        assert mask_metadata == "mask_metadata" and isinstance(scores["score"], np.ndarray)
        return scores


def test_score_variant_batches_both_alleles(dm):
    seq, _ = make_dna_model_inputs(DNA_MODEL_CASE)
    ref = DM.as_token_tensor(seq)
    alt = ref.clone()
    alt[0, 1000] = (alt[0, 1000] + 1) % 4

    class Settings:
        requested_output = _Out.DNASE

    class Meta(dict):
        padding = {_Out.DNASE: np.array([False] * 380 + [True] * 4)}
        strand_reindexing = {_Out.DNASE: np.arange(380)[::-1].copy()}

    class Interval:
        negative_strand = True

    meta = Meta({_Out.DNASE: np.arange(384)})
    sc = _FakeScorer()
    n0 = dm.trunk_forwards
    res = dm.score_variant(ref, alt, organism="mouse", scorer=sc, settings=Settings, variant=None, interval=Interval, track_metadata=meta,
                           requested_outputs={"dnase"})
    assert dm.trunk_forwards - n0 == 1 and res["score"].shape == (1, 380)
    assert list(sc.ref) == [_Out.DNASE] and tuple(sc.ref[_Out.DNASE].shape) == (1, 2048, 380)
    assert list(sc.track_metadata) == [_Out.DNASE] and len(sc.track_metadata[_Out.DNASE]) == 380
    # against two separate forwards + the same masking / reverse-complement by hand
    org = torch.ones(1, dtype=torch.long, device="cuda")
    for tok, got in ((ref, sc.ref), (alt, sc.alt)):
        o = dm.model.inference(tok.cuda(), org, requested_heads={"dnase"})["mouse"]["dnase"]["scaled_predictions_1bp"]
        want = o[..., :380].flip(1)[..., torch.arange(379, -1, -1, device="cuda")]
        assert torch.allclose(got[_Out.DNASE], want, rtol=1e-3, atol=1e-3 * float(want.abs().max()))
    assert float(res["score"].max()) > 0
    with pytest.raises(ImportError):
        dm.predict_interval(None)
