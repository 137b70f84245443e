"""CPU: host-side logic of the DNAModel callers (no kernels): sequence encoding, splice-site inference from class
probabilities (dna_model.py:336-373), output-name resolution, result flattening for HostPipeline, and — when the
reference tree is present (build container) — agreement with the reference's own helpers."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from alphagenome_b200 import dna_model as DM
from alphagenome_b200.model import AlphaGenome, Embeds, EmbedsWithOperands
from alphagenome_b200.pipeline import flatten_result

ROOT = Path(__file__).resolve().parent.parent


def test_token_conversion():
    t = DM.as_token_tensor("ACGTNacgtX")
    assert t.tolist() == [[0, 1, 2, 3, -1, 0, 1, 2, 3, -1]]
    assert DM.as_token_tensor(["AC", "GT"]).tolist() == [[0, 1], [2, 3]]
    assert DM.as_token_tensor(np.array([0, 1, 2])).shape == (1, 3)
    onehot = torch.eye(4)[torch.tensor([[3, 0, 2]])]
    assert DM.as_token_tensor(onehot).tolist() == [[3, 0, 2]]
    with pytest.raises(ValueError):
        DM.as_token_tensor(["ACG", "AC"])
    with pytest.raises(ValueError):
        DM.as_token_tensor([])
    with pytest.raises(ValueError):
        DM.as_token_tensor(torch.zeros(2, 3))
    with pytest.raises(TypeError):
        DM.as_token_tensor(12)


def test_output_name_resolution():
    assert DM.resolve_output_name("rna_seq") == "RNA_SEQ" and DM.resolve_output_name("DNASE") == "DNASE"
    assert DM.resolve_output_name(None) is None

    class E:
        name = "CHIP_TF"

    assert DM.resolve_output_name(E()) == "CHIP_TF"
    d = {"scaled_predictions_1bp": 1, "scaled_predictions_128bp": 2}
    assert DM.select_head_output(d, "DNASE") == 1 and DM.select_head_output(d, "CHIP_TF") == 2 and DM.select_head_output(5, "ATAC") == 5


def _dm(num_sites, thr):
    return DM.DNAModel(AlphaGenome(dims=(128, 256), transformer_kwargs=dict(depth=1)), settings=DM.ModelSettings(num_splice_sites=num_sites, splice_site_threshold=thr))


def test_splice_site_inference_known_answer():
    dm = _dm(4, 0.3)
    probs = torch.zeros(1, 6, 5)
    probs[0, :, 0] = torch.tensor([0.9, 0.1, 0.5, 0.31, 0.2, 0.8])
    probs[0, :, 1] = torch.tensor([0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
    probs[0, :, 2] = torch.tensor([0.4, 0.4, 0.4, 0.4, 0.4, 0.4])
    pos = dm._infer_splice_site_positions_from_probs(probs)
    assert pos.shape == (1, 4, 4)
    assert pos[0, 0].tolist() == [0, 2, 3, 5]            # top-4 above 0.3, sorted
    assert pos[0, 1].tolist() == [-1, -1, -1, -1]        # nothing above the threshold
    assert sorted(pos[0, 2].tolist()) == pos[0, 2].tolist() and (pos[0, 2] >= 0).all()
    dm2 = _dm(8, 0.0)                                    # more sites wanted than positions: -1 padded
    pos2 = dm2._infer_splice_site_positions_from_probs(probs)
    assert pos2.shape == (1, 4, 8) and pos2[0, 0, :6].tolist() == [0, 1, 2, 3, 4, 5] and pos2[0, 0, 6:].tolist() == [-1, -1]


def test_splice_site_inference_matches_reference_helper():
    if not Path("/root/reference/alphagenome_pytorch/dna_model.py").exists():
        pytest.skip("reference tree not present")
    sys.path.insert(0, str(ROOT / "oracle" / "shims"))
    sys.path.append("/root/reference")
    from alphagenome_pytorch import dna_model as ref

    rng = np.random.default_rng(3)
    logits = torch.from_numpy(rng.standard_normal((2, 300, 5)).astype(np.float32))
    for n, thr in ((16, 0.1), (512, 0.0), (7, 0.25)):
        mine = _dm(n, thr)._infer_splice_site_positions(logits)
        theirs = ref.DNAModel(torch.nn.Linear(1, 1), settings=ref.ModelSettings(num_splice_sites=n, splice_site_threshold=thr))._infer_splice_site_positions(logits)
        assert torch.equal(mine, theirs)
    for seqs in ("ACGTN", ["ACG", "TTT"], np.array([[0, 1], [2, 3]])):
        assert torch.equal(DM.as_token_tensor(seqs), ref._as_token_tensor(seqs))


def test_organism_resolution_and_errors():
    dm = _dm(4, 0.1)
    assert dm._resolve_organism_index("mouse", None, 3).tolist() == [1, 1, 1]
    assert dm._resolve_organism_index(None, None, 2).tolist() == [0, 0]
    assert dm._resolve_organism_index(None, torch.tensor([1, 0]), 2).tolist() == [1, 0]
    with pytest.raises(KeyError):
        dm._resolve_organism_index("zebrafish", None, 1)
    assert DM.DNAModel._pick_organism_key({"human": 1}, None, False) == "human"
    assert DM.DNAModel._pick_organism_key({"human": 1, "mouse": 2}, None, False) is None
    assert DM.DNAModel._pick_organism_key({"human": 1, "mouse": 2}, "mouse", False) == "mouse"
    with pytest.raises(ImportError):
        dm.predict_variant()
    with pytest.raises(ValueError):
        dm.ism(["AC", "GT"], output_key="dnase")          # a single sequence only (dna_model.py:546-547)


def test_flatten_result_and_embeds_with_operands():
    e = EmbedsWithOperands(torch.zeros(1), torch.ones(2), torch.ones(3), operands=dict(embeds_1bp_b16=torch.zeros(1)))
    a, b, c = e
    assert isinstance(e, Embeds) and b.numel() == 2 and "embeds_1bp_b16" in e.operands
    assert [n for n, _ in flatten_result(e)] == ["embeds_1bp", "embeds_128bp", "embeds_pair"]
    res = {"human": {"dnase": {"scaled_predictions_1bp": torch.zeros(1), "scaled_predictions_128bp": torch.zeros(1)}, "x": None, "contact_maps": torch.zeros(2)}}
    assert [n for n, _ in flatten_result(res)] == ["human/dnase/scaled_predictions_1bp", "human/dnase/scaled_predictions_128bp", "human/contact_maps"]


def test_split_batch():
    out = {"a": torch.arange(8).view(4, 2), "d": {"p": torch.arange(4), "q": None}}
    r, a = DM._split_batch(out, 2)
    assert r["a"].tolist() == [[0, 1], [2, 3]] and a["a"].tolist() == [[4, 5], [6, 7]] and a["d"]["p"].tolist() == [2, 3] and a["d"]["q"] is None
