"""`DNAModel`: the caller-side wrapper of the reference (alphagenome_pytorch/dna_model.py: create :72-113, embeddings
:375-389, predict :391-483, score_variant :485-527, ism :529-673) on top of the B200 model — same signatures, same return
values — with the scheduling the reference leaves on the table:

  * predict(): the two-pass splice flow (classification -> inferred splice sites -> junction head) runs the trunk ONCE:
    the first pass keeps its embeddings and bf16 head operands, the second pass runs only the junction head on them
    (the reference recomputes the whole forward, dna_model.py:431-470);
  * ism(): mutant batches are built on the device, replayed from one CUDA graph per batch shape, and only the scored
    rows [mutants, tracks] leave the device (the reference copies every [B, S, T] output to the host and indexes there,
    :630-661);
  * score_variant(): reference and alternate allele go through ONE batched forward (AG:2091-2104 runs two).

`predict_interval` / `predict_variant` / `output_metadata` (dna_model.py:711-938) are formatting layers over the
`alphagenome` / `alphagenome_research` SDK types (FASTA extractors, TrackData, AnnData); those packages are not part of
this framework, so these entry points raise ImportError naming what is missing.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Mapping, Optional, Sequence

import numpy as np
import torch

from .model import AlphaGenome

_DNA_TO_INDEX = {"A": 0, "C": 1, "G": 2, "T": 3, "N": -1}
_INDEX_TO_DNA = np.array(["A", "C", "G", "T"])
_PAD_VALUE = -1
_NUM_SPLICE_CLASSES = 4

# dna_model.py:35-62
OUTPUT_NAME_TO_RESOLUTION = {"ATAC": 1, "CAGE": 1, "DNASE": 1, "RNA_SEQ": 1, "PROCAP": 1, "CHIP_HISTONE": 128, "CHIP_TF": 128,
                             "SPLICE_SITES": 1, "SPLICE_SITE_USAGE": 1, "SPLICE_JUNCTIONS": 1, "CONTACT_MAPS": 2048}
HEAD_TO_OUTPUT_NAME = {"atac": "ATAC", "cage": "CAGE", "dnase": "DNASE", "rna_seq": "RNA_SEQ", "procap": "PROCAP",
                       "chip_histone": "CHIP_HISTONE", "chip_tf": "CHIP_TF", "contact_maps": "CONTACT_MAPS",
                       "splice_sites_classification": "SPLICE_SITES", "splice_sites_usage": "SPLICE_SITE_USAGE",
                       "splice_sites_junction": "SPLICE_JUNCTIONS"}
JUNCTION_HEAD = "splice_sites_junction"
CLASSIFICATION_HEAD = "splice_sites_classification"


@dataclass
class ModelSettings:
    """dna_model.py:64-69."""

    num_splice_sites: int = 512
    splice_site_threshold: float = 0.1


def create(*, model: Optional[AlphaGenome] = None, model_kwargs: Optional[Mapping] = None, settings: Optional[ModelSettings] = None,
           device=None, add_reference_heads: bool = False, organisms: Sequence[str] = ("human", "mouse"),
           checkpoint_path: Optional[str] = None, strict: bool = False, compile: bool = False, compile_kwargs=None,
           cuda_graphs: bool = True, freeze_norms: bool = True) -> "DNAModel":
    """dna_model.py:72-113.  `compile` is accepted for signature compatibility and ignored (the forward is hand-written
    kernels; `cuda_graphs` is the launch-overhead remedy here).  `freeze_norms` runs BatchRMSNorm with its stored
    statistics (inference); pass False to keep the reference's default of re-estimating them on every forward."""
    from .model import set_update_running_var

    if model is None:
        model = AlphaGenome(**(model_kwargs or {}))
    if add_reference_heads:
        for organism in organisms:
            model.add_reference_heads(organism)
    if checkpoint_path is not None:
        ckpt = torch.load(checkpoint_path, map_location="cpu")
        model.load_state_dict(ckpt.get("model_state_dict", ckpt), strict=strict)
    if freeze_norms:
        set_update_running_var(model, False)
    model.enable_cuda_graphs(cuda_graphs)
    return DNAModel(model, settings=settings, device=device)


def encode_sequence(seq: str) -> torch.Tensor:
    return torch.tensor([_DNA_TO_INDEX.get(b, -1) for b in seq.upper()], dtype=torch.long)


def as_token_tensor(sequences) -> torch.Tensor:
    """dna_model.py:122-149: str | list of equal-length str | int array [S] / [B, S] | one-hot float [.., S, 4] -> int64 [B, S]."""
    if isinstance(sequences, np.ndarray):
        sequences = torch.from_numpy(sequences)
    if torch.is_tensor(sequences):
        if sequences.dtype.is_floating_point:
            if sequences.ndim >= 2 and sequences.shape[-1] == 4:
                return sequences.argmax(dim=-1).long()
            raise ValueError("Floating-point sequences must be one-hot with last dim=4.")
        return (sequences.unsqueeze(0) if sequences.ndim == 1 else sequences).long()
    if isinstance(sequences, str):
        return encode_sequence(sequences).unsqueeze(0)
    if isinstance(sequences, Sequence):
        if not sequences:
            raise ValueError("Empty sequence list.")
        enc = [encode_sequence(s) for s in sequences]
        if len({e.shape[0] for e in enc}) != 1:
            raise ValueError("All sequences must have the same length.")
        return torch.stack(enc, dim=0)
    raise TypeError("Sequences must be a string, a list of strings, a numpy array, or a torch Tensor.")


def resolve_output_name(output_key) -> Optional[str]:
    if output_key is None:
        return None
    if hasattr(output_key, "name"):
        return output_key.name
    if isinstance(output_key, str):
        return HEAD_TO_OUTPUT_NAME.get(output_key, output_key.upper())
    return None


def select_head_output(head_output, output_name: str):
    """dna_model.py:164-173: pick the resolution of a GenomeTracksHead dict that the output type is defined at."""
    res = OUTPUT_NAME_TO_RESOLUTION.get(output_name)
    if isinstance(head_output, dict):
        key = f"scaled_predictions_{res}bp"
        if key in head_output:
            return head_output[key]
    return head_output


def _split_batch(x, parts: int):
    """A head output computed on `parts` concatenated batches -> list of per-batch outputs (tensors split on dim 0)."""
    if torch.is_tensor(x):
        return list(x.chunk(parts, dim=0))
    if isinstance(x, dict):
        cols = {k: _split_batch(v, parts) for k, v in x.items()}
        return [{k: v[i] for k, v in cols.items()} for i in range(parts)]
    return [x] * parts


class DNAModel:
    """dna_model.py:176-673 on the B200 model."""

    def __init__(self, model: AlphaGenome, *, settings: Optional[ModelSettings] = None, device=None,
                 organism_map: Optional[Mapping[str, int]] = None, track_ontology: Optional[Mapping[str, Mapping[str, Sequence[str]]]] = None):
        """track_ontology: {organism: {head name: ontology CURIE of every track}} — the plain-data form of the SDK
        metadata the reference reads its `ontology_terms` filter from (dna_model.py:262-334)."""
        self.model = model
        self.settings = settings or ModelSettings()
        self.organism_map = dict(organism_map or {"human": 0, "mouse": 1})
        self.track_ontology = {k: dict(v) for k, v in (track_ontology or {}).items()}
        self.trunk_forwards = 0     # how many times a trunk forward was launched through this wrapper
        if device is not None:
            self.model.to(device)

    @property
    def device(self) -> torch.device:
        p = next(iter(self.model.parameters()), None)
        return p.device if p is not None else torch.device("cpu")

    def to(self, device) -> "DNAModel":
        self.model.to(device)
        return self

    def eval(self) -> "DNAModel":
        self.model.eval()
        return self

    def train(self, mode: bool = True) -> "DNAModel":
        self.model.train(mode)
        return self

    # ------------------------------------------------------------------------------------------ helpers
    def _resolve_organism_index(self, organism, organism_index, batch_size: int) -> torch.Tensor:
        if organism_index is None:
            if organism is None:
                organism_index = 0
            elif organism not in self.organism_map:
                raise KeyError(f"Unknown organism {organism!r}. Available: {list(self.organism_map.keys())}")
            else:
                organism_index = self.organism_map[organism]
        if torch.is_tensor(organism_index):
            return organism_index.to(device=self.device)
        return torch.full((batch_size,), int(organism_index), dtype=torch.long, device=self.device)

    @staticmethod
    def _pick_organism_key(predictions: Mapping, organism, return_all_organisms: bool):
        if return_all_organisms:
            return None
        if organism is not None:
            return organism
        if len(predictions) == 1:
            return next(iter(predictions.keys()))
        return None

    def _resolve_track_masks(self, organism, requested_outputs, ontology_terms):
        if ontology_terms is None:
            return None
        meta = self.track_ontology.get(organism or "human")
        if not meta:
            return None
        wanted = set(ontology_terms)
        heads = self.model.heads[organism or "human"] if (organism or "human") in self.model.heads else {}
        names = set(requested_outputs) if requested_outputs else set(heads.keys())
        masks = {h: torch.tensor([str(c) in wanted for c in meta[h]], dtype=torch.bool, device=self.device)
                 for h in names if h in meta and h in HEAD_TO_OUTPUT_NAME}
        return masks or None

    def _infer_splice_site_positions_from_probs(self, probs: torch.Tensor) -> torch.Tensor:
        """dna_model.py:336-369: per class the top-k positions above the threshold, sorted, -1 padded -> [B, 4, K]."""
        probs = probs.float()
        batch, seq_len, _ = probs.shape
        k = min(self.settings.num_splice_sites, seq_len)
        pad = self.settings.num_splice_sites - k
        thr = self.settings.splice_site_threshold
        out = []
        for c in range(_NUM_SPLICE_CLASSES):
            vals, idx = probs[:, :, c].topk(k, dim=1)
            if thr > 0:
                idx = idx.masked_fill(vals < thr, seq_len)
            idx, _ = idx.sort(dim=1)
            if thr > 0:
                idx = idx.masked_fill(idx == seq_len, _PAD_VALUE)
            if pad > 0:
                idx = torch.cat((idx, torch.full((batch, pad), _PAD_VALUE, dtype=idx.dtype, device=idx.device)), dim=1)
            out.append(idx)
        return torch.stack(out, dim=1)

    def _infer_splice_site_positions(self, splice_logits: torch.Tensor) -> torch.Tensor:
        return self._infer_splice_site_positions_from_probs(torch.softmax(splice_logits.float(), dim=-1))

    def _inference(self, tokens, organism_index, **kw):
        if kw.get("embeds") is None:
            self.trunk_forwards += 1
        return self.model.inference(tokens, organism_index, **kw)

    # ------------------------------------------------------------------------------------------ public
    @torch.no_grad()
    def embeddings(self, sequences, *, organism=None, organism_index=None, no_grad: bool = True):
        tokens = as_token_tensor(sequences).to(self.device)
        org = self._resolve_organism_index(organism, organism_index, tokens.shape[0])
        self.trunk_forwards += 1
        return self.model(tokens, org, return_embeds=True)

    @torch.no_grad()
    def predict(self, sequences, *, organism=None, organism_index=None, include_splice_junctions: bool = True,
                splice_site_positions=None, return_all_organisms: bool = False, no_grad: bool = True,
                requested_outputs: Optional[Iterable[str]] = None, ontology_terms: Optional[Iterable[str]] = None, **head_kwargs):
        """dna_model.py:391-483.  One trunk forward; when the splice sites have to be inferred from the classification
        head first, the junction head then runs on the embeddings of that same forward."""
        tokens = as_token_tensor(sequences).to(self.device)
        org = self._resolve_organism_index(organism, organism_index, tokens.shape[0])
        if splice_site_positions is None:
            splice_site_positions = head_kwargs.get("splice_site_positions")
        if splice_site_positions is not None and "splice_site_positions" not in head_kwargs:
            head_kwargs = {**head_kwargs, "splice_site_positions": splice_site_positions}
        requested = set(requested_outputs) if requested_outputs else None
        two_pass = include_splice_junctions and splice_site_positions is None
        if two_pass and requested is not None and JUNCTION_HEAD in requested:
            requested = requested | {CLASSIFICATION_HEAD}          # needed to infer the splice sites (:410-418)
        masks = self._resolve_track_masks(organism, requested_outputs, ontology_terms)

        predictions, embeds = self._inference(tokens, org, requested_heads=requested, track_masks=masks, return_embeds=True, **head_kwargs)
        if two_pass and isinstance(predictions, dict) and len(self.model.heads) != 0:
            key = self._pick_organism_key(predictions, organism, return_all_organisms)
            if key is None:
                raise ValueError("Multiple organisms present. Provide organism=... or splice_site_positions=... when include_splice_junctions=True.")
            organism_preds = predictions.get(key, {})
            logits = organism_preds.get(CLASSIFICATION_HEAD)
            if logits is None:
                return predictions if return_all_organisms else organism_preds
            positions = self._infer_splice_site_positions(logits)
            if requested is None or JUNCTION_HEAD in requested:
                # second pass (:451-470): only the head that consumes the inferred sites, on the first pass's embeddings
                second = self._inference(tokens, org, requested_heads={JUNCTION_HEAD}, track_masks=masks, embeds=embeds,
                                         **{**head_kwargs, "splice_site_positions": positions})
                for o, heads in second.items():
                    predictions.setdefault(o, {}).update(heads)
        if not isinstance(predictions, dict) or len(self.model.heads) == 0:
            return predictions
        key = self._pick_organism_key(predictions, organism, return_all_organisms)
        if key is None:
            return predictions
        if key not in predictions:
            raise KeyError(f"Organism {key!r} not found in model output. Available: {list(predictions.keys())}")
        return predictions[key]

    @torch.no_grad()
    def score_variant(self, seq_ref, seq_alt, *, organism=None, organism_index=None, scorer, settings, variant, interval, track_metadata,
                      no_grad: bool = True, requested_outputs: Optional[Iterable[str]] = None, ontology_terms: Optional[Iterable[str]] = None):
        """dna_model.py:485-527 -> AlphaGenome.score_variant (AG:2003-2346)."""
        ref = as_token_tensor(seq_ref).to(self.device)
        alt = as_token_tensor(seq_alt).to(self.device)
        org = self._resolve_organism_index(organism, organism_index, ref.shape[0])
        self.trunk_forwards += 1 if ref.shape == alt.shape else 2
        return self.model.score_variant(ref, alt, org, scorer=scorer, settings=settings, variant=variant, interval=interval,
                                        track_metadata=track_metadata, organism=organism,
                                        requested_heads=set(requested_outputs) if requested_outputs else None,
                                        track_masks=self._resolve_track_masks(organism, requested_outputs, ontology_terms))

    @torch.no_grad()
    def ism(self, sequence, *, output_key, organism=None, organism_index=None, positions: Optional[Iterable[int]] = None,
            alt_bases: str = "ACGT", batch_size: int = 32, no_grad: bool = True, return_predictions: bool = False,
            requested_outputs: Optional[Iterable[str]] = None, ontology_terms: Optional[Iterable[str]] = None):
        """In-silico mutagenesis (dna_model.py:529-673): delta[m] = alt prediction - reference prediction at the mutated
        position (or its 128 bp bin), for every (position, alternative base)."""
        tokens = as_token_tensor(sequence).to(self.device)
        if tokens.shape[0] != 1:
            raise ValueError("ISM currently supports a single sequence.")
        org = self._resolve_organism_index(organism, organism_index, batch_size=1)
        output_name = resolve_output_name(output_key)
        if output_name is None:
            raise ValueError("output_key must be a string or enum with a name attribute.")
        head_name = next((h for h, o in HEAD_TO_OUTPUT_NAME.items() if o == output_name), None)
        if requested_outputs is not None:
            requested = set(requested_outputs)
        else:
            requested = {head_name} if head_name is not None else None
        masks = self._resolve_track_masks(organism, requested_outputs, ontology_terms)

        def run(tok, o):
            preds = self._inference(tok, o, requested_heads=requested, track_masks=masks)
            if isinstance(preds, dict):
                key = self._pick_organism_key(preds, organism, return_all_organisms=False)
                if key is None:
                    raise ValueError("Provide organism=... when multiple organisms exist.")
                preds = preds[key]
            if head_name is None or head_name not in preds:
                raise KeyError(f"Output {output_name!r} not found in model predictions.")
            return select_head_output(preds[head_name], output_name)

        ref_output = run(tokens, org)
        if ref_output.ndim != 3:
            raise ValueError("ISM currently supports 1D track outputs with shape [B, S, T].")
        ref_output = ref_output.clone()      # under CUDA graphs the next replay overwrites the head's static output
        seq_len = tokens.shape[1]
        pos_list = list(positions) if positions is not None else list(range(seq_len))
        bases = [b for b in alt_bases.upper() if b in _DNA_TO_INDEX]
        ref_tokens = tokens[0].cpu().numpy()
        ref_bases = _INDEX_TO_DNA[ref_tokens]                    # N (-1) wraps to 'T', exactly like the reference (:607-608)
        mutations = [(p, b) for p in pos_list if 0 <= p < seq_len for b in bases if b != ref_bases[p]]
        if not mutations:
            return {"positions": np.array([], dtype=np.int64), "alt_bases": np.array([], dtype="<U1"),
                    "delta": np.zeros((0, ref_output.shape[-1]), dtype=np.float32)}
        resolution = OUTPUT_NAME_TO_RESOLUTION.get(output_name, 1)
        m_pos = torch.tensor([p for p, _ in mutations], dtype=torch.long, device=self.device)
        m_base = torch.tensor([_DNA_TO_INDEX[b] for _, b in mutations], dtype=torch.long, device=self.device)
        m_row = m_pos if resolution == 1 else m_pos // resolution
        deltas, alts = [], []
        for s in range(0, len(mutations), batch_size):
            pos, base, row = m_pos[s:s + batch_size], m_base[s:s + batch_size], m_row[s:s + batch_size]
            nb = pos.shape[0]
            batch = tokens.repeat(nb, 1)
            ar = torch.arange(nb, device=self.device)
            batch[ar, pos] = base                                  # mutant batch built on the device
            out = run(batch, org.repeat(nb))
            alt_at = out[ar, row, :]                               # only the scored rows leave the forward's output
            deltas.append(alt_at - ref_output[0, row, :])
            if return_predictions:
                alts.append(alt_at.clone())
        res = {"positions": np.array([p for p, _ in mutations], dtype=np.int64), "alt_bases": np.array([b for _, b in mutations]),
               "delta": torch.cat(deltas, dim=0).float().cpu().numpy()}
        if return_predictions:
            res["predictions"] = torch.cat(alts, dim=0).float().cpu().numpy()
        return res

    # ------------------------------------------------------------------------------------------ SDK formatting layers
    def _needs_sdk(self, what: str):
        raise ImportError(f"DNAModel.{what} formats results with the `alphagenome` / `alphagenome_research` SDK (FASTA extractors, TrackData, "
                          "AnnData metadata); those packages are outside this framework — use predict() / score_variant() / ism() on sequences")

    def output_metadata(self, *a, **k):
        self._needs_sdk("output_metadata")

    def predict_interval(self, *a, **k):
        self._needs_sdk("predict_interval")

    def predict_variant(self, *a, **k):
        self._needs_sdk("predict_variant")


# ------------------------------------------------------------------------------------------------ variant scoring glue
def _looks_one_hot(seq: torch.Tensor) -> bool:
    if seq.dim() < 2 or seq.shape[-1] != 4:
        return False
    if seq.dtype.is_floating_point or seq.dtype == torch.bool:
        return True
    if seq.dtype in (torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64):
        return seq.min().item() >= 0 and seq.max().item() <= 1
    return False


def _normalize_seq(seq) -> torch.Tensor:
    if not torch.is_tensor(seq):
        raise TypeError("Sequence inputs must be torch.Tensors.")
    if _looks_one_hot(seq):
        seq = seq.argmax(dim=-1)
    if seq.dim() == 1:
        seq = seq.unsqueeze(0)
    if seq.dim() != 2:
        raise ValueError(f"Expected tokenized sequences with shape [B, S]. Got {tuple(seq.shape)}.")
    return seq


def _resolve_mapping_key(mapping, key):
    """AG:2197-2219: find `key` (enum member or string, any case) among a mapping's keys."""
    if key in mapping:
        return key
    name = key.name if hasattr(key, "name") else (key if isinstance(key, str) else None)
    if name is None:
        return None
    for cand in (name, name.lower(), name.upper()):
        if cand in mapping:
            return cand
    for cand in mapping:
        if hasattr(cand, "name") and cand.name == name.upper():
            return cand
    return None


def _mask_tracks(prediction, mask: np.ndarray):
    if prediction is None or mask is None:
        return prediction
    if isinstance(prediction, dict) and "predictions" in prediction:      # junctions: [.., 2 * tissues]
        p = dict(prediction)
        if torch.is_tensor(p.get("predictions")):
            p["predictions"] = p["predictions"][..., torch.from_numpy(mask).to(p["predictions"].device, torch.bool).repeat(2)]
        elif p.get("predictions") is not None:
            p["predictions"] = p["predictions"][..., np.tile(mask, 2)]
        return p
    if torch.is_tensor(prediction):
        return prediction[..., torch.from_numpy(mask).to(prediction.device, torch.bool)]
    return prediction[..., mask]


def _reverse_complement(prediction, reindex):
    if prediction is None or reindex is None or (isinstance(prediction, dict) and "predictions" in prediction):
        return prediction
    if torch.is_tensor(prediction):
        rc = prediction.flip(1) if prediction.dim() == 3 else prediction.flip(1).flip(2) if prediction.dim() == 4 else prediction
        return rc.index_select(-1, torch.as_tensor(reindex, device=rc.device, dtype=torch.long))
    rc = prediction[:, ::-1, :] if prediction.ndim == 3 else prediction[:, ::-1, ::-1, :] if prediction.ndim == 4 else prediction
    return rc[..., reindex]


def ref_alt_outputs(model, seq_ref, seq_alt, organism_index, *, requested_heads=None, track_masks=None, **head_kwargs):
    """Predictions for the reference and the alternate allele (AG:2091-2104).  With equal shapes and frozen norms both
    alleles share ONE batched forward; updating BatchRMSNorm statistics depend on the batch, so that case keeps the
    reference's two sequential calls."""
    seq_ref, seq_alt = _normalize_seq(seq_ref), _normalize_seq(seq_alt)
    updating = any(type(m).__name__ == "BatchRMSNorm" and (m.training if getattr(m, "update_running_var", None) is None else m.update_running_var)
                   for m in model.modules())
    tensor_kw = all(torch.is_tensor(v) or v is None for v in head_kwargs.values())
    if seq_ref.shape != seq_alt.shape or updating or not tensor_kw:
        kw = dict(requested_heads=requested_heads, track_masks=track_masks, **head_kwargs)
        return model.inference(seq_ref, organism_index, **kw), model.inference(seq_alt, organism_index, **kw)
    B = seq_ref.shape[0]
    org = organism_index if isinstance(organism_index, int) else torch.cat((organism_index, organism_index))
    kw2 = {k: (None if v is None else torch.cat((v, v))) for k, v in head_kwargs.items()}
    both = model.inference(torch.cat((seq_ref, seq_alt)), org, requested_heads=requested_heads, track_masks=track_masks, **kw2)
    halves = ({}, {})
    for organism, heads in both.items():
        for name, pred in heads.items():
            r, a = _split_batch(pred, 2)
            halves[0].setdefault(organism, {})[name] = r
            halves[1].setdefault(organism, {})[name] = a
    assert all(torch.is_tensor(v) is False or v.shape[0] == B for h in halves[0].values() for v in h.values())
    return halves


def run_score_variant(model, seq_ref, seq_alt, organism_index, scorer, settings, variant, interval, track_metadata, organism=None,
                      requested_heads=None, track_masks=None, **head_kwargs):
    """AlphaGenome.score_variant AG:2003-2346: predictions for both alleles, then the scorer protocol
    (get_masks_and_metadata -> score_variant -> finalize_variant)."""
    ref_out, alt_out = ref_alt_outputs(model, seq_ref, seq_alt, organism_index, requested_heads=requested_heads, track_masks=track_masks, **head_kwargs)
    if isinstance(ref_out, dict) and ref_out and isinstance(next(iter(ref_out.values())), dict):
        if organism is None:
            if len(ref_out) != 1:
                raise ValueError('Multiple organisms found in model output. Pass organism="..." to select the correct outputs.')
            organism = next(iter(ref_out.keys()))
        if organism not in ref_out:
            raise KeyError(f"Organism {organism!r} not found in model output. Available organisms: {list(ref_out.keys())}")
        ref_out, alt_out = ref_out[organism], alt_out[organism]
    requested_output = getattr(settings, "requested_output", None)
    if requested_output is not None:
        from enum import Enum

        enum_t = type(requested_output) if isinstance(requested_output, Enum) else None

        def to_enum_keys(outputs):
            if enum_t is None or not isinstance(outputs, dict):
                return outputs
            mapped = {}
            for head, pred in outputs.items():
                name = HEAD_TO_OUTPUT_NAME.get(head)
                if name is None or name not in OUTPUT_NAME_TO_RESOLUTION:
                    continue
                try:
                    mapped[enum_t[name]] = select_head_output(pred, name)
                except Exception:  # noqa: BLE001 — an enum without this member
                    continue
            return mapped or outputs

        ref_out, alt_out = to_enum_keys(ref_out), to_enum_keys(alt_out)
        if isinstance(ref_out, dict) and requested_output not in ref_out:
            raise KeyError(f"Requested output {requested_output!r} not found in model predictions. Ensure reference heads are added "
                           "(e.g. add_reference_heads) or provide outputs keyed by the requested output enum.")
    meta_for_masks = track_metadata
    if requested_output is not None and isinstance(ref_out, dict):
        if hasattr(meta_for_masks, "padding"):            # drop padding tracks from predictions and metadata (AG:2262-2287)
            pkey = _resolve_mapping_key(meta_for_masks.padding, requested_output)
            if pkey is not None:
                mask = ~np.asarray(meta_for_masks.padding[pkey])
                okey = _resolve_mapping_key(ref_out, requested_output) or requested_output
                for outs in (ref_out, alt_out):
                    if okey in outs:
                        outs[okey] = _mask_tracks(outs[okey], mask)
                if hasattr(track_metadata, "get"):
                    meta = track_metadata.get(pkey)
                    if meta is not None:
                        try:
                            meta = meta[mask]
                        except Exception:  # noqa: BLE001
                            pass
                        track_metadata = {okey: meta}
        if getattr(interval, "negative_strand", False) and hasattr(meta_for_masks, "strand_reindexing"):   # AG:2289-2309
            rkey = _resolve_mapping_key(meta_for_masks.strand_reindexing, requested_output)
            if rkey is not None:
                reindex = meta_for_masks.strand_reindexing[rkey]
                okey = _resolve_mapping_key(ref_out, requested_output) or requested_output
                for outs in (ref_out, alt_out):
                    if okey in outs:
                        outs[okey] = _reverse_complement(outs[okey], reindex)
    masks, mask_metadata = scorer.get_masks_and_metadata(interval, variant, settings=settings, track_metadata=track_metadata)
    scores = scorer.score_variant(ref_out, alt_out, masks=masks, settings=settings, variant=variant, interval=interval)
    scores_np = {k: v.cpu().numpy() if torch.is_tensor(v) else v for k, v in scores.items()}
    return scorer.finalize_variant(scores_np, track_metadata=track_metadata, mask_metadata=mask_metadata, settings=settings)
