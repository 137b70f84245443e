"""Prediction heads on the B200: the modules `AlphaGenome.add_heads` registers in the reference
(AG:1653-1682) — parameter holders with the reference's names — and the engine that runs them on libagb200.

    TracksScaledPrediction  AG:1378-1393   softplus(Linear(x)) * softplus(scale)          -> ag_gemm_fwd, softplus epilogue
    ContactMapHead          AG:1273-1299   symmetrize, RMSNorm, + organism, GELU, Linear  -> ag_rmsnorm_act_fwd + ag_gemm_fwd
    SpliceSiteClassifier    AG:1301-1310   Linear(1536 -> 5)                              -> ag_gemm_fwd
    SpliceSiteUsage         AG:1312-1321   Linear(...).sigmoid()                          -> ag_gemm_fwd, sigmoid epilogue
    SpliceJunctionHead      AG:1322-1376   project, gather sites, per-context scale/offset + rotary, donor x acceptor
                                           einsum, softplus        -> ag_gemm_fwd (gathered rows) + ag_splice_rope_fwd +
                                                                      batched ag_gemm_fwd with the softplus epilogue

As in the reference, EVERY organism's heads run on every sample (AG:1823-1850).  Outputs whose channel count is not a
multiple of 16 are computed into a 16-padded buffer and returned as the `[..., :N]` view (same values, row stride
padded).  A head that is not one of these classes (add_head with a user module) is simply called, like the reference does.
"""
from __future__ import annotations

import inspect
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .ops import ACT_GELU1702, ACT_NONE, ACT_SIGMOID, ACT_SOFTPLUS_MUL, Act


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _gemm_n_tile(n: int) -> int:
    """The N tile ag_gemm_fwd picks for an output width n (multiple of 16): the largest multiple of 16 <= 256 dividing n."""
    for t in range(256, 15, -16):
        if n % t == 0:
            return t
    return 0


def _padded_width(n: int) -> int:
    """Output width the head GEMM is run at: n rounded up to 16, or further to a multiple of 64 / 128 / 256 when that buys
    a wider MMA tile (e.g. 730 tracks: 736 = 23 x 32 would run 32-wide tiles, 768 runs 256-wide ones; 3165: 3168 =
    18 x 176 vs 3328 = 13 x 256).  Cost model (operand ingest): a t-wide tile of a CTA pair needs
    (16384 + 64 t) bytes per 2 t tensor cycles and levels off at ~53 B/clk of L2 -> smem ingest; 256-wide tiles run at
    ~0.77 of the full tensor rate."""
    def eff(t):
        return min(1.0, 53.0 / (0.77 * (16384 + 64 * t) / (2 * t)))

    best = _round_up(n, 16)
    best_cost = best / eff(_gemm_n_tile(best))
    for m in (64, 128, 256):
        cand = _round_up(n, m)
        cost = cand / eff(_gemm_n_tile(cand))
        if cost < 0.97 * best_cost:          # only pad further for a clear win (the padding also costs output memory)
            best, best_cost = cand, cost
    return best


# --------------------------------------------------------------------------------------------- parameter holders
class TracksScaledPrediction(nn.Module):
    """AG:1378-1393."""

    def __init__(self, dim, num_tracks):
        super().__init__()
        self.to_pred = nn.Linear(dim, num_tracks)
        self.scale = nn.Parameter(torch.zeros(num_tracks))

    def forward(self, x):  # the argument name selects the input embedding (AG:1640)
        raise RuntimeError("run through AlphaGenome.forward (alphagenome_b200.heads.run_heads)")


class ContactMapHead(nn.Module):
    """AG:1273-1299."""

    def __init__(self, input_dim, num_tracks, num_organisms):
        super().__init__()
        from .model import RMSNormWithBias, _NoParams

        self.norm = RMSNormWithBias(input_dim)
        self.embed = nn.Embedding(num_organisms, input_dim)
        self.gelu = _NoParams()
        self.proj = nn.Linear(input_dim, num_tracks)

    def forward(self, embeds_pair, organism_index):
        raise RuntimeError("run through AlphaGenome.forward (alphagenome_b200.heads.run_heads)")


class SpliceSiteClassifier(nn.Module):
    """AG:1301-1310."""

    def __init__(self, input_dim):
        super().__init__()
        self.linear = nn.Linear(input_dim, 5)

    def forward(self, embeds_1bp):
        raise RuntimeError("run through AlphaGenome.forward (alphagenome_b200.heads.run_heads)")


class SpliceSiteUsage(nn.Module):
    """AG:1312-1321."""

    def __init__(self, input_dim, n_contexts):
        super().__init__()
        self.linear = nn.Linear(input_dim, n_contexts)

    def forward(self, embeds_1bp):
        raise RuntimeError("run through AlphaGenome.forward (alphagenome_b200.heads.run_heads)")


class _Rope(nn.Module):
    """RotaryEmbedding AG:534-557 (only its inv_freq buffer is needed here)."""

    def __init__(self, dim, max_positions=8192):
        super().__init__()
        from .model import geomspace

        nf = dim // 2
        self.register_buffer("inv_freq", 1.0 / (torch.arange(nf).float() + geomspace(1, max_positions - nf + 1, nf)))


class SpliceJunctionHead(nn.Module):
    """AG:1322-1376."""

    def __init__(self, input_dim, hidden_dim, n_contexts, rope_max_position=2 ** 20):
        super().__init__()
        self.project = nn.Linear(input_dim, hidden_dim)
        self.scale = nn.Parameter(torch.ones(n_contexts, hidden_dim))
        self.offset = nn.Parameter(torch.zeros(n_contexts, hidden_dim))
        self.rope = _Rope(hidden_dim, max_positions=rope_max_position)
        self.n_contexts = n_contexts

    def forward(self, embeds_1bp, splice_donor_idx, splice_acceptor_idx):
        raise RuntimeError("run through AlphaGenome.forward (alphagenome_b200.heads.run_heads)")


class GenomeTracksHead(nn.Module):
    """AG:1395-1423 (JAX-style genome tracks head with per-resolution submodules)."""

    def __init__(self, dim_1bp, dim_128bp, num_tracks, resolutions=(1, 128)):
        super().__init__()
        self.resolutions = nn.ModuleDict()
        if 1 in resolutions:
            self.resolutions["resolution_1"] = TracksScaledPrediction(dim_1bp, num_tracks)
        if 128 in resolutions:
            self.resolutions["resolution_128"] = TracksScaledPrediction(dim_128bp, num_tracks)

    def forward(self, embeds_1bp=None, embeds_128bp=None):
        raise RuntimeError("run through AlphaGenome.forward (alphagenome_b200.heads.run_heads)")


class ContactMapsHead(nn.Module):
    """AG:1425-1442 (symmetrize + Linear)."""

    def __init__(self, input_dim, num_tracks):
        super().__init__()
        self.to_pred = nn.Linear(input_dim, num_tracks)

    def forward(self, embeds_pair, organism_index=None):
        raise RuntimeError("run through AlphaGenome.forward (alphagenome_b200.heads.run_heads)")


class SpliceSitesJunctionHead(nn.Module):
    """AG:1444-1545 (JAX-style splice junction head: per-site rotary embeddings, positive / negative strand)."""

    def __init__(self, input_dim, hidden_dim, max_num_tissues, rope_max_position=2 ** 20):
        super().__init__()
        self.project = nn.Linear(input_dim, hidden_dim)
        self.hidden_dim, self.max_num_tissues = hidden_dim, max_num_tissues
        n = 2 * max_num_tissues * hidden_dim
        self.pos_donor_embeddings = nn.Parameter(torch.zeros(n))
        self.pos_acceptor_embeddings = nn.Parameter(torch.zeros(n))
        self.neg_donor_embeddings = nn.Parameter(torch.zeros(n))
        self.neg_acceptor_embeddings = nn.Parameter(torch.zeros(n))
        self.rope = _Rope(hidden_dim, max_positions=rope_max_position)

    def forward(self, embeds_1bp, splice_site_positions=None):
        raise RuntimeError("run through AlphaGenome.forward (alphagenome_b200.heads.run_heads)")


# AG:77-90 — the head specs of the official (JAX) checkpoint
jax_genome_track_heads = {
    "rna_seq": dict(num_tracks=768, resolutions=(1, 128)), "cage": dict(num_tracks=640, resolutions=(1, 128)),
    "dnase": dict(num_tracks=384, resolutions=(1, 128)), "procap": dict(num_tracks=128, resolutions=(1, 128)),
    "atac": dict(num_tracks=256, resolutions=(1, 128)), "chip_tf": dict(num_tracks=1664, resolutions=(128,)),
    "chip_histone": dict(num_tracks=1152, resolutions=(128,)),
}
jax_contact_maps_tracks = 28
jax_splice_usage_tracks = 734
jax_splice_junction_hidden_dim = 768
jax_splice_junction_max_tissues = 367


def get_function_arg_names(fn):
    """AG:1262-1265."""
    return [p.name for p in inspect.signature(fn).parameters.values()]


# --------------------------------------------------------------------------------------------- engine
class HeadsEngine:
    """Packs the head weights once per parameter version (bf16, rows zero-padded to a multiple of 16) and runs them."""

    def __init__(self, model):
        self.model = model
        self._pack: Optional[Dict] = None
        self._key = None

    def _version_key(self, device):
        return (str(device),) + tuple((p.data_ptr(), p._version) for p in self.model.heads.parameters())

    @staticmethod
    def _lin(lin: nn.Linear, device):
        """Linear -> (bf16 [1, N, K], fp32 bias padded to a multiple of 16, N, N16)."""
        N, K = lin.weight.shape
        N16 = _padded_width(N)
        W = lin.weight.detach().to(device).to(torch.bfloat16).contiguous().unsqueeze(0)
        b = torch.zeros(N16, device=device, dtype=torch.float32)
        if lin.bias is not None:
            b[:N] = lin.bias.detach().to(device).float()
        return dict(W=W, b=b, N=N, N16=N16)

    def _tracks_pack(self, head, device):
        e = self._lin(head.to_pred, device)
        mul = torch.ones(e["N16"], device=device, dtype=torch.float32)
        mul[: e["N"]] = F.softplus(head.scale.detach().to(device).float())      # AG:1393
        e["mul"] = mul
        return e

    def pack(self, device):
        key = self._version_key(device)
        if self._pack is not None and self._key == key:
            return self._pack
        P: Dict = {}
        for organism, heads in self.model.heads.items():
            for name, head in heads.items():
                e: Dict = dict(kind=type(head).__name__)
                if isinstance(head, TracksScaledPrediction):
                    e.update(self._tracks_pack(head, device))
                elif isinstance(head, GenomeTracksHead):
                    e["res"] = {k: self._tracks_pack(h, device) for k, h in head.resolutions.items()}
                elif isinstance(head, ContactMapsHead):
                    e.update(self._lin(head.to_pred, device))
                elif isinstance(head, SpliceSitesJunctionHead):
                    e.update(self._lin(head.project, device))
                    T, Hd = head.max_num_tissues, head.hidden_dim
                    e.update(T=T, Hd=Hd, inv_freq=head.rope.inv_freq.detach().to(device).float().contiguous())
                    for nm in ("pos_donor", "pos_acceptor", "neg_donor", "neg_acceptor"):
                        par = getattr(head, nm + "_embeddings").detach().to(device).float().view(2, T, Hd)     # AG:1489-1491
                        e[nm] = (par[0].contiguous(), par[1].contiguous())
                elif isinstance(head, ContactMapHead):
                    e.update(self._lin(head.proj, device))
                    e.update(nw=head.norm.weight.detach().to(device).float().contiguous(), nb=head.norm.bias.detach().to(device).float().contiguous(),
                             emb=head.embed.weight.detach().to(device).float().contiguous())
                elif isinstance(head, (SpliceSiteClassifier, SpliceSiteUsage)):
                    e.update(self._lin(head.linear, device))
                elif isinstance(head, SpliceJunctionHead):
                    e.update(self._lin(head.project, device))
                    T, Hd = head.scale.shape
                    ang = torch.arange(T, dtype=torch.float32)[:, None] * head.rope.inv_freq.detach().float().cpu()[None, :]   # AG:1356, 555-556
                    e.update(scale=head.scale.detach().to(device).float().contiguous(), offset=head.offset.detach().to(device).float().contiguous(),
                             cos=ang.cos().to(device).contiguous(), sin=ang.sin().to(device).contiguous(), T=T, Hd=Hd)
                P[(organism, name)] = e
        self._pack, self._key = P, key
        return P

    # ---------------------------------------------------------------------------------------- head runners
    @staticmethod
    def _gemm_out(A_b16, e, act=ACT_NONE, mul=None):
        """[B, rows, K] bf16 -> fp32 view [B, rows, N] of a 16-padded buffer; bias via pre_shift, act in the epilogue."""
        B, rows, _ = A_b16.shape
        buf = torch.empty(B, rows, e["N16"], device=A_b16.device, dtype=torch.float32)
        if act == ACT_NONE and mul is None:
            ops.gemm(A_b16, e["W"], n_pad=e["N16"], pre_shift=e["b"], out=buf)
        else:
            ops.gemm(A_b16, e["W"], n_pad=e["N16"], pre_shift=e["b"], actA=Act(buf, mul, None, act))
        return buf[..., : e["N"]]

    def tracks(self, e, x_b16):
        return self._gemm_out(x_b16, e, ACT_SOFTPLUS_MUL, e["mul"])

    def contact(self, e, embeds_pair, organism_index):
        # symmetrize(embeds_pair) (AG:1292) is the identity: OutputPairEmbedding already made it symmetric bit for bit
        B, rows, Pn, C = embeds_pair.shape
        xb = torch.empty(B, rows * Pn, C, device=embeds_pair.device, dtype=torch.bfloat16)
        add = e["emb"].index_select(0, organism_index).contiguous()
        ops.rmsnorm_act(embeds_pair, e["nw"], e["nb"], xb, add=add, rows_per_batch=rows * Pn, act=ACT_GELU1702)
        return self._gemm_out(xb, e).view(B, rows, Pn, e["N"])

    def splice_juncs(self, e, e1_b16, donor_idx, acceptor_idx, sp=None, plan=None):
        dev = e1_b16.device
        B = e1_b16.shape[0]
        T, Hd = e["T"], e["Hd"]

        def site_rows(idx):
            # project(embeds_1bp)[b, idx] == project(embeds_1bp[b, idx]): gather the P site rows first (AG:1349, 1370)
            idx = idx.to(dev).long()
            if sp is not None and plan is not None and plan.world > 1:
                own = (idx >= plan.start) & (idx < plan.end)
                loc = (idx - plan.start).clamp(0, plan.L_loc - 1)
            else:
                own, loc = None, idx
            rows = torch.gather(e1_b16, 1, loc[:, :, None].expand(-1, -1, e1_b16.shape[2])).contiguous()
            x = torch.empty(B, idx.shape[1], e["N16"], device=dev, dtype=torch.float32)
            ops.gemm(rows, e["W"], n_pad=e["N16"], pre_shift=e["b"], out=x)
            if own is not None:   # each site row lives on exactly one rank: mask and sum over ranks
                x = sp.all_reduce_sum(x * own[:, :, None].to(x.dtype))
            return x

        def rope(x):
            P = x.shape[1]
            out = torch.zeros(B, T, _round_up(P, 16), Hd, device=dev, dtype=torch.bfloat16)
            ops.splice_rope(x, e["scale"], e["offset"], e["cos"], e["sin"], out)
            return out, P

        d, Pd = rope(site_rows(donor_idx))
        a, Pa = rope(site_rows(acceptor_idx))
        Pa16 = a.shape[2]
        scores = torch.empty(B * T, Pd, Pa16, device=dev, dtype=torch.float32)
        # per (batch, context): donors [Pd, hidden] x acceptors [Pa, hidden]^T, softplus in the epilogue (AG:1375-1376)
        ops.gemm(d.view(B * T, d.shape[2], Hd), a.view(B * T, Pa16, Hd), M=Pd, w_batched=True, actA=Act(scores, None, None, ACT_SOFTPLUS_MUL))
        return scores.view(B, T, Pd, Pa16)[..., :Pa].permute(0, 2, 3, 1)   # 'b d a t'

    def splice_sites_junction(self, e, e1_b16, positions, sp=None, plan=None):
        """SpliceSitesJunctionHead.forward AG:1501-1545."""
        if positions is None:
            return None
        dev = e1_b16.device
        B = e1_b16.shape[0]
        T, Hd = e["T"], e["Hd"]
        positions = positions.to(dev).long()
        sharded = sp is not None and plan is not None and plan.world > 1

        def site_embed(idx, par):
            safe = idx.clamp_min(0)                                                      # AG:1485
            if sharded:
                own = (safe >= plan.start) & (safe < plan.end)
                loc = (safe - plan.start).clamp(0, plan.L_loc - 1)
            else:
                own, loc = None, safe
            rows = torch.gather(e1_b16, 1, loc[:, :, None].expand(-1, -1, e1_b16.shape[2])).contiguous()
            x = torch.empty(B, idx.shape[1], e["N16"], device=dev, dtype=torch.float32)
            ops.gemm(rows, e["W"], n_pad=e["N16"], pre_shift=e["b"], out=x)
            if own is not None:
                x = sp.all_reduce_sum(x * own[:, :, None].to(x.dtype))
            ang = safe.to(torch.float32)[:, :, None] * e["inv_freq"][None, None, :]       # rope(safe_indices), AG:1495
            P = idx.shape[1]
            out = torch.zeros(B, T, _round_up(P, 16), Hd, device=dev, dtype=torch.bfloat16)
            ops.splice_rope(x[..., :Hd].contiguous() if e["N16"] != Hd else x, par[0], par[1], ang.cos().contiguous(), ang.sin().contiguous(), out)
            return out, P

        def counts(didx, aidx, dpar, apar):
            d, Pd = site_embed(didx, dpar)
            a, Pa = site_embed(aidx, apar)
            Pa16 = a.shape[2]
            scores = torch.empty(B * T, Pd, Pa16, device=dev, dtype=torch.float32)
            ops.gemm(d.view(B * T, d.shape[2], Hd), a.view(B * T, Pa16, Hd), M=Pd, w_batched=True, actA=Act(scores, None, None, ACT_SOFTPLUS_MUL))
            c = scores.view(B, T, Pd, Pa16)[..., :Pa].permute(0, 2, 3, 1)                # 'b d a t'
            mask = (didx >= 0)[:, :, None] & (aidx >= 0)[:, None, :]                     # AG:1523-1524
            return c * mask[..., None], mask

        pd, pa, nd, na = (positions[:, i] for i in range(4))
        pos_c, pos_m = counts(pd, pa, e["pos_donor"], e["pos_acceptor"])
        neg_c, neg_m = counts(nd, na, e["neg_donor"], e["neg_acceptor"])
        pred = torch.cat((pos_c, neg_c), dim=-1)
        mask = torch.cat((pos_m[..., None].expand(*pos_m.shape, T), neg_m[..., None].expand(*neg_m.shape, T)), dim=-1)
        return dict(predictions=pred, splice_site_positions=positions, splice_junction_mask=mask)

    # ---------------------------------------------------------------------------------------- forward
    @torch.no_grad()
    def run(self, out: Dict, organism_index, head_kwargs: Dict, sp=None, plan=None, requested_heads=None, track_masks=None):
        """out: the engine's dict (embeddings fp32 + bf16 copies `embeds_1bp_b16` / `embeds_128bp_b16`; a missing bf16
        copy — embeddings handed back by the caller, AG:1911 — is made by rounding the fp32 tensor)."""
        model = self.model
        dev = out["embeds_1bp"].device
        P = self.pack(dev)
        head_inputs = {k: out[k] for k in ("embeds_1bp", "embeds_128bp", "embeds_pair", "unet_output", "single_repr", "pairwise_repr") if k in out}
        head_inputs["organism_index"] = organism_index
        assert not set(head_inputs) & set(head_kwargs)                      # AG:1818
        head_inputs.update(head_kwargs)
        head_inputs.setdefault("splice_site_positions", None)
        b16 = {}

        def operand(name):
            if name not in b16:
                t = out.get(name + "_b16")
                b16[name] = t if t is not None else out[name].to(torch.bfloat16)
            return b16[name]

        result = {}
        for organism, heads in model.heads.items():
            organism_out = {}
            for name, head in heads.items():
                if requested_heads is not None and name not in requested_heads:      # AG:1941
                    continue
                arg_names = model.head_forward_arg_names[organism][name]
                arg_map = model.head_forward_arg_maps[organism][name]
                kwargs = {arg_map.get(k, k): head_inputs[k] for k in arg_names}    # AG:1838-1842
                e = P[(organism, name)]
                if isinstance(head, TracksScaledPrediction):
                    (src,) = arg_names
                    if src not in ("embeds_1bp", "embeds_128bp"):
                        raise ops._lib.AgbError(f"TracksScaledPrediction on {src!r}: only embeds_1bp / embeds_128bp inputs are supported")
                    pred = self.tracks(e, operand(src))
                elif isinstance(head, GenomeTracksHead):                              # AG:1412-1423
                    pred = {}
                    if "resolution_1" in e["res"]:
                        pred["scaled_predictions_1bp"] = self.tracks(e["res"]["resolution_1"], operand("embeds_1bp"))
                    if "resolution_128" in e["res"]:
                        pred["scaled_predictions_128bp"] = self.tracks(e["res"]["resolution_128"], operand("embeds_128bp"))
                elif isinstance(head, ContactMapsHead):
                    # symmetrize (AG:1440) is the identity on embeds_pair; Linear on its bf16 rounding
                    ep = kwargs["embeds_pair"]
                    Bq, rows, Pn, Cq = ep.shape
                    pred = self._gemm_out(ep.to(torch.bfloat16).view(Bq, rows * Pn, Cq), e).view(Bq, rows, Pn, e["N"])
                elif isinstance(head, SpliceSitesJunctionHead):
                    pred = self.splice_sites_junction(e, operand("embeds_1bp"), kwargs.get("splice_site_positions"), sp, plan)
                elif isinstance(head, ContactMapHead):
                    pred = self.contact(e, kwargs["embeds_pair"], organism_index)
                elif isinstance(head, SpliceSiteClassifier):
                    pred = self._gemm_out(operand("embeds_1bp"), e)
                elif isinstance(head, SpliceSiteUsage):
                    pred = self._gemm_out(operand("embeds_1bp"), e, ACT_SIGMOID)
                elif isinstance(head, SpliceJunctionHead):
                    pred = self.splice_juncs(e, operand("embeds_1bp"), kwargs["splice_donor_idx"], kwargs["splice_acceptor_idx"], sp, plan)
                else:
                    pred = head(**kwargs)   # a user module: called like the reference does (AG:1846)
                if track_masks and name in track_masks:                              # AG:1958-1959
                    pred = apply_track_mask(pred, track_masks[name])
                organism_out[name] = pred
            if organism_out or requested_heads is None:
                result[organism] = organism_out
        return result


def apply_track_mask(head_output, mask):
    """AlphaGenome._apply_track_mask AG:1967-1987."""
    if torch.is_tensor(head_output):
        return head_output[..., mask]
    if isinstance(head_output, dict):
        return {k: (v[..., mask] if torch.is_tensor(v) and v.dim() >= 2 else v) for k, v in head_output.items()}
    return head_output
