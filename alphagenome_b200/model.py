"""Host-side mirror of the reference module tree (lucidrains/alphagenome, alphagenome_pytorch/alphagenome.py = "AG").

These classes carry the parameters under EXACTLY the reference's names, shapes and dtypes (the 559-key
state_dict of a head-less `AlphaGenome()`; tests/golden/state_dict_keys.txt is the contract), keep the
constructor and forward signatures of the classes they replace, and do no arithmetic of their own: every
forward calls the sm_100a kernels in libagb200.so through alphagenome_b200.ops.  There is no CPU path —
calling a forward without a B200 and the built library raises.

    model = AlphaGenome()                      # same ctor signature as AG:1572-1588
    model.load_state_dict(reference_sd)        # strict=True works with a reference checkpoint
    embeds = model(dna, organism_index)        # Embeds(embeds_1bp, embeds_128bp, embeds_pair), AG:1780-1801

`install(reference_model)` swaps the trunk forward of an already constructed reference model instead.
"""
from __future__ import annotations

from collections import defaultdict, namedtuple

import torch
from torch import nn

TransformerUnetOutput = namedtuple("TransformerUnetOutput", ["unet_output", "single_repr", "pairwise_repr"])
Embeds = namedtuple("Embeds", ["embeds_1bp", "embeds_128bp", "embeds_pair"])

class EmbedsWithOperands(Embeds):
    """Embeds (same three fields, unpacks the same) that also carries the bf16 tensor-core operands of the embeddings."""

    def __new__(cls, embeds_1bp, embeds_128bp, embeds_pair, operands=None):
        self = super().__new__(cls, embeds_1bp, embeds_128bp, embeds_pair)
        self.operands = operands or {}
        return self


# AG:57-74 — kept verbatim as data so add_heads(**publication_heads_config[organism]) keeps working
publication_heads_config = {
    "human": dict(organism="human", num_tracks_1bp=3165, num_tracks_128bp=2733, num_tracks_contacts=28,
                  num_splicing_contexts=282, hidden_dim_splice_juncs=512),
    "mouse": dict(organism="mouse", num_tracks_1bp=730, num_tracks_128bp=310, num_tracks_contacts=8,
                  num_splicing_contexts=75, hidden_dim_splice_juncs=512),
}


def geomspace(start: float, stop: float, num: int) -> torch.Tensor:
    """numpy.geomspace semantics in fp32 (AG:139-148)."""
    if num == 1:
        return torch.tensor([float(start)])
    return start * (stop / start) ** torch.linspace(0, 1, num)


# --------------------------------------------------------------------------------------------- parameter holders
class BatchRMSNorm(nn.Module):
    """AG:302-361.  Frozen (update_running_var False, or None in eval mode): a per-channel affine folded into the
    producing kernel's epilogue.  Updating (the reference's default, AG:331-340): the engine re-estimates running_var
    from the batch in this buffer before normalising (ag_bn_stats_fwd / ag_bn_update_fwd / ag_affine_act_fwd)."""

    def __init__(self, dim_feat, channel_first=False, distributed=True, momentum=0.9, eps=1e-5, update_running_var=True):
        super().__init__()
        self.eps = eps
        self.momentum = 1.0 - momentum      # AG:315: the lerp weight of the new batch variance
        self.distributed = distributed
        self.channel_first = channel_first
        self.gamma = nn.Parameter(torch.ones(dim_feat))
        self.beta = nn.Parameter(torch.zeros(dim_feat))
        self.update_running_var = update_running_var
        self.register_buffer("running_var", torch.ones(dim_feat))

    def affine(self):
        return self.gamma / torch.sqrt(self.running_var + self.eps), self.beta


class RMSNormWithBias(nn.Module):
    """AG:393-405."""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))


class _NoParams(nn.Module):
    """Placeholder for parameter-free reference submodules (GELU1702, nn.GELU, ReLU, Dropout, Rearrange) so that
    nn.Sequential indices — and therefore state_dict keys — match the reference."""


class ConvBlock(nn.Module):
    """AG:431-452: net = Sequential(BatchRMSNorm(channel_first), GELU1702, Conv1d(width, padding=width//2))."""

    def __init__(self, dim, width=5, dim_out=None, batch_rmsnorm=True):
        super().__init__()
        assert width % 2 == 1 and batch_rmsnorm
        dim_out = dim if dim_out is None else dim_out
        self.net = nn.Sequential(BatchRMSNorm(dim, channel_first=True), _NoParams(), nn.Conv1d(dim, dim_out, width, padding=width // 2))

    @property
    def norm(self):
        return self.net[0]

    @property
    def conv(self):
        return self.net[2]


class DNAEmbed(nn.Module):
    """AG:1140-1181."""

    def __init__(self, dim, dim_input=4, width=15):
        super().__init__()
        assert width % 2 == 1
        self.dim_input = dim_input
        self.conv = nn.Conv1d(dim_input, dim, width, padding=width // 2)
        self.pointwise = ConvBlock(dim, width=5)


class DownresBlock(nn.Module):
    """AG:454-481."""

    def __init__(self, dim, channels_to_add=128):
        super().__init__()
        self.pad = channels_to_add
        self.conv = ConvBlock(dim, width=5, dim_out=dim + channels_to_add)
        self.conv_out = ConvBlock(dim + channels_to_add, width=5)


class UpresBlock(nn.Module):
    """AG:483-519."""

    def __init__(self, dim, channels_to_remove=128, residual_scale_init=1.0, has_skip=True):
        super().__init__()
        dim_out = dim - channels_to_remove
        self.pad = channels_to_remove
        self.conv = ConvBlock(dim, width=5, dim_out=dim_out)
        self.unet_conv = ConvBlock(dim_out, width=1) if has_skip else None
        self.conv_out = ConvBlock(dim_out, width=5)
        self.residual_scale = nn.Parameter(torch.ones(1) * residual_scale_init)


class Attention(nn.Module):
    """AG:643-786 (parameters only; the arithmetic is ag_gemm_fwd + ag_qkv_post_fwd + ag_attention_fwd)."""

    def __init__(self, dim, dim_head=64, heads=8, kv_heads=1, dim_head_qk=128, dim_head_v=192, dim_pairwise=None,
                 softclamp_value=5.0, use_qk_rmsnorm=False, attn_bias_use_batch_rmsnorm=True):
        super().__init__()
        if kv_heads != 1 or use_qk_rmsnorm or not attn_bias_use_batch_rmsnorm:
            raise NotImplementedError("B200 path implements the default multi-query / LayerNorm / BatchRMSNorm configuration")
        dim_pairwise = dim if dim_pairwise is None else dim_pairwise
        self.heads, self.dim_head_qk, self.dim_head_v = heads, dim_head_qk, dim_head_v
        self.scale = dim_head_qk ** -0.5
        self.softclamp_value = softclamp_value
        self.to_qkv = nn.Linear(dim, dim_head_qk * heads + dim_head_qk + dim_head_v, bias=False)
        self.to_out = nn.Linear(dim_head_v * heads, dim)
        self.q_norm = nn.LayerNorm(dim_head_qk)
        self.k_norm = nn.LayerNorm(dim_head_qk)
        self.v_norm = nn.LayerNorm(dim_head_v)
        self.to_attn_bias = nn.Sequential(BatchRMSNorm(dim_pairwise), _NoParams(), nn.Linear(dim_pairwise, heads, bias=False), _NoParams())


class NormWrapper(nn.Module):
    """AG:612-639."""

    def __init__(self, dim, block, dropout=0.0, sandwich=False, use_batch_rmsnorm=True):
        super().__init__()
        self.block = block
        self.pre_rmsnorm = BatchRMSNorm(dim) if use_batch_rmsnorm else RMSNormWithBias(dim)
        self.post_block_dropout = _NoParams()
        self.post_rmsnorm = (BatchRMSNorm(dim) if use_batch_rmsnorm else RMSNormWithBias(dim)) if sandwich else _NoParams()


def FeedForward(dim, *, dropout=0.0, expansion_factor=2.0):
    """AG:893-906: Sequential(Linear, ReLU, Dropout, Linear) -> keys block.0.*, block.3.*"""
    inner = int(dim * expansion_factor)
    return nn.Sequential(nn.Linear(dim, inner), _NoParams(), _NoParams(), nn.Linear(inner, dim))


class SingleToPairwise(nn.Module):
    """AG:790-856."""

    def __init__(self, dim, pool_size=16, dim_pairwise=128, heads=32, rel_pos_features_dim=None):
        super().__init__()
        self.pool_size, self.heads, self.dim_pairwise = pool_size, heads, dim_pairwise
        rel_pos_features_dim = heads * 2 if rel_pos_features_dim is None else rel_pos_features_dim
        self.norm = RMSNormWithBias(dim)
        self.to_outer_sum = nn.Sequential(_NoParams(), nn.Linear(dim, dim_pairwise * 2, bias=False))
        self.to_qk = nn.Linear(dim, heads * dim_pairwise * 2, bias=False)
        self.qk_to_pairwise = nn.Linear(heads, dim_pairwise)
        self.to_rel_pos_encoding = nn.Linear(rel_pos_features_dim, heads * dim_pairwise)
        self.qk_rel_pos_bias = nn.Parameter(torch.zeros(2, 1, 1, heads, dim_pairwise))


class PairwiseRowAttention(nn.Module):
    """AG:860-889."""

    def __init__(self, dim):
        super().__init__()
        self.scale = dim ** -0.5
        self.to_qk = nn.Linear(dim, dim * 2, bias=False)
        self.to_v = nn.Linear(dim, dim)


class RelativePosFeatures(nn.Module):
    """AG:570-604: data-independent features, rows d = -P..P-1, [|d| < w_c] and sign(d)*[|d| < w_c]."""

    def __init__(self, pool_size=16, num_features=32, max_distance=8192):
        super().__init__()
        self.pool_size, self.num_features, self.max_distance = pool_size, num_features, max_distance
        self.register_buffer("dummy", torch.tensor(0))

    def features(self, P: int) -> torch.Tensor:
        max_len = self.max_distance // self.pool_size
        d = torch.arange(-P, P)
        widths = torch.arange(self.num_features).float() + geomspace(1, max_len - self.num_features + 1, self.num_features + 1)[:-1]
        f = (widths[None, :] > d.abs()[:, None].float()).float()
        return torch.cat((f, d.sign()[:, None].float() * f), dim=-1)


class RotaryEmbedding(nn.Module):
    """AG:534-557."""

    def __init__(self, dim, max_positions=8192):
        super().__init__()
        nf = dim // 2
        self.register_buffer("inv_freq", 1.0 / (torch.arange(nf).float() + geomspace(1, max_positions - nf + 1, nf)))


class TransformerTower(nn.Module):
    """AG:910-1027."""

    def __init__(self, dim, *, depth=9, heads=8, dim_head_qk=128, dim_head_v=192, dropout=0.0, ff_expansion_factor=2.0,
                 polar_pos_emb=False, max_positions=8192, dim_pairwise=128, pairwise_every_num_single_blocks=2,
                 single_to_pairwise_heads=32, pool_size=16, attn_kwargs: dict = dict(), ff_kwargs: dict = dict()):
        super().__init__()
        if polar_pos_emb:
            raise NotImplementedError("polar_pos_emb (PoPE) is non-default and outside the B200 hot path")
        self.dim, self.depth, self.heads = dim, depth, heads
        self.dim_pairwise = dim if dim_pairwise is None else dim_pairwise
        self.pairwise_every = pairwise_every_num_single_blocks
        self.pool_size = pool_size
        self.max_positions = max_positions
        self.rel_pos_features = RelativePosFeatures(pool_size, num_features=single_to_pairwise_heads, max_distance=max_positions)
        self.polar_pos_emb = False
        self.rotary_emb = RotaryEmbedding(dim_head_qk, max_positions=max_positions)
        layers = []
        for li in range(depth):
            attn = NormWrapper(dim, Attention(dim=dim, dim_head_qk=dim_head_qk, dim_head_v=dim_head_v, heads=heads,
                                              dim_pairwise=self.dim_pairwise), dropout=dropout, sandwich=True)
            ff = NormWrapper(dim, FeedForward(dim=dim, expansion_factor=ff_expansion_factor), dropout=dropout, sandwich=True)
            s2p = pattn = pff = None
            if li % self.pairwise_every == 0:
                s2p = SingleToPairwise(dim=dim, dim_pairwise=self.dim_pairwise, heads=single_to_pairwise_heads,
                                       pool_size=pool_size, rel_pos_features_dim=single_to_pairwise_heads * 2)
                pattn = NormWrapper(self.dim_pairwise, PairwiseRowAttention(self.dim_pairwise), dropout=dropout, use_batch_rmsnorm=False)
                pff = NormWrapper(self.dim_pairwise, FeedForward(dim=self.dim_pairwise, expansion_factor=ff_expansion_factor),
                                  dropout=dropout, use_batch_rmsnorm=False)
            layers.append(nn.ModuleList([attn, ff, s2p, pattn, pff]))
        self.layers = nn.ModuleList(layers)


class TransformerUnet(nn.Module):
    """AG:1031-1136. forward(seq, pre_attend_embed) -> TransformerUnetOutput, computed by the fused trunk engine."""

    def __init__(self, dims=(768, 896, 1024, 1152, 1280, 1408, 1536), basepairs=4, dna_embed_width=15, transformer_kwargs: dict = dict()):
        super().__init__()
        assert dna_embed_width % 2 == 1 and len(dims) >= 2
        self.dims = tuple(dims)
        self.basepairs = basepairs
        self.dna_embed = DNAEmbed(dims[0], dim_input=basepairs, width=dna_embed_width)
        self.downs = nn.ModuleList([DownresBlock(a, channels_to_add=b - a) for a, b in zip(dims[:-1], dims[1:])])
        dims_rev = list(reversed(dims))
        input_dims = [dims[-1]] + dims_rev[:-1]
        self.ups = nn.ModuleList([UpresBlock(i, channels_to_remove=i - o, has_skip=True) for i, o in zip(input_dims, dims_rev)])
        self.transformer = TransformerTower(dim=dims[-1], **transformer_kwargs)

    def forward(self, seq, pre_attend_embed=None):
        from .trunk import engine_for

        out = engine_for(self).run_trunk(seq, pre_attend_embed)
        return TransformerUnetOutput(out["unet_output"], out["single_repr"], out["pairwise_repr"])


class OrganismEmbedding(nn.Module):
    """AG:1183-1196."""

    def __init__(self, dim, num_organisms):
        super().__init__()
        self.embed = nn.Embedding(num_organisms, dim)

    def forward(self, organism_index):
        return self.embed(organism_index)


class OutputEmbedding(nn.Module):
    """AG:1198-1235."""

    def __init__(self, input_dim, num_organisms=2, skip_dim=None, use_batch_rmsnorm=True):
        super().__init__()
        self.double_features = nn.Linear(input_dim, 2 * input_dim)
        self.skip_proj = nn.Linear(skip_dim, 2 * input_dim, bias=False) if skip_dim is not None else None
        self.norm = BatchRMSNorm(2 * input_dim)
        self.embed = nn.Embedding(num_organisms, 2 * input_dim)
        self.activation = _NoParams()


class OutputPairEmbedding(nn.Module):
    """AG:1237-1259."""

    def __init__(self, pair_dim=128, num_organisms=2):
        super().__init__()
        self.norm = RMSNormWithBias(pair_dim)
        self.embed = nn.Embedding(num_organisms, pair_dim)
        self.activation = _NoParams()


def set_update_running_var(root: nn.Module, update_running_var):
    """AG:365-371."""
    for mod in root.modules():
        if isinstance(mod, BatchRMSNorm):
            mod.update_running_var = update_running_var


class AlphaGenome(nn.Module):
    """AG:1572-1852 — same constructor, same `forward(seq, organism_index)` -> Embeds, same state_dict keys."""

    def __init__(self, dims=(768, 896, 1024, 1152, 1280, 1408, 1536), basepairs=4, dna_embed_width=15, num_organisms=2,
                 transformer_kwargs: dict = dict()):
        super().__init__()
        self.transformer_unet = TransformerUnet(dims=dims, basepairs=basepairs, dna_embed_width=dna_embed_width,
                                                transformer_kwargs=transformer_kwargs)
        first_dim, last_dim = dims[0], dims[-1]
        self.organism_embed = OrganismEmbedding(last_dim, num_organisms)
        self.outembed_128bp = OutputEmbedding(last_dim, num_organisms)
        self.outembed_1bp = OutputEmbedding(first_dim, num_organisms, skip_dim=2 * last_dim)
        self.outembed_pair = OutputPairEmbedding(self.transformer_unet.transformer.dim_pairwise, num_organisms)
        self.num_organisms = num_organisms
        self.dim_1bp = last_dim
        self.dim_128bp = 2 * last_dim
        self.dim_contacts = self.transformer_unet.transformer.dim_pairwise
        self.heads = nn.ModuleDict()
        self.head_forward_arg_names = defaultdict(dict)
        self.head_forward_arg_maps = defaultdict(dict)
        self.cuda_graphs = False  # see enable_cuda_graphs()

    # ---- prediction heads: same registration API as the reference (AG:1621-1682) ----
    def add_head(self, organism, head_name, head: nn.Module, head_input_kwarg_names=None):
        """AG:1621-1651."""
        from .heads import get_function_arg_names

        if isinstance(head_input_kwarg_names, str):
            head_input_kwarg_names = (head_input_kwarg_names,)
        if organism not in self.heads:
            self.heads[organism] = nn.ModuleDict()
        self.heads[organism][head_name] = head
        head_forward_arg_names = get_function_arg_names(head.forward)
        self.head_forward_arg_names[organism][head_name] = head_forward_arg_names if head_input_kwarg_names is None else head_input_kwarg_names
        arg_map = dict()
        if head_input_kwarg_names is not None:
            arg_map = dict(zip(head_input_kwarg_names, head_forward_arg_names))
        self.head_forward_arg_maps[organism][head_name] = arg_map

    def add_heads(self, organism, num_tracks_1bp, num_tracks_128bp, num_tracks_contacts, num_splicing_contexts,
                  hidden_dim_splice_juncs=512):
        """AG:1653-1682: the publication heads of one organism (`add_heads(**publication_heads_config['human'])`)."""
        from .heads import ContactMapHead, SpliceJunctionHead, SpliceSiteClassifier, SpliceSiteUsage, TracksScaledPrediction

        organism_heads = (
            ("1bp_tracks", TracksScaledPrediction(self.dim_1bp, num_tracks_1bp), ("embeds_1bp",)),
            ("128bp_tracks", TracksScaledPrediction(self.dim_128bp, num_tracks_128bp), ("embeds_128bp",)),
            ("contact_head", ContactMapHead(self.dim_contacts, num_tracks_contacts, self.num_organisms)),
            ("splice_logits", SpliceSiteClassifier(self.dim_1bp)),
            ("splice_usage", SpliceSiteUsage(self.dim_1bp, num_splicing_contexts)),
            ("splice_juncs", SpliceJunctionHead(self.dim_1bp, hidden_dim_splice_juncs, num_splicing_contexts)),
        )
        for args in organism_heads:
            self.add_head(organism, *args)

    def enable_cuda_graphs(self, enabled: bool = True):
        """Replay the forward from a CUDA graph (per input shape).  Outputs are then static buffers that the next
        call with the same shape overwrites — copy them if you need to keep them."""
        self.cuda_graphs = bool(enabled)
        return self

    @property
    def total_parameters(self):
        return sum(p.numel() for p in self.parameters())

    def _organism_tensor(self, seq, organism_index):
        if isinstance(organism_index, int):
            organism_index = torch.full((seq.shape[0],), organism_index, device=seq.device, dtype=torch.long)
        return organism_index

    @torch.no_grad()
    def get_embeds(self, seq, organism_index, return_unet_transformer_outputs=False):
        """AG:1754-1778."""
        from .trunk import engine_for

        organism_index = self._organism_tensor(seq, organism_index)
        eng = engine_for(self.transformer_unet, self)
        out = eng.run_embeds_graphed(seq, organism_index) if self.cuda_graphs else eng.run_embeds(seq, organism_index)
        embeds = Embeds(out["embeds_1bp"], out["embeds_128bp"], out["embeds_pair"])
        if not return_unet_transformer_outputs:
            return embeds
        return embeds, TransformerUnetOutput(out["unet_output"], out["single_repr"], out["pairwise_repr"])

    def add_reference_heads(self, organism, genome_track_heads=None, num_tracks_contacts=None, num_splice_usage_tracks=None,
                            splice_junction_hidden_dim=None, splice_junction_max_tissues=None):
        """AG:1684-1731: the JAX-aligned heads of the official checkpoint (rna_seq, cage, dnase, procap, atac, chip_tf,
        chip_histone, contact_maps, splice_sites_{classification,usage,junction})."""
        from . import heads as H

        genome_track_heads = H.jax_genome_track_heads if genome_track_heads is None else genome_track_heads
        for head_name, spec in genome_track_heads.items():
            self.add_head(organism, head_name, H.GenomeTracksHead(dim_1bp=self.dim_1bp, dim_128bp=self.dim_128bp,
                                                                  num_tracks=spec["num_tracks"], resolutions=spec["resolutions"]))
        self.add_head(organism, "contact_maps", H.ContactMapsHead(self.dim_contacts, H.jax_contact_maps_tracks if num_tracks_contacts is None else num_tracks_contacts))
        self.add_head(organism, "splice_sites_classification", H.SpliceSiteClassifier(self.dim_1bp))
        self.add_head(organism, "splice_sites_usage", H.SpliceSiteUsage(self.dim_1bp, H.jax_splice_usage_tracks if num_splice_usage_tracks is None else num_splice_usage_tracks))
        self.add_head(organism, "splice_sites_junction", H.SpliceSitesJunctionHead(
            self.dim_1bp, H.jax_splice_junction_hidden_dim if splice_junction_hidden_dim is None else splice_junction_hidden_dim,
            H.jax_splice_junction_max_tissues if splice_junction_max_tissues is None else splice_junction_max_tissues))

    def load_from_official_jax_model(self, model_version="all_folds", use_gpu=True, strict=False, params=None, state=None):
        """AG:1733-1752: load the official JAX/Haiku checkpoint.  `params` / `state` (flat or nested Haiku trees) may be
        handed in directly; otherwise they are fetched with alphagenome_research's loader like the reference does
        (needs jax + alphagenome_research + network — none of which this package depends on)."""
        from .convert import convert_checkpoint, flatten

        if params is None:
            try:
                import jax
                from alphagenome_research.model.dna_model import create_from_huggingface
            except ImportError as e:
                print(f"Error: Missing required dependency for conversion. Details: {e}")
                return
            device = jax.devices("gpu" if use_gpu else "cpu")[0]
            jm = create_from_huggingface(model_version=model_version, device=device)
            params, state = flatten(jm._params), flatten(jm._state)
        self.load_state_dict(convert_checkpoint(params, state or {}), strict=strict)
        return self

    def _heads(self):
        from .heads import HeadsEngine

        if getattr(self, "_heads_engine", None) is None:
            object.__setattr__(self, "_heads_engine", HeadsEngine(self))
        return self._heads_engine

    @torch.no_grad()
    def forward(self, seq, organism_index, return_embeds=False, **head_kwargs):
        """AG:1780-1852: Embeds when no heads are registered (or return_embeds), else {organism: {head: prediction}}
        with every organism's heads run on every sample."""
        if len(self.heads) == 0 or return_embeds:
            return self.get_embeds(seq, organism_index)
        from .trunk import engine_for

        organism_index = self._organism_tensor(seq, organism_index).to(seq.device).long()
        eng = engine_for(self.transformer_unet, self)

        def trunk_and_heads(seq, org, **kw):
            out = eng.run_embeds(seq, org, want_head_operands=True)
            plan = eng.sp.plan(seq.shape[1]) if eng.sp is not None else None
            return self._heads().run(out, org, kw, sp=eng.sp, plan=plan)

        if self.cuda_graphs and all(torch.is_tensor(v) for v in head_kwargs.values()):
            # trunk + every registered head as ONE graph replay (static inputs: tokens, organism index, head kwargs)
            self._heads().pack(seq.device)
            kw = {k: v.to(seq.device) for k, v in head_kwargs.items()}
            return eng.run_graphed(("forward+heads", self._heads()._version_key(seq.device)), lambda seq, org, **kw: trunk_and_heads(seq, org, **kw),
                                   dict(seq=seq.to(torch.int64), org=organism_index, **kw))
        return trunk_and_heads(seq, organism_index, **head_kwargs)

    @torch.no_grad()
    def inference(self, seq, organism_index, *, requested_heads=None, track_masks=None, embeds=None, return_embeds=False,
                  **head_kwargs):
        """AG:1853-1965: selective head execution (`requested_heads`), per-head track masks, and reuse of embeddings
        computed earlier with get_embeds() (`embeds=`; their bf16 MMA operands are then made by rounding)."""
        from .trunk import engine_for

        organism_index = self._organism_tensor(seq, organism_index).to(seq.device).long()
        eng = engine_for(self.transformer_unet, self)
        if embeds is None:
            want = len(self.heads) != 0
            if self.cuda_graphs:
                out = eng.run_graphed(("embeds", want), lambda seq, org: eng.run_embeds(seq, org, want_head_operands=want),
                                      dict(seq=seq.to(torch.int64), org=organism_index))
            else:
                out = eng.run_embeds(seq, organism_index, want_head_operands=want)
            # the bf16 head operands the trunk's last epilogues wrote travel with the embeddings, so handing them back
            # (embeds=...) reproduces this call's head outputs bit for bit
            embeds = EmbedsWithOperands(out["embeds_1bp"], out["embeds_128bp"], out["embeds_pair"],
                                        operands={k: out[k] for k in ("embeds_1bp_b16", "embeds_128bp_b16") if k in out})
        else:
            out = dict(embeds_1bp=embeds[0], embeds_128bp=embeds[1], embeds_pair=embeds[2])
            out.update(getattr(embeds, "operands", None) or {})
        if len(self.heads) == 0:
            return embeds if return_embeds else {}
        plan = eng.sp.plan(seq.shape[1]) if eng.sp is not None else None
        res = self._heads().run(out, organism_index, head_kwargs, sp=eng.sp, plan=plan, requested_heads=requested_heads,
                                track_masks=track_masks)
        res = {o: h for o, h in res.items() if h}                                   # AG:1963
        return (res, embeds) if return_embeds else res

    @torch.no_grad()
    def score_variant(self, seq_ref, seq_alt, organism_index, scorer, settings, variant, interval, track_metadata, organism=None,
                      requested_heads=None, track_masks=None, **head_kwargs):
        """AG:2003-2346: both alleles through the model (one batched forward), then scorer.get_masks_and_metadata ->
        scorer.score_variant -> scorer.finalize_variant."""
        from .dna_model import run_score_variant

        return run_score_variant(self, seq_ref, seq_alt, organism_index, scorer, settings, variant, interval, track_metadata,
                                 organism=organism, requested_heads=requested_heads, track_masks=track_masks, **head_kwargs)

    def _apply_track_mask(self, head_output, mask):
        from .heads import apply_track_mask

        return apply_track_mask(head_output, mask)
