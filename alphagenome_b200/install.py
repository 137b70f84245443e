"""`install(reference_model)`: route an already constructed lucidrains/alphagenome `AlphaGenome` through libagb200.

The engine only reads attributes that the reference's own classes already have (same names: AG:1054-1093,
1602-1606), plus three tiny accessors that are added here without touching the reference source:
`BatchRMSNorm.affine()` (the eval-mode per-channel affine, AG:344-361), `ConvBlock.norm/.conv` (AG:445-449),
and the channel list `unet.dims`.  Everything above `get_embeds` (forward, inference, heads, DNAModel) is untouched.
"""
from __future__ import annotations

import types
from collections import namedtuple

import torch

from .trunk import engine_for


def _affine(self):
    return self.gamma / torch.sqrt(self.running_var + self.eps), self.beta


def _adapt(unet) -> None:
    for mod in unet.modules():
        name = type(mod).__name__
        cls = type(mod)
        if name == "BatchRMSNorm" and not hasattr(cls, "affine"):
            cls.affine = _affine
        if name == "ConvBlock" and not hasattr(cls, "norm"):
            cls.norm = property(lambda s: s.net[0])
            cls.conv = property(lambda s: s.net[2])
    # the reference keeps `heads`, `pool_size`, ... as constructor locals in a few places: recover them from shapes
    for layer in unet.transformer.layers:
        attn = layer[0].block
        if not hasattr(attn, "heads"):
            attn.heads = attn.to_attn_bias[2].weight.shape[0]
            attn.dim_head_qk = attn.q_norm.weight.numel()
            attn.dim_head_v = attn.v_norm.weight.numel()
        s2p = layer[2]
        if s2p is not None and not hasattr(s2p, "heads"):
            s2p.heads = s2p.qk_to_pairwise.weight.shape[1]
            s2p.dim_pairwise = s2p.qk_to_pairwise.weight.shape[0]
            s2p.pool_size = getattr(unet.transformer.rel_pos_features, "pool_size", 16)
    tw = unet.transformer
    if not hasattr(tw, "max_positions"):
        tw.max_positions = tw.rel_pos_features.max_distance
    if not hasattr(tw.rel_pos_features, "features"):
        from .model import RelativePosFeatures

        tw.rel_pos_features.features = types.MethodType(RelativePosFeatures.features, tw.rel_pos_features)
    if not hasattr(unet, "dims"):
        chans = [unet.dna_embed.conv.weight.shape[0]] + [d.conv_out.net[2].weight.shape[0] for d in unet.downs]
        unet.dims = tuple(chans)


def install(model):
    """Swap `model.get_embeds` (AG:1754-1778) for the B200 engine.  `model` is a reference `AlphaGenome` (or the
    mirror in alphagenome_b200.model).  BatchRMSNorm behaves as the modules say: `update_running_var` true (the
    reference's default) re-estimates the variance on every forward in the module's own buffer, false is the frozen path."""
    unet = model.transformer_unet
    _adapt(unet)
    for oe in (model.outembed_128bp, model.outembed_1bp):
        if not hasattr(type(oe.norm), "affine"):
            type(oe.norm).affine = _affine
    eng = engine_for(unet, model)
    Embeds = namedtuple("Embeds", ["embeds_1bp", "embeds_128bp", "embeds_pair"])
    TUO = namedtuple("TransformerUnetOutput", ["unet_output", "single_repr", "pairwise_repr"])

    @torch.no_grad()
    def get_embeds(self, seq, organism_index, return_unet_transformer_outputs=False):
        if isinstance(organism_index, int):
            organism_index = torch.full((seq.shape[0],), organism_index, device=seq.device, dtype=torch.long)
        out = eng.run_embeds(seq, organism_index)
        emb = Embeds(out["embeds_1bp"], out["embeds_128bp"], out["embeds_pair"])
        if not return_unet_transformer_outputs:
            return emb
        return emb, TUO(out["unet_output"], out["single_repr"], out["pairwise_repr"])

    model.get_embeds = types.MethodType(get_embeds, model)
    model._agb200_engine = eng
    return model
