"""JAX/Haiku checkpoint -> state_dict of `alphagenome_b200.AlphaGenome` (and of the reference model: same keys).

Host-side counterpart of the reference's `alphagenome_pytorch/convert/` (convert_checkpoint.py:108-233 walks a table built
by mapping_spec.py:130-984; shape_transforms.py holds the layout rules).  Written from the checkpoint's layout rules, not
from the reference's table: every Haiku module kind has ONE schema below and the table is generated from the model
structure (7 resolutions, 9 tower layers, a pair block every second layer, two organisms).  tests/test_convert_cpu.py
checks the generated table entry by entry against a dump of the reference's (tests/golden/convert_mapping.json,
tools/make_golden_convert.py) and a converted synthetic checkpoint tensor by tensor against the reference converter.

Layout rules (Haiku -> PyTorch):
    Linear  w (in, out)            -> weight (out, in)                      "linear"
    Conv1D  w (width, in, out)     -> weight (out, in, width)               "conv"
    Linear used as a width-1 conv  -> weight (out, in, 1)                   "pointwise"
    q / k (/ v) projections        -> one concatenated weight, rows q|k|v   "concat"
    q_r_bias, k_r_bias             -> stacked on a new first axis           "stack"
    RMSBatchNorm scale/offset/var_ema [1, 1, C] -> [C]                      "squeeze"
    scalar residual_scale []       -> [1]                                   "scalar"
    StandardizedConv1D (w, scale, bias): centre w over (width, in) per output channel, multiply by
        scale * rsqrt(max(fan_in * var, 1e-4)), then the conv rule          "standardized"
    multi-organism head parameters [organisms, ...] -> one organism's slice (index), then the rule above.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Mapping, Optional, Tuple

import torch

ROOT = "alphagenome"
ORGANISMS = ("human", "mouse")
# the heads of the official checkpoint (AG:77-90): name -> resolutions
GENOME_TRACK_HEADS = {"rna_seq": (1, 128), "cage": (1, 128), "dnase": (1, 128), "procap": (1, 128), "atac": (1, 128),
                      "chip_tf": (128,), "chip_histone": (128,)}


@dataclass(frozen=True)
class Rule:
    """state_dict[torch_key] = op(checkpoint[src...]);  `state` rules read the Haiku state (var_ema), the rest params."""

    torch_key: str
    src: Tuple[str, ...]
    op: str = "copy"
    extra: Tuple[Tuple[str, str], ...] = ()      # named side inputs ("scale", "bias" of a standardized conv)
    index: Optional[int] = None                  # organism slice of a multi-organism head parameter
    state: bool = False


# ---------------------------------------------------------------------------------------------- layout operations
def standardize_conv(w: torch.Tensor, scale: torch.Tensor, fan_in: Optional[int] = None, floor: float = 1e-4) -> torch.Tensor:
    """StandardizedConv1D's effective kernel, still in Haiku layout (width, in, out)."""
    width, cin, _ = w.shape
    fan_in = width * cin if fan_in is None else fan_in
    c = w - w.mean(dim=(0, 1), keepdim=True)
    var = c.var(dim=(0, 1), keepdim=True, unbiased=False)
    denom = torch.maximum((fan_in * var).to(c.dtype), torch.tensor(floor, dtype=c.dtype, device=c.device))
    return c * (scale.reshape(-1) if scale.ndim == 1 else scale) * torch.rsqrt(denom)


def apply_rule(rule: Rule, tensors: List[torch.Tensor], extra: Mapping[str, torch.Tensor]):
    if rule.index is not None:
        tensors = [t[rule.index] for t in tensors]
        extra = {k: v[rule.index] for k, v in extra.items()}
    op = rule.op
    if op == "copy":
        (t,) = tensors
        return t
    if op == "linear":
        (t,) = tensors
        _want(t, 2, rule)
        return t.t()
    if op == "conv":
        (t,) = tensors
        _want(t, 3, rule)
        return t.permute(2, 1, 0).contiguous()
    if op == "pointwise":
        (t,) = tensors
        _want(t, 2, rule)
        return t.t().unsqueeze(-1).contiguous()
    if op == "concat":
        for t in tensors:
            _want(t, 2, rule)
        if len({t.shape[0] for t in tensors}) != 1:
            raise ValueError(f"{rule.torch_key}: projections to concatenate disagree on the input width")
        return torch.cat([t.t() for t in tensors], dim=0).contiguous()
    if op == "stack":
        return torch.stack(tensors, dim=0)
    if op == "squeeze":
        (t,) = tensors
        return t.squeeze()
    if op == "scalar":
        (t,) = tensors
        return t.unsqueeze(0)
    if op == "standardized":
        (w,) = tensors
        _want(w, 3, rule)
        bias = extra["bias"]
        if bias.ndim != 1 or bias.shape[0] != w.shape[2]:
            raise ValueError(f"{rule.torch_key}: bias shape {tuple(bias.shape)} does not match {w.shape[2]} output channels")
        return standardize_conv(w, extra["scale"]).permute(2, 1, 0).contiguous(), bias.contiguous()
    raise ValueError(f"unknown op {op!r}")


def _want(t, ndim, rule):
    if t.ndim != ndim:
        raise ValueError(f"{rule.torch_key}: expected a {ndim}-d array, got shape {tuple(t.shape)}")


# ---------------------------------------------------------------------------------------------- table construction
def _nth(name: str, i: int) -> str:
    """Haiku numbers repeated modules `name`, `name_1`, `name_2`, ..."""
    return name if i == 0 else f"{name}_{i}"


class _Table:
    def __init__(self):
        self.rules: List[Rule] = []

    def add(self, torch_key, src, op="copy", **kw):
        self.rules.append(Rule(torch_key, (src,) if isinstance(src, str) else tuple(src), op, **kw))

    # --- module schemas ---
    def batch_rms(self, t, j):
        """RMSBatchNorm: params scale / offset, state var_ema."""
        self.add(t + ".gamma", j + "/scale", "squeeze")
        self.add(t + ".beta", j + "/offset", "squeeze")
        self.add(t + ".running_var", j + "/var_ema", "squeeze", state=True)

    def affine_norm(self, t, j):
        """LayerNorm / RMSNorm with bias: scale, offset kept as they are."""
        self.add(t + ".weight", j + "/scale")
        self.add(t + ".bias", j + "/offset")

    def linear(self, t, j, bias=True, **kw):
        self.add(t + ".weight", j + "/w", "linear", **kw)
        if bias:
            self.add(t + ".bias", j + "/b", **kw)

    def conv_block(self, t, j):
        """ConvBlock = RMSBatchNorm -> GELU -> StandardizedConv1D (Sequential indices 0, 1, 2)."""
        sc = j + "/standardized_conv1_d"
        self.add(t + ".net.2.weight", sc + "/w", "standardized", extra=(("scale", sc + "/scale"), ("bias", sc + "/bias")))
        self.batch_rms(t + ".net.0", j + "/rms_batch_norm")

    def sandwich(self, t, j):
        """NormWrapper with BatchRMSNorm before and after the block."""
        self.batch_rms(t + ".pre_rmsnorm", j + "/rms_batch_norm")
        self.batch_rms(t + ".post_rmsnorm", j + "/rms_batch_norm_1")

    def mlp(self, t, j):
        """Sequential(Linear, ReLU, Dropout, Linear): indices 0 and 3."""
        self.linear(t + ".0", j + "/linear")
        self.linear(t + ".3", j + "/linear_1")


def build_rules(n_down: int = 6, depth: int = 9, pair_every: int = 2, heads: bool = True) -> List[Rule]:
    """The whole table for the published architecture (trunk + output embeddings + the checkpoint's heads)."""
    T = _Table()
    enc, dec, tower = f"{ROOT}/sequence_encoder", f"{ROOT}/sequence_decoder", f"{ROOT}/transformer_tower"
    # DNA embedder: plain width-15 conv, then a ConvBlock
    T.add("transformer_unet.dna_embed.conv.weight", f"{enc}/dna_embedder/conv1_d/w", "conv")
    T.add("transformer_unet.dna_embed.conv.bias", f"{enc}/dna_embedder/conv1_d/b")
    T.conv_block("transformer_unet.dna_embed.pointwise", f"{enc}/dna_embedder/conv_block")
    # encoder: downres_block_{i} numbers from 0 explicitly
    for i in range(n_down):
        T.conv_block(f"transformer_unet.downs.{i}.conv", f"{enc}/downres_block_{i}/conv_block")
        T.conv_block(f"transformer_unet.downs.{i}.conv_out", f"{enc}/downres_block_{i}/conv_block_1")
    # tower
    n_pair = 0
    for li in range(depth):
        t = f"transformer_unet.transformer.layers.{li}"
        mha, bias_blk, mlp_blk = (f"{tower}/{_nth(n, li)}" for n in ("mha_block", "attention_bias_block", "mlp_block"))
        T.add(t + ".0.block.to_qkv.weight", [f"{mha}/{p}_layer/w" for p in "qkv"], "concat")
        for p in "qkv":
            T.affine_norm(f"{t}.0.block.{p}_norm", f"{mha}/norm_{p}")
        T.linear(t + ".0.block.to_out", mha + "/linear_embedding")
        T.sandwich(t + ".0", mha)
        T.linear(t + ".0.block.to_attn_bias.2", bias_blk + "/linear", bias=False)
        T.batch_rms(t + ".0.block.to_attn_bias.0", bias_blk + "/rms_batch_norm")
        T.mlp(t + ".1.block", mlp_blk)
        T.sandwich(t + ".1", mlp_blk)
        if li % pair_every == 0:
            pu = f"{tower}/{_nth('pair_update_block', n_pair)}"
            n_pair += 1
            s2p, row, pmlp = pu + "/sequence_to_pair_block", pu + "/row_attention_block", pu + "/pair_mlp_block"
            T.add(t + ".2.to_qk.weight", [s2p + "/linear_q/w", s2p + "/linear_k/w"], "concat")
            T.linear(t + ".2.qk_to_pairwise", s2p + "/linear_pair")
            T.add(t + ".2.to_outer_sum.1.weight", [s2p + "/linear_y_q/w", s2p + "/linear_y_k/w"], "concat")
            T.linear(t + ".2.to_rel_pos_encoding", s2p + "/linear_pos_features")
            T.add(t + ".2.qk_rel_pos_bias", [s2p + "/q_r_bias", s2p + "/k_r_bias"], "stack")
            T.affine_norm(t + ".2.norm", s2p + "/norm_seq2pair")
            T.add(t + ".3.block.to_qk.weight", [row + "/linear_q/w", row + "/linear_k/w"], "concat")
            T.linear(t + ".3.block.to_v", row + "/linear_v")
            T.affine_norm(t + ".3.pre_rmsnorm", row + "/layer_norm")
            T.mlp(t + ".4.block", pmlp)
            T.affine_norm(t + ".4.pre_rmsnorm", pmlp + "/layer_norm")
    # decoder
    for j in range(n_down + 1):
        t, up = f"transformer_unet.ups.{j}", f"{dec}/{_nth('up_res_block', j)}"
        T.conv_block(t + ".conv", up + "/conv_in")
        T.add(t + ".residual_scale", up + "/residual_scale", "scalar")
        T.add(t + ".unet_conv.net.2.weight", up + "/pointwise_conv_unet_skip/linear/w", "pointwise")
        T.add(t + ".unet_conv.net.2.bias", up + "/pointwise_conv_unet_skip/linear/b")
        T.batch_rms(t + ".unet_conv.net.0", up + "/pointwise_conv_unet_skip/rms_batch_norm")
        T.conv_block(t + ".conv_out", up + "/conv_out")
    # organism + output embeddings
    T.add("organism_embed.embed.weight", f"{ROOT}/embed/embeddings")
    for t, j, skip in (("outembed_128bp", f"{ROOT}/output_embedder", False), ("outembed_1bp", f"{ROOT}/output_embedder_1", True)):
        T.linear(t + ".double_features", j + "/linear")
        if skip:
            T.linear(t + ".skip_proj", j + "/linear_1", bias=False)
        T.batch_rms(t + ".norm", j + "/rms_batch_norm")
        T.add(t + ".embed.weight", j + "/embed/embeddings")
    T.affine_norm("outembed_pair.norm", f"{ROOT}/output_pair/layer_norm")
    T.add("outembed_pair.embed.weight", f"{ROOT}/output_pair/embed/embeddings")
    # heads of the official checkpoint (add_reference_heads, AG:1684-1731): parameters carry an organism axis
    if heads:
        hd = f"{ROOT}/head"
        for oi, org in enumerate(ORGANISMS):
            for name, resolutions in GENOME_TRACK_HEADS.items():
                for r in resolutions:
                    t, j = f"heads.{org}.{name}.resolutions.resolution_{r}", f"{hd}/{name}/resolution_{r}"
                    T.linear(t + ".to_pred", j + "/multi_organism_linear", index=oi)
                    T.add(t + ".scale", j + "/learnt_scale", index=oi)
            T.linear(f"heads.{org}.contact_maps.to_pred", f"{hd}/contact_maps/multi_organism_linear", index=oi)
            T.linear(f"heads.{org}.splice_sites_classification.linear", f"{hd}/splice_sites_classification/multi_organism_linear", index=oi)
            T.linear(f"heads.{org}.splice_sites_usage.linear", f"{hd}/splice_sites_usage/multi_organism_linear", index=oi)
            T.linear(f"heads.{org}.splice_sites_junction.project", f"{hd}/splice_sites_junction/multi_organism_linear", index=oi)
            for side in ("pos_donor", "pos_acceptor", "neg_donor", "neg_acceptor"):
                T.add(f"heads.{org}.splice_sites_junction.{side}_embeddings", f"{hd}/splice_sites_junction/{side}_logits/embeddings", index=oi)
    return T.rules


# ---------------------------------------------------------------------------------------------- conversion
def flatten(tree: Mapping, prefix: str = "", sep: str = "/") -> Dict[str, object]:
    """Nested Haiku {module: {param: array}} -> {'module/param': array}."""
    out: Dict[str, object] = {}
    for k, v in tree.items():
        key = f"{prefix}{sep}{k}" if prefix else str(k)
        if isinstance(v, Mapping):
            out.update(flatten(v, key, sep))
        else:
            out[key] = v
    return out


def _tensor(a) -> torch.Tensor:
    if torch.is_tensor(a):
        return a
    import numpy as np

    return torch.from_numpy(np.array(a.numpy() if hasattr(a, "numpy") else a, copy=True))


def convert_checkpoint(params: Mapping, state: Mapping, rules: Optional[Iterable[Rule]] = None, require_heads: bool = True,
                       report: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """Flat (or nested) Haiku params + state -> state_dict.  Like the reference converter (convert_checkpoint.py:146-160)
    a missing array is a KeyError; require_heads=False skips head entries whose arrays are absent (a trunk-only
    checkpoint).  `report` (optional dict) receives the unused checkpoint keys."""
    params = flatten(params) if any(isinstance(v, Mapping) for v in params.values()) else dict(params)
    state = flatten(state) if any(isinstance(v, Mapping) for v in state.values()) else dict(state)
    rules = build_rules() if rules is None else list(rules)
    sd: Dict[str, torch.Tensor] = {}
    used = set()
    for rule in rules:
        src = state if rule.state else params
        names = list(rule.src) + [k for _, k in rule.extra]
        missing = [k for k in names if k not in src]
        if missing:
            if rule.torch_key.startswith("heads.") and not require_heads:
                continue
            raise KeyError(f"{rule.torch_key}: checkpoint has no {missing[0]!r}")
        used.update(names)
        res = apply_rule(rule, [_tensor(src[k]) for k in rule.src], {n: _tensor(src[k]) for n, k in rule.extra})
        if isinstance(res, tuple):            # a standardized conv yields (weight, bias)
            sd[rule.torch_key], sd[rule.torch_key[: -len("weight")] + "bias"] = res
        else:
            sd[rule.torch_key] = res
    if report is not None:
        report["unused"] = sorted((set(params) | set(state)) - used)
    return sd


def load_converted(model, params: Mapping, state: Mapping, strict: bool = True):
    """convert_checkpoint + load_state_dict.  Entries the conversion does not produce but the model owns as constants
    (rotary inv_freq, the relative-position dummy buffer) are taken from the model itself."""
    sd = convert_checkpoint(params, state)
    own = model.state_dict()
    for k, v in own.items():
        if k not in sd and (k.endswith("inv_freq") or k.endswith("rel_pos_features.dummy")):
            sd[k] = v
    return model.load_state_dict(sd, strict=strict)
