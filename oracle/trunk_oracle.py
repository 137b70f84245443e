"""CPU oracle for the AlphaGenome trunk forward — TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import this
module; the product path (alphagenome_b200/) never does and has no CPU fallback.

This is an independent fp32 restatement (plain torch on the CPU, channels-last, functional over a state_dict)
of the reference algorithm in lucidrains/alphagenome `alphagenome_pytorch/alphagenome.py` ("AG").  It is
PINNED: tests/golden/*.npz hold outputs of the unmodified reference (imported in the build container through
oracle/shims, see tools/make_golden.py) for the same synthetic weights and inputs, and
tests/test_oracle_cpu.py checks this file against them to ~1e-5.  Against the *published* pretrained model
parity is unpinned: the reference's own value-level goldens (tests/regression_data, produced by the JAX
model with HuggingFace weights) cannot be regenerated offline.

Layout: activations are [B, length, C] (channels-last); the pair tensor is [B, P, P, 128].
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
EPS = 1e-5


# ---------------------------------------------------------------------------------------------- primitives
def gelu1702(x):
    """AG:130-132."""
    return x * torch.sigmoid(1.702 * x)


def batch_rms_affine(sd: SD, p: str):
    """Frozen BatchRMSNorm (AG:344-361) is a per-channel affine: x / sqrt(running_var + eps) * gamma + beta."""
    return sd[p + "gamma"] / torch.sqrt(sd[p + "running_var"] + EPS), sd[p + "beta"]


UPDATE_KEY = "__update_running_var__"   # sd[UPDATE_KEY] = True: BatchRMSNorm re-estimates running_var (the reference's default)


def batch_rms(sd: SD, p: str, x):
    """BatchRMSNorm.forward AG:324-361.  With sd[UPDATE_KEY] set, the population variance of the batch (over every
    non-channel element, AG:276-284) is first lerped into running_var (weight 1 - momentum = 0.1, AG:315, 340) — the
    entry of `sd` is REPLACED, so the caller sees the updated statistics, as the reference's buffer does."""
    if sd.get(UPDATE_KEY, False):
        var = x.reshape(-1, x.shape[-1]).var(dim=0, unbiased=False)
        sd[p + "running_var"] = torch.lerp(sd[p + "running_var"], var, 0.1)
    s, t = batch_rms_affine(sd, p)
    return x * s + t


def rms_norm_bias(sd: SD, p: str, x):
    """RMSNormWithBias AG:400-405."""
    inv = torch.rsqrt(x.float().square().mean(-1, keepdim=True) + EPS)
    return x * inv * sd[p + "weight"] + sd[p + "bias"]


def layer_norm(sd: SD, p: str, x):
    """nn.LayerNorm over the last dim, eps 1e-5 (AG:681-684)."""
    mu = x.mean(-1, keepdim=True)
    var = (x - mu).square().mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + EPS) * sd[p + "weight"] + sd[p + "bias"]


def conv1d_cl(x, w, b):
    """'same' zero-padded Conv1d on channels-last input; w is the reference's [C_out, C_in, taps]."""
    return F.conv1d(x.transpose(1, 2), w, b, padding=w.shape[-1] // 2).transpose(1, 2)


def conv_block(sd: SD, p: str, x):
    """ConvBlock AG:445-452: BatchRMSNorm -> GELU1702 -> Conv1d."""
    return conv1d_cl(gelu1702(batch_rms(sd, p + "net.0.", x)), sd[p + "net.2.weight"], sd[p + "net.2.bias"])


def maxpool2(x):
    B, L, C = x.shape
    return x.reshape(B, L // 2, 2, C).amax(2)


def linear(sd: SD, p: str, x):
    return F.linear(x, sd[p + "weight"], sd.get(p + "bias"))


# ---------------------------------------------------------------------------------------------- encoder
def dna_embed(sd: SD, p: str, seq):
    """DNAEmbed.forward AG:1167-1181. Negative tokens are all-zero one-hot rows."""
    w = sd[p + "conv.weight"]  # [C, n_base, 15]
    n_base = w.shape[1]
    onehot = F.one_hot(seq.clamp(min=0), n_base).float() * (seq >= 0).unsqueeze(-1).float()
    x = conv1d_cl(onehot, w, sd[p + "conv.bias"])
    x = x + conv_block(sd, p + "pointwise.", x)
    return maxpool2(x), x


def down_block(sd: SD, p: str, x):
    """DownresBlock.forward AG:470-481: zero-padded channel residual, second residual conv, maxpool2."""
    y = conv_block(sd, p + "conv.", x)
    y = y + F.pad(x, (0, y.shape[-1] - x.shape[-1]))
    y = conv_block(sd, p + "conv_out.", y) + y
    return maxpool2(y), y


def up_block(sd: SD, p: str, x, skip):
    """UpresBlock.forward AG:503-519."""
    y = conv_block(sd, p + "conv.", x)
    y = y + x[..., : y.shape[-1]]
    y = y.repeat_interleave(2, dim=1) * sd[p + "residual_scale"]
    y = y + conv_block(sd, p + "unet_conv.", skip)
    return conv_block(sd, p + "conv_out.", y) + y


# ---------------------------------------------------------------------------------------------- tower
def rotary_table(sd: SD, p: str, n: int):
    """RotaryEmbedding.forward AG:545-557: angles [n, d] with each frequency repeated for the (even, odd) pair."""
    inv = sd.get(p + "inv_freq")
    if inv is None:
        inv = rotary_inv_freq(128)
    return (torch.arange(n, dtype=inv.dtype)[:, None] * inv[None, :]).repeat_interleave(2, dim=-1)


def geomspace(start: float, stop: float, num: int):
    """AG:139-153 geomspace: `num` log-spaced points including both ends (computed in fp32 like the reference)."""
    return torch.logspace(math.log10(start), math.log10(stop), num, base=10.0, dtype=torch.float32)


def rotary_inv_freq(dim: int, max_positions: int = 8192):
    nf = dim // 2
    return 1.0 / (torch.arange(nf).float() + geomspace(1, max_positions - nf + 1, nf))


def apply_rope(ang, t):
    """apply_rotary_pos_emb / rotate_half AG:559-566 (interleaved pairs: (x0, x1) -> (-x1, x0))."""
    t2 = t.reshape(*t.shape[:-1], -1, 2)
    rot = torch.stack((-t2[..., 1], t2[..., 0]), dim=-1).reshape(t.shape)
    return t * ang.cos() + rot * ang.sin()


def attention_bias(sd: SD, p: str, pair):
    """to_attn_bias AG:688-693: BatchRMSNorm -> erf GELU -> Linear(128 -> heads); returns [B, heads, P, P]."""
    h = F.gelu(batch_rms(sd, p + "to_attn_bias.0.", pair))
    return F.linear(h, sd[p + "to_attn_bias.2.weight"]).permute(0, 3, 1, 2)


def attention(sd: SD, p: str, x, pair, ang, softcap: float = 5.0):
    """Attention.forward AG:699-786, fp32 (CPU) path: MQA with 8 query heads, one K/V head."""
    B, n, _ = x.shape
    dq = sd[p + "q_norm.weight"].numel()
    dv = sd[p + "v_norm.weight"].numel()
    qkv = F.linear(x, sd[p + "to_qkv.weight"])
    H = (qkv.shape[-1] - dq - dv) // dq
    q, k, v = qkv.split((H * dq, dq, dv), dim=-1)
    q = layer_norm(sd, p + "q_norm.", q.reshape(B, n, H, dq))
    k = layer_norm(sd, p + "k_norm.", k)
    v = layer_norm(sd, p + "v_norm.", v)
    q = apply_rope(ang[None, :, None, :], q)
    k = apply_rope(ang[None], k)
    sim = torch.einsum("bihd,bjd->bhij", q, k) * dq ** -0.5
    bias = attention_bias(sd, p, pair)  # [B, H, P, P]
    r = n // bias.shape[-1]
    bias = bias.repeat_interleave(r, dim=2).repeat_interleave(r, dim=3)
    sim = torch.tanh((sim + bias) / softcap) * softcap
    o = torch.einsum("bhij,bjd->bihd", sim.softmax(-1), v).reshape(B, n, H * dv)
    return linear(sd, p + "to_out.", o)


def feed_forward(sd: SD, p: str, x):
    """FeedForward AG:893-906: Linear -> ReLU -> Linear."""
    return linear(sd, p + "3.", torch.relu(linear(sd, p + "0.", x)))


def rel_pos_features(P: int, num_features: int = 32, max_distance: int = 8192, pool: int = 16):
    """RelativePosFeatures.forward AG:587-604: rows d = -P..P-1, columns [|d| < w_c] then sign(d)*[|d| < w_c]."""
    max_len = max_distance // pool
    d = torch.arange(-P, P)
    widths = torch.arange(num_features).float() + geomspace(1, max_len - num_features + 1, num_features + 1)[:-1]
    f = (widths[None, :] > d.abs()[:, None].float()).float()
    return torch.cat((f, d.sign()[:, None].float() * f), dim=-1)


def single_to_pairwise(sd: SD, p: str, single, feats, pool: int = 16):
    """SingleToPairwise.forward AG:822-856.

    sim[i,j,h] = q_i.k_j + 0.5 * ((q_i + bq).R[P + j - i] + (k_j + bk).R[P + i - j]), with R the encoded
    relative-position features (relative_shift AG:523-530 picks column P + (j - i) of row i).
    """
    B, n, D = single.shape
    P = n // pool
    x = rms_norm_bias(sd, p + "norm.", single.reshape(B, P, pool, D).mean(2))
    Hd = sd[p + "to_qk.weight"].shape[0] // 2
    H = sd[p + "qk_to_pairwise.weight"].shape[1]
    d = Hd // H
    qk = F.linear(x, sd[p + "to_qk.weight"])
    q = qk[..., :Hd].reshape(B, P, H, d)
    k = qk[..., Hd:].reshape(B, P, H, d)
    R = linear(sd, p + "to_rel_pos_encoding.", feats).reshape(2 * P, H, d)
    bq, bk = sd[p + "qk_rel_pos_bias"][0].reshape(H, d), sd[p + "qk_rel_pos_bias"][1].reshape(H, d)
    sim = torch.einsum("bihd,bjhd->bijh", q, k)
    rq = torch.einsum("bihd,phd->bihp", q + bq, R)  # [B, P, H, 2P]
    rk = torch.einsum("bjhd,phd->bjhp", k + bk, R)
    i = torch.arange(P)
    col = P + (i[None, :] - i[:, None])  # [i, j] -> P + j - i
    rel_q = rq.permute(0, 2, 1, 3).gather(3, col[None, None].expand(B, H, P, P))  # [B, H, i, j]
    rel_k = rk.permute(0, 2, 1, 3).gather(3, col[None, None].expand(B, H, P, P))  # [B, H, j, i] indexed (row=j, col=i)
    sim = sim + 0.5 * (rel_q.permute(0, 2, 3, 1) + rel_k.permute(0, 3, 2, 1))
    y = linear(sd, p + "qk_to_pairwise.", sim)
    outer = F.linear(F.gelu(x), sd[p + "to_outer_sum.1.weight"])
    dp = outer.shape[-1] // 2
    return y + outer[..., :dp][:, :, None, :] + outer[..., dp:][:, None, :, :]


def pair_row_attention(sd: SD, p: str, x):
    """PairwiseRowAttention.forward AG:871-889: for each first index, attend over the second index."""
    D = x.shape[-1]
    qk = F.linear(x, sd[p + "to_qk.weight"])
    q, k = qk[..., :D], qk[..., D:]
    v = linear(sd, p + "to_v.", x)
    att = (torch.einsum("bnid,bnjd->bnij", q, k) * D ** -0.5).softmax(-1)
    return torch.einsum("bnij,bnjd->bnid", att, v)


def tower(sd: SD, p: str, single, collect: Optional[Callable] = None):
    """TransformerTower.forward AG:993-1027 (RoPE variant)."""
    B, n, D = single.shape
    ang = rotary_table(sd, p + "rotary_emb.", n)
    n_layers = 1 + max(int(k[len(p + "layers."):].split(".")[0]) for k in sd if k.startswith(p + "layers."))
    pair = None
    feats = None
    for li in range(n_layers):
        lp = f"{p}layers.{li}."
        if (lp + "2.to_qk.weight") in sd:
            pool = 16
            if feats is None:
                H = sd[lp + "2.qk_to_pairwise.weight"].shape[1]
                feats = rel_pos_features(n // pool, num_features=H)
            new = single_to_pairwise(sd, lp + "2.", single, feats, pool)
            pair = new if pair is None else new + pair
            pair = pair_row_attention(sd, lp + "3.block.", rms_norm_bias(sd, lp + "3.pre_rmsnorm.", pair)) + pair
            pair = feed_forward(sd, lp + "4.block.", rms_norm_bias(sd, lp + "4.pre_rmsnorm.", pair)) + pair
        a = attention(sd, lp + "0.block.", batch_rms(sd, lp + "0.pre_rmsnorm.", single), pair, ang)
        single = batch_rms(sd, lp + "0.post_rmsnorm.", a) + single
        f = feed_forward(sd, lp + "1.block.", batch_rms(sd, lp + "1.pre_rmsnorm.", single))
        single = batch_rms(sd, lp + "1.post_rmsnorm.", f) + single
        if collect:
            collect(f"tower.layer{li}.single", single)
            collect(f"tower.layer{li}.pair", pair)
    return single, pair


# ---------------------------------------------------------------------------------------------- trunk + embeds
def transformer_unet(sd: SD, seq, organism_embed, p: str = "transformer_unet.", collect: Optional[Callable] = None):
    """TransformerUnet.forward AG:1095-1136 -> (unet_output [B,L,C0], single [B,n,D], pair [B,P,P,128])."""
    x, skip = dna_embed(sd, p + "dna_embed.", seq)
    skips = [skip]
    if collect:
        collect("dna_embed.skip", skip)
        collect("embed.pooled", x)
    n_down = 1 + max(int(k[len(p + "downs."):].split(".")[0]) for k in sd if k.startswith(p + "downs."))
    for i in range(n_down):
        x, skip = down_block(sd, f"{p}downs.{i}.", x)
        skips.append(skip)
        if collect:
            collect(f"downs.{i}.pooled", x)
    if organism_embed is not None:
        x = x + organism_embed[:, None, :]
    single, pair = tower(sd, p + "transformer.", x, collect)
    x = single
    n_up = 1 + max(int(k[len(p + "ups."):].split(".")[0]) for k in sd if k.startswith(p + "ups."))
    for j in range(n_up):
        x = up_block(sd, f"{p}ups.{j}.", x, skips.pop())
        if collect:
            collect(f"ups.{j}", x)
    return x, single, pair


def output_embedding(sd: SD, p: str, x, organism_index, skip=None):
    """OutputEmbedding.forward AG:1213-1235."""
    y = linear(sd, p + "double_features.", x)
    if skip is not None:
        s = F.linear(skip, sd[p + "skip_proj.weight"])
        y = y + s.repeat_interleave(y.shape[1] // s.shape[1], dim=1)
    y = batch_rms(sd, p + "norm.", y) + sd[p + "embed.weight"][organism_index][:, None, :]
    return gelu1702(y)


def output_pair_embedding(sd: SD, p: str, pair, organism_index):
    """OutputPairEmbedding.forward AG:1248-1259 (symmetrize AG:121-122)."""
    y = 0.5 * (pair + pair.transpose(1, 2))
    y = rms_norm_bias(sd, p + "norm.", y) + sd[p + "embed.weight"][organism_index][:, None, None, :]
    return gelu1702(y)


@torch.no_grad()
def alphagenome_embeds(sd: SD, seq, organism_index, collect: Optional[Callable] = None):
    """AlphaGenome.get_embeds AG:1754-1778 -> dict with the three embeddings and the trunk outputs."""
    if isinstance(organism_index, int):
        organism_index = torch.full((seq.shape[0],), organism_index, dtype=torch.long)
    org = sd["organism_embed.embed.weight"][organism_index]
    unet_out, single, pair = transformer_unet(sd, seq, org, collect=collect)
    e128 = output_embedding(sd, "outembed_128bp.", single, organism_index)
    e1 = output_embedding(sd, "outembed_1bp.", unet_out, organism_index, skip=e128)
    epair = output_pair_embedding(sd, "outembed_pair.", pair, organism_index)
    return dict(embeds_1bp=e1, embeds_128bp=e128, embeds_pair=epair, unet_output=unet_out, single_repr=single,
                pairwise_repr=pair)


# ---------------------------------------------------------------------------------------------- prediction heads
# (`AlphaGenome.add_heads`, AG:1653-1682 — every organism's heads run on every sample, AG:1823-1850)
def tracks_scaled_prediction(sd: SD, p: str, x):
    """TracksScaledPrediction.forward AG:1378-1393: softplus(Linear(x)) * softplus(scale)."""
    return F.softplus(linear(sd, p + "to_pred.", x)) * F.softplus(sd[p + "scale"])


def contact_map_head(sd: SD, p: str, embeds_pair, organism_index):
    """ContactMapHead.forward AG:1286-1299: symmetrize -> RMSNormWithBias -> + organism embedding -> GELU1702 -> Linear."""
    x = 0.5 * (embeds_pair + embeds_pair.transpose(1, 2))
    x = rms_norm_bias(sd, p + "norm.", x) + sd[p + "embed.weight"][organism_index][:, None, None, :]
    return linear(sd, p + "proj.", gelu1702(x))


def splice_rotary_table(hidden: int, n_contexts: int, max_positions: int = 2 ** 20):
    """RotaryEmbedding(hidden, max_positions = 2**20)(n_contexts) AG:541-557: [contexts, hidden] angles."""
    inv = rotary_inv_freq(hidden, max_positions)
    ang = torch.arange(n_contexts, dtype=torch.float32)[:, None] * inv[None, :]
    return ang.repeat_interleave(2, dim=-1)


def splice_junction_head(sd: SD, p: str, embeds_1bp, donor_idx, acceptor_idx):
    """SpliceJunctionHead.forward AG:1322-1376 -> softplus scores [B, donors, acceptors, contexts]."""
    x_proj = linear(sd, p + "project.", embeds_1bp)                     # [B, L, hidden]
    scale, offset = sd[p + "scale"], sd[p + "offset"]                   # [contexts, hidden]
    ang = splice_rotary_table(scale.shape[1], scale.shape[0])           # the context index is the rotary position
    b = torch.arange(x_proj.shape[0])[:, None]

    def tissue_scaled_rope(idx):
        x = x_proj[b, idx]                                              # [B, P, hidden]
        x = x[:, :, None, :] * scale + offset                           # [B, P, contexts, hidden]
        return apply_rope(ang, x)

    d, a = tissue_scaled_rope(donor_idx), tissue_scaled_rope(acceptor_idx)
    return F.softplus(torch.einsum("bdth,bath->bdat", d, a))


@torch.no_grad()
def alphagenome_heads(sd: SD, embeds: dict, organism_index, organisms, splice_donor_idx=None, splice_acceptor_idx=None):
    """AlphaGenome.forward AG:1780-1852 after get_embeds, for heads registered with add_heads."""
    out = {}
    for org in organisms:
        hp = f"heads.{org}."
        o = dict()
        o["1bp_tracks"] = tracks_scaled_prediction(sd, hp + "1bp_tracks.", embeds["embeds_1bp"])
        o["128bp_tracks"] = tracks_scaled_prediction(sd, hp + "128bp_tracks.", embeds["embeds_128bp"])
        o["contact_head"] = contact_map_head(sd, hp + "contact_head.", embeds["embeds_pair"], organism_index)
        o["splice_logits"] = linear(sd, hp + "splice_logits.linear.", embeds["embeds_1bp"])                 # AG:1301-1310
        o["splice_usage"] = torch.sigmoid(linear(sd, hp + "splice_usage.linear.", embeds["embeds_1bp"]))    # AG:1312-1321
        if splice_donor_idx is not None:
            o["splice_juncs"] = splice_junction_head(sd, hp + "splice_juncs.", embeds["embeds_1bp"], splice_donor_idx, splice_acceptor_idx)
        out[org] = o
    return out


# ---------------------------------------------------------------------------------------------- JAX-aligned heads
# (`AlphaGenome.add_reference_heads`, AG:1684-1731: the heads of the official checkpoint)
JAX_GENOME_TRACK_HEADS = {  # AG:77-85
    "rna_seq": (768, (1, 128)), "cage": (640, (1, 128)), "dnase": (384, (1, 128)), "procap": (128, (1, 128)),
    "atac": (256, (1, 128)), "chip_tf": (1664, (128,)), "chip_histone": (1152, (128,)),
}


def genome_tracks_head(sd: SD, p: str, embeds_1bp, embeds_128bp):
    """GenomeTracksHead.forward AG:1395-1423: one TracksScaledPrediction per resolution present in the state_dict."""
    out = {}
    if p + "resolutions.resolution_1.scale" in sd:
        out["scaled_predictions_1bp"] = tracks_scaled_prediction(sd, p + "resolutions.resolution_1.", embeds_1bp)
    if p + "resolutions.resolution_128.scale" in sd:
        out["scaled_predictions_128bp"] = tracks_scaled_prediction(sd, p + "resolutions.resolution_128.", embeds_128bp)
    return out


def contact_maps_head(sd: SD, p: str, embeds_pair):
    """ContactMapsHead.forward AG:1425-1442: symmetrize -> Linear."""
    return linear(sd, p + "to_pred.", 0.5 * (embeds_pair + embeds_pair.transpose(1, 2)))


def splice_sites_junction_head(sd: SD, p: str, embeds_1bp, splice_site_positions):
    """SpliceSitesJunctionHead.forward AG:1444-1545: rotary position = the genomic index of the site."""
    x_proj = linear(sd, p + "project.", embeds_1bp)
    Hd = x_proj.shape[-1]
    T = sd[p + "pos_donor_embeddings"].numel() // (2 * Hd)
    inv = rotary_inv_freq(Hd, 2 ** 20)
    b = torch.arange(x_proj.shape[0])[:, None]

    def apply(idx, param):
        safe = idx.clamp_min(0)
        x = x_proj[b, safe]                                              # [B, P, H]
        par = sd[p + param].view(2, T, Hd)
        x = x[:, :, None, :] * par[0] + par[1]                           # [B, P, T, H]
        ang = (safe.float()[..., None] * inv).repeat_interleave(2, dim=-1)[:, :, None, :]
        return apply_rope(ang, x)

    pd, pa, nd, na = (splice_site_positions[:, i] for i in range(4))
    pos = F.softplus(torch.einsum("bdth,bath->bdat", apply(pd, "pos_donor_embeddings"), apply(pa, "pos_acceptor_embeddings")))
    neg = F.softplus(torch.einsum("bdth,bath->bdat", apply(nd, "neg_donor_embeddings"), apply(na, "neg_acceptor_embeddings")))
    pmask = (pd >= 0)[:, :, None] & (pa >= 0)[:, None, :]
    nmask = (nd >= 0)[:, :, None] & (na >= 0)[:, None, :]
    pred = torch.cat((pos * pmask[..., None], neg * nmask[..., None]), dim=-1)
    mask = torch.cat((pmask[..., None].expand(*pmask.shape, T), nmask[..., None].expand(*nmask.shape, T)), dim=-1)
    return dict(predictions=pred, splice_site_positions=splice_site_positions, splice_junction_mask=mask)


@torch.no_grad()
def alphagenome_reference_heads(sd: SD, embeds: dict, organism_index, organisms, splice_site_positions=None):
    """AlphaGenome.forward AG:1780-1852 for heads registered with add_reference_heads."""
    out = {}
    for org in organisms:
        hp = f"heads.{org}."
        o = {name: genome_tracks_head(sd, hp + name + ".", embeds["embeds_1bp"], embeds["embeds_128bp"]) for name in JAX_GENOME_TRACK_HEADS}
        o["contact_maps"] = contact_maps_head(sd, hp + "contact_maps.", embeds["embeds_pair"])
        o["splice_sites_classification"] = linear(sd, hp + "splice_sites_classification.linear.", embeds["embeds_1bp"])
        o["splice_sites_usage"] = torch.sigmoid(linear(sd, hp + "splice_sites_usage.linear.", embeds["embeds_1bp"]))
        o["splice_sites_junction"] = (None if splice_site_positions is None else
                                      splice_sites_junction_head(sd, hp + "splice_sites_junction.", embeds["embeds_1bp"], splice_site_positions))
        out[org] = o
    return out
