"""Fused trunk engine: schedules the libagb200 kernels for TransformerUnet.forward (AG:1095-1136) and the
output embeddings (AG:1754-1778).  AG = reference alphagenome_pytorch/alphagenome.py.

Data layout in HBM (B = batch, len = positions at the current resolution, C = channels):
  * activations are channels-last [B, len, C];
  * the residual stream is fp32; every tensor-core operand is a bf16 tensor that was produced *already
    normalised and activated* (BatchRMSNorm affine + GELU1702 of the consumer) by the epilogue of the kernel
    that produced the residual — so a conv's zero padding (AG:448) is TMA out-of-bounds fill and no
    activation is ever re-read just to normalise it;
  * the U-Net skips are never stored raw: the encoder epilogue writes gelu(norm_unet(skip)) in bf16, the
    only form UpresBlock.unet_conv consumes (AG:517);
  * the pair tensor is fp32 [B, P, P, 128]; attention logits are never materialised.

Weights are packed once per parameter version (bf16, [taps][C_out][C_in]; norms folded to per-channel
(scale, shift) vectors) and cached on the engine.
"""
from __future__ import annotations

import math
import os
import weakref
from typing import Dict, Optional

import torch

from . import ops
from .ops import ACT_GELU1702, ACT_GELU_ERF, ACT_NONE, Act

_ENGINES: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def engine_for(unet, model=None) -> "TrunkEngine":
    eng = _ENGINES.get(unet)
    if eng is None:
        eng = TrunkEngine(unet)
        _ENGINES[unet] = eng
    if model is not None:
        eng.model = weakref.ref(model)
    return eng


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _Norm:
    """One packed BatchRMSNorm (AG:302-361): `s` = gamma / sqrt(running_var + eps) and `t` = beta are the vectors the
    fused kernels read; `mod` is the module whose running_var an updating forward re-estimates in place."""

    __slots__ = ("mod", "s", "t", "fold_bias", "fold_shift")

    def __init__(self, mod, s, t):
        self.mod, self.s, self.t = mod, s, t
        self.fold_bias = self.fold_shift = None   # sandwich post-norms: shift = bias * s + t folded behind the Linear

    def updating(self) -> bool:
        u = getattr(self.mod, "update_running_var", None)       # default(arg, self.update_running_var, self.training), AG:331
        return bool(self.mod.training if u is None else u)


class TrunkEngine:
    def __init__(self, unet):
        self.unet = weakref.ref(unet)
        self.model = None
        self._pack: Optional[Dict] = None
        self._pack_key = None
        self._tensors = None   # (model, [parameters and buffers]) the version key walks
        self._norms: Dict = {}  # data_ptr of a packed scale vector -> _Norm
        self._own = None       # (first owned bp, end, window bp) of a sequence-parallel conv window
        self._rope: Dict = {}
        self._rel: Dict = {}
        self.collect = None  # optional callback(name, tensor) for per-stage parity tests
        self.sp = None       # alphagenome_b200.parallel.SeqParallel when the context is sharded over ranks
        self._graphs: Dict = {}          # insertion order = least recently used first
        self.graph_pool_budget = 0.6     # fraction of device memory the kept graphs' pools may hold

    # ------------------------------------------------------------------------------------------ packing
    def _version_key(self, device):
        """(data_ptr, _version) of every parameter / buffer: repacks after load_state_dict, .to(), or an in-place update
        of a parameter.  Updates made through `.data` do not bump `_version` — call invalidate() after those."""
        model = self.model() if self.model is not None else None
        if self._tensors is None or self._tensors[0] is not model:
            mods = [self.unet()] + ([model] if model is not None else [])
            self._tensors = (model, [t for m in mods for t in list(m.parameters()) + list(m.buffers())])
        return (str(device),) + tuple((t.data_ptr(), t._version) for t in self._tensors[1])

    def pack(self, device) -> Dict:
        key = self._version_key(device)
        if self._pack is not None and self._pack_key == key:
            return self._pack
        unet = self.unet()
        tw0 = unet.transformer
        at0 = tw0.layers[0][0].block
        s2p0 = next((lyr[2] for lyr in tw0.layers if lyr[2] is not None), None)
        bad = []
        if getattr(tw0, "polar_pos_emb", False) or getattr(tw0, "rotary_emb", None) is None:
            bad.append("polar_pos_emb=True (PoPE)")
        if not hasattr(at0.q_norm, "bias") or at0.q_norm.bias is None:
            bad.append("use_qk_rmsnorm=True")
        if at0.to_attn_bias[2].weight.shape != (8, 128):
            bad.append(f"attention heads / pair dim {tuple(at0.to_attn_bias[2].weight.shape)} (kernels are built for 8 heads x 128)")
        if at0.q_norm.weight.numel() != 128 or at0.v_norm.weight.numel() != 192:
            bad.append("dim_head_qk / dim_head_v other than 128 / 192")
        if at0.to_qkv.weight.shape[0] != 8 * 128 + 128 + 192:
            bad.append("kv_heads > 1")
        if s2p0 is not None and (getattr(s2p0, "pool_size", 16) != 16 or s2p0.qk_to_pairwise.weight.shape != (128, 32)):
            bad.append("SingleToPairwise pool_size / heads / dim_pairwise other than 16 / 32 / 128")
        if bad:
            raise ops._lib.AgbError("alphagenome_b200 implements the default trunk configuration only; unsupported: " + "; ".join(bad))
        if device.type != "cuda" and not getattr(self, "_allow_cpu_pack", False):  # (tests pack on CPU to compare layouts)
            raise ops._lib.AgbError("alphagenome_b200 runs only on a CUDA (sm_100a) device; move the model and inputs to cuda")
        f32 = dict(device=device, dtype=torch.float32)

        def vec(t):
            return t.detach().to(**f32).contiguous()

        def wconv(conv):  # [Co, Ci, taps] -> bf16 [taps, Co, Ci]
            return conv.weight.detach().to(device).permute(2, 0, 1).contiguous().to(torch.bfloat16)

        def wlin(lin):  # [N, K] -> bf16 [1, N, K]
            return lin.weight.detach().to(device).to(torch.bfloat16).contiguous().unsqueeze(0)

        self._norms = {}

        def aff(norm, s_into=None, t_into=None):
            s, t = norm.affine()
            s, t = vec(s), vec(t)
            if s_into is not None:      # a row of a stacked parameter set the kernel reads as one tensor
                s_into.copy_(s)
                t_into.copy_(t)
                s, t = s_into, t_into
            self._norms[s.data_ptr()] = _Norm(norm, s, t)
            return s, t

        def convblock(cb):
            s, t = aff(cb.norm)
            return dict(W=wconv(cb.conv), b=vec(cb.conv.bias), s=s, t=t)

        P: Dict = {}
        de = unet.dna_embed
        C0, nb, width = de.conv.weight.shape
        k_emb = _round_up(width * nb, 64)
        w15 = torch.zeros(C0, k_emb, device=device, dtype=torch.float32)
        # im2col column t*n_base + base  <-  conv.weight[c, base, t]
        w15[:, : width * nb] = de.conv.weight.detach().to(device).permute(0, 2, 1).reshape(C0, width * nb)
        P["embed"] = dict(W=w15.to(torch.bfloat16).unsqueeze(0).contiguous(), b=vec(de.conv.bias), width=width, n_base=nb, K=k_emb)
        P["pointwise"] = convblock(de.pointwise)
        P["pointwise"]["b_embed"] = (P["pointwise"]["b"] + P["embed"]["b"]).contiguous()   # both biases, for the fused residual
        P["downs"] = [dict(conv=convblock(d.conv), conv_out=convblock(d.conv_out)) for d in unet.downs]
        P["ups"] = [dict(conv=convblock(u.conv), unet_conv=convblock(u.unet_conv), conv_out=convblock(u.conv_out),
                         scale=vec(u.residual_scale)) for u in unet.ups]

        tw = unet.transformer
        # to_attn_bias of a pair-updating layer and of the following layer that reuses the same pair tensor are projected
        # in one pass over it (ag_pair_bias_fwd takes up to two parameter sets): their (scale, shift, W) live stacked
        has_pair = [lyr[2] is not None for lyr in tw.layers]
        groups = []
        for li in range(len(tw.layers)):
            if has_pair[li] or li == 0:
                groups.append([li] + ([li + 1] if li + 1 < len(tw.layers) and not has_pair[li + 1] else []))
            elif not any(li in g for g in groups):
                groups.append([li])
        bias_slot = {li: (gi, k) for gi, g in enumerate(groups) for k, li in enumerate(g)}
        at0 = tw.layers[0][0].block
        Hb, Cp = at0.to_attn_bias[2].weight.shape
        bias_sets = [dict(layers=g, s=torch.empty(len(g), Cp, **f32), t=torch.empty(len(g), Cp, **f32), W=torch.empty(len(g), Hb, Cp, **f32))
                     for g in groups]
        layers = []
        for li, (attn_w, ff_w, s2p, pattn_w, pff_w) in enumerate(tw.layers):
            L: Dict = {}
            at = attn_w.block
            s_pre, t_pre = aff(attn_w.pre_rmsnorm)
            s_post, t_post = aff(attn_w.post_rmsnorm)
            gi, gk = bias_slot[li]
            ab_s, ab_t = aff(at.to_attn_bias[0], bias_sets[gi]["s"][gk], bias_sets[gi]["t"][gk])
            bias_sets[gi]["W"][gk].copy_(vec(at.to_attn_bias[2].weight))
            nrm = self._norms[s_post.data_ptr()]
            nrm.fold_bias, nrm.fold_shift = vec(at.to_out.bias), (vec(at.to_out.bias) * s_post + t_post).contiguous()
            L["attn"] = dict(
                pre=(s_pre, t_pre), Wqkv=wlin(at.to_qkv), Wout=wlin(at.to_out),
                out_scale=s_post, out_shift=nrm.fold_shift,
                qw=vec(at.q_norm.weight), qb=vec(at.q_norm.bias), kw=vec(at.k_norm.weight), kb=vec(at.k_norm.bias),
                vw=vec(at.v_norm.weight), vb=vec(at.v_norm.bias), ab_s=ab_s, ab_t=ab_t, ab_W=bias_sets[gi]["W"][gk],
                heads=at.heads, d=at.dim_head_qk, dv=at.dim_head_v, scale=at.scale, softcap=at.softclamp_value)
            fs_pre, ft_pre = aff(ff_w.pre_rmsnorm)
            fs_post, ft_post = aff(ff_w.post_rmsnorm)
            nrm = self._norms[fs_post.data_ptr()]
            nrm.fold_bias, nrm.fold_shift = vec(ff_w.block[3].bias), (vec(ff_w.block[3].bias) * fs_post + ft_post).contiguous()
            L["ff"] = dict(pre=(fs_pre, ft_pre), W1=wlin(ff_w.block[0]), b1=vec(ff_w.block[0].bias), W2=wlin(ff_w.block[3]),
                           out_scale=fs_post, out_shift=nrm.fold_shift)
            if s2p is not None:
                H, dp = s2p.heads, s2p.dim_pairwise
                L["s2p"] = dict(
                    nw=vec(s2p.norm.weight), nb=vec(s2p.norm.bias), Wqk=wlin(s2p.to_qk), Wouter=wlin(s2p.to_outer_sum[1]),
                    relbias=vec(s2p.qk_rel_pos_bias.reshape(-1)), Wp=vec(s2p.qk_to_pairwise.weight), bp=vec(s2p.qk_to_pairwise.bias),
                    Wrel=wlin(s2p.to_rel_pos_encoding), brel=vec(s2p.to_rel_pos_encoding.bias), heads=H, dp=dp, pool=s2p.pool_size)
                pa = pattn_w.block
                L["pattn"] = dict(nw=vec(pattn_w.pre_rmsnorm.weight), nb=vec(pattn_w.pre_rmsnorm.bias), Wqk=wlin(pa.to_qk),
                                  Wv=wlin(pa.to_v), bv=vec(pa.to_v.bias), scale=pa.scale)
                L["pff"] = dict(nw=vec(pff_w.pre_rmsnorm.weight), nb=vec(pff_w.pre_rmsnorm.bias), W1=wlin(pff_w.block[0]),
                                b1=vec(pff_w.block[0].bias), W2=wlin(pff_w.block[3]), b2=vec(pff_w.block[3].bias))
            layers.append(L)
        for g in bias_sets:
            if len(g["layers"]) == 2 or "s2p" in layers[g["layers"][0]] or g["layers"][0] == 0:
                layers[g["layers"][0]]["bias_group"] = g
        P["layers"] = layers
        P["inv_freq"] = tw.rotary_emb.inv_freq.detach().float().cpu()

        model = self.model() if self.model is not None else None
        if model is not None:
            def outemb(oe):
                s, t = aff(oe.norm)
                # per-organism shift table: norm shift + organism embedding (AG:1231-1233)
                table = (t[None, :] + vec(oe.embed.weight)).contiguous()
                return dict(W=wlin(oe.double_features), b=vec(oe.double_features.bias), s=s, table=table,
                            Wskip=wlin(oe.skip_proj) if oe.skip_proj is not None else None)

            P["out128"] = outemb(model.outembed_128bp)
            P["out1"] = outemb(model.outembed_1bp)
            P["outpair"] = dict(w=vec(model.outembed_pair.norm.weight), b=vec(model.outembed_pair.norm.bias),
                                emb=vec(model.outembed_pair.embed.weight))
            P["org"] = vec(model.organism_embed.embed.weight)
        self._pack, self._pack_key = P, key
        self._rel.clear()
        self._rope.clear()
        return P

    def rope_tables(self, n: int, device, offset: int = 0):
        key = (n, offset, str(device))
        if key not in self._rope:
            inv = self._pack["inv_freq"]
            ang = torch.arange(offset, offset + n, dtype=torch.float32)[:, None] * inv[None, :]  # AG:555-556
            self._rope[key] = (ang.cos().to(device).contiguous(), ang.sin().to(device).contiguous())
        return self._rope[key]

    def rel_encoding(self, li: int, Pn: int, device):
        """R = to_rel_pos_encoding(features) for 2P relative distances (AG:838-839): bf16 [2P, H*dp]; data-independent."""
        key = (li, Pn)
        if key not in self._rel:
            sp = self._pack["layers"][li]["s2p"]
            feats = self.unet().transformer.rel_pos_features.features(Pn).to(device).to(torch.bfloat16).contiguous()  # exact 0/+-1
            R = torch.empty(1, 2 * Pn, sp["heads"] * sp["dp"], device=device, dtype=torch.bfloat16)
            ops.gemm(feats.unsqueeze(0), sp["Wrel"], pre_shift=sp["brel"], actA=Act(R))
            self._rel[key] = R[0]
        return self._rel[key]

    # ------------------------------------------------------------------------------------------ updating BatchRMSNorm
    def _updating(self, scale) -> Optional[_Norm]:
        """The packed norm behind `scale` if its module re-estimates running_var on this forward (AG:331), else None."""
        if scale is None:
            return None
        nrm = self._norms.get(scale.data_ptr())
        return nrm if nrm is not None and nrm.updating() else None

    def _own_rows(self, x):
        """Rows of a conv-stack tensor this rank owns (statistics exclude the recomputed halo rows of a shard)."""
        if self._own is None:
            return x
        lo, hi, win = self._own
        f = win // x.shape[1]
        return x[:, lo // f: hi // f]

    def _bn_update(self, nrm: _Norm, src):
        """running_var.lerp_(var(src), momentum) in the module's buffer and the packed scale (AG:331-340): one pass of
        per-channel sums over this rank's rows, ONE all-reduce under sequence parallelism (the reference: three)."""
        rv = nrm.mod.running_var
        if rv.device != src.device or rv.dtype != torch.float32 or not rv.is_contiguous():
            raise ops._lib.AgbError("updating BatchRMSNorm needs the module's fp32 running_var on the forward's device")
        src = self._own_rows(src)
        B, rows, Cc = src.shape
        sums = torch.empty(2, Cc, device=src.device, dtype=torch.float64)
        ops.bn_stats(src, sums)
        count = B * rows
        if self.sp is not None and self.sp.world > 1 and getattr(nrm.mod, "distributed", True):
            sums = self.sp.all_reduce_sum(sums)
            count *= self.sp.world
        gamma = nrm.mod.gamma.detach()
        if gamma.device != src.device or gamma.dtype != torch.float32:
            gamma = gamma.to(device=src.device, dtype=torch.float32)
        ops.bn_update(sums, count, rv, gamma.contiguous(), getattr(nrm.mod, "momentum", 0.1), nrm.mod.eps, nrm.s,
                      fold_bias=nrm.fold_bias, fold_beta=nrm.t if nrm.fold_shift is not None else None, fold_shift_out=nrm.fold_shift)

    def _gemm(self, A, W, **kw):
        """ops.gemm, except that activated outputs whose BatchRMSNorm re-estimates its variance are produced after the
        GEMM: the fp32 value is written first, its per-channel statistics update the norm, then the output is made."""
        deferred = [(k, kw[k]) for k in ("actA", "actB", "actP") if kw.get(k) is not None and self._updating(kw[k].scale) is not None]
        if not deferred:
            return ops.gemm(A, W, **kw)
        for k, _ in deferred:
            kw[k] = None
        Arows = A.shape[-2] if kw.get("M") is None else kw["M"]
        for slot, keys in (("out", ("actA", "actB")), ("pool", ("actP",))):
            acts = [a for k, a in deferred if k in keys]
            if not acts:
                continue
            src = kw.get(slot)
            if src is None:     # an fp32 destination of the same norm doubles as the raw buffer (normalised in place)
                f32_dst = [a.dst for a in acts if a.dst.dtype == torch.float32]
                src = f32_dst[0] if f32_dst else torch.empty(acts[0].dst.shape, device=A.device, dtype=torch.float32)
                kw[slot] = src
        ops.gemm(A, W, **kw)
        done = set()
        for slot, keys in (("out", ("actA", "actB")), ("pool", ("actP",))):
            acts = [a for k, a in deferred if k in keys]
            src = kw.get(slot)
            acts.sort(key=lambda a: a.dst.data_ptr() == src.data_ptr())   # the destination that aliases the raw buffer goes last
            for a in acts:
                nrm = self._norms[a.scale.data_ptr()]
                if id(nrm) not in done:
                    self._bn_update(nrm, src)
                    done.add(id(nrm))
                ops.affine_act(src, a.scale, a.shift, out_f32=a.dst if a.dst.dtype == torch.float32 else None,
                               out_bf16=a.dst if a.dst.dtype == torch.bfloat16 else None, act=a.act)

    def _post_norm_gemm(self, A, W, blk, single, act: Optional[Act], actB: Optional[Act] = None):
        """Sandwich block tail (AG:630-639, 1024-1025): single += post_norm(A @ W^T + bias), then the next block's
        pre-normalised operand(s).  Frozen: one GEMM with the norm folded into its epilogue."""
        post = self._updating(blk["out_scale"])
        if post is None and (act is None or self._updating(act.scale) is None):
            ops.gemm(A, W, pre_scale=blk["out_scale"], pre_shift=blk["out_shift"], res=single, out=single, actA=act, actB=actB)
            return
        if post is not None:
            raw = torch.empty_like(single)
            ops.gemm(A, W, pre_shift=post.fold_bias, out=raw)
            self._bn_update(post, raw)
            ops.affine_act(raw, post.s, post.t, res=single, out_f32=single)
        else:
            ops.gemm(A, W, pre_scale=blk["out_scale"], pre_shift=blk["out_shift"], res=single, out=single)
        for a in (act, actB):
            if a is None:
                continue
            nrm = self._updating(a.scale)
            if nrm is not None:
                self._bn_update(nrm, single)
            ops.affine_act(single, a.scale, a.shift, out_f32=a.dst if a.dst.dtype == torch.float32 else None,
                           out_bf16=a.dst if a.dst.dtype == torch.bfloat16 else None, act=a.act)

    # ------------------------------------------------------------------------------------------ forward pieces
    def _note(self, name, t):
        if self.collect is not None:
            self.collect(name, t)

    def _encoder(self, P, seq):
        B, L = seq.shape
        dev = seq.device
        unet = self.unet()
        n_down = len(P["downs"])
        if L % (2 ** (n_down + 1) * 16) != 0:
            raise ops._lib.AgbError(f"sequence length {L} must be a multiple of {2 ** (n_down + 1) * 16} (pool 2^{n_down + 1}, then 16)")
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        e = P["embed"]
        C0 = unet.dims[0]
        X = torch.empty(B, L, e["K"], **bf)
        ops.onehot_im2col(seq, X, e["width"], e["n_base"])
        pw = P["pointwise"]
        a0 = torch.empty(B, L, C0, **bf)
        # x0 = conv15(one_hot) (AG:1177) is needed twice: normalised + activated as pointwise's operand (a0), and as the
        # residual of x0 + pointwise(x0) (AG:1178).  The residual is re-derived on the tensor core — the K = 64 im2col
        # product rides along as a second operand pair of the width-5 GEMM (+1.7 % MMA work), both biases summed —
        # so at 1 Mbp the 3.2 GB fp32 x0 is neither written nor read back.  A pointwise norm that re-estimates its
        # variance needs x0's statistics, hence x0 itself: that (default-constructed, updating) case keeps the plain form.
        fuse_x0 = self._updating(pw["s"]) is None and os.environ.get("AGB_EMBED_FUSE", "1") != "0"   # (=0: A/B timing)
        x0 = None if fuse_x0 else torch.empty(B, L, C0, **f32)
        self._gemm(X, e["W"], pre_shift=e["b"], out=x0, actA=Act(a0, pw["s"], pw["t"], ACT_GELU1702))
        ups = P["ups"]
        skips = [None] * (n_down + 1)  # skips[j] feeds ups[j]
        un = ups[n_down]["unet_conv"]
        skips[n_down] = torch.empty(B, L, C0, **bf)
        x = torch.empty(B, L // 2, C0, **f32)
        a = torch.empty(B, L // 2, C0, **bf)
        d0 = P["downs"][0]["conv"]
        if fuse_x0:
            self._gemm(a0, pw["W"], pre_shift=pw["b_embed"], A2=X, W2=e["W"], actA=Act(skips[n_down], un["s"], un["t"], ACT_GELU1702),
                       pool=x, actP=Act(a, d0["s"], d0["t"], ACT_GELU1702))
        else:
            self._gemm(a0, pw["W"], pre_shift=pw["b"], res=x0, actA=Act(skips[n_down], un["s"], un["t"], ACT_GELU1702),
                       pool=x, actP=Act(a, d0["s"], d0["t"], ACT_GELU1702))
        del x0, a0, X
        self._note("embed.pooled", x)
        length = L // 2
        for i, d in enumerate(P["downs"]):
            c, co = d["conv"], d["conv_out"]
            C_in, C_out = c["W"].shape[2], c["W"].shape[1]
            y1 = torch.empty(B, length, C_out, **f32)
            ay1 = torch.empty(B, length, C_out, **bf)
            self._gemm(a, c["W"], pre_shift=c["b"], res=x, res_ch=C_in, out=y1, actA=Act(ay1, co["s"], co["t"], ACT_GELU1702))
            j = n_down - 1 - i
            un = ups[j]["unet_conv"]
            skips[j] = torch.empty(B, length, C_out, **bf)
            xn = torch.empty(B, length // 2, C_out, **f32)
            actP = None
            an = None
            if i + 1 < n_down:
                nx = P["downs"][i + 1]["conv"]
                an = torch.empty(B, length // 2, C_out, **bf)
                actP = Act(an, nx["s"], nx["t"], ACT_GELU1702)
            self._gemm(ay1, co["W"], pre_shift=co["b"], res=y1, actA=Act(skips[j], un["s"], un["t"], ACT_GELU1702), pool=xn, actP=actP)
            x, a = xn, an
            length //= 2
            self._note(f"downs.{i}.pooled", x)
        return x, skips

    def _pool_rows(self, P, li, single, plan=None):
        """First step of the pair block: 16-token average pooling + RMSNorm (+ its erf-GELU) of `single` (AG:828-829, 852).
        Sharded: the pooled rows of every rank are needed; the all-gather is started here and waited for in _pair_block,
        so it travels while the attention projections of the same layer are computed."""
        sp = P["layers"][li]["s2p"]
        B, n_loc, D = single.shape
        Pl = n_loc // sp["pool"]
        xsg = torch.empty(2, B, Pl, D, device=single.device, dtype=torch.bfloat16)    # [0] normalised pooled rows, [1] their erf-GELU
        ops.pool_rmsnorm(single, sp["nw"], sp["nb"], xsg[0], xsg[1], sp["pool"])
        if plan is not None and plan.world > 1:
            if self.sp.peer:
                # one-launch push into every rank's symmetric [world, 2B, Pl, D] buffer; made visible by the layer's barrier
                buf, hdl, ptrs = self.sp.symm_buffer(("rows", li & 2, B, Pl, D), (plan.world, 2 * B, Pl, D), torch.bfloat16, single.device)
                ops.peer_push(xsg, ptrs, plan.rank * xsg.numel() * 2)
                return ("peer", buf, hdl)
            return self.sp.gather_rows_async(xsg.view(2 * B, Pl, D))
        return xsg

    def _pair_block(self, P, li, single, pair, plan=None, pooled=None):
        """SingleToPairwise + row attention + pair FF (AG:1019-1022).  With a ShardPlan, `single` holds this rank's
        tokens and `pair` its rows [B, P_loc, P, 128]; pooled rows are all-gathered, everything else is row-local."""
        L = P["layers"][li]
        sp, pa, pf = L["s2p"], L["pattn"], L["pff"]
        B, n_loc, D = single.shape
        dev = single.device
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        pool, H, dp = sp["pool"], sp["heads"], sp["dp"]
        Pl = n_loc // pool                      # local rows
        if pooled is None:
            pooled = self._pool_rows(P, li, single, plan)
        if plan is not None and plan.world > 1:
            if isinstance(pooled, tuple):     # peer path: [world, 2B, Pl, D], already behind the layer's barrier
                g = pooled[1].view(plan.world, 2, B, Pl, D).permute(1, 2, 0, 3, 4).reshape(2, B, plan.P, D)
            else:
                g = pooled.wait().view(2, B, plan.P, D)
            xs, xg, Pn, i0 = g[0], g[1], plan.P, plan.row_off
        else:
            xs, xg, Pn, i0 = pooled[0], pooled[1], Pl, 0
        qk_plain = torch.empty(B, Pn, 2 * H * dp, **bf)
        qk_bias = torch.empty(B, Pn, 2 * H * dp, **bf)
        ops.gemm(xs, sp["Wqk"], actA=Act(qk_plain), actB=Act(qk_bias, None, sp["relbias"]))
        outer = torch.empty(B, Pn, 2 * dp, **f32)
        ops.gemm(xg, sp["Wouter"], out=outer)
        R = self.rel_encoding(li, Pn, dev).view(2 * Pn, H, dp).permute(1, 0, 2)  # [H, 2P, dp] view
        Np, N2 = _round_up(Pn, 16), _round_up(2 * Pn, 16)
        qk_out = torch.empty(B, H, Pl, Np, **f32)
        relq = torch.empty(B, H, Pl, N2, **f32)
        relk = torch.empty(B, H, Pn, N2, **f32)
        for b in range(B):
            qp = qk_plain[b, i0:i0 + Pl, : H * dp].view(Pl, H, dp).permute(1, 0, 2)
            kp = qk_plain[b, :, H * dp:].view(Pn, H, dp).permute(1, 0, 2)
            qb = qk_bias[b, i0:i0 + Pl, : H * dp].view(Pl, H, dp).permute(1, 0, 2)
            kb = qk_bias[b, :, H * dp:].view(Pn, H, dp).permute(1, 0, 2)
            ops.gemm(qp, kp, w_batched=True, n_pad=Np, out=qk_out[b])       # q_i . k_j            (AG:835)
            ops.gemm(qb, R, w_batched=True, n_pad=N2, out=relq[b])          # (q_i + bq) . R[p]    (AG:843)
            ops.gemm(kb, R, w_batched=True, n_pad=N2, out=relk[b])          # (k_j + bk) . R[p]    (AG:844)
        new_pair = torch.empty(B, Pl, Pn, dp, **f32)
        cells = B * Pl * Pn
        pn = torch.empty(1, cells, dp, **bf)   # RMSNormWithBias(pair): the row attention's pre-norm, fused into the combine
        ops.pair_combine(qk_out, relq, relk, sp["Wp"], sp["bp"], outer, pair, new_pair, P=Pn, i_off=i0,
                         norm_w=pa["nw"], norm_b=pa["nb"], norm_out=pn)
        pair = new_pair
        del qk_out, relq, relk, qk_plain, qk_bias
        self._note(f"tower.layer{li}.pair_s2p", pair)
        # --- row attention (AG:1021): attends over the second index, so a row shard is self-contained ---
        qkp = torch.empty(1, cells, 2 * dp, **bf)
        ops.gemm(pn, pa["Wqk"], actA=Act(qkp))
        # V^T straight out of the GEMM: with Wv as the row operand (shared by every pair row) and the normalised pair
        # row as the "weight" of batch i, out[i] = Wv . pn[i]^T = [dp, Pn] with keys contiguous — the layout the
        # attention's P.V operand wants; no transpose pass.  (to_v's bias is added after the softmax, see out_bias.)
        Pk = _round_up(Pn, 16)
        vpt = torch.empty(B * Pl, dp, Pk, **bf)
        ops.gemm(pa["Wv"], pn.view(B * Pl, Pn, dp), w_batched=True, n_pad=Pk, actA=Act(vpt))
        q4 = qkp.view(B, Pl, Pn, 2 * dp)
        # the epilogue also emits RMSNormWithBias(pair) for the pair feed-forward (its pre-norm, AG:972-975) into pn
        ops.attention(q4[..., :dp], q4[..., dp:], vpt.view(B, Pl, dp, Pk), pair, mode=1, scale=pa["scale"],
                      out_bias=pa["bv"], res=pair, norm_w=pf["nw"], norm_b=pf["nb"], norm_out=pn.view(B, Pl, Pn, dp))
        self._note(f"tower.layer{li}.pair_attn", pair)
        # --- pair feed-forward (AG:1022) ---
        hid = torch.empty(1, cells, pf["W1"].shape[1], **bf)
        ops.gemm(pn, pf["W1"], pre_shift=pf["b1"], relu=True, actA=Act(hid))
        pflat = pair.view(1, cells, dp)
        ops.gemm(hid, pf["W2"], pre_shift=pf["b2"], res=pflat, out=pflat)
        return pair

    def _tower(self, P, x, org_embed, final_act: Act, final_plain: Optional[torch.Tensor], plan=None):
        """x: fp32 [B, n_loc, D] pooled encoder output (this rank's tokens).  Returns (single fp32, pair fp32)."""
        B, n, D = x.shape
        dev = x.device
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        layers = P["layers"]
        tw = self.unet().transformer
        sharded = plan is not None and plan.world > 1
        n_all = plan.n if sharded else n
        if n_all > tw.max_positions:
            raise ops._lib.AgbError(f"{n_all} tokens exceed max_positions={tw.max_positions} (AG:538, 922)")
        single = torch.empty(B, n, D, **f32)
        a = torch.empty(B, n, D, **bf)
        s0, t0 = layers[0]["attn"]["pre"]
        if org_embed is None:
            org_embed = torch.zeros(B, D, **f32)
        ops.add_embed_norm(x, org_embed.to(torch.float32).contiguous(), s0, t0, single, a)
        nrm0 = self._updating(s0)
        if nrm0 is not None:     # the first pre-norm sees x + organism embedding (AG:1119-1120, 636)
            self._bn_update(nrm0, single)
            ops.affine_act(single, s0, t0, out_bf16=a)
        cos_t, sin_t = self.rope_tables(n, dev, offset=plan.tok_off if sharded else 0)
        pair = None
        bias_of = {}
        a_ff = torch.empty(B, n, D, **bf)
        for li, L in enumerate(layers):
            # Order inside a layer: the pair block (AG:1019-1022) and the attention projections both read only the layer's
            # input, so the projections go first and — sequence parallel — their K/V all-gather, like the pooled-row
            # gather before it, is in flight on NCCL's stream while the pair block runs on this one.
            pooled = self._pool_rows(P, li, single, plan) if "s2p" in L else None
            at = L["attn"]
            Hh, d, dv = at["heads"], at["d"], at["dv"]
            qkv = torch.empty(B, n, Hh * d + d + dv, **f32)
            ops.gemm(a, at["Wqkv"], out=qkv)
            Q = torch.empty(B, n, Hh * d, **bf)
            kv_pending = None
            if sharded and self.sp.peer:
                # K/V of every rank (AG:748 attends over the whole context): the LayerNorm + RoPE kernel stores this rank's
                # K | V rows straight into every rank's gathered buffer over NVLink (fused compute + all-gather); one
                # barrier per layer then covers them and the pooled rows pushed above.  Two buffers alternate between
                # layers: a rank can only be writing layer l+2's rows after every rank has passed layer l+1's barrier,
                # i.e. after every rank's layer-l attention has finished reading.
                KV, hdl, ptrs = self.sp.symm_buffer(("kv", li & 1, B, n_all, d + dv), (B, n_all, d + dv), torch.bfloat16, dev)
                ops.qkv_post_gather(qkv, Q, ptrs, n_all, plan.tok_off, cos_t, sin_t, at["qw"], at["qb"], at["kw"], at["kb"], at["vw"], at["vb"], Hh, d, dv)
                self.sp.peer_barrier(hdl)
                K, V = KV[..., :d], KV[..., d:]
            else:
                K = torch.empty(B, n, d, **bf)
                V = torch.empty(B, n, dv, **bf)
                ops.qkv_post(qkv, Q, K, V, cos_t, sin_t, at["qw"], at["qb"], at["kw"], at["kb"], at["vw"], at["vb"], Hh, d, dv)
                if sharded:  # NCCL fallback: ONE all-gather per layer of [K | V], in flight while the pair block runs
                    kv_pending = self.sp.gather_rows_async(torch.cat((K, V), dim=-1))
            if "s2p" in L:
                pair = self._pair_block(P, li, single, pair, plan, pooled=pooled)
            if kv_pending is not None:
                KV = kv_pending.wait()
                K, V = KV[..., :d], KV[..., d:]                         # strided views (TMA / transpose take a row stride)
            nk = _round_up(n_all, 8)
            Vt = torch.empty(B, dv, nk, **bf)
            ops.transpose_bf16(V, Vt)
            if "bias_group" in L or li not in bias_of:
                # to_attn_bias' BatchRMSNorm (AG:688-693) of the layer(s) projected in this pass
                for j in (L["bias_group"]["layers"] if "bias_group" in L else [li]):
                    nrm = self._updating(layers[j]["attn"]["ab_s"])
                    if nrm is not None:
                        self._bn_update(nrm, pair.view(B, -1, pair.shape[-1]))
            if "bias_group" in L:
                bg = L["bias_group"]
                bias_sets = torch.empty(len(bg["layers"]), B, Hh, pair.shape[1], pair.shape[2], **f32)
                if len(bg["layers"]) == 2:
                    ops.pair_bias(pair, bg["s"], bg["t"], bg["W"], bias_sets)
                else:
                    ops.pair_bias(pair, bg["s"][0], bg["t"][0], bg["W"][0], bias_sets[0])
                bias_of = {j: bias_sets[k] for k, j in enumerate(bg["layers"])}
            elif li not in bias_of:   # a layer without its own pair block that follows another such layer
                bias_of = {li: torch.empty(B, Hh, pair.shape[1], pair.shape[2], **f32)}
                ops.pair_bias(pair, at["ab_s"], at["ab_t"], at["ab_W"], bias_of[li])
            bias = bias_of[li]
            O = torch.empty(B, n, Hh * dv, **bf)
            ops.attention(Q.view(B, n, Hh, d).permute(0, 2, 1, 3), K.unsqueeze(1), Vt.view(B, 1, dv, nk),
                          O.view(B, n, Hh, dv).permute(0, 2, 1, 3), mode=0, scale=at["scale"], bias=bias, softcap=at["softcap"])
            ff = L["ff"]
            self._post_norm_gemm(O, at["Wout"], at, single, Act(a_ff, ff["pre"][0], ff["pre"][1]))
            hid = torch.empty(B, n, ff["W1"].shape[1], **bf)
            ops.gemm(a_ff, ff["W1"], pre_shift=ff["b1"], relu=True, actA=Act(hid))
            last = li + 1 == len(layers)
            if not last:
                nxt = layers[li + 1]["attn"]["pre"]
                self._post_norm_gemm(hid, ff["W2"], ff, single, Act(a, nxt[0], nxt[1]))
            else:
                self._post_norm_gemm(hid, ff["W2"], ff, single, final_act, Act(final_plain) if final_plain is not None else None)
            self._note(f"tower.layer{li}.single", single)
            self._note(f"tower.layer{li}.pair", pair)
        return single, pair

    def _decoder(self, P, x, a, skips, want_plain: bool):
        B, length, _ = x.shape
        dev = x.device
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        ups = P["ups"]
        plain = None
        for j, u in enumerate(ups):
            c, un, co = u["conv"], u["unet_conv"], u["conv_out"]
            C_out = c["W"].shape[1]
            low = torch.empty(B, length, C_out, **f32)
            self._gemm(a, c["W"], pre_shift=c["b"], res=x, res_ch=C_out, post_scale=u["scale"], out=low)
            y = torch.empty(B, 2 * length, C_out, **f32)
            ay = torch.empty(B, 2 * length, C_out, **bf)
            self._gemm(skips[j], un["W"], pre_shift=un["b"], res=low, res_shift=1, out=y, actA=Act(ay, co["s"], co["t"], ACT_GELU1702))
            skips[j] = None
            xn = torch.empty(B, 2 * length, C_out, **f32)
            if j + 1 < len(ups):
                nx = ups[j + 1]["conv"]
                an = torch.empty(B, 2 * length, C_out, **bf)
                self._gemm(ay, co["W"], pre_shift=co["b"], res=y, out=xn, actA=Act(an, nx["s"], nx["t"], ACT_GELU1702))
            else:
                an = None
                if want_plain:
                    plain = torch.empty(B, 2 * length, C_out, **bf)
                self._gemm(ay, co["W"], pre_shift=co["b"], res=y, out=xn, actA=Act(plain) if want_plain else None)
            x, a = xn, an
            length *= 2
            self._note(f"ups.{j}", x)
        return x, plain

    # ------------------------------------------------------------------------------------------ public
    @torch.no_grad()
    def run(self, seq, org_embed, organism_index=None, with_embeds=False, want_head_operands=False):
        """seq: int64 [B, L] — the WHOLE context on every rank.  With `self.sp` set (sequence parallel) each rank
        computes and returns only its shard: bp [r L/G, (r+1) L/G), the matching tokens, and its pair rows."""
        if not seq.is_cuda:
            raise ops._lib.AgbError("alphagenome_b200 has no CPU path: inputs must be CUDA tensors on a B200")
        dev = seq.device
        P = self.pack(dev)
        bf = dict(device=dev, dtype=torch.bfloat16)
        plan = self.sp.plan(seq.shape[1]) if self.sp is not None else None
        sharded = plan is not None and plan.world > 1
        if sharded:
            seq = seq[:, plan.win_start:plan.win_end]
        seq = seq.to(torch.int64).contiguous()
        self._own = (plan.halo_left, plan.halo_left + plan.L_loc, seq.shape[1]) if sharded else None
        x, skips = self._encoder(P, seq)
        self._own = None
        hl, hr = (plan.halo_tok_left, plan.halo_tok_right) if sharded else (0, 0)
        if sharded:
            x = x[:, hl:hl + plan.n_loc].contiguous()  # the halo tokens only existed to make the owned ones exact
        B, n, D = x.shape
        u0 = P["ups"][0]["conv"]
        a_dec = torch.empty(B, n, D, **bf)
        single_b16 = torch.empty(B, n, D, **bf) if with_embeds else None
        single, pair = self._tower(P, x, org_embed, Act(a_dec, u0["s"], u0["t"], ACT_GELU1702), single_b16, plan)
        x_dec = single
        halo_pending = pairT_pending = None
        if sharded:
            # the decoder's receptive field reaches 6 tokens into the neighbours: fetch halo tokens of the tower
            # output (fp32 residual + the pre-activated bf16 operand), one collective for both; the pair blocks needed
            # by the symmetrize of the pair output embedding (AG:1253) travel at the same time.  Both are in flight on
            # NCCL's stream while the 128 bp output embedding (and, for the pair blocks, the whole decoder) runs here.
            ht = self.sp.halo_bp // 128
            halo_pending = self.sp.exchange_halos_async([single, a_dec], ht)
            if with_embeds:
                pairT_pending = self.sp.transpose_blocks_async(pair)
        if with_embeds:
            f32 = dict(device=dev, dtype=torch.float32)
            o128, o1, op = P["out128"], P["out1"], P["outpair"]
            t128 = o128["table"].index_select(0, organism_index)
            e128 = torch.empty(B, n, o128["W"].shape[1], **f32)
            e128_b16 = torch.empty(B, n, o128["W"].shape[1], **bf)
            self._gemm(single_b16, o128["W"], pre_shift=o128["b"], actA=Act(e128, o128["s"], t128, ACT_GELU1702),
                       actB=Act(e128_b16, o128["s"], t128, ACT_GELU1702))
            skp = torch.empty(B, n, o1["Wskip"].shape[1], **f32)
            ops.gemm(e128_b16, o1["Wskip"], out=skp)
        if sharded:
            (l32, r32), (l16, r16) = halo_pending.wait()
            x_dec = torch.cat([t for t in (l32, single, r32) if t is not None], dim=1)
            a_dec = torch.cat([t for t in (l16, a_dec, r16) if t is not None], dim=1)
        self._own = (plan.halo_left, plan.halo_left + plan.L_loc, seq.shape[1]) if sharded else None
        unet_out, unet_b16 = self._decoder(P, x_dec, a_dec, skips, want_plain=with_embeds)
        self._own = None
        if sharded:
            unet_out = unet_out[:, plan.halo_left:plan.halo_left + plan.L_loc]
            if unet_b16 is not None:
                unet_b16 = unet_b16[:, plan.halo_left:plan.halo_left + plan.L_loc]
        out = dict(unet_output=unet_out, single_repr=single, pairwise_repr=pair)
        if with_embeds:
            Lfull = unet_out.shape[1]
            shift = int(math.log2(Lfull // n))
            assert (n << shift) == Lfull
            t1 = o1["table"].index_select(0, organism_index)
            e1 = torch.empty(B, Lfull, o1["W"].shape[1], **f32)
            # prediction heads take embeds_1bp as a bf16 MMA operand: emitted by the same epilogue when asked for
            e1_b16 = torch.empty(B, Lfull, o1["W"].shape[1], **bf) if want_head_operands else None
            self._gemm(unet_b16, o1["W"], pre_shift=o1["b"], res=skp, res_shift=shift, actA=Act(e1, o1["s"], t1, ACT_GELU1702),
                       actB=Act(e1_b16, o1["s"], t1, ACT_GELU1702) if want_head_operands else None)
            epair = torch.empty_like(pair)
            pairT = pairT_pending.wait() if sharded else None
            ops.pair_out_embed(pair, op["w"], op["b"], op["emb"].index_select(0, organism_index).contiguous(), epair, pairT=pairT)
            out.update(embeds_1bp=e1, embeds_128bp=e128, embeds_pair=epair, embeds_128bp_b16=e128_b16)
            if want_head_operands:
                out.update(embeds_1bp_b16=e1_b16)
        return out

    def run_graphed(self, tag, fn, tensors: Dict[str, torch.Tensor]):
        """Replay `fn(**tensors)` from a CUDA graph.  One graph per (tag, input shapes / dtypes, parameter version);
        the most recently used graphs are kept (alternating shapes — ISM windows, two-pass splice flows — do not
        re-capture) until their private pools exceed `graph_pool_budget` of the device memory.  The returned tensors are
        the graph's static outputs: the next call with the same key overwrites them."""
        dev = next(iter(tensors.values())).device
        key = (tag, tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(tensors.items())), str(dev), self.sp is not None,
               self._version_key(dev))
        ent = self._graphs.pop(key, None)
        if ent is None:
            if self.collect is not None:
                raise ops._lib.AgbError("stage collection is not available under CUDA graphs")
            static = {k: v.clone() for k, v in tensors.items()}
            stream = torch.cuda.Stream(device=dev)
            stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(stream):
                fn(**static)  # warm-up on the side stream: builds the pack / rope / rel caches
            torch.cuda.current_stream().wait_stream(stream)
            torch.cuda.synchronize()
            torch.cuda.empty_cache()   # the warm-up's blocks would otherwise sit cached next to the graph's private pool
            before = torch.cuda.memory_reserved(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = fn(**static)
            ent = dict(graph=g, static=static, out=out, bytes=max(torch.cuda.memory_reserved(dev) - before, 0))
            budget = self.graph_pool_budget * torch.cuda.get_device_properties(dev).total_memory
            while self._graphs and sum(e["bytes"] for e in self._graphs.values()) + ent["bytes"] > budget:
                self._graphs.pop(next(iter(self._graphs)))   # least recently used first
        self._graphs[key] = ent      # (re-)inserted last = most recently used
        for k, v in tensors.items():
            ent["static"][k].copy_(v, non_blocking=True)
        ent["graph"].replay()
        return ent["out"]

    def run_embeds_graphed(self, seq, organism_index):
        """run_embeds replayed from a CUDA graph: removes the ~190 per-launch host round trips, which matter once a
        rank's share of the forward is only a few ms."""
        return self.run_graphed("embeds", lambda seq, org: self.run_embeds(seq, org),
                                dict(seq=seq.to(torch.int64), org=organism_index.to(seq.device).long()))

    def invalidate(self):
        """Drop the packed weights, tables and captured graphs (call after changing weights through `.data`, which
        does not bump the parameter version the caches are keyed on)."""
        self._pack = self._pack_key = self._tensors = None
        self._rope.clear()
        self._rel.clear()
        self._graphs.clear()

    def run_trunk(self, seq, pre_attend_embed=None):
        return self.run(seq, pre_attend_embed, with_embeds=False)

    def run_embeds(self, seq, organism_index, want_head_operands=False):
        P = self.pack(seq.device)
        if "org" not in P:
            raise ops._lib.AgbError("run_embeds needs the AlphaGenome shell (organism / output embeddings)")
        organism_index = organism_index.to(seq.device).long()
        org = P["org"].index_select(0, organism_index)
        return self.run(seq, org, organism_index=organism_index, with_embeds=True, want_head_operands=want_head_operands)
