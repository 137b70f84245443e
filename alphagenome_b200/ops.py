"""Thin Python wrappers over the C ABI: torch tensors in, raw pointers + sizes out.

No arithmetic happens here; each function fills the C struct the library expects and launches on the
current CUDA stream.  Tensors are `[batch, rows, channels]` views whose last dimension is contiguous.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_GELU1702, ACT_GELU_ERF, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SOFTPLUS_MUL, AgAttnArgs, AgEpiOut, AgGemmArgs  # noqa: F401


# bench.py sets this to a list to collect (start_event, end_event, algorithmic_flops) per ag_gemm_fwd launch
PROFILE_GEMM = None


@dataclass
class Act:
    """One fused activated output: dst = act(v * scale + shift)."""

    dst: torch.Tensor  # [batch, rows, N] bf16 or fp32
    scale: Optional[torch.Tensor] = None  # [N] fp32
    shift: Optional[torch.Tensor] = None  # [N] or [batch, N] fp32
    act: int = ACT_NONE


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(t: Optional[torch.Tensor] = None) -> int:
    """The current stream OF THE TENSOR'S DEVICE (a model on cuda:1 launches on cuda:1's current stream even when the
    process's current device is cuda:0)."""
    return torch.cuda.current_stream(t.device if t is not None else None).cuda_stream


def _require(cond: bool, msg: str) -> None:
    if not cond:
        raise _lib.AgbError(msg)


def _as3(t: torch.Tensor) -> torch.Tensor:
    if t.dim() == 2:
        t = t.unsqueeze(0)
    _require(t.dim() == 3, f"expected a [batch, rows, channels] tensor, got {tuple(t.shape)}")
    _require(t.stride(2) == 1, "last dimension must be contiguous")
    return t


def _fill_act(e: AgEpiOut, a: Optional[Act], batch: int, rows: int, N: int) -> None:
    if a is None:
        e.ptr = None
        return
    d = _as3(a.dst)
    _require(d.is_cuda and d.dtype in (torch.bfloat16, torch.float32), "Act.dst must be a CUDA bf16/fp32 tensor")
    _require(d.shape[0] == batch and d.shape[1] >= rows and d.shape[2] == N, f"Act.dst shape {tuple(d.shape)} vs ({batch},{rows},{N})")
    e.ptr = d.data_ptr()
    e.bstride = d.stride(0)
    e.ld = d.stride(1)
    e.dtype = 0 if d.dtype == torch.bfloat16 else 1
    e.act = int(a.act)
    if a.scale is not None:
        _require(a.scale.dtype == torch.float32 and a.scale.numel() == N and a.scale.is_contiguous(), "Act.scale must be fp32 [N]")
    e.scale = _ptr(a.scale)
    e.shift_bstride = 0
    if a.shift is not None:
        s = a.shift
        _require(s.dtype == torch.float32 and s.is_contiguous() and s.shape[-1] == N, "Act.shift must be fp32 [N] or [batch, N]")
        if s.dim() == 2:
            _require(s.shape[0] == batch, "Act.shift batch mismatch")
            e.shift_bstride = N
    e.shift = _ptr(a.shift)


def gemm(
    A: torch.Tensor,
    W: torch.Tensor,
    *,
    M: Optional[int] = None,
    a_row0: int = 0,
    w_batched: bool = False,
    n_pad: Optional[int] = None,
    pre_scale: Optional[torch.Tensor] = None,
    pre_shift: Optional[torch.Tensor] = None,
    relu: bool = False,
    res: Optional[torch.Tensor] = None,
    res_ch: Optional[int] = None,
    res_shift: int = 0,
    post_scale: Optional[torch.Tensor] = None,
    out: Optional[torch.Tensor] = None,
    actA: Optional[Act] = None,
    actB: Optional[Act] = None,
    pool: Optional[torch.Tensor] = None,
    actP: Optional[Act] = None,
    cta_group: int = 0,
    A2: Optional[torch.Tensor] = None,
    W2: Optional[torch.Tensor] = None,
) -> None:
    """ag_gemm_fwd: conv1d (taps = W.shape[0]) / linear / batched matmul with the fused epilogue of agb200.h.

    A: bf16 [batch, a_rows, K];  W: bf16 [taps, N, K] (shared) or [batch, N, K] (w_batched).
    A2 / W2: optional second operand pair, bf16 [batch, M, K2] / [N, K2], accumulated into the same tile.
    """
    lib = _lib.load()
    A = _as3(A)
    W = _as3(W)
    _require(A.is_cuda and A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16, "A and W must be CUDA bf16 tensors")
    batch, a_rows, K = A.shape
    N = W.shape[1]
    _require(W.shape[2] == K, f"K mismatch: A {tuple(A.shape)} W {tuple(W.shape)}")
    taps = 1 if w_batched else W.shape[0]
    a_shared = False
    if w_batched:
        if batch == 1 and W.shape[0] > 1:   # one A (e.g. a weight matrix as the row operand) against every W[b]
            a_shared, batch = True, W.shape[0]
        _require(W.shape[0] == batch, "batched W needs W.shape[0] == batch")
    if M is None:
        M = a_rows - 2 * a_row0
    w_rows = 0
    if n_pad is not None and n_pad != N:
        # W has fewer rows than the (16-aligned) N the kernel computes; the missing rows read as zero
        _require(n_pad > N and n_pad % 16 == 0, "n_pad must be a multiple of 16 and >= W rows")
        w_rows, N = N, n_pad

    g = AgGemmArgs()
    g.A = A.data_ptr()
    g.a_bstride = 0 if a_shared else (A.stride(0) if batch > 1 else A.stride(1) * a_rows)  # stride of a size-1 dim is arbitrary
    g.a_ld = A.stride(1)
    g.a_rows = a_rows
    g.a_row0 = a_row0
    g.w_batched = 1 if w_batched else 0
    g.W = W.data_ptr()
    g.w_zstride = W.stride(0) if W.shape[0] > 1 else W.stride(1) * W.shape[1]
    g.w_ld = W.stride(1)
    g.batch, g.M, g.N, g.K, g.taps = batch, M, N, K, taps
    g.w_rows = w_rows
    g.cta_group = int(cta_group)
    for name, v in (("pre_scale", pre_scale), ("pre_shift", pre_shift)):
        if v is not None:
            _require(v.dtype == torch.float32 and v.numel() == N and v.is_contiguous(), f"{name} must be fp32 [N]")
    g.pre_scale = _ptr(pre_scale)
    g.pre_shift = _ptr(pre_shift)
    g.relu = 1 if relu else 0
    if res is not None:
        r = _as3(res)
        _require(r.dtype == torch.float32 and r.shape[0] == batch, "res must be fp32 [batch, rows, C]")
        g.res = r.data_ptr()
        g.res_bstride = r.stride(0)
        g.res_ld = r.stride(1)
        g.res_ch = min(N, r.shape[2]) if res_ch is None else res_ch
        g.res_shift = res_shift
        _require(((M - 1) >> res_shift) < r.shape[1], "res has too few rows")
    else:
        g.res = None
    if post_scale is not None:
        _require(post_scale.dtype == torch.float32 and post_scale.numel() == 1, "post_scale must be an fp32 scalar tensor")
    g.post_scale = _ptr(post_scale)
    if out is not None:
        o = _as3(out)
        _require(o.dtype == torch.float32 and o.shape[0] == batch and o.shape[1] >= M and o.shape[2] == N, f"out shape {tuple(o.shape)}")
        g.out = o.data_ptr()
        g.out_bstride = o.stride(0)
        g.out_ld = o.stride(1)
    else:
        g.out = None
    if pool is not None:
        pz = _as3(pool)
        _require(pz.dtype == torch.float32 and pz.shape[0] == batch and pz.shape[1] >= M // 2 and pz.shape[2] == N, f"pool shape {tuple(pz.shape)}")
        g.pool = pz.data_ptr()
        g.pool_bstride = pz.stride(0)
        g.pool_ld = pz.stride(1)
    else:
        g.pool = None
    _fill_act(g.actA, actA, batch, M, N)
    _fill_act(g.actB, actB, batch, M, N)
    _fill_act(g.actP, actP, batch, M // 2, N)
    if A2 is not None or W2 is not None:
        _require(A2 is not None and W2 is not None, "A2 and W2 must be given together")
        A2 = _as3(A2)
        W2 = _as3(W2)
        _require(A2.is_cuda and A2.dtype == torch.bfloat16 and W2.dtype == torch.bfloat16, "A2 and W2 must be CUDA bf16 tensors")
        _require(A2.shape[0] == batch and A2.shape[1] >= M and W2.shape[0] == 1 and W2.shape[1] == N and W2.shape[2] == A2.shape[2],
                 f"second operand pair: A2 {tuple(A2.shape)} W2 {tuple(W2.shape)} vs batch {batch}, M {M}, N {N}")
        g.A2, g.W2 = A2.data_ptr(), W2.data_ptr()
        g.a2_bstride = A2.stride(0) if batch > 1 else A2.stride(1) * A2.shape[1]
        g.a2_ld, g.a2_rows, g.a2_row0 = A2.stride(1), A2.shape[1], 0
        g.w2_ld, g.K2 = W2.stride(1), A2.shape[2]
    if PROFILE_GEMM is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.ag_gemm_fwd(C.byref(g), A.device.index or 0, _stream(A)))
        e1.record()
        # algorithmic bytes of this launch: every operand / output once at its storage dtype
        def _out_bytes(a, rows):
            return 0 if a is None else batch * rows * N * a.dst.element_size()

        nbytes = (batch if not (w_batched and A.shape[0] == 1) else 1) * a_rows * K * 2 + W.numel() * 2
        if res is not None:
            nbytes += batch * (((M - 1) >> res_shift) + 1) * g.res_ch * 4
        if out is not None:
            nbytes += batch * M * N * 4
        if pool is not None:
            nbytes += batch * (M // 2) * N * 4
        nbytes += _out_bytes(actA, M) + _out_bytes(actB, M) + _out_bytes(actP, M // 2)
        if A2 is not None:   # its 2*M*N*K2 products are a recomputation, not algorithmic work: bytes counted, FLOPs not
            nbytes += batch * M * A2.shape[2] * 2 + W2.numel() * 2
        PROFILE_GEMM.append((e0, e1, 2.0 * batch * M * (w_rows or N) * K * taps, float(nbytes)))
        return
    _lib.check(lib.ag_gemm_fwd(C.byref(g), A.device.index or 0, _stream(A)))


def set_gemm_cta_group(cta_group: int) -> int:
    """ag_gemm_set_cta_group: 0 = automatic, 1 = single-CTA MMAs, 2 = CTA pairs.  Returns the previous setting."""
    prev = _lib.load().ag_gemm_set_cta_group(int(cta_group))
    _require(prev >= 0, f"bad cta_group {cta_group}")
    return prev


def _dev(t: torch.Tensor) -> int:
    return t.device.index or 0


def _f32(t: Optional[torch.Tensor], name: str) -> None:
    if t is not None:
        _require(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), f"{name} must be a contiguous CUDA fp32 tensor")


def attention(
    Q: torch.Tensor,
    K: torch.Tensor,
    Vt: torch.Tensor,
    out: torch.Tensor,
    *,
    mode: int,
    scale: float,
    bias: Optional[torch.Tensor] = None,
    softcap: float = 5.0,
    out_bias: Optional[torch.Tensor] = None,
    res: Optional[torch.Tensor] = None,
    norm_w: Optional[torch.Tensor] = None,
    norm_b: Optional[torch.Tensor] = None,
    norm_out: Optional[torch.Tensor] = None,
) -> None:
    """ag_attention_fwd.  Q [B, H, n_q, d], K [B, Hk, n_k, d], Vt [B, Hk, dv, n_k(+pad)] (bf16 views, last dim
    contiguous), out [B, H, n_q, dv] (bf16 or fp32 view).  n_k is taken from K."""
    lib = _lib.load()
    for t, nm in ((Q, "Q"), (K, "K"), (Vt, "Vt")):
        _require(t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 4 and t.stride(3) == 1, f"{nm} must be a 4-d CUDA bf16 view")
    B, H, n_q, d = Q.shape
    Hk, n_k = K.shape[1], K.shape[2]
    dv = Vt.shape[2]
    _require(K.shape[0] == B and K.shape[3] == d and Vt.shape[0] == B and Vt.shape[1] == Hk and Vt.shape[3] >= n_k, "K/Vt shape mismatch")
    _require(out.dim() == 4 and tuple(out.shape) == (B, H, n_q, dv) and out.stride(3) == 1, f"out must be [B,H,n_q,dv], got {tuple(out.shape)}")
    a = AgAttnArgs()
    a.Q, a.K, a.Vt = Q.data_ptr(), K.data_ptr(), Vt.data_ptr()
    def st(t, i):  # the stride of a size-1 dim is arbitrary: give TMA a well-formed one
        return t.stride(i) if t.shape[i] > 1 else max(t.stride(j) * t.shape[j] for j in range(i + 1, 4))

    a.q_bstride, a.q_hstride, a.q_ld = st(Q, 0), st(Q, 1), Q.stride(2)
    a.k_bstride, a.k_hstride, a.k_ld = st(K, 0), st(K, 1), K.stride(2)
    a.v_bstride, a.v_hstride, a.v_ld = st(Vt, 0), st(Vt, 1), Vt.stride(2)
    a.batch, a.heads, a.kv_heads, a.n_q, a.n_k, a.d, a.dv = B, H, Hk, n_q, n_k, d, dv
    a.mode = mode
    a.scale, a.softcap = float(scale), float(softcap)
    if bias is not None:
        _f32(bias, "bias")
        _require(bias.dim() == 4 and bias.shape[0] == B and bias.shape[1] == H, "bias must be [B,H,rows,cols]")
        a.bias, a.bias_rows, a.bias_p = bias.data_ptr(), bias.shape[2], bias.shape[3]
    a.out = out.data_ptr()
    a.out_bstride, a.out_hstride, a.out_ld = st(out, 0), st(out, 1), out.stride(2)
    a.out_dtype = 0 if out.dtype == torch.bfloat16 else 1
    _f32(out_bias, "out_bias")
    a.out_bias = _ptr(out_bias)
    if res is not None:
        _require(res.dtype == torch.float32 and res.shape == out.shape and res.stride() == out.stride(), "res must match out's layout")
    a.res = _ptr(res)
    if norm_out is not None:
        _require(out.dtype == torch.float32 and norm_out.dtype == torch.bfloat16 and norm_out.shape == out.shape and
                 norm_out.stride() == out.stride(), "norm_out must be bf16 with out's layout (fp32 out only)")
        _f32(norm_w, "norm_w")
        _f32(norm_b, "norm_b")
        _require(norm_w is not None and norm_b is not None and norm_w.numel() == dv and norm_b.numel() == dv, "norm_w/norm_b must be fp32 [dv]")
    a.norm_w, a.norm_b, a.norm_out = _ptr(norm_w), _ptr(norm_b), _ptr(norm_out)
    _lib.check(lib.ag_attention_fwd(C.byref(a), _dev(Q), _stream(Q)))


def onehot_im2col(tokens: torch.Tensor, out: torch.Tensor, width: int, n_base: int) -> None:
    _require(tokens.is_cuda and tokens.dtype == torch.int64 and tokens.is_contiguous() and tokens.dim() == 2, "tokens must be CUDA int64 [B, L]")
    B, L = tokens.shape
    _require(out.dtype == torch.bfloat16 and out.is_contiguous() and tuple(out.shape[:2]) == (B, L), "out must be bf16 [B, L, ld]")
    _lib.check(_lib.load().ag_onehot_im2col_fwd(tokens.data_ptr(), out.data_ptr(), B, L, width, n_base, out.shape[2], _dev(out), _stream(out)))


def qkv_post(qkv, Q, K, V, cos_tab, sin_tab, qw, qb, kw, kb, vw, vb, heads: int, d: int, dv: int) -> None:
    B, n, Ctot = qkv.shape
    _require(Ctot == heads * d + d + dv, "qkv channel count mismatch")
    for t, nm in ((qkv, "qkv"), (cos_tab, "cos"), (sin_tab, "sin"), (qw, "qw"), (qb, "qb"), (kw, "kw"), (kb, "kb"), (vw, "vw"), (vb, "vb")):
        _f32(t, nm)
    _require(cos_tab.shape[0] >= n and cos_tab.shape[1] == d // 2, "rope tables must be [>=n, d/2]")
    for t in (Q, K, V):
        _require(t.dtype == torch.bfloat16 and t.is_contiguous(), "Q/K/V outputs must be contiguous bf16")
    _lib.check(_lib.load().ag_qkv_post_fwd(qkv.data_ptr(), Q.data_ptr(), K.data_ptr(), V.data_ptr(), cos_tab.data_ptr(), sin_tab.data_ptr(),
                                           qw.data_ptr(), qb.data_ptr(), kw.data_ptr(), kb.data_ptr(), vw.data_ptr(), vb.data_ptr(),
                                           B, n, heads, d, dv, _dev(qkv), _stream(qkv)))


def transpose_bf16(src: torch.Tensor, dst: torch.Tensor) -> None:
    """dst[b, c, r] = src[b, r, c]; src [nb, R, C] bf16 view, dst [nb, C, >=R] bf16 view (last dims contiguous)."""
    _require(src.dtype == torch.bfloat16 and dst.dtype == torch.bfloat16 and src.dim() == 3 and dst.dim() == 3, "bf16 3-d views expected")
    nb, R, Cc = src.shape
    _require(dst.shape[0] == nb and dst.shape[1] == Cc and dst.shape[2] >= R and src.stride(2) == 1 and dst.stride(2) == 1, "transpose shape mismatch")
    _lib.check(_lib.load().ag_transpose_bf16(src.data_ptr(), dst.data_ptr(), nb, R, Cc, src.stride(0), src.stride(1), dst.stride(0), dst.stride(1), _dev(src), _stream(src)))


def pair_bias(pair, scale, shift, W, bias_out) -> None:
    """ag_pair_bias_fwd.  One layer: scale/shift [C], W [heads, C], bias_out [B, heads, rows, cols].  Two layers sharing
    the pair tensor: scale/shift [2, C], W [2, heads, C], bias_out [2, B, heads, rows, cols]."""
    B, rows, cols, Cc = pair.shape
    for t, nm in ((pair, "pair"), (scale, "scale"), (shift, "shift"), (W, "W"), (bias_out, "bias_out")):
        _f32(t, nm)
    sets = W.shape[0] if W.dim() == 3 else 1
    heads = W.shape[-2]
    _require(scale.numel() == sets * Cc and shift.numel() == sets * Cc and W.shape[-1] == Cc, "scale/shift/W shape mismatch")
    want = (sets, B, heads, rows, cols) if W.dim() == 3 else (B, heads, rows, cols)
    _require(tuple(bias_out.shape) == want, f"bias_out must be {want}")
    _lib.check(_lib.load().ag_pair_bias_fwd(pair.data_ptr(), scale.data_ptr(), shift.data_ptr(), W.data_ptr(), bias_out.data_ptr(), B, rows,
                                            cols, Cc, heads, sets, B * heads * rows * cols, _dev(pair), _stream(pair)))


def pool_rmsnorm(single, w, b, x_out, xg_out, pool: int) -> None:
    B, n, Cc = single.shape
    for t, nm in ((single, "single"), (w, "w"), (b, "b")):
        _f32(t, nm)
    for t in (x_out, xg_out):
        _require(t.dtype == torch.bfloat16 and t.is_contiguous() and tuple(t.shape) == (B, n // pool, Cc), "pooled outputs must be bf16 [B, n/pool, C]")
    _lib.check(_lib.load().ag_pool_rmsnorm_fwd(single.data_ptr(), w.data_ptr(), b.data_ptr(), x_out.data_ptr(), xg_out.data_ptr(), B, n, Cc, pool, _dev(single), _stream(single)))


def pair_combine(qk, relq, relk, Wp, bp, outer, prev, pair_out, P: int, i_off: int = 0, norm_w=None, norm_b=None, norm_out=None) -> None:
    """qk/relq/prev/pair_out hold this rank's rows [i_off, i_off+rows); relk and outer hold all P rows.
    norm_out (bf16, pair_out's shape) optionally receives RMSNormWithBias(pair_out) with (norm_w, norm_b)."""
    B, H, rows = qk.shape[0], qk.shape[1], qk.shape[2]
    D = Wp.shape[0]
    for t, nm in ((qk, "qk"), (relq, "relq"), (relk, "relk"), (Wp, "Wp"), (bp, "bp"), (outer, "outer"), (prev, "prev"), (pair_out, "pair_out"),
                  (norm_w, "norm_w"), (norm_b, "norm_b")):
        _f32(t, nm)
    _require(tuple(relq.shape[:3]) == (B, H, rows) and tuple(relk.shape[:3]) == (B, H, P) and relk.shape[3] == relq.shape[3], "relq [B,H,rows,ld], relk [B,H,P,ld]")
    _require(tuple(outer.shape) == (B, P, 2 * D) and tuple(pair_out.shape) == (B, rows, P, D), "outer/pair_out shape mismatch")
    _require(prev is None or prev.shape == pair_out.shape, "prev must match pair_out")
    if norm_out is not None:
        _require(norm_out.dtype == torch.bfloat16 and norm_out.is_contiguous() and norm_out.numel() == pair_out.numel(), "norm_out must be contiguous bf16 of pair_out's size")
        _require(norm_w is not None and norm_b is not None and norm_w.numel() == D and norm_b.numel() == D, "norm_w/norm_b must be fp32 [D]")
    _lib.check(_lib.load().ag_pair_combine_fwd(qk.data_ptr(), relq.data_ptr(), relk.data_ptr(), Wp.data_ptr(), bp.data_ptr(), outer.data_ptr(),
                                               _ptr(prev), pair_out.data_ptr(), _ptr(norm_w), _ptr(norm_b), _ptr(norm_out), B, P, H, D, qk.shape[3],
                                               relq.shape[3], rows, i_off, _dev(qk), _stream(qk)))


def rmsnorm(x, w, b, out) -> None:
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    for t, nm in ((x, "x"), (w, "w"), (b, "b")):
        _f32(t, nm)
    _require(out.dtype == torch.bfloat16 and out.is_contiguous() and out.numel() == x.numel(), "out must be contiguous bf16 of x's size")
    _lib.check(_lib.load().ag_rmsnorm_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), rows, Cc, _dev(x), _stream(x)))


def add_embed_norm(x, emb, scale, shift, single_out, a_out) -> None:
    B, rows, Cc = x.shape
    for t, nm in ((x, "x"), (emb, "emb"), (scale, "scale"), (shift, "shift"), (single_out, "single_out")):
        _f32(t, nm)
    _require(tuple(emb.shape) == (B, Cc) and a_out.dtype == torch.bfloat16 and a_out.is_contiguous(), "emb must be [B, C]; a_out bf16")
    _lib.check(_lib.load().ag_add_embed_norm_fwd(x.data_ptr(), emb.data_ptr(), scale.data_ptr(), shift.data_ptr(), single_out.data_ptr(), a_out.data_ptr(), B, rows, Cc, _dev(x), _stream(x)))


def pair_out_embed(pair, w, b, emb, out, pairT=None) -> None:
    """pair/out: this rank's rows [B, rows, P, C]; pairT: transposed blocks [P/rows, B, rows, rows, C] (None on one rank)."""
    B, rows, P, Cc = pair.shape
    if pairT is None:
        _require(rows == P, "pairT is required for a row shard")
        pairT = pair
    for t, nm in ((pair, "pair"), (pairT, "pairT"), (w, "w"), (b, "b"), (emb, "emb"), (out, "out")):
        _f32(t, nm)
    _require(pairT.numel() == pair.numel(), "pairT size mismatch")
    _lib.check(_lib.load().ag_pair_out_embed_fwd(pair.data_ptr(), pairT.data_ptr(), w.data_ptr(), b.data_ptr(), emb.data_ptr(), out.data_ptr(), B, rows, P, Cc, _dev(pair), _stream(pair)))


def rmsnorm_act(x, w, b, out, add=None, rows_per_batch: Optional[int] = None, act: int = ACT_NONE) -> None:
    """ag_rmsnorm_act_fwd: out = act(RMSNormWithBias(x) + add[batch]) in bf16; x [..., C] fp32, add [batches, C] or None."""
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    for t, nm in ((x, "x"), (w, "w"), (b, "b"), (add, "add")):
        _f32(t, nm)
    _require(out.dtype == torch.bfloat16 and out.is_contiguous() and out.numel() == x.numel(), "out must be contiguous bf16 of x's size")
    if add is not None:
        _require(rows_per_batch is not None and add.shape[-1] == Cc and add.numel() // Cc * rows_per_batch == rows, "add must be [rows / rows_per_batch, C]")
    _lib.check(_lib.load().ag_rmsnorm_act_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), _ptr(add), out.data_ptr(), rows,
                                              rows_per_batch or rows, Cc, int(act), _dev(x), _stream(x)))


def splice_rope(x, scale, offset, cos_tab, sin_tab, out) -> None:
    """ag_splice_rope_fwd: x fp32 [B, P, H]; scale/offset [T, H]; out bf16 [B, T, >=P, H]; cos/sin [T, H/2] (rotary
    position = context index) or [B, P, H/2] (rotary position = the site)."""
    B, P, Hd = x.shape
    T = scale.shape[0]
    for t, nm in ((x, "x"), (scale, "scale"), (offset, "offset"), (cos_tab, "cos"), (sin_tab, "sin")):
        _f32(t, nm)
    per_site = cos_tab.dim() == 3
    want = (B, P, Hd // 2) if per_site else (T, Hd // 2)
    _require(tuple(scale.shape) == (T, Hd) and tuple(offset.shape) == (T, Hd) and tuple(cos_tab.shape) == want and tuple(sin_tab.shape) == want,
             "scale/offset [T, H]; cos/sin [T, H/2] or [B, P, H/2]")
    _require(out.dtype == torch.bfloat16 and out.is_contiguous() and out.dim() == 4 and out.shape[0] == B and out.shape[1] == T and out.shape[2] >= P and out.shape[3] == Hd, "out must be bf16 [B, T, >=P, H]")
    _lib.check(_lib.load().ag_splice_rope_fwd(x.data_ptr(), scale.data_ptr(), offset.data_ptr(), cos_tab.data_ptr(), sin_tab.data_ptr(), out.data_ptr(),
                                              B, P, T, Hd, out.shape[2], 1 if per_site else 0, _dev(x), _stream(x)))


# ------------------------------------------------------------------------------------------------ updating BatchRMSNorm
def bn_stats(x: torch.Tensor, sums: torch.Tensor) -> None:
    """ag_bn_stats_fwd: x fp32 [B, rows, C] view (last dim contiguous) -> sums fp64 [2, C] (sum, sum of squares)."""
    x = _as3(x)
    _require(x.is_cuda and x.dtype == torch.float32, "x must be a CUDA fp32 tensor")
    B, rows, Cc = x.shape
    _require(sums.dtype == torch.float64 and sums.is_contiguous() and tuple(sums.shape) == (2, Cc), "sums must be fp64 [2, C]")
    bs = x.stride(0) if B > 1 else x.stride(1) * rows
    _lib.check(_lib.load().ag_bn_stats_fwd(x.data_ptr(), sums.data_ptr(), B, rows, bs, x.stride(1), Cc, _dev(x), _stream(x)))


def bn_update(sums, count: float, running_var, gamma, momentum: float, eps: float, scale_out, fold_bias=None, fold_beta=None,
              fold_shift_out=None) -> None:
    """ag_bn_update_fwd: running_var.lerp_(batch variance, momentum) in place; scale_out = gamma / sqrt(running_var + eps)."""
    Cc = running_var.numel()
    for t, nm in ((running_var, "running_var"), (gamma, "gamma"), (scale_out, "scale_out"), (fold_bias, "fold_bias"), (fold_beta, "fold_beta"),
                  (fold_shift_out, "fold_shift_out")):
        _f32(t, nm)
        _require(t is None or t.numel() == Cc, f"{nm} must have C elements")
    _require(sums.dtype == torch.float64 and sums.is_contiguous() and sums.numel() == 2 * Cc, "sums must be fp64 [2, C]")
    _lib.check(_lib.load().ag_bn_update_fwd(sums.data_ptr(), float(count), running_var.data_ptr(), gamma.data_ptr(), float(momentum), float(eps),
                                            scale_out.data_ptr(), _ptr(fold_bias), _ptr(fold_beta), _ptr(fold_shift_out), Cc, _dev(running_var), _stream(running_var)))


def affine_act(x, scale=None, shift=None, *, res=None, out_f32=None, out_bf16=None, act: int = ACT_NONE) -> None:
    """ag_affine_act_fwd: y = x * scale[c] + shift[(b,) c] (+ res); out_f32 / out_bf16 = act(y).  [B, rows, C] views."""
    x = _as3(x)
    B, rows, Cc = x.shape
    _require(x.is_cuda and x.dtype == torch.float32, "x must be a CUDA fp32 tensor")

    def bst(t):
        return t.stride(0) if B > 1 else t.stride(1) * rows

    a = _lib.AgAffineArgs()
    a.x, a.x_bstride, a.x_ld = x.data_ptr(), bst(x), x.stride(1)
    _f32(scale, "scale")
    a.scale = _ptr(scale)
    a.shift_bstride = 0
    if shift is not None:
        _f32(shift, "shift")
        _require(shift.shape[-1] == Cc, "shift must be [C] or [B, C]")
        if shift.dim() == 2:
            _require(shift.shape[0] == B, "shift batch mismatch")
            a.shift_bstride = Cc
    a.shift = _ptr(shift)
    if res is not None:
        r = _as3(res)
        _require(r.dtype == torch.float32 and tuple(r.shape) == (B, rows, Cc), "res must be fp32 with x's shape")
        a.res, a.res_bstride, a.res_ld = r.data_ptr(), bst(r), r.stride(1)
    if out_f32 is not None:
        o = _as3(out_f32)
        _require(o.dtype == torch.float32 and tuple(o.shape) == (B, rows, Cc), "out_f32 must be fp32 with x's shape")
        a.out_f32, a.out_f32_bstride, a.out_f32_ld = o.data_ptr(), bst(o), o.stride(1)
    if out_bf16 is not None:
        o = _as3(out_bf16)
        _require(o.dtype == torch.bfloat16 and tuple(o.shape) == (B, rows, Cc), "out_bf16 must be bf16 with x's shape")
        a.out_bf16, a.out_bf16_bstride, a.out_bf16_ld = o.data_ptr(), bst(o), o.stride(1)
    a.B, a.rows, a.C, a.act = B, rows, Cc, int(act)
    _lib.check(_lib.load().ag_affine_act_fwd(C.byref(a), _dev(x), _stream(x)))


# ------------------------------------------------------------------------------------------------ sequence parallel, peer memory
def _peer_array(ptrs):
    return (C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def qkv_post_gather(qkv, Q, kv_peer_ptrs, n_all: int, row0: int, cos_tab, sin_tab, qw, qb, kw, kb, vw, vb, heads: int, d: int, dv: int) -> None:
    """ag_qkv_post_gather_fwd: qkv_post whose K / V rows go straight into every rank's gathered [B, n_all, d + dv] buffer
    (kv_peer_ptrs: device pointers of the symmetric allocation, one per rank)."""
    B, n, Ctot = qkv.shape
    _require(Ctot == heads * d + d + dv, "qkv channel count mismatch")
    for t, nm in ((qkv, "qkv"), (cos_tab, "cos"), (sin_tab, "sin"), (qw, "qw"), (qb, "qb"), (kw, "kw"), (kb, "kb"), (vw, "vw"), (vb, "vb")):
        _f32(t, nm)
    _require(cos_tab.shape[0] >= n and cos_tab.shape[1] == d // 2, "rope tables must be [>=n, d/2]")
    _require(Q.dtype == torch.bfloat16 and Q.is_contiguous(), "Q must be contiguous bf16")
    _lib.check(_lib.load().ag_qkv_post_gather_fwd(qkv.data_ptr(), Q.data_ptr(), _peer_array(kv_peer_ptrs), len(kv_peer_ptrs), n_all, row0,
                                                  cos_tab.data_ptr(), sin_tab.data_ptr(), qw.data_ptr(), qb.data_ptr(), kw.data_ptr(), kb.data_ptr(),
                                                  vw.data_ptr(), vb.data_ptr(), B, n, heads, d, dv, _dev(qkv), _stream(qkv)))


def peer_push(src: torch.Tensor, peer_ptrs, dst_offset_bytes: int) -> None:
    """ag_peer_push_fwd: src (contiguous) -> byte offset dst_offset_bytes of every rank's buffer."""
    _require(src.is_cuda and src.is_contiguous(), "src must be a contiguous CUDA tensor")
    _lib.check(_lib.load().ag_peer_push_fwd(src.data_ptr(), _peer_array(peer_ptrs), len(peer_ptrs), src.numel() * src.element_size(),
                                            int(dst_offset_bytes), _dev(src), _stream(src)))
