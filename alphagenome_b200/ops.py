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
from ._lib import ACT_GELU1702, ACT_GELU_ERF, ACT_NONE, ACT_RELU, AgEpiOut, AgGemmArgs  # noqa: F401


@dataclass
class Act:
    """One fused activated output: dst = act(v * scale + shift)."""

    dst: torch.Tensor  # [batch, rows, N] bf16 or fp32
    scale: Optional[torch.Tensor] = None  # [N] fp32
    shift: Optional[torch.Tensor] = None  # [N] or [batch, N] fp32
    act: int = ACT_NONE


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


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
) -> None:
    """ag_gemm_fwd: conv1d (taps = W.shape[0]) / linear / batched matmul with the fused epilogue of agb200.h.

    A: bf16 [batch, a_rows, K];  W: bf16 [taps, N, K] (shared) or [batch, N, K] (w_batched).
    """
    lib = _lib.load()
    A = _as3(A)
    W = _as3(W)
    _require(A.is_cuda and A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16, "A and W must be CUDA bf16 tensors")
    batch, a_rows, K = A.shape
    N = W.shape[1]
    _require(W.shape[2] == K, f"K mismatch: A {tuple(A.shape)} W {tuple(W.shape)}")
    taps = 1 if w_batched else W.shape[0]
    if w_batched:
        _require(W.shape[0] == batch, "batched W needs W.shape[0] == batch")
    if M is None:
        M = a_rows - 2 * a_row0

    g = AgGemmArgs()
    g.A = A.data_ptr()
    g.a_bstride = A.stride(0)
    g.a_ld = A.stride(1)
    g.a_rows = a_rows
    g.a_row0 = a_row0
    g.w_batched = 1 if w_batched else 0
    g.W = W.data_ptr()
    g.w_zstride = W.stride(0)
    g.w_ld = W.stride(1)
    g.batch, g.M, g.N, g.K, g.taps = batch, M, N, K, taps
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
    _lib.check(lib.ag_gemm_fwd(C.byref(g), A.device.index or 0, _stream()))
