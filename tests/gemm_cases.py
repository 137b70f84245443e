"""Shared GEMM/conv parity cases: the CUDA kernel (through the C ABI) against a plain fp32 torch restatement
of agb200.h's ag_gemm_fwd formula on the same bf16-rounded operands."""
from __future__ import annotations

import math

import torch

from alphagenome_b200 import ops


def act_ref(x, mode):
    if mode == ops.ACT_GELU1702:
        return torch.sigmoid(1.702 * x) * x
    if mode == ops.ACT_GELU_ERF:
        return torch.nn.functional.gelu(x)
    if mode == ops.ACT_RELU:
        return torch.relu(x)
    return x


def ref_acc(A, W, M, a_row0, w_batched):
    """fp64 accumulate of sum_t sum_k A[b, r + t - taps/2 (+row0), k] W[z, n, k]; zero outside the buffer."""
    Af, Wf = A.double(), W.double()
    b, rows, K = A.shape
    N = W.shape[1]
    taps = 1 if w_batched else W.shape[0]
    acc = torch.zeros(b, M, N, dtype=torch.float64, device=A.device)
    r = torch.arange(M, device=A.device)
    for t in range(taps):
        idx = r + a_row0 + t - taps // 2
        ok = (idx >= 0) & (idx < rows)
        X = torch.zeros(b, M, K, dtype=torch.float64, device=A.device)
        X[:, ok] = Af[:, idx[ok]]
        if w_batched:
            acc += torch.einsum("bmk,bnk->bmn", X, Wf)
        else:
            acc += X @ Wf[t].T
    return acc


def run_case(name, *, batch, M, K, N, taps=1, a_row0=0, w_batched=False, epilogue="plain", seed=0, strided=False, cta_group=0, a_shared=False, K2=0):
    """Returns dict(name, ok, max_err, details...). Tolerance: fp32-accumulate of bf16 products, so the only
    error is summation order -> 2e-3 relative to the output scale is generous and still catches any
    descriptor/layout bug (those give O(1) errors)."""
    dev = torch.device("cuda")
    gen = torch.Generator(device="cpu").manual_seed(seed)
    rows = M + 2 * a_row0
    if strided:
        # head-batched layout used by the pair kernels: storage [rows, batch, K], batch stride = K
        A_store = (torch.randn(rows, batch, K, generator=gen) * 0.5).to(dev).bfloat16()
        A = A_store.permute(1, 0, 2)
        W_store = (torch.randn(N, batch, K, generator=gen) * 0.5).to(dev).bfloat16()
        W = W_store.permute(1, 0, 2)
    else:
        A = (torch.randn(1 if a_shared else batch, rows, K, generator=gen) * 0.5).to(dev).bfloat16()
        nz = batch if w_batched else taps
        W = (torch.randn(nz, N, K, generator=gen) / math.sqrt(K * taps)).to(dev).bfloat16()
    # a_shared: ONE A for every batch of a batched W (the pair row attention's V^T = Wv . pn[i]^T)
    acc = ref_acc(A.expand(batch, -1, -1) if a_shared else A, W, M, a_row0, w_batched)

    res = {}
    kw = {}
    if K2:
        # second operand pair accumulated into the same tile (DNAEmbed: the one-hot im2col product rides along with
        # the width-5 conv); the A2 buffer holds exactly the M output rows
        A2 = (torch.randn(batch, M, K2, generator=gen) * 0.5).to(dev).bfloat16()
        W2 = (torch.randn(1, N, K2, generator=gen) / math.sqrt(K2)).to(dev).bfloat16()
        acc = acc + A2.double() @ W2[0].double().T
        kw.update(A2=A2, W2=W2)
    checks = []
    f32 = dict(dtype=torch.float32, device=dev)
    if epilogue == "plain":
        out = torch.full((batch, M, N), float("nan"), **f32)
        kw["out"] = out
        checks.append(("out", out, acc))
    elif epilogue == "full":
        pre_s = (torch.rand(N, generator=gen) + 0.5).to(dev)
        pre_t = torch.randn(N, generator=gen).to(dev)
        res_ch = N - 16 if N > 16 else N
        resid = torch.randn(batch, (M + 1) // 2, res_ch, generator=gen).to(dev)
        post = torch.tensor([0.75], **f32)
        sA = (torch.rand(N, generator=gen) + 0.5).to(dev)
        tA = torch.randn(batch, N, generator=gen).to(dev)
        sB = (torch.rand(N, generator=gen) + 0.5).to(dev)
        tB = torch.randn(N, generator=gen).to(dev)
        sP = (torch.rand(N, generator=gen) + 0.5).to(dev)
        tP = torch.randn(N, generator=gen).to(dev)
        out = torch.full((batch, M, N), float("nan"), **f32)
        dA = torch.zeros(batch, M, N, dtype=torch.bfloat16, device=dev)
        dB = torch.full((batch, M, N), float("nan"), **f32)
        pool = torch.full((batch, M // 2, N), float("nan"), **f32)
        dP = torch.zeros(batch, M // 2, N, dtype=torch.bfloat16, device=dev)
        kw.update(pre_scale=pre_s, pre_shift=pre_t, res=resid, res_ch=res_ch, res_shift=1, post_scale=post, out=out,
                  actA=ops.Act(dA, sA, tA, ops.ACT_GELU1702), actB=ops.Act(dB, sB, tB, ops.ACT_GELU1702),
                  pool=pool, actP=ops.Act(dP, sP, tP, ops.ACT_GELU1702))
        v = acc * pre_s.double() + pre_t.double()
        rr = torch.arange(M, device=dev) >> 1
        v[:, :, :res_ch] += resid.double()[:, rr]
        v = v * 0.75
        checks.append(("out", out, v))
        checks.append(("actA", dA, act_ref(v * sA.double() + tA.double()[:, None], ops.ACT_GELU1702)))
        checks.append(("actB", dB, act_ref(v * sB.double() + tB.double(), ops.ACT_GELU1702)))
        pv = torch.maximum(v[:, 0::2], v[:, 1::2])
        checks.append(("pool", pool, pv))
        checks.append(("actP", dP, act_ref(pv * sP.double() + tP.double(), ops.ACT_GELU1702)))
    elif epilogue == "relu_bf16":
        bias = torch.randn(N, generator=gen).to(dev)
        dA = torch.zeros(batch, M, N, dtype=torch.bfloat16, device=dev)
        kw.update(pre_shift=bias, relu=True, actA=ops.Act(dA))
        checks.append(("actA", dA, torch.relu(acc + bias.double())))
    else:
        raise ValueError(epilogue)

    prev = ops.set_gemm_cta_group(cta_group)   # 0 auto, 1 single-CTA MMAs, 2 CTA pairs (cta_group::2)
    try:
        ops.gemm(A, W, M=M, a_row0=a_row0, w_batched=w_batched, **kw)
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_cta_group(prev)

    ok = True
    worst = 0.0
    for nm, got, want in checks:
        scale = max(want.abs().max().item(), 1e-6)
        tol = 2e-3 if got.dtype == torch.float32 else 1.2e-2  # bf16 outputs: half-ulp 2^-9 relative
        err = (got.double() - want).abs().max().item() / scale
        bad = not math.isfinite(err) or err > tol
        res[nm] = err
        worst = max(worst, err if math.isfinite(err) else float("inf"))
        ok = ok and not bad
    return dict(name=name, ok=ok, max_rel_err=worst, per_output=res)


CASES = [
    dict(name="lin_1kb", batch=1, M=256, K=64, N=128),
    dict(name="lin_k256_n256", batch=1, M=384, K=256, N=256),
    dict(name="lin_ragged_m", batch=2, M=300, K=128, N=192),
    dict(name="lin_tiny_m16", batch=1, M=16, K=1536, N=1344),
    dict(name="lin_n176", batch=1, M=128, K=192, N=1408),
    dict(name="lin_k_tail", batch=1, M=130, K=72, N=64),
    dict(name="conv5", batch=2, M=300, K=128, N=192, taps=5),
    dict(name="conv5_halo", batch=2, M=256, K=128, N=128, taps=5, a_row0=2),
    dict(name="conv5_full_epi", batch=2, M=512, K=256, N=384, taps=5, epilogue="full"),
    dict(name="lin_relu_bf16", batch=1, M=200, K=256, N=512, epilogue="relu_bf16"),
    dict(name="batched_w", batch=4, M=160, K=128, N=320, w_batched=True),
    dict(name="batched_heads_strided", batch=8, M=96, K=128, N=96, w_batched=True, strided=True),
    dict(name="persistent_many_tiles", batch=1, M=128 * 700, K=128, N=256),
    dict(name="conv5_big", batch=1, M=16384, K=768, N=768, taps=5, epilogue="full"),
    # CTA-pair specifics: second CTA of the last pair entirely / partly outside M, N tiles 224 and 176 split in halves,
    # a deep K loop that wraps the 6-stage ring many times, N = 16 (8-row W halves)
    dict(name="pair_m_tail_empty_peer", batch=1, M=128 * 5 + 7, K=256, N=256, epilogue="relu_bf16"),
    dict(name="pair_n224_conv5", batch=1, M=1024, K=768, N=896, taps=5),
    dict(name="pair_n176_full_epi", batch=2, M=768, K=320, N=1408, epilogue="full"),
    dict(name="pair_deep_k", batch=1, M=512, K=3072, N=1536),
    dict(name="pair_n16", batch=1, M=4096, K=128, N=16),
    # mixed-width N tiling (k x 256 + one remainder tile): remainders 128 / 64 / 32 / 16 (the last one only tiles that way
    # with single-CTA MMAs; CTA pairs fall back to the divisor tiling), with every epilogue output and with conv taps
    dict(name="mixed_n1152_conv5_full_epi", batch=1, M=640, K=256, N=1152, taps=5, epilogue="full"),
    dict(name="mixed_n288_relu", batch=2, M=300, K=128, N=288, epilogue="relu_bf16"),
    dict(name="mixed_n272", batch=1, M=260, K=64, N=272),
    dict(name="mixed_n832_batched", batch=3, M=200, K=128, N=832, w_batched=True),
    dict(name="shared_a_batched_w", batch=6, M=128, K=128, N=48, w_batched=True, a_shared=True, epilogue="relu_bf16"),
    # second operand pair (AgGemmArgs.A2 / W2): the DNAEmbed shape (conv5 768 -> 768 + a K2 = 64 im2col product, every
    # epilogue output), a ragged M with a K2 tail block, and a mixed-width N with two batches
    dict(name="second_pair_embed_shape", batch=1, M=4096, K=768, N=768, taps=5, epilogue="full", K2=64),
    dict(name="second_pair_ragged_k2tail", batch=2, M=300, K=128, N=192, taps=5, K2=72),
    dict(name="second_pair_mixed_n1152", batch=2, M=640, K=256, N=1152, epilogue="relu_bf16", K2=128),
]
