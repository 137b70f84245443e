"""-m gpu: the conversion path end to end on the B200 kernels.

  * the reference's known-answer standardized-conv bake (tests/test_conversion_shape_transforms.py:49-71: seed 0,
    width 5, 4 -> 7 channels) pushed through ag_gemm_fwd as a width-5 conv;
  * a synthetic JAX-layout checkpoint -> alphagenome_b200.convert -> load_state_dict(strict) -> B200 forward with the
    checkpoint's heads, against the CPU oracle on the same converted state_dict.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from alphagenome_b200 import convert as CV
from alphagenome_b200 import ops
from alphagenome_b200.model import AlphaGenome, set_update_running_var
from oracle import trunk_oracle as O
from tests.convert_util import synth_checkpoint

pytestmark = pytest.mark.gpu


def test_baked_standardized_conv_through_ag_gemm():
    torch.manual_seed(0)
    width, cin, cout = 5, 4, 7
    w = torch.randn(width, cin, cout)
    scale = torch.randn(1, 1, cout)
    bias = torch.randn(cout)
    bw, bb = CV.apply_rule(CV.Rule("c.weight", ("w",), "standardized", extra=(("scale", "s"), ("bias", "b"))), [w], dict(scale=scale, bias=bias))
    L = 300
    x = torch.randn(1, L, cin)
    # ag_gemm_fwd wants K % 8 == 0 and N % 16 == 0: zero input channels 4..7, weight rows 7..15 read as zero (n_pad)
    A = torch.zeros(1, L, 8, dtype=torch.bfloat16, device="cuda")
    A[..., :cin] = x.cuda().to(torch.bfloat16)
    W = torch.zeros(width, cout, 8, dtype=torch.bfloat16, device="cuda")
    W[..., :cin] = bw.permute(2, 0, 1).cuda().to(torch.bfloat16)           # [taps, C_out, C_in]
    shift = torch.zeros(16, device="cuda")
    shift[:cout] = bb.cuda()
    out = torch.empty(1, L, 16, device="cuda")
    ops.gemm(A, W, n_pad=16, pre_shift=shift, out=out)
    torch.cuda.synchronize()
    xr = A[..., :cin].float().cpu()
    wr = W[..., :cin].float().cpu().permute(1, 2, 0)                        # the same bf16-rounded operands, fp32 math
    want = F.conv1d(xr.transpose(1, 2), wr, bb, padding=width // 2).transpose(1, 2)
    got = out[..., :cout].cpu()
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-4), (got - want).abs().max()
    assert torch.equal(out[..., cout:].cpu(), torch.zeros(1, L, 16 - cout))


def test_converted_checkpoint_runs_and_matches_oracle():
    """A JAX-layout checkpoint (the inverse image of a well-behaved synthetic state_dict, so the weight distribution is
    the one the north-star tolerance is stated for) -> convert -> load -> B200 forward + the checkpoint's heads."""
    from alphagenome_b200.init_weights import apply_synth_weights
    from tests.convert_util import checkpoint_from_state_dict

    src = AlphaGenome()
    for organism in ("human", "mouse"):
        src.add_reference_heads(organism)
    apply_synth_weights(src, seed=3)
    params, state = checkpoint_from_state_dict(src.state_dict())
    model = AlphaGenome()
    for organism in ("human", "mouse"):
        model.add_reference_heads(organism)
    assert model.load_from_official_jax_model(params=params, state=state, strict=False) is model
    # the conversion is the exact inverse except that baked conv kernels are centred per output channel
    a, b = src.state_dict(), model.state_dict()
    for k in a:
        if k.endswith("net.2.weight") and a[k].shape[-1] > 1:
            assert torch.allclose(b[k], a[k] - a[k].mean(dim=(1, 2), keepdim=True), atol=1e-6), k
        elif a[k].is_floating_point():
            assert torch.allclose(b[k], a[k], atol=1e-6), k
    model = model.cuda().eval()
    set_update_running_var(model, False)
    L = 4096
    rng = np.random.default_rng(12)
    seq = torch.from_numpy(rng.integers(0, 4, size=(2, L)).astype(np.int64))
    org = torch.tensor([0, 1])
    pos = torch.from_numpy(np.sort(rng.choice(L, size=(2, 4, 6)), axis=-1).astype(np.int64))
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    emb = model(seq.cuda(), org.cuda(), return_embeds=True)
    heads = model(seq.cuda(), org.cuda(), splice_site_positions=pos.cuda())
    torch.cuda.synchronize()
    ref = O.alphagenome_embeds(sd, seq, org)
    for name, got in zip(("embeds_1bp", "embeds_128bp", "embeds_pair"), emb):
        err = ((got.float().cpu() - ref[name]).abs().max() / ref[name].abs().max()).item()
        assert err <= 1e-2, (name, err)
    ref_heads = O.alphagenome_reference_heads(sd, ref, org, ("human", "mouse"), pos)
    n = 0
    for organism in ("human", "mouse"):
        for name, want in ref_heads[organism].items():
            got = heads[organism][name]
            pairs = [(got, want)] if torch.is_tensor(want) else [(got[k], want[k]) for k in want if torch.is_tensor(want[k]) and want[k].is_floating_point()]
            for g, w in pairs:
                g, w = g.float().cpu().double().flatten(), w.double().flatten()
                g, w = g - g.mean(), w - w.mean()
                r = float((g @ w) / (g.norm() * w.norm()).clamp(min=1e-30))
                assert r >= 0.999, (organism, name, r)
                n += 1
    assert n >= 20
