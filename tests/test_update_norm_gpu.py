"""-m gpu: BatchRMSNorm with UPDATING statistics (AG:324-361 with update_running_var = the reference's
default, eval included; get_maybe_dist_var AG:276-300) — the three kernels against torch, and the whole forward of a
default-constructed model against a golden of the unmodified reference (tools/make_golden_update.py): embeddings of
the second of two consecutive forwards and all 81 running_var buffers after each."""
from pathlib import Path

import numpy as np
import pytest
import torch

from alphagenome_b200 import ops
from alphagenome_b200.init_weights import apply_synth_weights
from alphagenome_b200.model import AlphaGenome, set_update_running_var
from tests.golden_util import UPDATE_CASE, make_update_inputs, sample_rows

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("shape,view", [((2, 1000, 768), None), ((1, 4099, 128), None), ((3, 640, 1536), (64, 513))])
def test_bn_stats_and_update_match_torch(shape, view):
    g = torch.Generator(device="cuda").manual_seed(1)
    B, R, C = shape
    x = torch.randn(B, R, C, device="cuda", generator=g) * 3.0 + 40.0 * torch.randn(C, device="cuda", generator=g)   # |mean| >> std
    xs = x if view is None else x[:, view[0]: view[0] + view[1]]
    sums = torch.empty(2, C, device="cuda", dtype=torch.float64)
    ops.bn_stats(xs, sums)
    flat = xs.reshape(-1, C).double()
    # fp32 partial sums over 8 rows per thread, then fp64: ~1e-7 relative
    assert torch.allclose(sums[0], flat.sum(0), rtol=2e-6, atol=1e-3)
    assert torch.allclose(sums[1], flat.square().sum(0), rtol=2e-6, atol=1e-3)
    rv = torch.rand(C, device="cuda", generator=g) + 0.5
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g)
    bias = torch.randn(C, device="cuda", generator=g)
    want_rv = torch.lerp(rv, xs.reshape(-1, C).var(dim=0, unbiased=False), 0.1)
    scale = torch.empty(C, device="cuda")
    fold = torch.empty(C, device="cuda")
    ops.bn_update(sums, flat.shape[0], rv, gamma, 0.1, 1e-5, scale, fold_bias=bias, fold_beta=beta, fold_shift_out=fold)
    assert torch.allclose(rv, want_rv, rtol=1e-4, atol=1e-6)
    want_s = gamma / torch.sqrt(want_rv + 1e-5)
    assert torch.allclose(scale, want_s, rtol=1e-4, atol=1e-6)
    assert torch.allclose(fold, bias * want_s + beta, rtol=1e-4, atol=1e-5)


def test_affine_act_matches_torch():
    g = torch.Generator(device="cuda").manual_seed(2)
    B, R, C = 2, 333, 896
    big = torch.randn(B, R + 10, C + 128, device="cuda", generator=g)
    x = big[:, 5:5 + R, :C]                                   # strided view
    s = torch.randn(C, device="cuda", generator=g)
    t = torch.randn(B, C, device="cuda", generator=g)
    res = torch.randn(B, R, C, device="cuda", generator=g)
    o32 = torch.empty(B, R, C, device="cuda")
    o16 = torch.empty(B, R, C, device="cuda", dtype=torch.bfloat16)
    ops.affine_act(x, s, t, res=res, out_f32=o32, out_bf16=o16, act=ops.ACT_GELU1702)
    y = x * s + t[:, None, :] + res
    want = y * torch.sigmoid(1.702 * y)
    assert torch.allclose(o32, want, rtol=1e-4, atol=1e-5)
    assert (o16.float() - want).abs().max() <= 8e-3 * want.abs().max()
    ops.affine_act(x, None, None, out_f32=o32)               # identity copy through the strided view
    assert torch.equal(o32, x)


def test_updating_forward_matches_reference_golden():
    g = np.load(GOLD / f"{UPDATE_CASE['name']}.npz")
    model = AlphaGenome().cuda().eval()                     # default-constructed: every BatchRMSNorm updates (AG:309, 331)
    apply_synth_weights(model, seed=UPDATE_CASE["weight_seed"])
    rv_keys = [k for k in model.state_dict() if k.endswith("running_var")]
    assert len(rv_keys) == 81
    worst_rv = 0.0
    for step, (seq, org) in enumerate(make_update_inputs(UPDATE_CASE)):
        emb = model(seq.cuda(), org.cuda())
        torch.cuda.synchronize()
        sd = model.state_dict()
        for k in rv_keys:
            got, want = sd[k].float().cpu().numpy(), g[f"rv{step}/{k}"]
            rel = float(np.abs(got - want).max() / np.abs(want).max())
            worst_rv = max(worst_rv, rel)
            assert rel <= 2e-2, (step, k, rel)               # bf16 operand noise in the activations the variance is taken of
    for name, got in zip(("embeds_1bp", "embeds_128bp", "embeds_pair"), emb):
        t = got.float().cpu()
        idx = sample_rows(name, t.shape)
        t = (t if idx is None else t[:, idx]).numpy()
        err = np.abs(t - g[name]).max() / float(g[name + "__absmax"])
        assert err <= 1e-2, (name, err)
    print("worst running_var rel diff", worst_rv)


def test_frozen_and_updating_differ_and_freeze_restores_fused_path():
    """set_update_running_var(model, False) after updating forwards: the fused (folded) path must see the statistics
    the updating forwards left in the buffers."""
    model = AlphaGenome().cuda().eval()
    apply_synth_weights(model, seed=1)
    seq = torch.from_numpy(np.random.default_rng(5).integers(0, 4, size=(1, 4096))).cuda()
    rv0 = model.transformer_unet.transformer.layers[3][0].post_rmsnorm.running_var.clone()
    a = model(seq, 0)
    rv1 = model.transformer_unet.transformer.layers[3][0].post_rmsnorm.running_var.clone()
    assert not torch.equal(rv0, rv1)
    set_update_running_var(model, False)
    b = model(seq, 0)                 # frozen, with the statistics of forward `a` (which normalised with them too)
    c = model(seq, 0)
    assert torch.equal(model.transformer_unet.transformer.layers[3][0].post_rmsnorm.running_var, rv1)
    for x, y in zip(b, c):
        assert torch.equal(x, y)
    for x, y in zip(a, b):            # same statistics, fused vs un-fused arithmetic: equal up to bf16 rounding noise
        assert ((x - y).abs().max() / y.abs().max()).item() < 1e-2


def test_updating_under_cuda_graph_updates_on_every_replay():
    model = AlphaGenome().cuda().eval()
    apply_synth_weights(model, seed=1)
    model.enable_cuda_graphs(True)
    seq = torch.from_numpy(np.random.default_rng(5).integers(0, 4, size=(1, 2048))).cuda()
    norm = model.transformer_unet.downs[2].conv_out.norm
    seen = [norm.running_var.clone()]
    for _ in range(3):
        model(seq, 0)
        torch.cuda.synchronize()
        seen.append(norm.running_var.clone())
    assert all(not torch.equal(seen[i], seen[i + 1]) for i in range(3))
