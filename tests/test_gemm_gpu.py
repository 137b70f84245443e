"""-m gpu: tcgen05 implicit-GEMM conv / linear kernel vs the fp64 restatement of ag_gemm_fwd's formula, with
single-CTA MMAs (cta_group::1) and with CTA pairs (cta_group::2); the two must also agree bit for bit."""
import pytest
import torch

from tests.gemm_cases import CASES, run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cta_group", [1, 2], ids=["cg1", "cg2"])
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_gemm_case(case, cta_group):
    r = run_case(**case, cta_group=cta_group)
    assert r["ok"], r


def test_cta_pair_matches_single_cta_bitwise():
    from alphagenome_b200 import ops

    torch.manual_seed(0)
    A = torch.randn(1, 1000, 384, device="cuda").bfloat16()
    W = (torch.randn(5, 448, 384, device="cuda") * 0.05).bfloat16()
    outs = []
    for cg in (1, 2):
        prev = ops.set_gemm_cta_group(cg)
        try:
            o = torch.empty(1, 1000, 448, device="cuda")
            ops.gemm(A, W, out=o)
            torch.cuda.synchronize()
        finally:
            ops.set_gemm_cta_group(prev)
        outs.append(o)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("cta_group", [1, 2], ids=["cg1", "cg2"])
def test_mixed_width_tiles_with_padded_weight_rows(cta_group):
    """n_pad: W has 282 rows, the kernel computes N = 288 = 256 + 32 columns; the rows the remainder tile's W box reads
    past the tensor are TMA zero fill (the heads' padded track counts)."""
    from alphagenome_b200 import ops

    torch.manual_seed(1)
    A = torch.randn(2, 500, 192, device="cuda").bfloat16()
    W = (torch.randn(1, 282, 192, device="cuda") * 0.07).bfloat16()
    out = torch.full((2, 500, 288), float("nan"), device="cuda")
    prev = ops.set_gemm_cta_group(cta_group)
    try:
        ops.gemm(A, W, n_pad=288, out=out)
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_cta_group(prev)
    want = A.double() @ W[0].double().T
    assert (out[..., :282].double() - want).abs().max() <= 2e-3 * want.abs().max()
    assert torch.equal(out[..., 282:], torch.zeros(2, 500, 6, device="cuda"))
