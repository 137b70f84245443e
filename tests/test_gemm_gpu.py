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
