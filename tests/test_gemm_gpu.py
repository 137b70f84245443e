"""-m gpu: tcgen05 implicit-GEMM conv / linear kernel vs the fp64 restatement of ag_gemm_fwd's formula."""
import pytest

from tests.gemm_cases import CASES, run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_gemm_case(case):
    r = run_case(**case)
    assert r["ok"], r
