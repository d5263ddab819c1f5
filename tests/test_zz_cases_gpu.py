"""The five activation-exchange patterns of test_activation_cases_cpu.py on the CUDA backend (device tensors, device
communication buffers, loop-back ranks).  Last-sorted file: first hardware run at round end."""
import pytest

from test_activation_cases_cpu import CASES, run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case,ptype,pdist,cdist", CASES, ids=["case%d" % c[0] for c in CASES])
def test_activation_exchange_case_device(case, ptype, pdist, cdist):
    run_case(case, ptype, pdist, cdist, "cuda")


@pytest.mark.parametrize("case,ptype,pdist,cdist", CASES, ids=["case%d" % c[0] for c in CASES])
def test_activation_exchange_case_fused_device(case, ptype, pdist, cdist):
    """start_comm_fused on the device: cases 4 / 5 are ONE strided all-to-all kernel (rectangles pulled straight out of the
    peers' unpacked tensors into the consumer's unpacked tensor), the others pack + exchange + unpack inside the library."""
    run_case(case, ptype, pdist, cdist, "cuda", fused=True)
