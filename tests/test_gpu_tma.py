"""Known-answer test of the TMA + SWIZZLE_128B operand path (csrc/tma.cuh, tma_probe.cu): all four combinations
of K-major / MN-major storage of A and B against torch matmul on bf16-rounded operands."""
import ctypes as C

import pytest
import torch

from montecarlopredictivecoding_b200 import _native as N

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("Ncols", [64, 256])
def test_tma_probe(a_mn, b_mn, Ncols):
    lib = N.load_probes()
    dev = torch.device("cuda:0")
    torch.manual_seed(a_mn * 10 + b_mn + Ncols)
    A = torch.randn(128, 64, device=dev)
    B = torch.randn(Ncols, 64, device=dev)
    A_st = A.t().contiguous() if a_mn else A
    B_st = B.t().contiguous() if b_mn else B
    D = torch.full((128, Ncols), float("nan"), device=dev)
    ws = torch.empty((128 + Ncols) * 64 * 2 + 1024, dtype=torch.uint8, device=dev)
    rc = lib.mcpc_debug_tma(A_st.data_ptr(), B_st.data_ptr(), Ncols, a_mn, b_mn, D.data_ptr(), ws.data_ptr(),
                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    N.check_probe(rc, "mcpc_debug_tma")
    torch.cuda.synchronize()
    ref = (A.bfloat16().double() @ B.bfloat16().double().t()).float()
    err = (D - ref).abs().max().item()
    print(f"a_mn={a_mn} b_mn={b_mn} N={Ncols}: max err {err:.3e}")
    assert err < 1e-3 * max(1.0, ref.abs().max().item()), err
