"""Known-answer test of the tcgen05 / TMEM / bulk-copy primitives (csrc/umma.cuh) against torch matmul on
bf16-rounded operands: K-major and MN-major (transposed) reads of one canonical no-swizzle weight tile."""
import ctypes as C

import pytest
import torch

from montecarlopredictivecoding_b200 import _native as N

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("Kin,Nrows", [(128, 16), (32, 16), (256, 32), (128, 32), (16, 16)])
def test_umma_probe(Kin, Nrows):
    lib = N.load_probes()
    dev = torch.device("cuda:0")
    torch.manual_seed(Kin * 100 + Nrows)
    Wt = torch.randn(128, Kin, device=dev)
    Bx = torch.randn(Nrows, Kin, device=dev)
    G = torch.randn(Nrows, 128, device=dev)
    D1 = torch.full((128, Nrows), float("nan"), device=dev)
    D2 = torch.full((128, Nrows), float("nan"), device=dev)
    ws = torch.empty(128 * Kin * 2 + 1024, dtype=torch.uint8, device=dev)
    rc = lib.mcpc_debug_umma(Wt.data_ptr(), Bx.data_ptr(), G.data_ptr(), Kin, Nrows, D1.data_ptr(), D2.data_ptr(),
                             ws.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    N.check_probe(rc, "mcpc_debug_umma")
    torch.cuda.synchronize()
    Wb, Bb, Gb = Wt.bfloat16().double(), Bx.bfloat16().double(), G.bfloat16().double()
    ref1 = (Wb @ Bb.T).float()
    ref2 = (Wb.T @ Gb.T).float()          # [Kin, N]
    m = min(Kin, 128)
    e1 = (D1 - ref1).abs().max().item()
    e2 = (D2[:m] - ref2[:m]).abs().max().item()
    print(f"Kin={Kin} N={Nrows}: |D1-ref|={e1:.3e} |D2-ref|={e2:.3e}")
    assert e1 < 2e-3 * max(1.0, ref1.abs().max().item()), e1
    assert e2 < 2e-3 * max(1.0, ref2.abs().max().item()), e2
