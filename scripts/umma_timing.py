"""tcgen05.mma issue-cost table (experiment, DESIGN.md "lessons"): needs the debug build of the probes library.
    python -m montecarlopredictivecoding_b200.build --debug && python scripts/umma_timing.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from montecarlopredictivecoding_b200 import build as B  # noqa: E402

os.environ["MCPC_UMMA_TIMING"] = "1"
lib = C.CDLL(B.PROBES_LIB.replace(".so", "_debug.so"))
t = torch.zeros(1 << 20, device="cuda")
lib.mcpc_debug_umma(*[C.c_void_p(t.data_ptr())] * 3, 128, 16, *[C.c_void_p(t.data_ptr())] * 3, None)
torch.cuda.synchronize()
