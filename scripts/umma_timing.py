import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from montecarlopredictivecoding_b200 import _native as N
os.environ["MCPC_UMMA_TIMING"] = "1"
lib = N.load()
t = torch.zeros(1 << 20, device="cuda")
lib.mcpc_debug_umma(t.data_ptr(), t.data_ptr(), t.data_ptr(), 128, 16, t.data_ptr(), t.data_ptr(), t.data_ptr(), None)
torch.cuda.synchronize()
