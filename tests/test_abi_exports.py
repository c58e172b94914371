"""The C-ABI library must load on a CPU-only box and export every symbol include/mcpc_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

from montecarlopredictivecoding_b200 import _native as N
from montecarlopredictivecoding_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header="mcpc_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcpc_[a-z_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = B.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_functions()
    assert set(declared) == set(N.EXPORTS), (declared, N.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    lib.mcpc_version.restype = ctypes.c_int
    assert lib.mcpc_version() == N.ABI_VERSION


def test_probe_library_is_separate_from_the_product_library():
    """The known-answer kernels of the tcgen05 / TMA primitives live in their own library."""
    B.build()
    probes = ctypes.CDLL(B.PROBES_LIB)
    declared = _declared_functions("mcpc_b200_probes.h")
    assert set(declared) == set(N.PROBE_EXPORTS), (declared, N.PROBE_EXPORTS)
    for name in declared:
        assert hasattr(probes, name), name
    product = ctypes.CDLL(B.LIB)
    for name in ("mcpc_debug_umma", "mcpc_debug_tma"):
        assert not hasattr(product, name), f"{name} must not ship in libmcpc_b200.so"


def test_struct_sizes_match_header_layout():
    # sizes computed from the header by hand: any drift between the C structs and the ctypes mirrors shows here
    assert ctypes.sizeof(N.McpcNet) == 4 * (2 + 8 + 1 + 8 + 8 + 1 + 1 + 1 + 1)
    assert ctypes.sizeof(N.McpcIO) == 8 * (9 + 9 + 8 + 3 + 8 + 8 + 8 + 2 + 8 + 1 + 9 + 9 + 2)
    assert ctypes.sizeof(N.McpcOpts) == 8 * 7 + 4 * 10
    assert ctypes.sizeof(N.McpcGradIO) == 8 * (3 + 9 + 9 + 2)
    assert ctypes.sizeof(N.McpcPStep) == 4 * 2 + 8 * (4 * 18) + 8 * 18 + 8 * 8 + 4 * 3 + 4     # 4 bytes of tail padding


def test_argument_validation_without_gpu():
    lib = N.load()
    net = N.McpcNet()
    net.n_layers = 0
    need = ctypes.c_size_t(0)
    rc = lib.mcpc_workspace_bytes(ctypes.byref(net), 4, 4, N.PREC_FP32, ctypes.byref(need))
    assert rc == -1 and b"n_layers" in lib.mcpc_last_error()
    net.n_layers, net.d_in, net.d_out = 1, 3, 2
    net.dims[0] = 3
    net.top = N.TOP_NONE
    rc = lib.mcpc_workspace_bytes(ctypes.byref(net), 4, 4, N.PREC_FP32, ctypes.byref(need))
    assert rc == 0 and need.value > 0
    net.dims[0] = 1 << 20          # far too wide for the resident fp32 kernel
    rc = lib.mcpc_workspace_bytes(ctypes.byref(net), 4, 4, N.PREC_FP32, ctypes.byref(need))
    assert rc == -2
