"""ctypes binding of libmcpc_b200.so (include/mcpc_b200.h).

There is no CPU fallback: if the library cannot be loaded (or built with nvcc) every call
raises.  The structures below mirror the C header field by field.
"""
import ctypes as C
import os
import threading

MAX_LAYERS = 8
ACT_IDENTITY, ACT_RELU, ACT_TANH = 0, 1, 2
TOP_NONE, TOP_ZERO, TOP_GAUSS, TOP_BERNOULLI = 0, 1, 2, 3
OPT_SGD, OPT_ADAM = 0, 1
NOISE_NONE, NOISE_SUPPLIED, NOISE_PHILOX = 0, 1, 2
PREC_FP32, PREC_BF16 = 0, 1
MODE_RESIDENT_FP32, MODE_RESIDENT_BF16, MODE_STREAMING_BF16 = 0, 1, 2
ABI_VERSION = 1

_FP = C.c_void_p   # device pointers travel as integers


class McpcNet(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32),
        ("d_in", C.c_int32),
        ("dims", C.c_int32 * MAX_LAYERS),
        ("d_out", C.c_int32),
        ("act", C.c_int32 * MAX_LAYERS),
        ("energy_scale", C.c_float * MAX_LAYERS),
        ("energy_coefficient", C.c_float),
        ("top", C.c_int32),
        ("top_inv_var", C.c_float),
        ("mask_start_col", C.c_int32),
    ]


class McpcIO(C.Structure):
    _fields_ = [
        ("W", _FP * (MAX_LAYERS + 1)),
        ("b", _FP * (MAX_LAYERS + 1)),
        ("x", _FP * MAX_LAYERS),
        ("inputs", _FP),
        ("target", _FP),
        ("noise", _FP),
        ("adam_m", _FP * MAX_LAYERS),
        ("adam_v", _FP * MAX_LAYERS),
        ("x_grad", _FP * MAX_LAYERS),
        ("energy", _FP),
        ("loss", _FP),
        ("traj_x", _FP * MAX_LAYERS),
        ("traj_out", _FP),
        ("gW", _FP * (MAX_LAYERS + 1)),
        ("gb", _FP * (MAX_LAYERS + 1)),
        ("save_g", _FP),
        ("save_f", _FP),
    ]


class McpcOpts(C.Structure):
    _fields_ = [
        ("lr", C.c_double),
        ("adam_beta1", C.c_double),
        ("adam_beta2", C.c_double),
        ("adam_eps", C.c_double),
        ("noise_scale", C.c_double),
        ("seed", C.c_uint64),
        ("chain_offset", C.c_uint64),
        ("n_steps", C.c_int32),
        ("t_begin", C.c_int32),
        ("optimizer", C.c_int32),
        ("update_x", C.c_int32),
        ("adam_step0", C.c_int32),
        ("noise_mode", C.c_int32),
        ("traj_every", C.c_int32),
        ("save_begin", C.c_int32),
        ("save_end", C.c_int32),
        ("precision", C.c_int32),
    ]


class McpcGradIO(C.Structure):
    _fields_ = [
        ("save_g", _FP),
        ("save_f", _FP),
        ("inputs", _FP),
        ("gW", _FP * (MAX_LAYERS + 1)),
        ("gb", _FP * (MAX_LAYERS + 1)),
        ("scratch", _FP),
        ("scratch_bytes", C.c_size_t),
    ]


MAX_PTENSORS = 2 * (MAX_LAYERS + 1)


class McpcPStep(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("n_tensors", C.c_int32),
        ("param", _FP * MAX_PTENSORS),
        ("grad", _FP * MAX_PTENSORS),
        ("state1", _FP * MAX_PTENSORS),
        ("state2", _FP * MAX_PTENSORS),
        ("numel", C.c_uint64 * MAX_PTENSORS),
        ("inv_norm", C.c_double),
        ("lr", C.c_double),
        ("weight_decay", C.c_double),
        ("momentum", C.c_double),
        ("dampening", C.c_double),
        ("beta1", C.c_double),
        ("beta2", C.c_double),
        ("eps", C.c_double),
        ("nesterov", C.c_int32),
        ("first_step", C.c_int32),
        ("step", C.c_int32),
    ]


class NativeError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()
LIB_NAME = "libmcpc_b200.so"
EXPORTS = ("mcpc_version", "mcpc_last_error", "mcpc_launch_count", "mcpc_workspace_bytes", "mcpc_save_layout", "mcpc_infer_mode", "mcpc_infer_fuses_weight_grad", "mcpc_infer", "mcpc_weight_grad",
           "mcpc_fill_noise", "mcpc_marginal_ll_workspace_bytes", "mcpc_marginal_ll_bernoulli", "mcpc_traj_stats_update",
           "mcpc_p_step")
PROBE_EXPORTS = ("mcpc_probes_last_error", "mcpc_debug_umma", "mcpc_debug_tma")


def lib_path():
    override = os.environ.get("MCPC_NATIVE_LIB")          # e.g. the -DMCPC_DEBUG_BUILD library (build.py --debug)
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)


def load():
    """Load (building in-tree with nvcc when the .so is missing) and type the entry points."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not os.environ.get("MCPC_NATIVE_LIB"):
            # never run kernels that do not match the sources in the tree: rebuild when the stamp disagrees (a no-op
            # otherwise); without nvcc a stale library is an error, not a silent fallback
            from . import build as _build
            if _build.is_stale():
                import shutil
                if shutil.which(_build.nvcc_path()) or os.path.exists(_build.nvcc_path()):
                    _build.build()
                elif os.path.exists(path):
                    raise NativeError(f"{LIB_NAME} was built from different sources and nvcc is not available to rebuild it")
        lib = C.CDLL(path)
        lib.mcpc_version.restype = C.c_int
        lib.mcpc_last_error.restype = C.c_char_p
        lib.mcpc_launch_count.restype = C.c_uint64
        lib.mcpc_workspace_bytes.restype = C.c_int
        lib.mcpc_workspace_bytes.argtypes = [C.POINTER(McpcNet), C.c_int32, C.c_int32, C.c_int32,
                                             C.POINTER(C.c_size_t)]
        lib.mcpc_save_layout.restype = C.c_int
        lib.mcpc_save_layout.argtypes = [C.POINTER(McpcNet), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int32)]
        lib.mcpc_infer_mode.restype = C.c_int
        lib.mcpc_infer_mode.argtypes = [C.POINTER(McpcNet), C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
        lib.mcpc_infer_fuses_weight_grad.restype = C.c_int
        lib.mcpc_infer_fuses_weight_grad.argtypes = [C.POINTER(McpcNet), C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
        lib.mcpc_infer.restype = C.c_int
        lib.mcpc_infer.argtypes = [C.POINTER(McpcNet), C.POINTER(McpcIO), C.POINTER(McpcOpts), C.c_int32,
                                   C.c_void_p, C.c_size_t, C.c_void_p]
        lib.mcpc_weight_grad.restype = C.c_int
        lib.mcpc_weight_grad.argtypes = [C.POINTER(McpcNet), C.POINTER(McpcGradIO), C.c_int32, C.c_int32, C.c_int32,
                                         C.c_void_p]
        lib.mcpc_fill_noise.restype = C.c_int
        lib.mcpc_fill_noise.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_uint64, C.c_int32, C.c_int32,
                                        C.c_float, C.c_void_p, C.c_void_p]
        lib.mcpc_marginal_ll_workspace_bytes.restype = C.c_int
        lib.mcpc_marginal_ll_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]
        lib.mcpc_marginal_ll_bernoulli.restype = C.c_int
        lib.mcpc_marginal_ll_bernoulli.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_float,
                                                   C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mcpc_traj_stats_update.restype = C.c_int
        lib.mcpc_traj_stats_update.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                               C.c_void_p]
        lib.mcpc_p_step.restype = C.c_int
        lib.mcpc_p_step.argtypes = [C.POINTER(McpcPStep), C.c_void_p]
        got = lib.mcpc_version()
        if got != ABI_VERSION:
            raise NativeError(f"{LIB_NAME} ABI version {got}, python binding expects {ABI_VERSION}")
        _lib = lib
    return _lib


_probes = None


def load_probes():
    """libmcpc_b200_probes.so (include/mcpc_b200_probes.h): validation-only known-answer tests of the tcgen05 / TMEM / TMA
    primitives, kept out of the product library."""
    global _probes
    if _probes is None:
        load()                              # builds both libraries when stale
        lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmcpc_b200_probes.so"))
        lib.mcpc_probes_last_error.restype = C.c_char_p
        lib.mcpc_debug_umma.restype = C.c_int
        lib.mcpc_debug_umma.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mcpc_debug_tma.restype = C.c_int
        lib.mcpc_debug_tma.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
        _probes = lib
    return _probes


def check_probe(rc, what):
    if rc != 0:
        raise NativeError(f"{what} failed (code {rc}): {load_probes().mcpc_probes_last_error().decode(errors='replace')}")


def check(rc, what):
    if rc != 0:
        msg = load().mcpc_last_error().decode(errors="replace")
        if rc == -2:
            raise NotImplementedError(f"{what}: {msg}")
        raise NativeError(f"{what} failed (code {rc}): {msg}")
