"""Tensor-level front of the C ABI: packs torch tensors into McpcNet/McpcIO/McpcOpts and calls
libmcpc_b200.so on torch's current CUDA stream.  This is the only engine the product ships;
it refuses anything that is not a contiguous fp32 CUDA tensor (no CPU fallback).
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from .. import _native as N
from .plan import NetPlan, TopPlan


@dataclass
class InferCall:
    plan: NetPlan
    top: TopPlan
    energy_coefficient: float
    B: int
    W: List[torch.Tensor]
    b: List[Optional[torch.Tensor]]
    x: List[torch.Tensor]
    inputs: Optional[torch.Tensor]
    target: Optional[torch.Tensor]
    energy: Optional[torch.Tensor]          # float64 [n_steps]
    loss: Optional[torch.Tensor]            # float64 [n_steps]
    n_steps: int
    t_begin: int = 0
    optimizer: int = N.OPT_SGD
    update_x: bool = True
    lr: float = 0.1
    betas: tuple = (0.9, 0.999)
    adam_eps: float = 1e-8
    adam_step0: int = 0
    adam_m: Optional[List[torch.Tensor]] = None
    adam_v: Optional[List[torch.Tensor]] = None
    noise_mode: int = N.NOISE_NONE
    noise: Optional[torch.Tensor] = None    # [n_steps, B, SD] raw gradient noise
    noise_scale: float = 0.0
    seed: int = 0
    chain_offset: int = 0
    x_grad: Optional[List[torch.Tensor]] = None
    traj_x: List[Optional[torch.Tensor]] = field(default_factory=list)
    traj_out: Optional[torch.Tensor] = None
    traj_every: int = 1
    save_g: Optional[torch.Tensor] = None
    save_f: Optional[torch.Tensor] = None
    save_begin: int = 0
    save_end: int = 0
    precision: int = N.PREC_FP32
    # MODE_STREAMING: accumulators of the weight update of steps [save_begin, save_end) (instead of save_g/save_f)
    gW: Optional[List[Optional[torch.Tensor]]] = None
    gb: Optional[List[Optional[torch.Tensor]]] = None


_net_cache = {}
_ENV_KNOBS = ("MCPC_FORCE_STREAMING", "MCPC_TC_ROWS", "MCPC_ROWS", "MCPC_WIDE_CTAS", "MCPC_TC_NOSPEC", "MCPC_WIDE_CG",
              "MCPC_WIDE_SLOTS")


_ENV_DATA = getattr(os.environ, "_data", None)       # CPython/posix: the dict behind os.environ (bytes keys)
_ENV_KNOBS_B = tuple(k.encode() for k in _ENV_KNOBS)
if not isinstance(_ENV_DATA, dict) or any(not isinstance(k, bytes) for k in list(_ENV_DATA)[:1]):
    _ENV_DATA = None


def _env_key():
    """Debug / test knobs the library reads with getenv(): part of every cache key that depends on them (read on
    every call -- tests flip them between calls -- so through the raw dict: 7 lookups instead of 7 encode + lookups)."""
    if _ENV_DATA is not None:
        get = _ENV_DATA.get
        return tuple(get(k) for k in _ENV_KNOBS_B)
    return tuple(os.environ.get(k) for k in _ENV_KNOBS)


def _raw_stream(dev):
    """cudaStream_t of torch's current stream on ``dev`` (the kernels are launched on it)."""
    try:
        return torch._C._cuda_getCurrentRawStream(dev.index if dev.index is not None else torch.cuda.current_device())
    except AttributeError:      # older / newer torch without the private accessor
        return torch.cuda.current_stream(dev).cuda_stream



def _net_key(plan: NetPlan, top: TopPlan, energy_coefficient: float):
    return (plan.d_in, tuple(plan.dims), plan.d_out, tuple(plan.act), tuple(plan.energy_scale), float(energy_coefficient),
            top.kind, float(top.inv_var), top.mask_start)


def net_struct(plan: NetPlan, top: TopPlan, energy_coefficient: float) -> N.McpcNet:
    """The McpcNet of a (plan, top, coefficient); built once per distinct network description (the structs are
    read-only for the library)."""
    key = _net_key(plan, top, energy_coefficient)
    hit = _net_cache.get(key)
    if hit is not None:
        return hit
    net = _build_net_struct(plan, top, energy_coefficient)
    if len(_net_cache) > 256:
        _net_cache.clear()
    _net_cache[key] = net
    return net


def _build_net_struct(plan: NetPlan, top: TopPlan, energy_coefficient: float) -> N.McpcNet:
    net = N.McpcNet()
    net.n_layers = plan.L
    net.d_in = plan.d_in
    net.d_out = plan.d_out
    for l in range(plan.L):
        net.dims[l] = plan.dims[l]
        net.act[l] = plan.act[l]
        net.energy_scale[l] = plan.energy_scale[l]
    net.energy_coefficient = energy_coefficient
    net.top = top.kind
    net.top_inv_var = top.inv_var
    net.mask_start_col = top.mask_start
    return net


def _ptr(t: Optional[torch.Tensor], what: str, dtype=torch.float32):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live on a CUDA device: the B200 build has no CPU path")
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f"{what} must be a contiguous {dtype} tensor (got {t.dtype}, contiguous={t.is_contiguous()})")
    return t.data_ptr()


class _OnDevice:
    """``torch.cuda.device(dev)`` only when ``dev`` is not already current (the context manager costs ~10 us)."""

    def __init__(self, dev):
        self._ctx = None if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)

    def __enter__(self):
        if self._ctx is not None:
            self._ctx.__enter__()

    def __exit__(self, *exc):
        if self._ctx is not None:
            return self._ctx.__exit__(*exc)
        return False


class NativeEngine:
    """Calls the sm_100a kernels through the C ABI."""

    name = "native"

    def __init__(self):
        self._lib = N.load()
        self._ws = {}
        self._ws_need = {}
        self._layout_cache = {}
        self._mode_cache = {}
        self._grad_scratch = {}
        self._fuse_cache = {}

    def _workspace(self, net, B, n_steps, precision, device, net_key=None):
        # keyed by the network DESCRIPTION (not the address of a cached struct, which can be reused after a cache flush)
        nkey = (net_key if net_key is not None else bytes(net), B, n_steps, precision, _env_key())
        need_v = self._ws_need.get(nkey)
        if need_v is None:
            need = C.c_size_t(0)
            N.check(self._lib.mcpc_workspace_bytes(C.byref(net), B, n_steps, precision, C.byref(need)),
                    "mcpc_workspace_bytes")
            need_v = need.value
            if len(self._ws_need) > 1024:
                self._ws_need.clear()
            self._ws_need[nkey] = need_v
        need = C.c_size_t(need_v)
        key = (device.index, )
        buf = self._ws.get(key)
        if buf is None or buf.numel() < need.value:
            buf = torch.empty(max(need.value, 1 << 20), dtype=torch.uint8, device=device)
            self._ws[key] = buf
        return buf

    def save_layout(self, plan: NetPlan, top: TopPlan, precision: int):
        """(g_width, f_width, torch dtype) of the operands saved for the weight update."""
        key = (_net_key(plan, top, 1.0), precision)
        hit = self._layout_cache.get(key)
        if hit is not None:
            return hit
        net = net_struct(plan, top, 1.0)
        gw, fw, eb = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        N.check(self._lib.mcpc_save_layout(C.byref(net), precision, C.byref(gw), C.byref(fw), C.byref(eb)),
                "mcpc_save_layout")
        out = (gw.value, fw.value, (torch.float32 if eb.value == 4 else torch.bfloat16))
        self._layout_cache[key] = out
        return out

    def infer_mode(self, plan: NetPlan, top: TopPlan, B: int, precision: int) -> int:
        """N.MODE_*: how mcpc_infer will execute this network (resident one-launch vs streaming per-step GEMMs)."""
        key = (_net_key(plan, top, 1.0), B, precision, _env_key())
        hit = self._mode_cache.get(key)
        if hit is not None:
            return hit
        net = net_struct(plan, top, 1.0)
        mode = C.c_int32(0)
        N.check(self._lib.mcpc_infer_mode(C.byref(net), B, precision, C.byref(mode)), "mcpc_infer_mode")
        self._mode_cache[key] = mode.value
        return mode.value

    def infer_fuses_weight_grad(self, plan: NetPlan, top: TopPlan, B: int, precision: int, has_inputs: bool) -> bool:
        """True when passing the ``.grad`` accumulators to ``infer`` (McpcIO.gW/gb) is the way to get the weight update:
        streaming mode, or the resident bf16 kernel with idle SMs for the concurrent weight-gradient kernel."""
        key = (_net_key(plan, top, 1.0), B, precision, bool(has_inputs), _env_key(), os.environ.get("MCPC_TC_DW_OVERLAP"))
        hit = self._fuse_cache.get(key)
        if hit is not None:
            return hit
        net = net_struct(plan, top, 1.0)
        out = C.c_int32(0)
        N.check(self._lib.mcpc_infer_fuses_weight_grad(C.byref(net), B, precision, 1 if has_inputs else 0, C.byref(out)),
                "mcpc_infer_fuses_weight_grad")
        self._fuse_cache[key] = bool(out.value)
        return bool(out.value)

    def infer(self, c: InferCall) -> None:
        plan = c.plan
        dev = c.x[0].device
        net = net_struct(plan, c.top, c.energy_coefficient)
        io = N.McpcIO()
        for i, (W, b) in enumerate(zip(c.W, c.b)):
            io.W[i] = _ptr(W, f"weight of Linear {i}")
            io.b[i] = _ptr(b, f"bias of Linear {i}")
        for l in range(plan.L):
            io.x[l] = _ptr(c.x[l], f"latent x[{l}]")
            if c.adam_m is not None:
                io.adam_m[l] = _ptr(c.adam_m[l], "adam_m")
                io.adam_v[l] = _ptr(c.adam_v[l], "adam_v")
            if c.x_grad is not None:
                io.x_grad[l] = _ptr(c.x_grad[l], "x_grad")
            if c.traj_x and c.traj_x[l] is not None:
                io.traj_x[l] = _ptr(c.traj_x[l], "traj_x")
        if c.gW is not None:
            for i in range(len(c.gW)):
                io.gW[i] = _ptr(c.gW[i], "gW")
                io.gb[i] = _ptr(c.gb[i], "gb")
        io.inputs = _ptr(c.inputs, "inputs")
        io.target = _ptr(c.target, "target")
        io.noise = _ptr(c.noise, "noise")
        io.energy = _ptr(c.energy, "energy", torch.float64)
        io.loss = _ptr(c.loss, "loss", torch.float64)
        io.traj_out = _ptr(c.traj_out, "traj_out")
        io.save_g = _ptr(c.save_g, "save_g", c.save_g.dtype if c.save_g is not None else torch.float32)
        io.save_f = _ptr(c.save_f, "save_f", c.save_f.dtype if c.save_f is not None else torch.float32)
        o = N.McpcOpts()
        o.lr = float(c.lr)
        o.adam_beta1, o.adam_beta2 = float(c.betas[0]), float(c.betas[1])
        o.adam_eps = float(c.adam_eps)
        o.noise_scale = float(c.noise_scale)
        o.seed = int(c.seed) & 0xFFFFFFFFFFFFFFFF
        o.chain_offset = int(c.chain_offset)
        o.n_steps = int(c.n_steps)
        o.t_begin = int(c.t_begin)
        o.optimizer = int(c.optimizer)
        o.update_x = 1 if c.update_x else 0
        o.adam_step0 = int(c.adam_step0)
        o.noise_mode = int(c.noise_mode)
        o.traj_every = int(c.traj_every)
        o.save_begin = int(c.save_begin)
        o.save_end = int(c.save_end)
        o.precision = int(c.precision)
        ws = self._workspace(net, c.B, c.n_steps, c.precision, dev, _net_key(plan, c.top, c.energy_coefficient))
        stream = _raw_stream(dev)
        with _OnDevice(dev):
            N.check(self._lib.mcpc_infer(C.byref(net), C.byref(io), C.byref(o), c.B, ws.data_ptr(), ws.numel(),
                                         C.c_void_p(stream)), "mcpc_infer")

    def weight_grad(self, plan: NetPlan, top: TopPlan, energy_coefficient: float, B: int, n_save: int,
                    save_g: torch.Tensor, save_f: torch.Tensor, inputs: Optional[torch.Tensor],
                    gW: List[Optional[torch.Tensor]], gb: List[Optional[torch.Tensor]], precision: int) -> None:
        dev = save_g.device
        net = net_struct(plan, top, energy_coefficient)
        io = N.McpcGradIO()
        io.save_g = _ptr(save_g, "save_g", save_g.dtype)
        io.save_f = _ptr(save_f, "save_f", save_f.dtype)
        io.inputs = _ptr(inputs, "inputs")
        for i in range(len(gW)):
            io.gW[i] = _ptr(gW[i], "gW")
            io.gb[i] = _ptr(gb[i], "gb")
        if inputs is not None and precision == N.PREC_BF16:
            # Linear_0 with non-zero inputs: scratch for the per-chain sum of its G operand over the saved steps
            need = B * plan.dims[0]
            buf = self._grad_scratch.get(dev.index)
            if buf is None or buf.numel() < need:
                buf = self._grad_scratch[dev.index] = torch.empty(need, dtype=torch.float32, device=dev)
            io.scratch = buf.data_ptr()
            io.scratch_bytes = buf.numel() * 4
        stream = _raw_stream(dev)
        with _OnDevice(dev):
            N.check(self._lib.mcpc_weight_grad(C.byref(net), C.byref(io), B, n_save, precision, C.c_void_p(stream)),
                    "mcpc_weight_grad")

    def traj_stats(self, traj: torch.Tensor, n_rec: int, count_before: int, mean: torch.Tensor, m2: torch.Tensor) -> None:
        """Fold ``traj[:n_rec]`` ([n_rec, ...] fp32 ring) into the running per-element (mean, m2) accumulators."""
        dev = traj.device
        n_elems = mean.numel()
        if traj[0].numel() != n_elems or m2.numel() != n_elems:
            raise RuntimeError("traj_stats: ring / accumulator shapes disagree")
        stream = _raw_stream(dev)
        with _OnDevice(dev):
            N.check(self._lib.mcpc_traj_stats_update(_ptr(traj, "trajectory ring"), int(n_rec), n_elems, int(count_before),
                                                     _ptr(mean, "mean"), _ptr(m2, "m2"), C.c_void_p(stream)),
                    "mcpc_traj_stats_update")

    def p_step(self, kind: int, params, grads, state1, state2, inv_norm: float, lr: float, weight_decay: float = 0.0,
               momentum: float = 0.0, dampening: float = 0.0, nesterov: bool = False, first_step: int = 0,
               beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, step: int = 0) -> None:
        """One ``mcpc_p_step`` launch: normalise the gradients and apply optim.SGD / optim.Adam to every tensor."""
        dev = params[0].device
        a = N.McpcPStep()
        a.kind = int(kind)
        a.n_tensors = len(params)
        for i, (q, g) in enumerate(zip(params, grads)):
            a.param[i] = _ptr(q.data, "parameter")
            a.grad[i] = _ptr(g, "parameter gradient")
            a.state1[i] = _ptr(state1[i], "optimizer state") if state1[i] is not None else None
            a.state2[i] = _ptr(state2[i], "optimizer state") if state2[i] is not None else None
            a.numel[i] = q.numel()
        a.inv_norm = float(inv_norm)
        a.lr, a.weight_decay, a.momentum, a.dampening = float(lr), float(weight_decay), float(momentum), float(dampening)
        a.beta1, a.beta2, a.eps = float(beta1), float(beta2), float(eps)
        a.nesterov, a.first_step, a.step = int(bool(nesterov)), int(first_step), int(step)
        stream = _raw_stream(dev)
        with _OnDevice(dev):
            N.check(self._lib.mcpc_p_step(C.byref(a), C.c_void_p(stream)), "mcpc_p_step")

    def fill_noise(self, seed: int, t_begin: int, n_steps: int, chain_offset: int, B: int, n_units: int,
                   noise_scale: float, device) -> torch.Tensor:
        out = torch.empty(n_steps, B, n_units, dtype=torch.float32, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        with torch.cuda.device(device):
            N.check(self._lib.mcpc_fill_noise(seed & 0xFFFFFFFFFFFFFFFF, t_begin, n_steps, chain_offset, B, n_units,
                                              float(noise_scale), out.data_ptr(), C.c_void_p(stream)),
                    "mcpc_fill_noise")
        return out
