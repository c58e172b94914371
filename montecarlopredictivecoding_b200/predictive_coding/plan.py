"""Plan compiler: turn a user ``nn.Module`` + ``loss_fn`` + callback into the flat description
the fused kernels take (``McpcNet``), or say precisely why that is impossible.

The reference executes arbitrary Python per step (autograd over any module graph).  The B200
path recognises the one graph family every script of the reference builds --
``[Linear, PCLayer, act?]* [Linear]?`` (utils/model.py:54-65, figure_2.py:40-44,
figure_3.py:50-55, figure_4.py:104-108, figure_6.py:45-49,85) -- and classifies the Python
callables numerically instead of by name, so user-written equivalents work too:

  * ``energy_fn``   must equal c * 0.5 * (mu - x)^2 elementwise   (pc_layer.py:17-18, figure_3.py:47-48)
  * ``loss_fn``     must be one of: constant (zero_fn), Gaussian (1/var)*0.5*sum(o-y)^2,
                    Bernoulli sum BCEWithLogits(o, y); each optionally restricted to the last
                    columns (the *_mask variants)                  (utils/model.py:17-33)
  * ``callback_after_t`` is folded into the kernel when it is ``random_step`` (utils/model.py:35-44)
"""
import inspect
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.nn as nn

from .. import _native as N
from .layer import PCLayer

_ACT_OF = {nn.ReLU: N.ACT_RELU, nn.Tanh: N.ACT_TANH, nn.Identity: N.ACT_IDENTITY}


class UnsupportedModel(NotImplementedError):
    """The model / callable is outside what the fused sm_100a kernels implement."""


@dataclass
class NetPlan:
    linears: List[nn.Linear]            # L (+1 when an output Linear exists)
    pc_layers: List[PCLayer]
    act: List[int]
    energy_scale: List[float]
    d_in: int
    dims: List[int]
    d_out: int
    signature: tuple = field(default=())

    @property
    def L(self):
        return len(self.pc_layers)

    @property
    def SD(self):
        return sum(self.dims)


@dataclass
class TopPlan:
    kind: int                 # N.TOP_*
    inv_var: float = 1.0
    mask_start: int = 0
    target_key: Optional[str] = None


_energy_cache = {}


def classify_energy_fn(fn) -> float:
    """Return c such that fn({'mu','x'}) == c*0.5*(mu-x)^2, or raise UnsupportedModel."""
    key = id(fn)
    hit = _energy_cache.get(key)
    if hit is not None and hit[0] is fn:
        return hit[1]
    mu = torch.tensor([[0.0, 1.0, -2.0, 0.5, 3.0]], dtype=torch.float64)
    x = torch.tensor([[0.25, -1.0, 1.5, 0.5, -0.75]], dtype=torch.float64)
    try:
        e = fn({"mu": mu, "x": x})
    except Exception as exc:  # noqa: BLE001
        raise UnsupportedModel(f"energy_fn could not be probed ({exc!r}); the fused kernel needs c*0.5*(mu-x)^2") from exc
    base = 0.5 * (mu - x) ** 2
    if not isinstance(e, torch.Tensor) or e.shape != base.shape:
        raise UnsupportedModel("energy_fn must be elementwise c*0.5*(mu-x)^2 for the fused kernel")
    nz = base > 0
    ratio = (e[nz] / base[nz])
    c = float(ratio[0])
    if not torch.allclose(ratio, torch.full_like(ratio, c), rtol=1e-6, atol=0) or \
            not torch.allclose(e[~nz], torch.zeros_like(e[~nz]), atol=1e-12) or not (c > 0):
        raise UnsupportedModel("energy_fn is not a positive multiple of 0.5*(mu-x)^2; only (scaled) quadratic "
                               "energies are implemented in the fused kernel")
    _energy_cache[key] = (fn, c)
    return c


import collections

_plan_cache = collections.OrderedDict()    # small LRU: a plan keeps its model (and the latents) alive
_PLAN_CACHE_SIZE = 8


def compile_net(model: nn.Module) -> NetPlan:
    """Walk the module list and recognise ``[Linear, PCLayer, act?]* [Linear]?``.

    The result is cached per module list (identity of the children); what a user can flip on a live PCLayer
    (masks, per-datapoint energies, held errors, the energy function) and the Linear shapes are re-checked on a hit."""
    if isinstance(model, (nn.Linear, PCLayer)):
        mods = [model]
    else:
        mods = list(model._modules.values())              # == list(model.children()) unless a child is None / shared
        if None in mods or len(set(map(id, mods))) != len(mods):
            mods = list(model.children())
    key = tuple(map(id, mods))
    hit = _plan_cache.get(key)
    if hit is not None:
        plan, efns, shapes = hit
        ok = all(m._S is None and m._M is None and not m.is_keep_energy_per_datapoint and not m.is_holding_error
                 for m in plan.pc_layers)
        ok = ok and all(pcl._energy_fn is f for pcl, f in zip(plan.pc_layers, efns))
        ok = ok and all((lin.in_features, lin.out_features, lin.bias is None) == sh for lin, sh in zip(plan.linears, shapes))
        if ok:
            _plan_cache.move_to_end(key)
            return plan
        del _plan_cache[key]
    plan = _compile_net(model, mods)
    while len(_plan_cache) >= _PLAN_CACHE_SIZE:
        _plan_cache.popitem(last=False)
    _plan_cache[key] = (plan, [pcl._energy_fn for pcl in plan.pc_layers],
                        [(lin.in_features, lin.out_features, lin.bias is None) for lin in plan.linears])
    return plan


def _compile_net(model: nn.Module, mods) -> NetPlan:
    if any(len(list(m.children())) > 0 for m in mods):
        raise UnsupportedModel("nested containers are not supported by the fused kernel; use a flat nn.Sequential of "
                               "Linear / PCLayer / activation modules")
    linears, pcs, acts = [], [], []
    i, n = 0, len(mods)
    pending_linear = None
    while i < n:
        m = mods[i]
        if isinstance(m, nn.Linear):
            if pending_linear is not None:
                raise UnsupportedModel("two Linear layers without a PCLayer between them")
            pending_linear = m
            i += 1
        elif isinstance(m, PCLayer):
            if pending_linear is None:
                raise UnsupportedModel("every PCLayer must directly follow its own nn.Linear")
            if m._S is not None or m._M is not None or m.is_keep_energy_per_datapoint or m.is_holding_error:
                raise UnsupportedModel("PCLayer S/M masks, per-datapoint energies and held errors are not implemented "
                                       "in the fused kernel")
            linears.append(pending_linear)
            pending_linear = None
            pcs.append(m)
            act = N.ACT_IDENTITY
            if i + 1 < n and type(mods[i + 1]) in _ACT_OF:
                act = _ACT_OF[type(mods[i + 1])]
                i += 1
            acts.append(act)
            i += 1
        else:
            raise UnsupportedModel(f"module {type(m).__name__} at position {i} is not part of the supported pattern "
                                   "[Linear, PCLayer, ReLU|Tanh|Identity]* [Linear]")
    if not pcs:
        raise UnsupportedModel("the model has no PCLayer")
    if len(pcs) > N.MAX_LAYERS:
        raise UnsupportedModel(f"more than {N.MAX_LAYERS} PCLayers")
    d_out = 0
    if pending_linear is not None:
        linears.append(pending_linear)
        d_out = pending_linear.out_features
    dims = [lin.out_features for lin in linears[:len(pcs)]]
    prev = linears[0].in_features
    for lin in linears:
        if lin.in_features != prev:
            raise UnsupportedModel("Linear in/out features do not chain")
        prev = lin.out_features
    scales = [classify_energy_fn(p._energy_fn) for p in pcs]
    sig = tuple(id(m) for m in mods)
    return NetPlan(linears=linears, pc_layers=pcs, act=acts, energy_scale=scales, d_in=linears[0].in_features,
                   dims=dims, d_out=d_out, signature=sig)


_loss_cache = {}


def classify_loss(loss_fn, loss_fn_kwargs: dict, B: int, d_out: int, device) -> TopPlan:
    """Probe ``loss_fn`` with a 2-row batch and match value + gradient against the closed forms."""
    if loss_fn is None:
        return TopPlan(kind=N.TOP_NONE)
    if d_out == 0:
        raise UnsupportedModel("loss_fn on a free output PCLayer is not supported")
    tkeys = [k for k, v in loss_fn_kwargs.items() if isinstance(v, torch.Tensor) and v.dim() == 2
             and tuple(v.shape) == (B, d_out)]
    scalars = tuple(sorted((k, repr(v)) for k, v in loss_fn_kwargs.items() if not isinstance(v, torch.Tensor)))
    key = (id(loss_fn), d_out, scalars, tuple(tkeys))
    hit = _loss_cache.get(key)
    if hit is not None and hit[0] is loss_fn:
        return hit[1]
    if len(tkeys) > 1:
        raise UnsupportedModel("loss_fn takes more than one [B, d_out] tensor; cannot tell which is the target")
    other_tensors = [k for k, v in loss_fn_kwargs.items() if isinstance(v, torch.Tensor) and k not in tkeys]
    if other_tensors:
        raise UnsupportedModel(f"loss_fn tensor kwargs {other_tensors} are not understood by the fused kernel")
    g = torch.Generator(device="cpu").manual_seed(1234)
    o = (torch.randn(2, d_out, generator=g, dtype=torch.float64) * 1.5).requires_grad_(True)
    kw = dict(loss_fn_kwargs)
    y = None
    if tkeys:
        y = (torch.rand(2, d_out, generator=g, dtype=torch.float64) < 0.5).double() * 0.75 + 0.125
        kw[tkeys[0]] = y
    try:
        val = loss_fn(o, **kw)
    except Exception as exc:  # noqa: BLE001
        raise UnsupportedModel(f"loss_fn could not be probed on a 2-row batch ({exc!r})") from exc
    plan = None
    if not isinstance(val, torch.Tensor) or not val.requires_grad:
        if float(val) == 0.0:
            plan = TopPlan(kind=N.TOP_ZERO)
        else:
            raise UnsupportedModel("loss_fn does not depend on the outputs but is not zero")
    else:
        if val.dim() != 0:
            raise UnsupportedModel("loss_fn must return a scalar")
        (grad,) = torch.autograd.grad(val, o)
        col_used = (grad != 0).any(dim=0)
        if not bool(col_used.any()):
            raise UnsupportedModel("loss_fn has zero gradient everywhere")
        ms = int(torch.nonzero(col_used)[0])
        if not bool(col_used[ms:].all()) or y is None:
            raise UnsupportedModel("loss_fn must act on a contiguous block of trailing output columns of a target")
        om, ym, gm = o.detach()[:, ms:], y[:, ms:], grad[:, ms:]
        # Bernoulli: grad = sigmoid(o) - y, value = sum BCE-with-logits
        bern_g = torch.sigmoid(om) - ym
        bern_v = (torch.clamp(om, min=0) - om * ym + torch.log1p(torch.exp(-om.abs()))).sum()
        ratio = gm / (om - ym)
        iv = float(ratio.flatten()[0])
        gauss_v = iv * 0.5 * ((om - ym) ** 2).sum()
        if torch.allclose(gm, bern_g, rtol=1e-9, atol=1e-12) and torch.allclose(val.detach(), bern_v, rtol=1e-9):
            plan = TopPlan(kind=N.TOP_BERNOULLI, mask_start=ms, target_key=tkeys[0])
        elif iv > 0 and torch.allclose(ratio, torch.full_like(ratio, iv), rtol=1e-9) and \
                torch.allclose(val.detach(), gauss_v, rtol=1e-9):
            plan = TopPlan(kind=N.TOP_GAUSS, inv_var=iv, mask_start=ms, target_key=tkeys[0])
        else:
            raise UnsupportedModel("loss_fn is neither Gaussian (1/var)*0.5*sum(o-y)^2 nor Bernoulli "
                                   "sum BCEWithLogits(o,y) (optionally on trailing columns); only those are fused")
    _loss_cache[key] = (loss_fn, plan)
    return plan


@dataclass
class LangevinPlan:
    var: float


_probe_cache = {}


def _cached_probe(fn, kind, probe):
    key = (kind, id(fn), id(getattr(fn, "__code__", None)))
    hit = _probe_cache.get(key)
    if hit is not None and hit[0] is fn:
        return hit[1]
    try:
        ok = bool(probe(fn))
    except Exception:  # noqa: BLE001 -- anything unexpected means "not the function we know"
        ok = False
    if len(_probe_cache) > 256:
        _probe_cache.clear()
    _probe_cache[key] = (fn, ok)
    return ok


def _behaves_like_random_step(cb) -> bool:
    """Behavioural probe of utils/model.py:35-44: on a stand-in trainer the callback must (a) overwrite every
    ``x.grad`` with torch's ``normal_(0, sqrt(var / optimizer.defaults['lr']))`` draws, layer by layer in order, (b) call
    ``optimizer.step()`` exactly once, (c) touch nothing else -- for two (var, lr) pairs.  The generator state of the
    caller is preserved."""
    import numpy as np

    class _Opt:
        def __init__(self, lr):
            self.defaults = {"lr": lr}
            self.steps = 0
            self.zeroed = 0

        def step(self):
            self.steps += 1

        def zero_grad(self, *a, **k):
            self.zeroed += 1

    class _Trainer:
        def __init__(self, xs, opt):
            self._xs, self._opt = xs, opt

        def get_model_xs(self):
            return iter(self._xs)

        def get_optimizer_x(self):
            return self._opt

    def probe(fn):
        state = torch.random.get_rng_state()
        try:
            for var, lr in ((2.0, 0.03), (0.7, 0.25)):
                xs = []
                for shape in ((3, 5), (3, 2)):
                    x = torch.full(shape, 1.25)
                    x.grad = torch.full(shape, -7.0)
                    xs.append(x)
                opt = _Opt(lr)
                torch.manual_seed(987654321)
                fn(0, _pc_trainer=_Trainer(xs, opt), var=var)
                torch.manual_seed(987654321)
                std = float(np.sqrt(var / lr))
                for x in xs:
                    want = torch.empty_like(x).normal_(0.0, std)
                    if not torch.equal(x.grad, want) or not torch.equal(x, torch.full_like(x, 1.25)):
                        return False
                if opt.steps != 1 or opt.zeroed != 0:
                    return False
            return True
        finally:
            torch.random.set_rng_state(state)
    return _cached_probe(cb, "langevin", probe)


def sampler_is_shape_only(fn) -> bool:
    """True when a ``sample_x_fn`` provably ignores the VALUES of ``mu`` / ``x``: tagged (``__mcpc_shape_only__``) or, for
    the functions named like the reference's utils/model.py:8-15 samplers, verified by a probe -- the same seed must give
    the same draw for two different ``mu`` contents (one of them NaN).  Such samplers can be called without a model
    forward; anything else gets the real t=0 forward of pc_trainer.py:717-733."""
    if getattr(fn, "__mcpc_shape_only__", False):
        return True
    named = getattr(fn, "__name__", "") in ("sample_x_fn", "sample_x_fn_normal", "sample_x_fn_cte") and \
        (getattr(fn, "__module__", "") or "").split(".")[-1] == "model"
    if not named:
        return False

    def probe(f):
        state = torch.random.get_rng_state()
        try:
            outs = []
            for fill in (0.5, float("nan")):
                torch.manual_seed(13579)
                mu = torch.full((4, 3), fill)
                out = f({"mu": mu, "x": None})
                if not torch.is_tensor(out) or out.shape != mu.shape or not torch.isfinite(out).all():
                    return False
                outs.append(out)
            return torch.equal(outs[0], outs[1])
        finally:
            torch.random.set_rng_state(state)
    return _cached_probe(fn, "sampler", probe)


_cb_default_var = {}


def classify_callback_after_t(cb, kwargs: dict, trainer) -> Optional[LangevinPlan]:
    """Recognise the Langevin ``random_step`` callback (utils/model.py:35-44).

    Returns the noise variance when the callback can be folded into the kernel, ``None`` when
    it must be executed as opaque Python (step-by-step mode).
    """
    if cb is None:
        return None
    tagged = getattr(cb, "__mcpc_langevin__", False)
    named = getattr(cb, "__name__", "") == "random_step" and \
        (getattr(cb, "__module__", "") or "").split(".")[-1] == "model"
    if not (tagged or named):
        return None
    # a name is not a contract: an untagged function is folded into the kernel only if it BEHAVES like the reference's
    # random_step (a forked utils/model.py with another noise law keeps running as opaque Python, step by step)
    if not tagged and not _behaves_like_random_step(cb):
        return None
    if kwargs.get("_pc_trainer", None) is not trainer:
        return None
    if set(kwargs) - {"_pc_trainer", "var"}:
        return None
    var = kwargs.get("var", None)
    if var is None:
        hit = _cb_default_var.get(id(cb))
        if hit is not None and hit[0] is cb:
            var = hit[1]
        else:
            try:
                var = inspect.signature(cb).parameters["var"].default
            except (KeyError, TypeError, ValueError):
                return None
            _cb_default_var[id(cb)] = (cb, var)
    try:
        var = float(var)
    except (TypeError, ValueError):
        return None
    if not (var >= 0.0):
        return None
    return LangevinPlan(var=var)
