"""The streaming bf16 path (per-step grouped tcgen05 GEMM kernels, csrc/infer_wide.cu) against the oracle run
with bf16-rounded contraction operands, and against the plain fp32 oracle under the stated bf16 bound.

Small layer widths are forced onto this path with MCPC_FORCE_STREAMING so the oracle finishes in seconds; the C5
shape (4 PCLayers, widths 1024 and 4096: long K rings, 3-D tensor maps, many tiles per CTA) runs WITHOUT the hook --
those nets do not fit one SM and take the streaming path on their own.  Every case runs on single CTAs
(MCPC_WIDE_CG=1) and on CTA pairs (cta_group::2, the default), and with 1 / several ring slots of the weight-gradient
operands (MCPC_WIDE_SLOTS)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.optim as optim

from golden_util import orc, rel_err

from montecarlopredictivecoding_b200 import _native as N
from montecarlopredictivecoding_b200 import mcpc_utils as mu
from montecarlopredictivecoding_b200 import predictive_coding as pc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KNOBS = ("MCPC_FORCE_STREAMING", "MCPC_WIDE_CTAS", "MCPC_WIDE_CG", "MCPC_WIDE_SLOTS", "MCPC_TC_NOSPEC")


@pytest.fixture(autouse=True)
def _clean_knobs():
    saved = {k: os.environ.get(k) for k in KNOBS}
    yield
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _model(dims, d_out, act, dev, seed=0, wide_init=False, d_in=None):
    torch.manual_seed(seed)
    A = {"relu": nn.ReLU, "tanh": nn.Tanh}[act]
    mods, prev = [], (dims[0] if d_in is None else d_in)
    for d in dims:
        mods += [nn.Linear(prev, d), pc.PCLayer(), A()]
        prev = d
    mods.append(nn.Linear(prev, d_out))
    m = nn.Sequential(*mods)
    m.train()
    if wide_init:                      # W ~ N(0, 1/fan_in), b = 0.1 N(0,1) (non-zero so the bias path is live)
        with torch.no_grad():
            for mod in m:
                if isinstance(mod, nn.Linear):
                    mod.weight.normal_(0, (1.0 / mod.in_features) ** 0.5)
                    mod.bias.normal_(0, 0.1)
    return m.to(dev)


def _run_case(dims, d_out, act, top, opt, B, mixing, sampling, lr=0.02, wide_init=False, bf16_oracle=True,
              want_outputs=True, seed=2024, d_in=None):
    """One learning call on the GPU + the same call through the oracle with the kernel's own noise; returns the
    dict of max-norm relative errors."""
    dev = torch.device(DEV)
    L = len(dims)
    T = mixing + sampling
    model = _model(dims, d_out, act, dev, wide_init=wide_init, d_in=d_in)
    opt_fn = optim.SGD if opt == "sgd" else optim.Adam
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=opt_fn, optimizer_x_kwargs={"lr": lr}, update_p_at="last",
                      accumulate_p_at=list(range(mixing, T)), optimizer_p_fn=optim.SGD, optimizer_p_kwargs={"lr": 0.0},
                      plot_progress_at=[])
    tr.set_precision("bf16")
    tr.set_noise_seed(seed)
    torch.manual_seed(3)
    y = (torch.rand(B, d_out, device=dev) < 0.5).float() if top == "bernoulli" else torch.randn(B, d_out, device=dev)
    x0 = [torch.randn(B, d, device=dev) for d in dims]
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    lins = [m for m in model if isinstance(m, nn.Linear)]
    for layer, v in zip(pcs, x0):
        layer._sample_x_fn = (lambda inputs, v=v: v.clone())
    loss_fn = mu.bernoulli_fn if top == "bernoulli" else mu.fe_fn
    kw = dict(callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr}) if opt == "sgd" else {}
    # d_in given: NON-ZERO inputs of that width (Linear_0 becomes a GEMM over a bf16 copy of them, and gets a weight gradient)
    inputs = torch.zeros(B, dims[0], device=dev) if d_in is None else torch.randn(B, d_in, device=dev)
    res = tr.train_on_batch(inputs, loss_fn=loss_fn, loss_fn_kwargs={"_target": y, "_var": 1.0},
                            is_log_progress=False, is_return_outputs=want_outputs, **kw)
    assert tr.last_call_info["mode"] == "fused"
    assert tr._get_engine().infer_mode(tr_plan(tr), tr_top(loss_fn, y, B, d_out), B, N.PREC_BF16) == N.MODE_STREAMING_BF16
    SD = sum(dims)
    noise = None
    if opt == "sgd":
        nz = tr._get_engine().fill_noise(seed, 0, T, 0, B, SD, float(np.sqrt(2.0 / lr)), dev).cpu().numpy()
        offs = np.cumsum([0] + list(dims))
        noise = [[nz[t][:, offs[l]:offs[l + 1]] for l in range(L)] for t in range(T)]
    net = orc.OracleNet(W=[l.weight.detach().cpu().numpy() for l in lins], b=[l.bias.detach().cpu().numpy() for l in lins],
                        n_layers=L, act=[orc.ACT_RELU if act == "relu" else orc.ACT_TANH] * L, energy_scale=[1.0] * L,
                        top=orc.TOP_BERNOULLI if top == "bernoulli" else orc.TOP_GAUSS, bf16_operands=bf16_oracle,
                        bf16_own_term=bf16_oracle and os.environ.get("MCPC_WIDE_G32", "0") == "0",
                        bf16_inputs=bf16_oracle)
    ref = orc.infer(net, [v.cpu().numpy() for v in x0], inputs.cpu().numpy(), y.cpu().numpy(), T,
                    optimizer=opt, lr=lr, noise=noise, acc_begin=mixing, acc_end=T, record_traj=want_outputs)
    errs = {f"x{l}": rel_err(pcs[l].get_x().detach().cpu().numpy(), ref.xs[l]) for l in range(L)}
    errs["energy"] = rel_err(res["energy"], ref.energy)
    errs["loss"] = rel_err(res["loss"], ref.loss)
    if want_outputs:
        errs["out"] = rel_err(torch.stack(res["outputs"]).cpu().numpy(), np.stack(ref.traj_out))
    div = sampling * B
    for i in range(1, L + 1):
        errs[f"gW_{i}"] = rel_err(lins[i].weight.grad.cpu().numpy(), ref.gW[i] / div)
    for i in range(0, L + 1):
        errs[f"gb_{i}"] = rel_err(lins[i].bias.grad.cpu().numpy(), ref.gb[i] / div)
    if d_in is None:
        assert float(lins[0].weight.grad.abs().max()) == 0.0
    else:
        errs["gW_0"] = rel_err(lins[0].weight.grad.cpu().numpy(), ref.gW[0] / div)
    return errs


def _check(errs, tol, label):
    print(label, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v < tol, (label, k, v)


@pytest.mark.parametrize("cg", [pytest.param(1, id="cg1"), pytest.param(2, id="cg2")])
@pytest.mark.parametrize("act,top,opt,B,ctas,slots", [
    ("tanh", "gauss", "sgd", 200, 0, 0), ("relu", "bernoulli", "sgd", 256, 0, 1), ("tanh", "gauss", "adam", 130, 0, 3),
    ("relu", "bernoulli", "sgd", 392, 2, 3), ("tanh", "gauss", "sgd", 300, 64, 2), ("tanh", "gauss", "sgd", 520, 4, 0)])
def test_streaming_path_vs_bf16_oracle(act, top, opt, B, ctas, slots, cg, monkeypatch):
    # ctas > 0: only that many persistent CTAs, so each walks several tiles (ring wrap-around, accumulator
    # double buffering across tiles); slots: ring depth of the weight-gradient operands (0 = default)
    monkeypatch.setenv("MCPC_FORCE_STREAMING", "1")
    monkeypatch.setenv("MCPC_WIDE_CG", str(cg))
    if ctas:
        monkeypatch.setenv("MCPC_WIDE_CTAS", str(ctas))
    if slots:
        monkeypatch.setenv("MCPC_WIDE_SLOTS", str(slots))
    dims, d_out = [128, 256, 144], 272
    if ctas == 64:          # all widths multiples of 64: MN-major operands come in through the 3-D tensor maps
        dims, d_out = [128, 320, 192], 576
    errs = _run_case(dims, d_out, act, top, opt, B, mixing=2, sampling=4)
    _check(errs, 5e-3 if opt == "adam" else 2e-3, f"{act}/{top}/{opt}/B={B}/cg={cg}")


@pytest.mark.parametrize("cg", [pytest.param(1, id="cg1"), pytest.param(2, id="cg2")])
def test_streaming_path_odd_widths(cg, monkeypatch):
    """Widths that are not multiples of 16 (or of 8): partial unit tiles, padded bf16 pitches, 2-D MN-major boxes."""
    monkeypatch.setenv("MCPC_FORCE_STREAMING", "1")
    monkeypatch.setenv("MCPC_WIDE_CG", str(cg))
    errs = _run_case([20, 130, 77], 101, "tanh", "gauss", "sgd", B=150, mixing=1, sampling=3)
    _check(errs, 2e-3, f"odd widths cg={cg}")


@pytest.mark.parametrize("cg", [pytest.param(1, id="cg1"), pytest.param(2, id="cg2")])
@pytest.mark.parametrize("dims,d_out,d_in,B,slots", [([128, 256, 144], 272, 50, 200, 0), ([128, 320, 192], 576, 64, 300, 2),
                                                     ([20, 130, 77], 101, 33, 150, 3), ([256, 256], 128, 300, 392, 1)])
def test_streaming_path_nonzero_inputs(dims, d_out, d_in, B, slots, cg, monkeypatch):
    """Non-zero ``inputs``: Linear_0 is a GEMM over the bf16 inputs block (PREDICT) and gets a weight gradient (WGRAD);
    widths with and without the 3-D tensor maps (all % 64 == 0), d_in narrower and wider than a unit tile."""
    monkeypatch.setenv("MCPC_FORCE_STREAMING", "1")
    monkeypatch.setenv("MCPC_WIDE_CG", str(cg))
    if slots:
        monkeypatch.setenv("MCPC_WIDE_SLOTS", str(slots))
    errs = _run_case(dims, d_out, "tanh", "gauss", "sgd", B, mixing=2, sampling=4, d_in=d_in)
    _check(errs, 2e-3, f"inputs d_in={d_in} dims={dims} cg={cg}")


@pytest.mark.parametrize("cg", [pytest.param(1, id="cg1"), pytest.param(2, id="cg2")])
@pytest.mark.parametrize("width,B,mixing,sampling", [(1024, 512, 2, 5), (4096, 256, 1, 3)])
def test_c5_shape_vs_bf16_oracle(width, B, mixing, sampling, cg, monkeypatch):
    """SURVEY §8d C5 shape: 4 PCLayers of `width` units + an output Linear of the same width, tanh, Gaussian top, SGD +
    Langevin noise, weight update accumulated over the sampling steps.  No MCPC_FORCE_STREAMING: these nets take the
    streaming path by themselves; 4096 is the benchmarked width (64 K-stages per tile, 3-D tensor maps, 5 ring
    slots > sampling steps)."""
    monkeypatch.setenv("MCPC_WIDE_CG", str(cg))
    errs = _run_case([width] * 4, width, "tanh", "gauss", "sgd", B, mixing, sampling, lr=0.01, wide_init=True,
                     want_outputs=False)
    _check(errs, 2e-3, f"C5 shape width={width} B={B} cg={cg}")


def test_c5_shape_bound_vs_fp32_oracle():
    """The stated bf16 bound of the streaming path against the plain fp32 oracle (no operand rounding) over the
    benchmark's T=100 Langevin steps (C5: lr 0.01, var 2, dW over all steps) at 4 x 1024: latents 5e-3, per-step
    energy / loss 1e-4, weight gradients 5e-3 (max-norm relative; measured r02: 1.1e-3 / 1.1e-5 / 1.3e-3)."""
    errs = _run_case([1024] * 4, 1024, "tanh", "gauss", "sgd", B=256, mixing=0, sampling=100, lr=0.01, wide_init=True,
                     bf16_oracle=False, want_outputs=False)
    print("C5 bf16-vs-fp32 bound, T=100:", {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        tol = 1e-4 if k in ("energy", "loss") else 5e-3
        assert v < tol, (k, v)


def tr_plan(tr):
    from montecarlopredictivecoding_b200.predictive_coding import plan as P
    return P.compile_net(tr.get_model())


def tr_top(loss_fn, y, B, d_out):
    from montecarlopredictivecoding_b200.predictive_coding import plan as P
    return P.classify_loss(loss_fn, {"_target": y, "_var": 1.0}, B, d_out, None)


@pytest.mark.parametrize("B", [256, 1000])
def test_specialised_update_kernel_equals_the_generic_one(B, monkeypatch):
    """wide_kernel<UPDATE, 1> (SGD + in-kernel Philox, no trajectories) only folds runtime flags into constants: the call
    must give the same latents, energies and weight gradients as the generic instantiation (MCPC_TC_NOSPEC)."""
    dev = torch.device(DEV)
    monkeypatch.setenv("MCPC_FORCE_STREAMING", "1")

    def run(nospec):
        if nospec:
            os.environ["MCPC_TC_NOSPEC"] = "1"
        else:
            os.environ.pop("MCPC_TC_NOSPEC", None)
        try:
            model = _model([128, 256, 192], 320, "tanh", dev, seed=3)
            T = 8
            tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.02}, update_p_at="last",
                              accumulate_p_at=list(range(2, T)), optimizer_p_fn=optim.SGD, optimizer_p_kwargs={"lr": 0.0},
                              plot_progress_at=[])
            tr.set_precision("bf16")
            tr.set_noise_seed(11)
            for layer in (m for m in model if isinstance(m, pc.PCLayer)):
                layer._sample_x_fn = mu.sample_x_fn_normal
            torch.manual_seed(9)
            y = torch.randn(B, 320, device=dev)
            res = tr.train_on_batch(torch.zeros(B, 128, device=dev), loss_fn=mu.fe_fn, loss_fn_kwargs={"_target": y, "_var": 1.0},
                                    callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr},
                                    is_log_progress=False, is_checking_after_callback_after_t=False)
            assert tr.last_call_info["mode"] == "fused"
            xs = [layer.get_x().detach().clone() for layer in model if isinstance(layer, pc.PCLayer)]
            grads = [p.grad.detach().clone() for p in model.parameters() if p.grad is not None]
            return xs, torch.tensor(res["energy"]), grads
        finally:
            os.environ.pop("MCPC_TC_NOSPEC", None)
    a, b = run(False), run(True)
    for xa, xb in zip(a[0], b[0]):
        assert torch.allclose(xa, xb, rtol=1e-5, atol=1e-5)
    assert torch.allclose(a[1], b[1], rtol=1e-5)
    for ga, gb in zip(a[2], b[2]):
        assert torch.allclose(ga, gb, rtol=1e-4, atol=1e-5)
