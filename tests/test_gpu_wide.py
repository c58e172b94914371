"""The streaming bf16 path (per-step grouped tcgen05 GEMM kernels, csrc/infer_wide.cu) against the oracle run
with bf16-rounded contraction operands.  Small layer widths are forced onto this path with
MCPC_FORCE_STREAMING so the oracle finishes in seconds; the full-size C5 shape is exercised by bench/scripts."""
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


@pytest.fixture(autouse=True)
def _force_streaming():
    os.environ["MCPC_FORCE_STREAMING"] = "1"
    yield
    os.environ.pop("MCPC_FORCE_STREAMING", None)


def _model(dims, d_out, act, dev, seed=0):
    torch.manual_seed(seed)
    A = {"relu": nn.ReLU, "tanh": nn.Tanh}[act]
    mods, prev = [], dims[0]
    for d in dims:
        mods += [nn.Linear(prev, d), pc.PCLayer(), A()]
        prev = d
    mods.append(nn.Linear(prev, d_out))
    m = nn.Sequential(*mods)
    m.train()
    return m.to(dev)


@pytest.mark.parametrize("act,top,opt,B,ctas", [("tanh", "gauss", "sgd", 200, 0), ("relu", "bernoulli", "sgd", 256, 0),
                                                ("tanh", "gauss", "adam", 130, 0), ("relu", "bernoulli", "sgd", 392, 2),
                                                ("tanh", "gauss", "sgd", 300, 64)])
def test_streaming_path_vs_bf16_oracle(act, top, opt, B, ctas, monkeypatch):
    # ctas > 0: only that many persistent CTAs, so each walks several tiles (ring wrap-around, accumulator
    # double buffering across tiles)
    if ctas:
        monkeypatch.setenv("MCPC_WIDE_CTAS", str(ctas))
    dev = torch.device(DEV)
    dims, d_out = [128, 256, 144], 272
    if ctas == 64:          # all widths multiples of 64: MN-major operands come in through the 3-D tensor maps
        dims, d_out = [128, 320, 192], 576
    mixing, sampling, lr = 2, 4, 0.02
    T = mixing + sampling
    model = _model(dims, d_out, act, dev)
    opt_fn = optim.SGD if opt == "sgd" else optim.Adam
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=opt_fn, optimizer_x_kwargs={"lr": lr}, update_p_at="last",
                      accumulate_p_at=list(range(mixing, T)), optimizer_p_fn=optim.SGD, optimizer_p_kwargs={"lr": 0.0},
                      plot_progress_at=[])
    tr.set_precision("bf16")
    tr.set_noise_seed(2024)
    torch.manual_seed(3)
    y = (torch.rand(B, d_out, device=dev) < 0.5).float() if top == "bernoulli" else torch.randn(B, d_out, device=dev)
    x0 = [torch.randn(B, d, device=dev) for d in dims]
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    lins = [m for m in model if isinstance(m, nn.Linear)]
    for layer, v in zip(pcs, x0):
        layer._sample_x_fn = (lambda inputs, v=v: v.clone())
    loss_fn = mu.bernoulli_fn if top == "bernoulli" else mu.fe_fn
    kw = dict(callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr}) if opt == "sgd" else {}
    res = tr.train_on_batch(torch.zeros(B, dims[0], device=dev), loss_fn=loss_fn, loss_fn_kwargs={"_target": y, "_var": 1.0},
                            is_log_progress=False, is_return_outputs=True, **kw)
    assert tr._get_engine().infer_mode(tr_plan(tr), tr_top(tr, loss_fn, y, B, d_out), B, N.PREC_BF16) == N.MODE_STREAMING_BF16
    SD = sum(dims)
    noise = None
    if opt == "sgd":
        nz = tr._get_engine().fill_noise(2024, 0, T, 0, B, SD, float(np.sqrt(2.0 / lr)), dev).cpu().numpy()
        offs = np.cumsum([0] + dims)
        noise = [[nz[t][:, offs[l]:offs[l + 1]] for l in range(3)] for t in range(T)]
    net = orc.OracleNet(W=[l.weight.detach().cpu().numpy() for l in lins], b=[l.bias.detach().cpu().numpy() for l in lins],
                        n_layers=3, act=[orc.ACT_RELU if act == "relu" else orc.ACT_TANH] * 3, energy_scale=[1.0] * 3,
                        top=orc.TOP_BERNOULLI if top == "bernoulli" else orc.TOP_GAUSS, bf16_operands=True)
    ref = orc.infer(net, [v.cpu().numpy() for v in x0], np.zeros((B, dims[0]), np.float32), y.cpu().numpy(), T,
                    optimizer=opt, lr=lr, noise=noise, acc_begin=mixing, acc_end=T, record_traj=True)
    errs = {f"x{l}": rel_err(pcs[l].get_x().detach().cpu().numpy(), ref.xs[l]) for l in range(3)}
    errs["energy"] = rel_err(res["energy"], ref.energy)
    errs["loss"] = rel_err(res["loss"], ref.loss)
    errs["out"] = rel_err(torch.stack(res["outputs"]).cpu().numpy(), np.stack(ref.traj_out))
    div = sampling * B
    for i in (1, 2, 3):
        errs[f"gW_{i}"] = rel_err(lins[i].weight.grad.cpu().numpy(), ref.gW[i] / div)
    for i in (0, 1, 3):
        errs[f"gb_{i}"] = rel_err(lins[i].bias.grad.cpu().numpy(), ref.gb[i] / div)
    print(act, top, opt, B, {k: f"{v:.2e}" for k, v in errs.items()})
    tol = 5e-3 if opt == "adam" else 2e-3
    for k, v in errs.items():
        assert v < tol, (k, v)
    assert float(lins[0].weight.grad.abs().max()) == 0.0


def tr_plan(tr):
    from montecarlopredictivecoding_b200.predictive_coding import plan as P
    return P.compile_net(tr.get_model())


def tr_top(tr, loss_fn, y, B, d_out):
    from montecarlopredictivecoding_b200.predictive_coding import plan as P
    return P.classify_loss(loss_fn, {"_target": y, "_var": 1.0}, B, d_out, None)


@pytest.mark.parametrize("B", [256, 1000])
def test_specialised_update_kernel_equals_the_generic_one(B):
    """wide_kernel<UPDATE, 1> (SGD + in-kernel Philox, no trajectories) only folds runtime flags into constants: the call
    must give the same latents, energies and weight gradients as the generic instantiation (MCPC_TC_NOSPEC)."""
    dev = torch.device(DEV)

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
