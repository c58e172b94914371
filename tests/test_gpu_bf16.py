"""MCPC_PREC_BF16 (tcgen05 tensor-core path) against (i) the golden vectors of the fp32 reference under the
stated bf16 bound and (ii) the oracle run with bf16-rounded contraction operands (tight).

Stated bound (north star: "a stated bf16/tf32 bound over T steps"): bf16 operands carry 2^-9 relative
rounding; with supplied noise and every call started from the reference's state, the latents of the recorded
cases stay within 5e-2 (max-norm relative; 1e-1 for the shipped-checkpoint case) of the fp32 reference over
their horizons (<= 60 steps), energies /
losses within 2e-2, normalised weight gradients within 3e-2 of their scale, parameters after the p-step within
5e-3 absolute.  Against the bf16-emulating oracle the same quantities agree to 2e-3 (differences come from
fp32 accumulation order flipping individual bf16 roundings)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.optim as optim

from golden_util import orc, rel_err
from trainer_replay import replay

from montecarlopredictivecoding_b200 import mcpc_utils as mu
from montecarlopredictivecoding_b200 import predictive_coding as pc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

BF16_CASES = ["mcpc_relu_bce_learn", "mcpc_ml_checkpoint", "pc_tanh_adam_mask", "fig2_linear", "free_output_layer",
              "gauss_mask_one_sample", "zero_fn_sampling", "update_p_all", "adam_carryover_batch_resize",
              "last_half_schedules"]


@pytest.mark.parametrize("name", BF16_CASES)
def test_golden_case_bf16_bound(name):
    # the shipped MNIST checkpoint has the largest weight norms: its Langevin call is the loosest case (1e-1)
    tol_x = 1e-1 if name == "mcpc_ml_checkpoint" else 5e-2
    worst = replay(name, torch.device(DEV), precision="bf16", tol_x=tol_x, tol_s=2e-2, tol_g=3e-2, tol_w=5e-3,
                   teacher_force=True, w_min_grad=0.05)
    print(name, {k: f"{v:.2e}" for k, v in worst.items()})


def test_nonzero_inputs_golden_case_in_bf16():
    """Non-zero ``inputs`` (mu_0 = W_0 inputs + b_0 per chain, gW_0 = (sum_t G_0)^T inputs) under the same bound."""
    worst = replay("mcpc_tanh_bce_learn_inputs", torch.device(DEV), precision="bf16", tol_x=5e-2, tol_s=2e-2, tol_g=3e-2,
                   tol_w=5e-3, teacher_force=True, w_min_grad=0.05)
    print({k: f"{v:.2e}" for k, v in worst.items()})


def test_auto_precision_takes_bf16_with_nonzero_inputs():
    dev = torch.device(DEV)
    cfg = {"input_size": 20, "hidden_size": 64, "hidden2_size": 64, "output_size": 32, "activation_fn": "tanh"}
    model = mu.get_model(cfg, use_cuda=False).to(dev)
    tr = pc.PCTrainer(model, T=6, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.05}, update_p_at="last",
                      optimizer_p_fn=optim.SGD, optimizer_p_kwargs={"lr": 0.01}, plot_progress_at=[])
    tr.set_precision("auto")
    tr.train_on_batch(torch.randn(16, 20, device=dev), loss_fn=mu.fe_fn, loss_fn_kwargs={"_target": torch.randn(16, 32, device=dev), "_var": 1.0},
                      is_log_progress=False)
    assert tr.last_call_info["precision"] == 1


# chains per CTA: B <= 1184 -> 8 (alternate-tile epilogue halves), up to 4,720 -> 16, beyond -> 32 (MMA N = 32)
@pytest.mark.parametrize("act,top,B,with_inputs", [("relu", "bernoulli", 1024, False), ("tanh", "gauss", 200, False),
                                                   ("relu", "zero", 4096, False), ("relu", "bernoulli", 2048, False),
                                                   ("tanh", "bernoulli", 8192, False), ("tanh", "bernoulli", 1000, True),
                                                   ("tanh", "gauss", 2048, True), ("relu", "bernoulli", 6000, True)])
def test_bf16_kernel_vs_bf16_oracle(act, top, B, with_inputs):
    _bf16_kernel_vs_bf16_oracle(act, top, B, with_inputs, (20, 128, 128, 784))


# Other tilings of the resident kernel: a single partial output tile (10 units: only its 16 valid rows are kept in shared
# memory and the M = 128 prediction MMA reads the following tiles), partial tiles of 72 and 2 valid units behind full ones,
# exactly 7 full output tiles, hidden layers of 2 and 3 unit tiles, a 12-unit input layer.
@pytest.mark.parametrize("dims,act,top,B", [((12, 64, 200, 10), "tanh", "gauss", 300), ((20, 128, 128, 200), "relu", "bernoulli", 1024),
                                            ((16, 300, 96, 130), "tanh", "bernoulli", 520), ((20, 128, 128, 896), "relu", "bernoulli", 1024),
                                            ((20, 128, 128, 10), "relu", "zero", 2048)])
def test_bf16_kernel_tilings_vs_bf16_oracle(dims, act, top, B):
    _bf16_kernel_vs_bf16_oracle(act, top, B, False, dims)


def _bf16_kernel_vs_bf16_oracle(act, top, B, with_inputs, dims):
    dev = torch.device(DEV)
    mixing, sampling, lr = 3, 5, 0.03
    T = mixing + sampling
    torch.manual_seed(1)
    d0, d1, d2, d_out = dims
    cfg = {"input_size": d0, "hidden_size": d1, "hidden2_size": d2, "output_size": d_out, "activation_fn": act}
    model = mu.get_model(cfg, use_cuda=False).to(dev)
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": lr}, update_p_at="last",
                      accumulate_p_at=list(range(mixing, T)), optimizer_p_fn=optim.SGD, optimizer_p_kwargs={"lr": 0.0},
                      plot_progress_at=[])
    tr.set_precision("bf16")
    tr.set_noise_seed(31337)
    y = (torch.rand(B, d_out, device=dev) < 0.5).float() if top != "gauss" else torch.randn(B, d_out, device=dev)
    x0 = [torch.randn(B, d, device=dev) for d in (d0, d1, d2)]
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    lins = [m for m in model if isinstance(m, nn.Linear)]
    for layer, v in zip(pcs, x0):
        layer._sample_x_fn = (lambda inputs, v=v: v.clone())
    loss_fn = {"bernoulli": mu.bernoulli_fn, "gauss": mu.fe_fn, "zero": mu.zero_fn}[top]
    kw = {} if top == "zero" else {"loss_fn_kwargs": {"_target": y, "_var": 1.0}}
    inputs = torch.randn(B, d0, device=dev) if with_inputs else torch.zeros(B, d0, device=dev)
    res = tr.train_on_batch(inputs, loss_fn=loss_fn, callback_after_t=mu.random_step,
                            callback_after_t_kwargs={"_pc_trainer": tr}, is_log_progress=False,
                            is_return_results_every_t=True, is_return_outputs=True, **kw)
    assert tr.last_call_info["precision"] == 1
    nz = tr._get_engine().fill_noise(31337, 0, T, 0, B, d0 + d1 + d2, float(np.sqrt(2.0 / lr)), dev).cpu().numpy()
    offs = [0, d0, d0 + d1, d0 + d1 + d2]
    noise = [[nz[t][:, offs[l]:offs[l + 1]] for l in range(3)] for t in range(T)]
    net = orc.OracleNet(W=[l.weight.detach().cpu().numpy() for l in lins], b=[l.bias.detach().cpu().numpy() for l in lins],
                        n_layers=3, act=[orc.ACT_RELU if act == "relu" else orc.ACT_TANH] * 3, energy_scale=[1.0] * 3,
                        top={"bernoulli": orc.TOP_BERNOULLI, "gauss": orc.TOP_GAUSS, "zero": orc.TOP_ZERO}[top],
                        bf16_operands=True)
    ref = orc.infer(net, [v.cpu().numpy() for v in x0], inputs.cpu().numpy(), y.cpu().numpy(), T,
                    optimizer="sgd", lr=lr, noise=noise, acc_begin=mixing, acc_end=T, record_traj=True)
    errs = {f"x{l}": rel_err(pcs[l].get_x().detach().cpu().numpy(), ref.xs[l]) for l in range(3)}
    errs["energy"] = rel_err(res["energy"], ref.energy)
    if top != "zero":
        errs["loss"] = rel_err(res["loss"], ref.loss)
    errs["out"] = rel_err(torch.stack(res["outputs"]).cpu().numpy(), np.stack(ref.traj_out))
    div = sampling * B
    if top != "zero":
        errs["gW_out"] = rel_err(lins[3].weight.grad.cpu().numpy(), ref.gW[3] / div)
    errs["gW_2"] = rel_err(lins[2].weight.grad.cpu().numpy(), ref.gW[2] / div)
    errs["gb_0"] = rel_err(lins[0].bias.grad.cpu().numpy(), ref.gb[0] / div)
    if with_inputs:
        errs["gW_0"] = rel_err(lins[0].weight.grad.cpu().numpy(), ref.gW[0] / div)
    print(dims, act, top, B, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v < 2e-3, (k, v)


def _run_plain_call(kind, B, nospec):
    """One call without trajectories (the shape the specialised kernel instantiations serve); returns final latents,
    energies and parameter gradients."""
    import os
    if nospec:
        os.environ["MCPC_TC_NOSPEC"] = "1"
    else:
        os.environ.pop("MCPC_TC_NOSPEC", None)
    try:
        torch.manual_seed(0)
        act = "tanh" if kind == "map" else "relu"
        cfg = {"input_size": 20, "hidden_size": 128, "hidden2_size": 128, "output_size": 784, "activation_fn": act}
        model = mu.get_model(cfg, use_cuda=False, sample_x_fn=mu.sample_x_fn_normal).to(DEV)
        y = (torch.rand(B, 784, device=DEV) < 0.5).float()
        z = torch.zeros(B, 20, device=DEV)
        if kind == "map":
            tr = pc.PCTrainer(model, T=40, optimizer_x_fn=optim.Adam, optimizer_x_kwargs={"lr": 0.1}, update_p_at="never",
                              plot_progress_at=[])
            kw = dict(loss_fn=mu.bernoulli_fn, loss_fn_kwargs={"_target": y, "_var": None})
        else:
            tr = pc.PCTrainer(model, T=30, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.03}, update_p_at="last",
                              accumulate_p_at=list(range(10, 30)), optimizer_p_fn=optim.SGD, optimizer_p_kwargs={"lr": 0.0},
                              plot_progress_at=[])
            kw = dict(callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr})
            if kind == "mcpc":
                kw.update(loss_fn=mu.bernoulli_fn, loss_fn_kwargs={"_target": y, "_var": None})
            else:
                kw.update(loss_fn=mu.zero_fn)
        tr.set_precision("bf16")
        tr.set_noise_seed(77)
        torch.manual_seed(5)
        res = tr.train_on_batch(inputs=z, is_log_progress=False, is_checking_after_callback_after_t=False, **kw)
        xs = [layer.get_x().detach().clone() for layer in model if isinstance(layer, pc.PCLayer)]
        grads = [p.grad.detach().clone() for p in model.parameters() if p.grad is not None]
        return xs, torch.tensor(res["energy"]), grads
    finally:
        os.environ.pop("MCPC_TC_NOSPEC", None)


@pytest.mark.parametrize("kind,B", [("mcpc", 1024), ("map", 512), ("sample", 1024), ("sample", 8192), ("mcpc", 2048)])
def test_specialised_instantiations_equal_the_generic_kernel(kind, B):
    """The mode-specialised instantiations (SGD+Philox+Bernoulli, Adam+Bernoulli, sampling without top gradient) only
    fold runtime flags into constants: same arithmetic, same results as the generic instantiation."""
    a = _run_plain_call(kind, B, nospec=False)
    b = _run_plain_call(kind, B, nospec=True)
    for xa, xb in zip(a[0], b[0]):
        assert torch.allclose(xa, xb, rtol=1e-5, atol=1e-5)
    assert torch.allclose(a[1], b[1], rtol=1e-5)
    assert len(a[2]) == len(b[2])
    for ga, gb in zip(a[2], b[2]):
        assert torch.allclose(ga, gb, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("B,T,mixing", [(1024, 150, 50), (1000, 40, 0), (333, 25, 5)])
def test_overlapped_weight_update_equals_the_sequential_one(B, T, mixing, monkeypatch):
    """Resident bf16 kernel, B <= ~1100: the weight-gradient kernel runs NEXT to the inference kernel on the idle SMs and
    consumes each saved step behind per-step flags.  Same call with the overlap off (weight update after the kernel) must
    give the same gradients; the saved-operand buffers are poisoned with NaN before the call, so a row that the consumer
    read before the producer had written it would show up as NaN."""
    dev = torch.device(DEV)

    def run(overlap):
        monkeypatch.setenv("MCPC_TC_DW_OVERLAP", "1" if overlap else "0")
        torch.manual_seed(0)
        cfg = {"input_size": 20, "hidden_size": 128, "hidden2_size": 128, "output_size": 784, "activation_fn": "relu"}
        model = mu.get_model(cfg, use_cuda=False).to(dev)
        tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.03}, update_p_at="last",
                          accumulate_p_at=list(range(mixing, T)), optimizer_p_fn=optim.SGD, optimizer_p_kwargs={"lr": 0.0},
                          plot_progress_at=[])
        tr.set_precision("bf16")
        tr.set_noise_seed(99)
        y = (torch.rand(B, 784, device=dev) < 0.5).float()
        x0 = [torch.randn(B, d, device=dev) for d in (20, 128, 128)]
        pcs = [m for m in model if isinstance(m, pc.PCLayer)]
        for layer, v in zip(pcs, x0):
            layer._sample_x_fn = (lambda inputs, v=v: v.clone())
        z = torch.zeros(B, 20, device=dev)
        out = None
        for rep in range(3):
            tr._noise_epoch = 0
            for role in ("save_g", "save_f"):
                if role in tr._buffers:
                    tr._buffers[role].fill_(float("nan"))
            res = tr.train_on_batch(z, loss_fn=mu.bernoulli_fn, loss_fn_kwargs={"_target": y, "_var": 1.0},
                                    callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr},
                                    is_log_progress=False, is_checking_after_callback_after_t=False)
            grads = [p.grad.detach().clone() for p in model.parameters() if p.grad is not None]
            assert all(bool(torch.isfinite(g).all()) for g in grads), f"NaN in the weight gradients (overlap={overlap}, call {rep})"
            out = (grads, torch.tensor(res["energy"]))
        eng = tr._get_engine()
        from montecarlopredictivecoding_b200.predictive_coding import plan as P
        netp = P.compile_net(model)
        top = P.classify_loss(mu.bernoulli_fn, {"_target": y, "_var": 1.0}, B, 784, dev)
        assert eng.infer_fuses_weight_grad(netp, top, B, 1, False) == overlap
        return out

    g_seq, e_seq = run(False)
    g_ovl, e_ovl = run(True)
    assert torch.equal(e_seq, e_ovl)
    for a, b in zip(g_seq, g_ovl):
        scale = float(a.abs().max()) + 1e-12
        # fp32 accumulation order differs (2 interleaved K slabs of 51,200 rows vs 15 contiguous ones)
        assert float((a - b).abs().max()) / scale < 2e-4, float((a - b).abs().max()) / scale


@pytest.mark.parametrize("B", [1024, 8192])
def test_sampling_instantiation_records_the_same_trajectories_as_the_generic_kernel(B, monkeypatch):
    """Instantiation 3 (SGD + Philox, no loss gradient) also serves sampling calls that record thinned read-outs
    (set_trajectory_stride): outputs, latents of every recorded step and the energies must equal the generic kernel's."""
    dev = torch.device(DEV)

    def run(nospec):
        if nospec:
            monkeypatch.setenv("MCPC_TC_NOSPEC", "1")
        else:
            monkeypatch.delenv("MCPC_TC_NOSPEC", raising=False)
        torch.manual_seed(0)
        cfg = {"input_size": 20, "hidden_size": 128, "hidden2_size": 128, "output_size": 784, "activation_fn": "relu"}
        model = mu.get_model(cfg, use_cuda=False, sample_x_fn=mu.sample_x_fn_normal).to(dev)
        tr = pc.PCTrainer(model, T=40, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.03}, update_p_at="never",
                          plot_progress_at=[])
        tr.set_precision("bf16")
        tr.set_noise_seed(5)
        tr.set_trajectory_stride(7, 3)
        tr.set_trajectories_on_device(True)
        torch.manual_seed(9)
        res = tr.train_on_batch(torch.zeros(B, 20, device=dev), loss_fn=mu.zero_fn, callback_after_t=mu.random_step,
                                callback_after_t_kwargs={"_pc_trainer": tr}, is_log_progress=False, is_return_outputs=True,
                                is_return_xs=True, is_checking_after_callback_after_t=False)
        return (torch.stack(res["outputs"]), [torch.stack([xs[l] for xs in res["xs"]]) for l in range(3)],
                torch.tensor(res["energy"]))

    out_g, xs_g, e_g = run(True)
    out_s, xs_s, e_s = run(False)
    assert out_s.shape[0] == len(range(3, 40, 7))
    assert torch.equal(out_g, out_s)
    for a, b in zip(xs_g, xs_s):
        assert torch.equal(a, b)
    assert torch.equal(e_g, e_s)
