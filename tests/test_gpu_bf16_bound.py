"""The stated bf16 bound, demonstrated WHERE THE NUMBERS ARE QUOTED, and statistical parity with
generated noise (SURVEY §8 north star; goldens recorded from the real reference by tests/golden/make_golden_stats.py).

Bound of MCPC_PREC_BF16 against the fp32 oracle (no operand rounding) on identical inputs and identical noise
(the kernel's Philox stream, replayed through mcpc_fill_noise), FREE-RUNNING over the whole horizon.  What every
consumer of the path reads -- per-step energy / loss, the weight gradients -- stays tight; individual chains of a relu
net do drift (a bf16 rounding that flips one relu gate sends that chain on another, equally valid, noisy trajectory),
so the latents are bounded in RMS over all chains and units, and their worst single element is reported:

  config                                        energy / loss    weight grads   latents RMS   latents max-norm
                                                 (per step)       (max-norm)     (relative)    (worst element; measured r02)
  C2  mcpc_ml learning call, B=1024, T=150         5e-4             1e-2           2e-2          0.21 / 0.10 (checkpoint)
  C3  sampling, 8,192 chains, T=1000, zero_fn      5e-4              --            3e-2          0.31
  C4  deterministic PC, Adam lr 0.3, T=250,        2e-2 free-running end energy / loss; latents 1e-1 max-norm teacher-forced
      masked BCE, B=1024                           every 25 steps (Adam trajectories are chaotic w.r.t. rounding)

The fp32 mode (MCPC_PREC_FP32) holds 1e-5 on all of these (tests/test_gpu_parity.py).  Statistical parity with generated
noise (posterior mean / variance, energy and loss levels, table_1 MSE) is identical for both modes, see below."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.optim as optim

from golden_util import GOLDEN_DIR, orc, rel_err

from montecarlopredictivecoding_b200 import _native as N
from montecarlopredictivecoding_b200 import mcpc_utils as mu
from montecarlopredictivecoding_b200 import predictive_coding as pc
from montecarlopredictivecoding_b200.predictive_coding import plan as P
from montecarlopredictivecoding_b200.predictive_coding.engine import InferCall, NativeEngine

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rms_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.sqrt(np.mean(b ** 2)), 1e-30))


DIMS = (20, 128, 128)
OFFS = [0, 20, 148, 276]


def _model(act="relu", dims=DIMS, checkpoint=False, seed=0):
    dev = torch.device(DEV)
    torch.manual_seed(seed)
    cfg = {"input_size": dims[0], "hidden_size": dims[1], "hidden2_size": dims[2], "output_size": 784, "activation_fn": act}
    model = mu.get_model(cfg, use_cuda=False)
    if checkpoint:      # the shipped models/mcpc_ml_1 weights, as recorded in the golden fixture
        z = np.load(os.path.join(GOLDEN_DIR, "stat_langevin_mcpc_ml.npz"))
        with torch.no_grad():
            for i, lin in enumerate(m for m in model if isinstance(m, nn.Linear)):
                lin.weight.copy_(torch.from_numpy(z[f"W{i}"]))
                lin.bias.copy_(torch.from_numpy(z[f"b{i}"]))
    return model.to(dev)


def _oracle_net(model, act, top):
    lins = [m for m in model if isinstance(m, nn.Linear)]
    return orc.OracleNet(W=[l.weight.detach().cpu().numpy() for l in lins], b=[l.bias.detach().cpu().numpy() for l in lins],
                         n_layers=3, act=[orc.ACT_RELU if act == "relu" else orc.ACT_TANH] * 3, energy_scale=[1.0] * 3, top=top)


@pytest.mark.parametrize("checkpoint", [False, True])
def test_c2_bound_full_size(checkpoint):
    """bench.py's headline config: B=1024, T=150 (mixing 50 + sampling 100), SGD lr 0.03, var 2, dW over the sampling steps."""
    dev = torch.device(DEV)
    B, mixing, sampling, lr = 1024, 50, 100, 0.03
    T = mixing + sampling
    model = _model("relu", checkpoint=checkpoint)
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": lr}, update_p_at="last",
                      accumulate_p_at=list(range(mixing, T)), optimizer_p_fn=optim.SGD, optimizer_p_kwargs={"lr": 0.0},
                      plot_progress_at=[])
    tr.set_precision("bf16")
    tr.set_noise_seed(2025)
    torch.manual_seed(3)
    y = (torch.rand(B, 784, device=dev) < 0.5).float()
    x0 = [torch.randn(B, d, device=dev) for d in DIMS]
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    lins = [m for m in model if isinstance(m, nn.Linear)]
    for layer, v in zip(pcs, x0):
        layer._sample_x_fn = (lambda inputs, v=v: v.clone())
    res = tr.train_on_batch(torch.zeros(B, 20, device=dev), loss_fn=mu.bernoulli_fn, loss_fn_kwargs={"_target": y, "_var": 1.0},
                            callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr},
                            is_log_progress=False, is_checking_after_callback_after_t=False)
    assert tr.last_call_info["precision"] == N.PREC_BF16 and tr.last_call_info["mode"] == "fused"
    nz = tr._get_engine().fill_noise(2025, 0, T, 0, B, 276, float(np.sqrt(2.0 / lr)), dev).cpu().numpy()
    noise = [[nz[t][:, OFFS[l]:OFFS[l + 1]] for l in range(3)] for t in range(T)]
    ref = orc.infer(_oracle_net(model, "relu", orc.TOP_BERNOULLI), [v.cpu().numpy() for v in x0], np.zeros((B, 20), np.float32),
                    y.cpu().numpy(), T, optimizer="sgd", lr=lr, noise=noise, acc_begin=mixing, acc_end=T)
    errs = {f"x{l}_rms": rms_err(pcs[l].get_x().detach().cpu().numpy(), ref.xs[l]) for l in range(3)}
    worst = {f"x{l}_max": rel_err(pcs[l].get_x().detach().cpu().numpy(), ref.xs[l]) for l in range(3)}
    errs["energy"] = rel_err(res["energy"], ref.energy)
    errs["loss"] = rel_err(res["loss"], ref.loss)
    div = sampling * B
    for i in (1, 2, 3):
        errs[f"gW_{i}"] = rel_err(lins[i].weight.grad.cpu().numpy(), ref.gW[i] / div)
    print("C2 bf16 vs fp32 oracle", "checkpoint" if checkpoint else "default-init", {k: f"{v:.2e}" for k, v in {**errs, **worst}.items()})
    for k, v in errs.items():
        tol = 5e-4 if k in ("energy", "loss") else (2e-2 if k.startswith("x") else 1e-2)
        assert v < tol, (k, v)
    assert max(worst.values()) < 0.5


def test_c3_bound_sampling_8192_chains_T1000():
    """C3: generative sampling (zero_fn), lr 0.1, 8,192 chains, the benchmark's T=1000 window; the oracle is chained in
    windows of 50 steps so that the replayed noise (9 GB in one piece) stays small."""
    dev = torch.device(DEV)
    B, T, lr, win = 8192, 1000, 0.1, 50
    model = _model("relu")
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": lr}, update_p_at="never", plot_progress_at=[])
    tr.set_precision("bf16")
    tr.set_noise_seed(99)
    torch.manual_seed(4)
    x0 = [torch.randn(B, d, device=dev) for d in DIMS]
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    for layer, v in zip(pcs, x0):
        layer._sample_x_fn = (lambda inputs, v=v: v.clone())
    res = tr.train_on_batch(torch.zeros(B, 20, device=dev), loss_fn=mu.zero_fn, callback_after_t=mu.random_step,
                            callback_after_t_kwargs={"_pc_trainer": tr}, is_log_progress=False,
                            is_checking_after_callback_after_t=False)
    net = _oracle_net(model, "relu", orc.TOP_ZERO)
    xs = [v.cpu().numpy() for v in x0]
    energy = []
    eng = tr._get_engine()
    for c0 in range(0, T, win):
        nz = eng.fill_noise(99, c0, win, 0, B, 276, float(np.sqrt(2.0 / lr)), dev).cpu().numpy()
        noise = [[nz[t][:, OFFS[l]:OFFS[l + 1]] for l in range(3)] for t in range(win)]
        r = orc.infer(net, xs, np.zeros((B, 20), np.float32), None, win, optimizer="sgd", lr=lr, noise=noise)
        xs = r.xs
        energy += list(r.energy)
    errs = {f"x{l}_rms": rms_err(pcs[l].get_x().detach().cpu().numpy(), xs[l]) for l in range(3)}
    worst = {f"x{l}_max": rel_err(pcs[l].get_x().detach().cpu().numpy(), xs[l]) for l in range(3)}
    errs["energy"] = rel_err(res["energy"], energy)
    print("C3 bf16 vs fp32 oracle, T=1000:", {k: f"{v:.2e}" for k, v in {**errs, **worst}.items()})
    for k, v in errs.items():
        assert v < (5e-4 if k == "energy" else 3e-2), (k, v)
    assert max(worst.values()) < 0.6


def test_c4_bound_adam_T250_teacher_forced():
    """C4: deterministic PC (pc_ml shape 25-128-128->784, tanh, Adam lr 0.3, masked BCE, B=1024, T=250).  Free-running
    Adam trajectories are chaotic w.r.t. rounding, so the latents are compared teacher-forced: every 25
    steps the kernel restarts from the oracle's latents and Adam state; the end-of-inference energy / loss are compared
    free-running."""
    dev = torch.device(DEV)
    B, T, lr, win = 1024, 250, 0.3, 25
    dims = (25, 128, 128)
    model = _model("tanh", dims=dims, seed=2)
    lins = [m for m in model if isinstance(m, nn.Linear)]
    torch.manual_seed(6)
    y = (torch.rand(B, 784, device=dev) < 0.5).float()
    x0 = [torch.empty(B, d, device=dev).uniform_(-10, 10) for d in dims]
    net = _oracle_net(model, "tanh", orc.TOP_BERNOULLI)
    net.mask_start_col = 392
    # free-running oracle with the state at every window boundary
    xs = [v.cpu().numpy() for v in x0]
    adam = orc.AdamState([np.zeros_like(a) for a in xs], [np.zeros_like(a) for a in xs], 0)    # updated in place by infer
    states, energy, loss = [], [], []
    for c0 in range(0, T, win):
        states.append(([a.copy() for a in xs], ([m.copy() for m in adam.m], [v.copy() for v in adam.v], adam.step)))
        r = orc.infer(net, xs, np.zeros((B, 25), np.float32), y.cpu().numpy(), win, optimizer="adam", lr=lr, adam_state=adam)
        xs = r.xs
        energy += list(r.energy)
        loss += list(r.loss)
    ends = [s[0] for s in states[1:]] + [xs]
    # teacher-forced kernel windows through the engine (same ABI call the trainer makes)
    eng = NativeEngine()
    netp = P.compile_net(model)
    top = P.classify_loss(mu.bernoulli_fn_mask, {"_target": y, "_var": 1.0}, B, 784, dev)
    worst = 0.0
    for wi, c0 in enumerate(range(0, T, win)):
        xs_w, ad = states[wi]
        xd = [torch.from_numpy(a).to(dev).contiguous() for a in xs_w]
        m = [torch.from_numpy(a).to(dev).contiguous() for a in ad[0]]
        v = [torch.from_numpy(a).to(dev).contiguous() for a in ad[1]]
        e = torch.zeros(win, dtype=torch.float64, device=dev)
        l_ = torch.zeros(win, dtype=torch.float64, device=dev)
        eng.infer(InferCall(plan=netp, top=top, energy_coefficient=1.0, B=B, W=[q.weight.detach() for q in lins],
                            b=[q.bias.detach() for q in lins], x=xd, inputs=None, target=y, energy=e, loss=l_, n_steps=win,
                            t_begin=c0, optimizer=N.OPT_ADAM, update_x=True, lr=lr, adam_step0=ad[2],
                            adam_m=m, adam_v=v, precision=N.PREC_BF16))
        for l in range(3):
            worst = max(worst, rel_err(xd[l].cpu().numpy(), ends[wi][l]))
        assert rel_err(e.cpu().numpy(), energy[c0:c0 + win]) < 2e-2
        assert rel_err(l_.cpu().numpy(), loss[c0:c0 + win]) < 2e-2
    print(f"C4 bf16 vs fp32 oracle, teacher-forced every {win} steps: worst latent error {worst:.2e}")
    assert worst < 1e-1
    # free-running through the trainer: end-of-inference energy / loss
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.Adam, optimizer_x_kwargs={"lr": lr}, update_p_at="never", plot_progress_at=[])
    tr.set_precision("bf16")
    pcs = [q for q in model if isinstance(q, pc.PCLayer)]
    for layer, v0 in zip(pcs, x0):
        layer._sample_x_fn = (lambda inputs, v0=v0: v0.clone())
    res = tr.train_on_batch(torch.zeros(B, 25, device=dev), loss_fn=mu.bernoulli_fn_mask, loss_fn_kwargs={"_target": y, "_var": 1.0},
                            is_log_progress=False)
    e_end = abs(res["energy"][-1] - energy[-1]) / abs(energy[-1])
    l_end = abs(res["loss"][-1] - loss[-1]) / abs(loss[-1])
    print(f"C4 free-running T=250: end energy rel err {e_end:.2e}, end loss rel err {l_end:.2e}")
    assert e_end < 2e-2 and l_end < 2e-2


# ---------------------------------------------------------------------------------------------------------------------
#  statistical parity with GENERATED noise against statistics recorded from the real reference
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_langevin_statistics_match_the_reference(precision):
    """tests/golden/stat_langevin_mcpc_ml.npz: 32 independent Langevin runs of the reference (stock random_step, torch
    RNG) from one start state; here the 32 replicas are 32 x 64 chains of ONE call with the in-kernel Philox noise.
    Compared: mean over the last 100 steps of the energy / loss per replica (difference of the two sample means within
    4 standard errors), the pooled posterior mean of every first-layer latent (z-scores with one effective sample per
    replica), and the posterior variance scale (10 %)."""
    dev = torch.device(DEV)
    z = np.load(os.path.join(GOLDEN_DIR, "stat_langevin_mcpc_ml.npz"))
    R, B0 = z["energy"].shape[0], z["data"].shape[0]
    mixing, sampling, lr = int(z["mixing"]), int(z["sampling"]), float(z["lr"])
    T = mixing + sampling
    model = _model("relu", checkpoint=True)
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": lr}, update_p_at="never", plot_progress_at=[])
    tr.set_precision(precision)
    B = R * B0
    y = torch.from_numpy(z["data"]).to(dev).repeat(R, 1)
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    for l, layer in enumerate(pcs):
        v = torch.from_numpy(z[f"x_start{l}"]).to(dev).repeat(R, 1)
        layer._sample_x_fn = (lambda inputs, v=v: v.clone())
    tr.set_trajectory_stride(1, start=mixing)
    tr.set_trajectories_on_device(True)
    tr.enable_trajectory_stats(start=mixing, stride=1, layers=[0])
    # per-replica energies need per-chain resolution: record the first-layer latents AND recompute nothing else --
    # the per-step scalars of the call are sums over all 2048 chains, i.e. the SUM over replicas
    n_calls = 3                                            # independent noise realisations: 3 x 32 replicas of ours
    e_calls, l_calls = [], []
    for ci in range(n_calls):
        tr.set_noise_seed(123456 + 1000 * ci)
        res = tr.train_on_batch(torch.zeros(B, 20, device=dev), loss_fn=mu.bernoulli_fn, loss_fn_kwargs={"_target": y, "_var": None},
                                callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr},
                                is_log_progress=False, is_checking_after_callback_after_t=False, is_return_representations=True)
        e_calls.append(np.mean(res["energy"][mixing:]) / R)   # sum over replicas / R = mean replica
        l_calls.append(np.mean(res["loss"][mixing:]) / R)
    # (1) energy / loss: mean over replicas of the last-100-step mean
    e_ref = z["energy"][:, mixing:].mean(1)
    l_ref = z["loss"][:, mixing:].mean(1)
    e_ours, l_ours = float(np.mean(e_calls)), float(np.mean(l_calls))
    print(f"[{precision}] per-call energy {np.round(e_calls, 1)} loss {np.round(l_calls, 1)}")
    se_e = e_ref.std(ddof=1) * np.sqrt(1.0 / R + 1.0 / (n_calls * R))
    se_l = l_ref.std(ddof=1) * np.sqrt(1.0 / R + 1.0 / (n_calls * R))
    print(f"[{precision}] energy: ours {e_ours:.1f} ref {e_ref.mean():.1f} (se {se_e:.1f}); loss: ours {l_ours:.1f} ref {l_ref.mean():.1f} (se {se_l:.1f})")
    assert abs(e_ours - e_ref.mean()) < 4.0 * se_e
    assert abs(l_ours - l_ref.mean()) < 4.0 * se_l
    # (2) posterior mean / variance of the first-layer latents, pooled over replicas and the last 100 steps
    reps = torch.stack(res["representations"])            # [100, R*B0, 20] on the device
    assert reps.is_cuda and reps.shape[0] == sampling
    pooled = reps.reshape(sampling, R, B0, 20).permute(0, 1, 2, 3).reshape(sampling * R, B0, 20)
    m_ours = pooled.mean(0).cpu().numpy()
    v_ours = pooled.var(0).cpu().numpy()
    m_ref, v_ref = z["post_mean"], z["post_var"]
    zscore = (m_ours - m_ref) / np.sqrt((v_ours + v_ref) / R)
    print(f"[{precision}] posterior mean z-scores: rms {np.sqrt(np.mean(zscore ** 2)):.2f}, max {np.abs(zscore).max():.2f}; "
          f"variance ratio {v_ours.mean() / v_ref.mean():.3f}")
    assert np.sqrt(np.mean(zscore ** 2)) < 1.5 and np.abs(zscore).max() < 6.0
    assert abs(v_ours.mean() / v_ref.mean() - 1.0) < 0.10
    # the on-device statistics (N2) see the same samples: mean over the 100 steps per chain
    st = tr.trajectory_stats()
    assert st["count"] == sampling
    assert torch.allclose(st["mean"][0], reps.mean(0), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-3), ("bf16", 5e-3)])
@pytest.mark.parametrize("tag,act,d0", [("pc", "tanh", 30), ("mcpc", "relu", 10)])
def test_table1_mse_metric_matches_the_reference(tag, act, d0, precision, tol):
    """table_1.py get_models_mse -> get_mse_rec (utils/training_evaluation.py:143-174): MAP inference (Adam lr 0.7, T=250,
    masked BCE) on 128 images, decode, threshold, MSE on the masked-out half.  Golden value from the real reference on the
    shipped checkpoints; stated tolerance: 2e-3 absolute (fp32 kernels), 5e-3 (bf16) on an MSE of 0.08-0.17."""
    dev = torch.device(DEV)
    z = np.load(os.path.join(GOLDEN_DIR, "stat_mse_rec.npz"))
    cfg = {"input_size": d0, "hidden_size": 256, "hidden2_size": 256, "output_size": 784, "activation_fn": act}
    model = mu.get_model(cfg, use_cuda=False)
    with torch.no_grad():
        for i, lin in enumerate(m for m in model if isinstance(m, nn.Linear)):
            lin.weight.copy_(torch.from_numpy(z[f"{tag}_W{i}"]))
            lin.bias.copy_(torch.from_numpy(z[f"{tag}_b{i}"]))
    model.to(dev)
    data = torch.from_numpy(z[f"{tag}_data"]).to(dev)
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    for l, layer in enumerate(pcs):
        v = torch.from_numpy(z[f"{tag}_x0_{l}"]).to(dev)
        layer._sample_x_fn = (lambda inputs, v=v: v.clone())
    tr = mu.get_pc_trainer(model, {"T_pc": 250, "optimizer_x_fn_pc": optim.Adam, "optimizer_x_kwargs_pc": {"lr": 0.7}},
                           is_mcpc=True, training=False)
    tr.set_precision(precision)
    tr.train_on_batch(inputs=torch.zeros(data.shape[0], d0, device=dev), loss_fn=mu.bernoulli_fn_mask,
                      loss_fn_kwargs={"_target": data, "_var": None}, is_log_progress=False, is_return_results_every_t=False,
                      is_checking_after_callback_after_t=False)
    with torch.no_grad():
        img = model[-1](model[-2](model[-3].get_x().detach()))
        img = (img > 0).type_as(img)
        half = round(data.shape[1] / 2)
        mse = float(((img[:, :-half] - data[:, :-half]) ** 2).mean(1).sum() / data.shape[0])
    print(f"[{tag} {precision}] MSE ours {mse:.5f} reference {float(z[f'{tag}_mse']):.5f}")
    assert abs(mse - float(z[f"{tag}_mse"])) < tol
