"""SURVEY §8(f) N2 / N3 on the GPU through the C ABI:
  * mcpc_traj_stats_update against torch reductions (single block and chunked merges);
  * thinned trajectory recording by the kernels (McpcOpts.traj_every) == slices of the every-step trajectory, for the
    resident fp32 / bf16 kernels and the streaming kernels;
  * mcpc_p_step (fused normalise + optimizer_p.step) against torch.optim.SGD / Adam on the same numbers, including the
    optimizer state it maintains in place."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.optim as optim

from montecarlopredictivecoding_b200 import _native as N
from montecarlopredictivecoding_b200 import mcpc_utils as mu
from montecarlopredictivecoding_b200 import predictive_coding as pc
from montecarlopredictivecoding_b200.predictive_coding.engine import NativeEngine

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n_rec,shape,chunks", [(1, (3, 5), 1), (37, (64, 20), 1), (100, (16, 128), 4), (10, (1, 1), 3)])
def test_traj_stats_kernel_vs_torch(n_rec, shape, chunks):
    dev = torch.device(DEV)
    eng = NativeEngine()
    torch.manual_seed(0)
    traj = (torch.randn(n_rec, *shape, device=dev) * 3.0 + 1.5).contiguous()
    mean = torch.full(shape, 7.0, device=dev)         # garbage: count_before = 0 must ignore it
    m2 = torch.full(shape, -3.0, device=dev)
    bounds = np.linspace(0, n_rec, chunks + 1).astype(int)
    count = 0
    for a, b in zip(bounds[:-1], bounds[1:]):
        if b > a:
            eng.traj_stats(traj[a:b].contiguous(), int(b - a), count, mean, m2)
            count += int(b - a)
    assert count == n_rec
    assert torch.allclose(mean, traj.mean(0), rtol=1e-5, atol=1e-5)
    if n_rec > 1:
        assert torch.allclose(m2 / (n_rec - 1), traj.var(0), rtol=1e-4, atol=1e-5)


def _trainer(precision, T, dims=(20, 128, 128), d_out=784, force_streaming=False):
    dev = torch.device(DEV)
    torch.manual_seed(0)
    cfg = {"input_size": dims[0], "hidden_size": dims[1], "hidden2_size": dims[2], "output_size": d_out, "activation_fn": "relu"}
    model = mu.get_model(cfg, use_cuda=False, sample_x_fn=mu.sample_x_fn_normal).to(dev)
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.03}, update_p_at="never", plot_progress_at=[])
    tr.set_precision(precision)
    return model, tr


@pytest.mark.parametrize("precision,streaming", [("fp32", False), ("bf16", False), ("bf16", True)])
def test_thinned_recording_and_device_stats_on_gpu(precision, streaming, monkeypatch):
    if streaming:
        monkeypatch.setenv("MCPC_FORCE_STREAMING", "1")
    dev = torch.device(DEV)
    T, B, stride, start = 40, 48, 6, 7
    dims = (32, 128, 64) if streaming else (20, 128, 128)
    d_out = 96 if streaming else 784
    model, tr = _trainer(precision, T, dims, d_out)
    y = (torch.rand(B, d_out, device=dev) < 0.5).float()

    def call(**kw):
        torch.manual_seed(5)
        tr.set_noise_seed(77)
        return tr.train_on_batch(torch.zeros(B, dims[0], device=dev), loss_fn=mu.bernoulli_fn, loss_fn_kwargs={"_target": y},
                                 callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr},
                                 is_log_progress=False, is_checking_after_callback_after_t=False, **kw)
    full = call(is_return_xs=True, is_return_outputs=True)
    tr.set_trajectory_stride(stride, start)
    tr.set_trajectories_on_device(True)
    thin = call(is_return_xs=True, is_return_outputs=True)
    steps = list(range(start, T, stride))
    assert len(thin["xs"]) == len(steps)
    for r, t in enumerate(steps):
        for l in range(3):
            assert thin["xs"][r][l].is_cuda
            assert torch.equal(thin["xs"][r][l].cpu(), full["xs"][t][l])
        assert torch.equal(thin["outputs"][r], full["outputs"][t])
    assert np.allclose(thin["energy"], full["energy"], rtol=1e-6)
    # statistics over a bounded ring (3 records per chunk) == torch over the full trajectory
    tr.set_trajectory_stride(1, 0)
    tr._traj_ring_bytes = 3 * 4 * B * max(dims)
    tr.enable_trajectory_stats(start=start, stride=stride)
    res = call()
    st = tr.trajectory_stats()
    assert st["count"] == len(steps) and "xs" not in res
    for l in range(3):
        ref = torch.stack([full["xs"][t][l] for t in steps]).to(dev)
        assert torch.allclose(st["mean"][l], ref.mean(0), rtol=1e-5, atol=1e-5)
        assert torch.allclose(st["var"][l], ref.var(0), rtol=1e-4, atol=1e-5)


def _rand_params(dev, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    shapes = [(20, 20), (20,), (128, 20), (128,), (784, 128), (784,), (3, 1)]
    return [torch.randn(*s, generator=g).to(dev) for s in shapes]


@pytest.mark.parametrize("name,cls,kw", [
    ("sgd", optim.SGD, dict(lr=0.07)),
    ("sgd_momentum", optim.SGD, dict(lr=0.07, momentum=0.2)),
    ("sgd_momentum_damp_wd", optim.SGD, dict(lr=0.05, momentum=0.9, dampening=0.1, weight_decay=0.01)),
    ("sgd_nesterov", optim.SGD, dict(lr=0.05, momentum=0.9, nesterov=True)),
    ("adam", optim.Adam, dict(lr=0.01)),
    ("adam_wd_betas", optim.Adam, dict(lr=0.15, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.02)),
])
def test_fused_p_step_matches_torch_optimizers(name, cls, kw):
    """Three successive parameter updates through PCTrainer._p_step (mcpc_p_step) against the plain torch path
    (flat.div_ + optimizer.step, pc_trainer.py:904-914) on clones: parameters, normalised .grad and optimizer state."""
    dev = torch.device(DEV)
    model_a = nn.Sequential(nn.Linear(6, 16), pc.PCLayer(), nn.ReLU(), nn.Linear(16, 12), pc.PCLayer(), nn.ReLU(), nn.Linear(12, 24)).to(dev)
    model_b = nn.Sequential(nn.Linear(6, 16), pc.PCLayer(), nn.ReLU(), nn.Linear(16, 12), pc.PCLayer(), nn.ReLU(), nn.Linear(12, 24)).to(dev)
    model_b.load_state_dict(model_a.state_dict())
    trs = []
    for m, fused in ((model_a, True), (model_b, False)):
        m.train()
        tr = pc.PCTrainer(m, T=4, update_p_at="last", accumulate_p_at=[2, 3], optimizer_p_fn=cls, optimizer_p_kwargs=kw, plot_progress_at=[])
        tr._fused_p_optimizer = fused
        trs.append(tr)
    from montecarlopredictivecoding_b200.predictive_coding import plan as P
    B = 8
    for it in range(3):
        flats = []
        for tr in trs:
            netp = P.compile_net(tr.get_model())
            flat, gW, gb = tr._ensure_flat_grads(netp, zero=True)
            g = torch.Generator(device="cpu").manual_seed(100 + it)
            flat.copy_(torch.randn(flat.numel(), generator=g).to(dev) * 5.0)
            flats.append(flat)
        launches0 = N.load().mcpc_launch_count()
        trs[0]._p_step(flats[0], B)
        assert N.load().mcpc_launch_count() == launches0 + 1, "the fused path must be ONE mcpc_p_step launch"
        trs[1]._p_step(flats[1], B)
        if not kw.get("nesterov", False):     # torch's foreach SGD adds momentum*buf INTO .grad for nesterov (in place)
            assert torch.allclose(flats[0], flats[1], rtol=1e-6, atol=1e-7), "normalised .grad"
        for pa, pb in zip(model_a.parameters(), model_b.parameters()):
            assert torch.allclose(pa, pb, rtol=2e-6, atol=2e-7), (name, it)
        sa, sb = trs[0].get_optimizer_p().state, trs[1].get_optimizer_p().state
        for pa, pb in zip(model_a.parameters(), model_b.parameters()):
            for key in sb[pb]:
                va, vb = sa[pa][key], sb[pb][key]
                if torch.is_tensor(vb):
                    assert torch.allclose(va.float().cpu(), vb.float().cpu(), rtol=2e-6, atol=1e-7), (name, it, key)
