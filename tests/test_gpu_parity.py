"""Parity tests proper: the CUDA path (through the C ABI, via the drop-in PCTrainer) against the golden
vectors recorded from the reference and against the oracle.  Tolerances (north star): 1e-5 relative on
per-step latents / energies with supplied noise in fp32 mode."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.optim as optim

from golden_util import ALL_CASES, GoldenCase, orc, rel_err
from trainer_replay import replay

from montecarlopredictivecoding_b200 import _native as N
from montecarlopredictivecoding_b200 import mcpc_utils as mu
from montecarlopredictivecoding_b200 import predictive_coding as pc
from montecarlopredictivecoding_b200.predictive_coding.engine import NativeEngine

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.mark.parametrize("name", ALL_CASES)
def test_golden_case_fp32(name):
    worst = replay(name, torch.device(DEV), precision="fp32")
    print(name, worst)


def test_stepwise_mode_on_gpu():
    from trainer_replay import build_model, make_trainer
    gc = GoldenCase("pc_tanh_adam_mask")
    dev = torch.device(DEV)
    model = build_model(gc, dev)
    trainer = make_trainer(model, gc.calls[0]["trainer"])
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    for l, layer in enumerate(pcs):
        layer._sample_x_fn = (lambda inputs, v=torch.from_numpy(gc.x0(0)[l]).to(dev): v.clone())
    res = trainer.train_on_batch(
        inputs=torch.from_numpy(gc.inputs).to(dev), loss_fn=mu.bernoulli_fn_mask,
        loss_fn_kwargs={"_target": torch.from_numpy(gc.target).to(dev), "_var": 1.0},
        callback_after_backward=lambda t: None, is_log_progress=False, is_return_xs=True)
    assert trainer.last_call_info["mode"] == "stepwise"
    T = gc.calls[0]["trainer"]["T"]
    for l in range(gc.L):
        got = np.stack([res["xs"][t][l].numpy() for t in range(T)])
        assert rel_err(got, gc.traj(0, l)) < 1e-5
    assert rel_err(res["energy"], gc.z["c0_energy"]) < 1e-5


def test_philox_noise_matches_cpu_restatement():
    eng = NativeEngine()
    got = eng.fill_noise(seed=0x1234ABCD5678, t_begin=3, n_steps=4, chain_offset=1021, B=37, n_units=29,
                         noise_scale=1.0, device=torch.device(DEV)).cpu().numpy()
    ref = np.stack([orc.langevin_normals(0x1234ABCD5678, 3 + s, 1021, 37, 29) for s in range(4)])
    # same Philox bits; Box-Muller runs with fast device intrinsics
    assert np.max(np.abs(got - ref)) < 2e-5
    big = eng.fill_noise(seed=7, t_begin=0, n_steps=8, chain_offset=0, B=4096, n_units=64, noise_scale=2.0,
                         device=torch.device(DEV))
    assert abs(float(big.mean())) < 0.01 and abs(float(big.var()) - 4.0) < 0.03


def _ml_model(dev, act="relu", dims=(20, 128, 128), d_out=784, seed=0):
    torch.manual_seed(seed)
    cfg = {"input_size": dims[0], "hidden_size": dims[1], "hidden2_size": dims[2], "output_size": d_out,
           "activation_fn": act}
    return mu.get_model(cfg, use_cuda=False).to(dev), cfg


def test_in_kernel_noise_equals_supplied_noise_run():
    """NOISE_PHILOX must apply exactly the tensor mcpc_fill_noise materialises (so a generated-noise
    run can be replayed by the reference / oracle with recorded noise)."""
    dev = torch.device(DEV)
    B, T, lr = 64, 9, 0.03
    finals = []
    for mode in ("philox", "supplied"):
        model, cfg = _ml_model(dev)
        tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": lr}, update_p_at="never",
                          plot_progress_at=[])
        tr.set_noise_seed(99)
        torch.manual_seed(5)
        y = (torch.rand(B, 784, device=dev) < 0.5).float()
        x0 = [torch.randn(B, d, device=dev) for d in (20, 128, 128)]
        for layer, v in zip([m for m in model if isinstance(m, pc.PCLayer)], x0):
            layer._sample_x_fn = (lambda inputs, v=v: v.clone())
        if mode == "supplied":
            nz = tr._get_engine().fill_noise(99, 0, T, 0, B, 276, float(np.sqrt(2.0 / lr)), dev)
            tr.set_supplied_noise(nz)
        tr.train_on_batch(torch.zeros(B, 20, device=dev), loss_fn=mu.bernoulli_fn, loss_fn_kwargs={"_target": y},
                          callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr},
                          is_log_progress=False, is_return_results_every_t=False)
        finals.append([m.get_x().detach().clone() for m in model if isinstance(m, pc.PCLayer)])
    for a, b in zip(*finals):
        assert torch.equal(a, b)


@pytest.mark.parametrize("act,top", [("relu", "bernoulli"), ("tanh", "gauss")])
def test_full_size_against_oracle(act, top):
    """mcpc_ml shape at the benchmark batch (B=1024): kernel vs oracle with kernel-generated noise,
    mixing 5 + sampling 10, including the accumulated weight gradient."""
    dev = torch.device(DEV)
    B, mixing, sampling, lr = 1024, 5, 10, 0.03
    T = mixing + sampling
    model, cfg = _ml_model(dev, act=act)
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": lr}, update_p_at="last",
                      accumulate_p_at=list(range(mixing, T)), optimizer_p_fn=optim.SGD,
                      optimizer_p_kwargs={"lr": 0.0}, plot_progress_at=[])
    tr.set_noise_seed(4242)
    torch.manual_seed(11)
    y = (torch.rand(B, 784, device=dev) < 0.5).float() if top == "bernoulli" else torch.randn(B, 784, device=dev)
    x0 = [torch.randn(B, d, device=dev) for d in (20, 128, 128)]
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    lins = [m for m in model if isinstance(m, nn.Linear)]
    for layer, v in zip(pcs, x0):
        layer._sample_x_fn = (lambda inputs, v=v: v.clone())
    loss_fn = mu.bernoulli_fn if top == "bernoulli" else mu.fe_fn
    res = tr.train_on_batch(torch.zeros(B, 20, device=dev), loss_fn=loss_fn, loss_fn_kwargs={"_target": y, "_var": 1.0},
                            callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr},
                            is_log_progress=False, is_return_results_every_t=True)
    nz = tr._get_engine().fill_noise(4242, 0, T, 0, B, 276, float(np.sqrt(2.0 / lr)), dev).cpu().numpy()
    offs = [0, 20, 148, 276]
    noise = [[nz[t][:, offs[l]:offs[l + 1]] for l in range(3)] for t in range(T)]
    net = orc.OracleNet(W=[l.weight.detach().cpu().numpy() for l in lins], b=[l.bias.detach().cpu().numpy() for l in lins],
                        n_layers=3, act=[orc.ACT_RELU if act == "relu" else orc.ACT_TANH] * 3, energy_scale=[1.0] * 3,
                        top=orc.TOP_BERNOULLI if top == "bernoulli" else orc.TOP_GAUSS)
    ref = orc.infer(net, [v.cpu().numpy() for v in x0], np.zeros((B, 20), np.float32), y.cpu().numpy(), T,
                    optimizer="sgd", lr=lr, noise=noise, acc_begin=mixing, acc_end=T)
    for l in range(3):
        assert rel_err(pcs[l].get_x().detach().cpu().numpy(), ref.xs[l]) < 1e-5
    assert rel_err(res["energy"], ref.energy) < 1e-5
    assert rel_err(res["loss"], ref.loss) < 1e-5
    div = sampling * B
    for i, lin in enumerate(lins):
        if i == 0:
            assert float(lin.weight.grad.abs().max()) == 0.0         # zero inputs: dead weight
        else:
            assert rel_err(lin.weight.grad.cpu().numpy(), ref.gW[i] / div) < 2e-5
        assert rel_err(lin.bias.grad.cpu().numpy(), ref.gb[i] / div) < 2e-5


def test_posterior_statistics_generated_noise():
    """figure_2.py:40-48,79: analytic posterior N(0.44, 0.2) of the linear-Gaussian model, in-kernel noise."""
    dev = torch.device(DEV)
    model = nn.Sequential(nn.Linear(1, 1), pc.PCLayer(sample_x_fn=mu.sample_x_fn_cte), nn.Linear(1, 1, bias=False))
    model.train()
    nn.init.constant_(model[0].bias, 0.2)
    nn.init.constant_(model[2].weight, 2.0)
    model.to(dev)
    B, T = 256, 4000
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.02}, update_p_at="never",
                      plot_progress_at=[])
    res = tr.train_on_batch(torch.zeros(B, 1, device=dev), loss_fn=mu.fe_fn,
                            loss_fn_kwargs={"_target": torch.ones(B, 1, device=dev), "_var": 1.0},
                            callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr},
                            is_log_progress=False, is_return_representations=True)
    s = torch.stack(res["representations"][500:])
    # Euler-Maruyama bias is O(lr): stationary variance 0.2/(1 - lr*5/2) = 0.2105
    assert abs(float(s.mean()) - 0.44) < 0.01
    assert abs(float(s.var()) - 0.2105) < 0.01
    assert len(res["energy"]) == T and len(res["loss"]) == T


@pytest.mark.parametrize("B,split", [(8192, 4096), (8192, 4099), (1000, 501)])
def test_idempotent_zero_lr_and_shard_invariance(B, split):
    """Size-independent properties at the sampling config's scale: (i) lr=0 without noise leaves the latents
    bit-identical; (ii) running rows [0,B) in one launch or as two shards with the right chain offsets gives
    bit-identical chains (the Philox stream is keyed by the global chain id) -- also when the second shard starts inside
    a group of four chains (the kernels' whole-quad fast path of the noise draw does not apply there) and on the 8-chain
    CTAs of a small batch."""
    dev = torch.device(DEV)
    T = 6
    model, cfg = _ml_model(dev)
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    torch.manual_seed(3)
    x0 = [torch.randn(B, d, device=dev) for d in (20, 128, 128)]

    def run(rows, chain_offset, lr):
        tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": lr}, update_p_at="never",
                          plot_progress_at=[])
        tr.set_noise_seed(77)
        tr._dp_chain_offset = chain_offset
        tr._chain_offset = lambda B_: chain_offset
        for layer, v in zip(pcs, x0):
            layer._sample_x_fn = (lambda inputs, v=v[rows]: v.clone())
        kw = {}
        if lr > 0:
            kw = dict(callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr})
        tr.train_on_batch(torch.zeros(len(x0[0][rows]), 20, device=dev), loss_fn=mu.zero_fn, is_log_progress=False,
                          is_return_results_every_t=False, **kw)
        return [p.get_x().detach().clone() for p in pcs]

    same = run(slice(0, B), 0, 0.0)
    for a, b in zip(same, x0):
        assert torch.equal(a, b)
    full = run(slice(0, B), 0, 0.1)
    lo = run(slice(0, split), 0, 0.1)
    hi = run(slice(split, B), split, 0.1)
    for f, a, b in zip(full, lo, hi):
        assert torch.equal(f, torch.cat([a, b]))


def test_ragged_and_tiny_batches():
    """B not a multiple of the row tile, B=1, and bias-free layers."""
    dev = torch.device(DEV)
    for B in (1, 3, 130):
        torch.manual_seed(B)
        model = nn.Sequential(nn.Linear(5, 5), pc.PCLayer(), nn.Tanh(), nn.Linear(5, 7, bias=False), pc.PCLayer(),
                              nn.Tanh(), nn.Linear(7, 6)).to(dev)
        model.train()
        pcs = [m for m in model if isinstance(m, pc.PCLayer)]
        lins = [m for m in model if isinstance(m, nn.Linear)]
        x0 = [torch.randn(B, 5, device=dev), torch.randn(B, 7, device=dev)]
        for layer, v in zip(pcs, x0):
            layer._sample_x_fn = (lambda inputs, v=v: v.clone())
        y = torch.randn(B, 6, device=dev)
        T = 7
        tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.Adam, optimizer_x_kwargs={"lr": 0.1}, update_p_at="never",
                          plot_progress_at=[])
        res = tr.train_on_batch(torch.zeros(B, 5, device=dev), loss_fn=mu.fe_fn, loss_fn_kwargs={"_target": y, "_var": 2.0},
                                is_log_progress=False)
        net = orc.OracleNet(W=[l.weight.detach().cpu().numpy() for l in lins],
                            b=[None if l.bias is None else l.bias.detach().cpu().numpy() for l in lins], n_layers=2,
                            act=[orc.ACT_TANH] * 2, energy_scale=[1.0, 1.0], top=orc.TOP_GAUSS, top_var=2.0)
        ref = orc.infer(net, [v.cpu().numpy() for v in x0], np.zeros((B, 5), np.float32), y.cpu().numpy(), T,
                        optimizer="adam", lr=0.1)
        for l in range(2):
            assert rel_err(pcs[l].get_x().detach().cpu().numpy(), ref.xs[l]) < 1e-5, B
        assert rel_err(res["energy"], ref.energy) < 1e-5
        assert rel_err(res["loss"], ref.loss) < 1e-5
