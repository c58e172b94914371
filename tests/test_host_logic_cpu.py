"""Host logic of the drop-in PCTrainer on CPU: the kernels are replaced by the oracle test double
(tests/oracle_engine.py), everything else -- plan compiler, segmentation at p-updates, zero_grad
windows, normalisation, optimizer_p, results dict -- is the shipped code.  Checked against the
golden vectors recorded from the reference."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from golden_util import ALL_CASES
from oracle_engine import OracleEngine
from trainer_replay import replay

from montecarlopredictivecoding_b200 import _native as N
from montecarlopredictivecoding_b200 import mcpc_utils as mu
from montecarlopredictivecoding_b200 import predictive_coding as pc
from montecarlopredictivecoding_b200.predictive_coding import plan as P


@pytest.mark.parametrize("name", ALL_CASES)
def test_trainer_host_logic_matches_reference(name):
    replay(name, torch.device("cpu"), engine_factory=OracleEngine)


def test_stepwise_mode_with_opaque_callback():
    """An unrecognised callback forces one launch per step with x.grad materialised."""
    def wrap(fn):
        def opaque(t, _pc_trainer, var=2.0):      # not named random_step, not tagged
            return fn(t, _pc_trainer, var)
        return opaque
    # the opaque callback draws its own torch noise, so only structure can be compared: run a
    # noise-free case (MAP, Adam) through step-by-step mode by adding a no-op backward callback
    from golden_util import GoldenCase
    from trainer_replay import build_model, make_trainer
    gc = GoldenCase("pc_tanh_adam_mask")
    model = build_model(gc, torch.device("cpu"))
    trainer = make_trainer(model, gc.calls[0]["trainer"])
    trainer._engine = OracleEngine()
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    for l, layer in enumerate(pcs):
        layer._sample_x_fn = (lambda inputs, v=torch.from_numpy(gc.x0(0)[l]): v.clone())
    seen = []
    res = trainer.train_on_batch(
        inputs=torch.from_numpy(gc.inputs), loss_fn=mu.bernoulli_fn_mask,
        loss_fn_kwargs={"_target": torch.from_numpy(gc.target), "_var": 1.0},
        callback_after_backward=lambda t: seen.append(t), is_log_progress=False, is_return_xs=True)
    assert trainer.last_call_info["mode"] == "stepwise"
    assert seen == list(range(gc.calls[0]["trainer"]["T"]))
    T = gc.calls[0]["trainer"]["T"]
    for l in range(gc.L):
        got = np.stack([res["xs"][t][l].numpy() for t in range(T)])
        assert np.max(np.abs(got - gc.traj(0, l))) / np.max(np.abs(gc.traj(0, l))) < 1e-5
    assert np.allclose(res["energy"], gc.z["c0_energy"], rtol=1e-5)
    for i, lin in enumerate(m for m in model if isinstance(m, nn.Linear)):
        assert np.max(np.abs(lin.weight.detach().numpy() - gc.weights(0, "after")[0][i])) < 2e-6


def test_plan_compiler_recognises_get_model():
    cfg = {"input_size": 20, "hidden_size": 128, "hidden2_size": 128, "output_size": 784, "activation_fn": "relu"}
    model = mu.get_model(cfg, use_cuda=False)
    netp = P.compile_net(model)
    assert netp.dims == [20, 128, 128] and netp.d_out == 784 and netp.d_in == 20
    assert netp.act == [N.ACT_RELU] * 3
    assert netp.energy_scale == [1.0, 1.0, 1.0]


def test_plan_compiler_rejects_unknown_graphs():
    with pytest.raises(P.UnsupportedModel):
        P.compile_net(nn.Sequential(nn.Linear(2, 2), nn.Linear(2, 2), pc.PCLayer()))
    with pytest.raises(P.UnsupportedModel):
        P.compile_net(nn.Sequential(nn.Linear(2, 2), pc.PCLayer(), nn.Sigmoid()))
    with pytest.raises(P.UnsupportedModel):
        P.compile_net(nn.Sequential(nn.Linear(2, 2), pc.PCLayer(energy_fn=lambda i: (i["mu"] - i["x"]).abs())))


def test_loss_classification():
    B, d = 4, 10
    y = torch.rand(B, d)
    assert P.classify_loss(None, {}, B, d, None).kind == N.TOP_NONE
    assert P.classify_loss(mu.zero_fn, {}, B, d, None).kind == N.TOP_ZERO
    t = P.classify_loss(mu.fe_fn, {"_target": y, "_var": 0.25}, B, d, None)
    assert t.kind == N.TOP_GAUSS and abs(t.inv_var - 4.0) < 1e-9 and t.mask_start == 0
    t = P.classify_loss(mu.bernoulli_fn_mask, {"_target": y, "_var": None, "perc": 0.3}, B, d, None)
    assert t.kind == N.TOP_BERNOULLI and t.mask_start == d - 3
    t = P.classify_loss(mu.fe_fn_mask, {"_target": y, "_var": 2.0}, B, d, None)
    assert t.kind == N.TOP_GAUSS and t.mask_start == 5 and abs(t.inv_var - 0.5) < 1e-9
    with pytest.raises(P.UnsupportedModel):
        P.classify_loss(lambda o, _target: (o - _target).abs().sum(), {"_target": y}, B, d, None)


def test_callback_recognition():
    model = nn.Sequential(nn.Linear(1, 1), pc.PCLayer(), nn.Linear(1, 1))
    model.train()
    tr = pc.PCTrainer(model, T=4, plot_progress_at=[])
    assert P.classify_callback_after_t(mu.random_step, {"_pc_trainer": tr}, tr).var == 2.0
    assert P.classify_callback_after_t(mu.random_step, {"_pc_trainer": tr, "var": 0.5}, tr).var == 0.5
    assert P.classify_callback_after_t(lambda t: None, {}, tr) is None
    other = pc.PCTrainer(model, T=4, plot_progress_at=[])
    assert P.classify_callback_after_t(mu.random_step, {"_pc_trainer": other}, tr) is None


def test_trainer_api_surface_and_asserts():
    model = nn.Sequential(nn.Linear(3, 3), pc.PCLayer(), nn.Tanh(), nn.Linear(3, 5))
    tr = pc.PCTrainer(model, T=6, update_p_at="last", accumulate_p_at=[3, 4, 5], plot_progress_at=[])
    assert tr.get_T() == 6 and tr.get_num_pc_layers() == 1 and tr.get_least_T() == 2
    assert tr._update_p_at == [5] and tr._accumulate_p_at == [3, 4, 5] and tr._update_x_at == list(range(6))
    assert len(list(tr.get_model_parameters())) == 4
    assert tr.get_is_model_training() is None          # Sequential defaults to train, PCLayer starts in eval
    with pytest.raises(AssertionError):
        tr.train_on_batch(torch.zeros(2, 3))
    model.train()
    assert tr.get_is_model_training() is True
    with pytest.raises(NotImplementedError):
        pc.PCTrainer(model, T=2, update_x_at="sometimes")
    with pytest.raises(AssertionError):
        pc.PCTrainer(model, T=0)
    layer = pc.PCLayer()
    assert layer.training is False and layer.get_x() is None
    assert layer(torch.ones(2, 2)).equal(torch.ones(2, 2))     # eval mode passes mu through


def test_product_refuses_cpu_tensors():
    """No CPU fallback: the shipped engine must fail loudly for a CPU model."""
    model = nn.Sequential(nn.Linear(3, 3), pc.PCLayer(), nn.Linear(3, 5))
    model.train()
    tr = pc.PCTrainer(model, T=3, update_p_at="never", plot_progress_at=[])
    with pytest.raises((RuntimeError, OSError)):
        tr.train_on_batch(torch.zeros(2, 3), is_log_progress=False)


def test_fast_t0_sampling_draws_the_same_latents_as_the_forward():
    """pc_trainer.py:717-733: the t=0 forward samples every PCLayer in module order.  For the library samplers (they
    only use the shape of mu) the trainer draws the latents directly; the generator must be consumed identically."""
    import torch
    import torch.optim as optim

    from montecarlopredictivecoding_b200 import mcpc_utils as mu
    from montecarlopredictivecoding_b200 import predictive_coding as pc
    from montecarlopredictivecoding_b200.predictive_coding import plan as P
    cfg = {"input_size": 5, "hidden_size": 7, "hidden2_size": 6, "output_size": 9, "activation_fn": "relu"}
    for sampler in (mu.sample_x_fn, mu.sample_x_fn_normal, mu.sample_x_fn_cte):
        torch.manual_seed(3)
        model = mu.get_model(cfg, use_cuda=False, sample_x_fn=sampler)
        pcs = [m for m in model if isinstance(m, pc.PCLayer)]
        inputs = torch.zeros(4, 5)
        torch.manual_seed(11)
        for layer in pcs:
            layer.set_is_sample_x(True)
        with torch.no_grad():
            model(inputs)
        want = [layer.get_x().detach().clone() for layer in pcs]
        trainer = pc.PCTrainer(model, T=2, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.1}, update_p_at="never",
                               plot_progress_at=[])
        torch.manual_seed(11)
        trainer._start_of_batch(P.compile_net(model), inputs, True, True, False)
        for layer, w in zip(pcs, want):
            assert torch.equal(layer.get_x().detach(), w)
            assert layer.get_x().requires_grad


def test_optimizer_x_reset_in_place_equals_a_fresh_optimizer():
    """pc_trainer.py:749-752 re-creates optimizer_x at every batch start; when the latents are the same Parameters the
    drop-in clears the state and restores the defaults instead.  A changed lr must not survive, state must be empty,
    and re-sampled latents (new Parameters) must lead to a really new optimizer."""
    import torch
    import torch.optim as optim

    from montecarlopredictivecoding_b200 import mcpc_utils as mu
    from montecarlopredictivecoding_b200 import predictive_coding as pc
    cfg = {"input_size": 4, "hidden_size": 6, "hidden2_size": 5, "output_size": 7, "activation_fn": "tanh"}
    torch.manual_seed(0)
    model = mu.get_model(cfg, use_cuda=False)
    trainer = pc.PCTrainer(model, T=2, optimizer_x_fn=optim.Adam, optimizer_x_kwargs={"lr": 0.1}, update_p_at="never",
                           plot_progress_at=[])
    for layer in (m for m in model if isinstance(m, pc.PCLayer)):
        layer.set_is_sample_x(True)
    with torch.no_grad():
        model(torch.zeros(3, 4))
    trainer.recreate_optimize_x()
    opt = trainer.get_optimizer_x()
    for x in trainer.get_model_xs():
        x.grad = torch.ones_like(x)
    opt.step()
    opt.param_groups[0]["lr"] = 7.0
    assert len(opt.state) > 0
    trainer._reset_optimizer_x()
    assert trainer.get_optimizer_x() is opt and len(opt.state) == 0 and opt.param_groups[0]["lr"] == 0.1
    for layer in (m for m in model if isinstance(m, pc.PCLayer)):
        layer.set_is_sample_x(True)
    with torch.no_grad():
        model(torch.zeros(3, 4))                     # new Parameters
    trainer._reset_optimizer_x()
    assert trainer.get_optimizer_x() is not opt


def test_compiled_plan_cache_rechecks_live_layer_flags():
    import torch

    from montecarlopredictivecoding_b200 import mcpc_utils as mu
    from montecarlopredictivecoding_b200 import predictive_coding as pc
    from montecarlopredictivecoding_b200.predictive_coding import plan as P
    cfg = {"input_size": 4, "hidden_size": 6, "hidden2_size": 5, "output_size": 7, "activation_fn": "relu"}
    model = mu.get_model(cfg, use_cuda=False)
    a = P.compile_net(model)
    assert P.compile_net(model) is a                 # cached
    layer = [m for m in model if isinstance(m, pc.PCLayer)][0]
    layer.is_holding_error = True
    try:
        P.compile_net(model)
        raise AssertionError("held errors must be refused even on a cache hit")
    except NotImplementedError:
        pass
    layer.is_holding_error = False
    b = P.compile_net(model)
    assert b.dims == a.dims
    _ = torch


def test_forked_reference_functions_are_not_folded_by_name_alone():
    """a function that merely carries the reference's name/module must BEHAVE like it to be folded into the
    kernel (random_step) or called without a forward (samplers)."""
    import types
    mod = types.ModuleType("model")

    def random_step(t, _pc_trainer, var=2.0):           # same name, uniform instead of Gaussian noise
        opt = _pc_trainer.get_optimizer_x()
        for x in _pc_trainer.get_model_xs():
            x.grad.uniform_(-1.0, 1.0)
        opt.step()

    def genuine(t, _pc_trainer, var=2.0):               # the reference's body under the reference's name
        xs = _pc_trainer.get_model_xs()
        optimizer = _pc_trainer.get_optimizer_x()
        for x in xs:
            x.grad.normal_(0., np.sqrt(var / optimizer.defaults['lr']))
        optimizer.step()

    def sample_x_fn(inputs):                            # same name, but uses the VALUES of mu
        return inputs["mu"].detach().clone() + torch.randn_like(inputs["mu"])

    def sample_x_fn_normal(inputs):
        return torch.randn_like(inputs["mu"])

    for f in (random_step, genuine, sample_x_fn, sample_x_fn_normal):
        f.__module__ = "utils.model"
    genuine.__name__ = "random_step"
    trainer = object()
    assert P.classify_callback_after_t(random_step, {"_pc_trainer": trainer}, trainer) is None
    plan = P.classify_callback_after_t(genuine, {"_pc_trainer": trainer, "var": 1.5}, trainer)
    assert plan is not None and plan.var == 1.5
    assert P.classify_callback_after_t(mu.random_step, {"_pc_trainer": trainer}, trainer).var == 2.0    # tagged
    assert not P.sampler_is_shape_only(sample_x_fn)
    assert P.sampler_is_shape_only(sample_x_fn_normal)
    assert P.sampler_is_shape_only(mu.sample_x_fn)
    state = torch.random.get_rng_state()
    P.sampler_is_shape_only(lambda inputs: torch.randn_like(inputs["mu"]))
    assert torch.equal(state, torch.random.get_rng_state())
