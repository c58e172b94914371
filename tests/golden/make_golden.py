"""Generate golden vectors by running the REAL reference implementation (CPU, fp32).

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.  The resulting ``tests/golden/*.npz`` files
are committed; tests never import the reference.

What is recorded per case: weights, inputs, target, the Langevin noise the reference drew
(recorded by a callback that is line-for-line ``random_step`` plus a clone),
per-step latents / outputs / energy / loss / overall, final latents, the parameter ``.grad``
left behind by the call and the parameters after ``optimizer_p.step()``.
"""
import json
import os
import sys
import types
import warnings

import numpy as np

REF = os.environ.get("MCPC_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

for _n in ("matplotlib", "matplotlib.pyplot", "seaborn"):
    sys.modules.setdefault(_n, types.ModuleType(_n))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, REF)
warnings.simplefilter("ignore")

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.optim as optim  # noqa: E402

import predictive_coding as pc  # noqa: E402  (the reference package)
from utils import model as ref_model  # noqa: E402

assert os.path.realpath(pc.__file__).startswith(os.path.realpath(REF)), pc.__file__

ACTS = {"identity": None, "relu": nn.ReLU, "tanh": nn.Tanh}
LOSSES = {
    "none": None,
    "zero": ref_model.zero_fn,
    "gauss": ref_model.fe_fn,
    "gauss_mask": ref_model.fe_fn_mask,
    "bernoulli": ref_model.bernoulli_fn,
    "bernoulli_mask": ref_model.bernoulli_fn_mask,
}
SAMPLERS = {
    "uniform": ref_model.sample_x_fn,
    "normal": ref_model.sample_x_fn_normal,
    "cte": ref_model.sample_x_fn_cte,
}


def build_model(spec):
    """[Linear, PCLayer, act?]* [Linear]? as in utils/model.py:54-65 / figure_2.py:40-44 / figure_3.py:50-55."""
    dims = spec["dims"]
    d_in = spec.get("d_in", dims[0])
    act = ACTS[spec["act"]]
    sampler = SAMPLERS[spec.get("sampler", "uniform")]
    mods = []
    prev = d_in
    scales = spec.get("energy_scale", [1.0] * len(dims))
    for l, d in enumerate(dims):
        mods.append(nn.Linear(prev, d, bias=spec.get("bias", [True] * (len(dims) + 1))[l]))
        if scales[l] == 1.0:
            mods.append(pc.PCLayer(sample_x_fn=sampler))
        else:
            c = scales[l]
            mods.append(pc.PCLayer(energy_fn=lambda inputs, c=c: c * 0.5 * (inputs["mu"] - inputs["x"]) ** 2,
                                   sample_x_fn=sampler))
        last_is_free_out = (l == len(dims) - 1) and spec.get("d_out", 0) == 0
        if act is not None and not last_is_free_out:
            mods.append(act())
        prev = d
    if spec.get("d_out", 0) > 0:
        mods.append(nn.Linear(prev, spec["d_out"], bias=spec.get("bias", [True] * (len(dims) + 1))[len(dims)]))
    m = nn.Sequential(*mods)
    m.train()
    return m


def linears(model):
    return [m for m in model if isinstance(m, nn.Linear)]


def pclayers(model):
    return [m for m in model if isinstance(m, pc.PCLayer)]


def run_call(model, trainer, spec, call, inputs, target, out, prefix):
    """One train_on_batch call; everything observable goes into ``out`` under ``prefix``."""
    noise_store = []

    def rec_random_step(t, _pc_trainer, var=2.0):
        xs = _pc_trainer.get_model_xs()
        optimizer = _pc_trainer.get_optimizer_x()
        row = []
        for x in xs:
            x.grad.normal_(0.0, np.sqrt(var / optimizer.defaults["lr"]))
            row.append(x.grad.clone())
        optimizer.step()
        noise_store.append(row)

    kwargs = dict(inputs=inputs, is_log_progress=False, is_return_results_every_t=True,
                  is_checking_after_callback_after_t=False, is_return_outputs=True, is_return_xs=True,
                  is_sample_x_at_batch_start=call.get("sample_x", True),
                  is_reset_optimizer_x_at_batch_start=call.get("reset_opt_x", True))
    loss_fn = LOSSES[call.get("loss", spec.get("loss", "none"))]
    if loss_fn is not None:
        kwargs["loss_fn"] = loss_fn
        if loss_fn is not ref_model.zero_fn:
            kwargs["loss_fn_kwargs"] = {"_target": target, "_var": spec.get("var", 1.0)}
            if "perc" in spec and "mask" in call.get("loss", spec.get("loss", "")):
                kwargs["loss_fn_kwargs"]["perc"] = spec["perc"]
    if call.get("langevin", False):
        kwargs["callback_after_t"] = rec_random_step
        kwargs["callback_after_t_kwargs"] = {"_pc_trainer": trainer}
        if "noise_var" in call:
            kwargs["callback_after_t_kwargs"]["var"] = call["noise_var"]
    for lin_i, lin in enumerate(linears(model)):
        if prefix != "c0_":
            break   # later calls start from the previous call's *_after
        out[f"{prefix}W{lin_i}_before"] = lin.weight.detach().numpy().copy()
        if lin.bias is not None:
            out[f"{prefix}b{lin_i}_before"] = lin.bias.detach().numpy().copy()
    res = trainer.train_on_batch(**kwargs)
    T = len(res["energy"])          # < trainer.get_T() when early_stop_condition fired (pc_trainer.py:979-981)
    L = len(pclayers(model))
    out[f"{prefix}energy"] = np.array(res["energy"], dtype=np.float64)
    out[f"{prefix}loss"] = np.array(res["loss"], dtype=np.float64)
    out[f"{prefix}overall"] = np.array(res["overall"], dtype=np.float64)
    for l in range(L):
        out[f"{prefix}traj_x{l}"] = np.stack([res["xs"][t][l].numpy() for t in range(T)])
        out[f"{prefix}x{l}_final"] = pclayers(model)[l].get_x().detach().numpy().copy()
        if noise_store:
            out[f"{prefix}noise{l}"] = np.stack([noise_store[t][l].numpy() for t in range(T)])
    out[f"{prefix}outputs"] = np.stack([res["outputs"][t].detach().numpy() for t in range(T)])
    for lin_i, lin in enumerate(linears(model)):
        if lin.weight.grad is not None:
            out[f"{prefix}gW{lin_i}"] = lin.weight.grad.detach().numpy().copy()
        if lin.bias is not None and lin.bias.grad is not None:
            out[f"{prefix}gb{lin_i}"] = lin.bias.grad.detach().numpy().copy()
        out[f"{prefix}W{lin_i}_after"] = lin.weight.detach().numpy().copy()
        if lin.bias is not None:
            out[f"{prefix}b{lin_i}_after"] = lin.bias.detach().numpy().copy()


OPTS = {"sgd": optim.SGD, "adam": optim.Adam}


def make_trainer(model, tr):
    kw = dict(T=tr["T"], update_x_at=tr.get("update_x_at", "all"),
              optimizer_x_fn=OPTS[tr["opt_x"]], optimizer_x_kwargs={"lr": tr["lr_x"]},
              update_p_at=tr.get("update_p_at", "never"), plot_progress_at=[],
              optimizer_p_fn=OPTS[tr.get("opt_p", "sgd")], optimizer_p_kwargs=tr.get("opt_p_kwargs", {"lr": 0.0}))
    if "accumulate_p_at" in tr:
        kw["accumulate_p_at"] = tr["accumulate_p_at"]
    if "energy_coefficient" in tr:
        kw["energy_coefficient"] = tr["energy_coefficient"]
    if "early_stop_condition" in tr:
        kw["early_stop_condition"] = tr["early_stop_condition"]
    return pc.PCTrainer(model, **kw)


def run_case(name, spec):
    torch.manual_seed(spec.get("seed", 30))
    np.random.seed(2)
    model = build_model(spec)
    if "checkpoint" in spec:
        sd = torch.load(os.path.join(REF, "models", spec["checkpoint"]), map_location="cpu", weights_only=True)
        model.load_state_dict({k: v for k, v in sd.items() if "_x" not in k}, strict=False)
    if "init_const" in spec:
        for lin, (wv, bv) in zip(linears(model), spec["init_const"]):
            if wv is not None:
                nn.init.constant_(lin.weight, wv)
            if bv is not None and lin.bias is not None:
                nn.init.constant_(lin.bias, bv)
    B = spec["B"]
    d_in = spec.get("d_in", spec["dims"][0])
    inputs = torch.zeros(B, d_in) if spec.get("zero_inputs", True) else torch.randn(B, d_in)
    d_t = spec["d_out"] if spec.get("d_out", 0) > 0 else spec["dims"][-1]
    if spec.get("target", "binary") == "binary":
        target = (torch.rand(B, d_t) < 0.5).float()
    elif spec["target"] == "ones":
        target = torch.ones(B, d_t)
    else:
        target = torch.randn(B, d_t)
    out = {"inputs": inputs.numpy().copy(), "target": target.numpy().copy()}
    # trainers are created up-front like the scripts do (figure_2.py:67-69)
    trainers = []
    for c in spec["calls"]:
        trainers.append(trainers[c["trainer_of"]] if "trainer_of" in c else make_trainer(model, c["trainer"]))
    for ci, call in enumerate(spec["calls"]):
        rows = call.get("rows", B)      # a call may use only the first `rows` chains (batch-size change)
        run_call(model, trainers[ci], spec, call, inputs[:rows], target[:rows], out, f"c{ci}_")
    out["spec_json"] = np.frombuffer(json.dumps(spec).encode(), dtype=np.uint8)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, keys={len(out)}")


CASES = {
    # MCPC learning step on a small relu net with Bernoulli top: mixing 3 + sampling 4, dW window, SGD p-step
    "mcpc_relu_bce_learn": dict(
        dims=[6, 16, 12], d_out=24, act="relu", loss="bernoulli", B=8, sampler="uniform",
        calls=[
            dict(trainer=dict(T=12, opt_x="adam", lr_x=0.1), sample_x=True),
            dict(trainer=dict(T=7, opt_x="sgd", lr_x=0.03, update_p_at="last",
                              accumulate_p_at=[3, 4, 5, 6], opt_p="sgd", opt_p_kwargs={"lr": 0.1}),
                 langevin=True, sample_x=False),
        ]),
    # same pattern, tanh + Adam p-optimizer, non-zero inputs (exercises the Linear_0 weight gradient)
    "mcpc_tanh_bce_learn_inputs": dict(
        dims=[6, 16, 12], d_out=24, act="tanh", loss="bernoulli", B=8, sampler="normal", zero_inputs=False,
        calls=[
            dict(trainer=dict(T=9, opt_x="sgd", lr_x=0.05, update_p_at="last",
                              accumulate_p_at=[4, 5, 6, 7, 8], opt_p="adam", opt_p_kwargs={"lr": 0.01}),
                 langevin=True, sample_x=True),
        ]),
    # shipped MNIST checkpoint, the mcpc_ml shape (table_1.py:195-212), MAP warm-up then Langevin
    "mcpc_ml_checkpoint": dict(
        dims=[20, 128, 128], d_out=784, act="relu", loss="bernoulli", B=16, sampler="uniform",
        checkpoint="mcpc_ml_1",
        calls=[
            dict(trainer=dict(T=10, opt_x="adam", lr_x=0.1), sample_x=True),
            dict(trainer=dict(T=12, opt_x="sgd", lr_x=0.03, update_p_at="last",
                              accumulate_p_at=[4, 5, 6, 7, 8, 9, 10, 11], opt_p="adam", opt_p_kwargs={"lr": 0.01}),
                 langevin=True, sample_x=False),
        ]),
    # deterministic PC path on the shipped pc_ml_1 checkpoint (table_1.py:214-225): [25,128,128] tanh, Adam lr 0.3,
    # masked BCE as get_mse_rec runs it (utils/training_evaluation.py:143-174); horizon kept below the chaotic
    # divergence of Adam trajectories
    "pc_ml_checkpoint_map": dict(
        dims=[25, 128, 128], d_out=784, act="tanh", loss="bernoulli_mask", B=16, sampler="uniform",
        checkpoint="pc_ml_1",
        calls=[
            dict(trainer=dict(T=40, opt_x="adam", lr_x=0.3), sample_x=True),
        ]),
    # deterministic PC path: tanh, Adam on x, masked BCE, PC training with p-step at last (table_1.py:214-225)
    "pc_tanh_adam_mask": dict(
        dims=[5, 16, 16], d_out=24, act="tanh", loss="bernoulli_mask", B=8, sampler="uniform",
        calls=[
            dict(trainer=dict(T=30, opt_x="adam", lr_x=0.3, update_p_at="last", opt_p="adam",
                              opt_p_kwargs={"lr": 0.01}), sample_x=True),
        ]),
    # figure_2.py:29-75 linear-Gaussian model: MAP (Adam) then Langevin, B=1
    "fig2_linear": dict(
        dims=[1], d_out=1, act="identity", loss="gauss", var=1.0, B=1, sampler="cte", target="ones",
        bias=[True, False], init_const=[(None, 0.2), (2.0, None)],
        calls=[
            dict(trainer=dict(T=60, opt_x="adam", lr_x=0.02), sample_x=True),
            dict(trainer=dict(T=50, opt_x="sgd", lr_x=0.02, accumulate_p_at=list(range(50))),
                 langevin=True, sample_x=True),
        ]),
    # figure_3.py:47-73 free output PCLayer with scaled energy, no loss, var kwarg of random_step
    "free_output_layer": dict(
        dims=[3, 5], d_out=0, act="tanh", loss="none", B=4, sampler="uniform", bias=[True, False],
        energy_scale=[1.0, 2.0],
        calls=[
            dict(trainer=dict(T=8, opt_x="adam", lr_x=0.5), sample_x=True),
            dict(trainer=dict(T=10, opt_x="sgd", lr_x=0.1), langevin=True, sample_x=False, noise_var=1.3),
        ]),
    # Gaussian top with mask and variance (utils/model.py:24-25), one-sample learning (/B normalisation)
    "gauss_mask_one_sample": dict(
        dims=[4, 8], d_out=10, act="tanh", loss="gauss_mask", var=0.5, perc=0.5, B=6, sampler="normal", target="normal",
        calls=[
            dict(trainer=dict(T=9, opt_x="sgd", lr_x=0.02, update_p_at="last", opt_p="sgd",
                              opt_p_kwargs={"lr": 0.07, "momentum": 0.2}), langevin=True, sample_x=True),
        ]),
    # zero_fn: sensory Linear is readout-only (figure_3.py:153-161)
    "zero_fn_sampling": dict(
        dims=[4, 8, 8], d_out=12, act="relu", loss="zero", B=5, sampler="uniform",
        calls=[
            dict(trainer=dict(T=6, opt_x="adam", lr_x=0.7), sample_x=True, loss="none"),
            dict(trainer=dict(T=9, opt_x="sgd", lr_x=0.1), langevin=True, sample_x=False),
        ]),
    # update_p_at='all' with an energy coefficient: weights move inside the T loop
    "update_p_all": dict(
        dims=[4, 6], d_out=8, act="tanh", loss="gauss", var=1.0, B=4, sampler="normal", target="normal",
        calls=[
            dict(trainer=dict(T=5, opt_x="sgd", lr_x=0.05, update_p_at="all", opt_p="sgd",
                              opt_p_kwargs={"lr": 0.02}, energy_coefficient=0.5), langevin=True, sample_x=True),
        ]),
    # schedules given as 'last_half' (pc_trainer.py:1094-1095): x frozen for the first half, p-steps inside the loop
    "last_half_schedules": dict(
        dims=[4, 10], d_out=8, act="tanh", loss="gauss", var=1.0, B=6, sampler="normal", target="normal",
        calls=[
            dict(trainer=dict(T=8, opt_x="sgd", lr_x=0.05, update_x_at="last_half", update_p_at="last_half",
                              opt_p="sgd", opt_p_kwargs={"lr": 0.02})),
        ]),
    # one trainer used for consecutive calls: Adam-on-x state carried over (is_reset_optimizer_x_at_batch_start=False),
    # then a smaller batch on the same trainer (latents re-sampled, optimizer_x re-created, pc_trainer.py:742-752)
    "adam_carryover_batch_resize": dict(
        dims=[5, 12], d_out=10, act="relu", loss="bernoulli", B=8, sampler="uniform",
        calls=[
            dict(trainer=dict(T=6, opt_x="adam", lr_x=0.1), sample_x=True),
            dict(trainer=dict(T=6, opt_x="adam", lr_x=0.1), trainer_of=0, sample_x=False, reset_opt_x=False),
            dict(trainer=dict(T=6, opt_x="adam", lr_x=0.1), trainer_of=0, sample_x=True, rows=5),
        ]),
    # early stop (pc_trainer.py:844-859, 904-914, 979-981): the p-step of the stopping step uses ONLY that step's gradient
    # (zero_grad fires because the step is outside accumulate_p_at); the second call checks nothing is left over
    "early_stop_p_update": dict(
        dims=[4, 8], d_out=6, act="tanh", loss="gauss", var=1.0, B=5, sampler="normal", target="normal",
        calls=[
            dict(trainer=dict(T=6, opt_x="sgd", lr_x=0.05, update_p_at="last", opt_p="sgd", opt_p_kwargs={"lr": 0.05},
                              early_stop_condition="t == 3"), sample_x=True),
            dict(trainer=dict(T=6, opt_x="sgd", lr_x=0.05, update_p_at="last", opt_p="sgd", opt_p_kwargs={"lr": 0.05},
                              early_stop_condition="t == 3"), trainer_of=0, sample_x=True),
        ]),
}


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        run_case(n, CASES[n])
