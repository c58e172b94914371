"""Replay a golden case (recorded from the reference) through the drop-in PCTrainer and compare
every observable.  Used on CPU with the oracle test double and on the GPU with the real kernels."""
import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from golden_util import GoldenCase, rel_err
from montecarlopredictivecoding_b200 import mcpc_utils as mu
from montecarlopredictivecoding_b200 import predictive_coding as pc

ACTS = {"identity": None, "relu": nn.ReLU, "tanh": nn.Tanh}
LOSSES = {"none": None, "zero": mu.zero_fn, "gauss": mu.fe_fn, "gauss_mask": mu.fe_fn_mask,
          "bernoulli": mu.bernoulli_fn, "bernoulli_mask": mu.bernoulli_fn_mask}
OPTS = {"sgd": optim.SGD, "adam": optim.Adam}


def build_model(gc: GoldenCase, device):
    spec = gc.spec
    dims = spec["dims"]
    act = ACTS[spec["act"]]
    scales = spec.get("energy_scale", [1.0] * len(dims))
    bias = spec.get("bias", [True] * (len(dims) + 1))
    mods = []
    prev = spec.get("d_in", dims[0])
    for l, d in enumerate(dims):
        mods.append(nn.Linear(prev, d, bias=bias[l]))
        if scales[l] == 1.0:
            mods.append(pc.PCLayer())
        else:
            mods.append(pc.PCLayer(energy_fn=lambda inputs, c=scales[l]: c * 0.5 * (inputs["mu"] - inputs["x"]) ** 2))
        free_out = (l == len(dims) - 1) and gc.d_out == 0
        if act is not None and not free_out:
            mods.append(act())
        prev = d
    if gc.d_out > 0:
        mods.append(nn.Linear(prev, gc.d_out, bias=bias[len(dims)]))
    model = nn.Sequential(*mods)
    W, b = gc.weights(0, "before")
    lins = [m for m in model if isinstance(m, nn.Linear)]
    with torch.no_grad():
        for lin, w, bb in zip(lins, W, b):
            lin.weight.copy_(torch.from_numpy(w))
            if bb is not None:
                lin.bias.copy_(torch.from_numpy(bb))
    model.train()
    return model.to(device)


def make_trainer(model, tr):
    kw = dict(T=tr["T"], update_x_at=tr.get("update_x_at", "all"), optimizer_x_fn=OPTS[tr["opt_x"]],
              optimizer_x_kwargs={"lr": tr["lr_x"]}, update_p_at=tr.get("update_p_at", "never"), plot_progress_at=[],
              optimizer_p_fn=OPTS[tr.get("opt_p", "sgd")], optimizer_p_kwargs=tr.get("opt_p_kwargs", {"lr": 0.0}))
    if "accumulate_p_at" in tr:
        kw["accumulate_p_at"] = tr["accumulate_p_at"]
    if "energy_coefficient" in tr:
        kw["energy_coefficient"] = tr["energy_coefficient"]
    if "early_stop_condition" in tr:
        kw["early_stop_condition"] = tr["early_stop_condition"]
    return pc.PCTrainer(model, **kw)


def replay(name, device, engine_factory=None, precision="fp32", tol_x=1e-5, tol_s=1e-5, tol_g=2e-5, tol_w=2e-6,
           callback_wrapper=None, teacher_force=False, w_min_grad=0.0):
    """Returns a dict of the worst errors seen; asserts against the tolerances."""
    gc = GoldenCase(name)
    model = build_model(gc, device)
    lins = [m for m in model if isinstance(m, nn.Linear)]
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    trainers = []
    for c in gc.calls:
        trainers.append(trainers[c["trainer_of"]] if "trainer_of" in c else make_trainer(model, c["trainer"]))
    inputs_all = torch.from_numpy(gc.inputs).to(device)
    target_all = torch.from_numpy(gc.target).to(device)
    worst = {"x": 0.0, "scalar": 0.0, "grad": 0.0, "w": 0.0}
    for ci, call in enumerate(gc.calls):
        tr = call["trainer"]
        trainer = trainers[ci]
        if engine_factory is not None:
            trainer._engine = engine_factory()
        trainer.set_precision(precision)
        T = len(gc.z[f"c{ci}_energy"])        # < tr["T"] when the reference stopped early
        inputs, target = inputs_all[:gc.rows(ci)], target_all[:gc.rows(ci)]
        if teacher_force and ci > 0:
            # start every call from the reference's own state (latents and parameters) so that the bound
            # measures ONE call, not the drift accumulated over the preceding (chaotic, Adam) calls
            with torch.no_grad():
                for l, layer in enumerate(pcs):
                    if not call.get("sample_x", True):      # (a sampling call installs its own start state below)
                        layer.get_x().copy_(torch.from_numpy(gc.x0(ci)[l]).to(device))
                Wb, bb = gc.weights(ci, "before")
                for lin, w, b_ in zip(lins, Wb, bb):
                    lin.weight.copy_(torch.from_numpy(w).to(device))
                    if b_ is not None:
                        lin.bias.copy_(torch.from_numpy(b_).to(device))
        if call.get("sample_x", True):
            for l, layer in enumerate(pcs):
                x0 = torch.from_numpy(gc.x0(ci)[l]).to(device)
                layer._sample_x_fn = (lambda inputs, v=x0: v.clone())
        kwargs = dict(inputs=inputs, is_log_progress=False, is_return_results_every_t=True,
                      is_checking_after_callback_after_t=False, is_return_outputs=True, is_return_xs=True,
                      is_sample_x_at_batch_start=call.get("sample_x", True),
                      is_reset_optimizer_x_at_batch_start=call.get("reset_opt_x", True))
        lk = gc.loss_kind(ci)
        if LOSSES[lk] is not None:
            kwargs["loss_fn"] = LOSSES[lk]
            if lk != "zero":
                kwargs["loss_fn_kwargs"] = {"_target": target, "_var": gc.spec.get("var", 1.0)}
                if "perc" in gc.spec and "mask" in lk:
                    kwargs["loss_fn_kwargs"]["perc"] = gc.spec["perc"]
        if call.get("langevin", False):
            cb = mu.random_step if callback_wrapper is None else callback_wrapper(mu.random_step)
            kwargs["callback_after_t"] = cb
            kwargs["callback_after_t_kwargs"] = {"_pc_trainer": trainer}
            if "noise_var" in call:
                kwargs["callback_after_t_kwargs"]["var"] = call["noise_var"]
            nz = np.concatenate([gc.z[f"c{ci}_noise{l}"] for l in range(gc.L)], axis=2)   # [T, B, SD]
            trainer.set_supplied_noise(torch.from_numpy(nz).to(device))
        res = trainer.train_on_batch(**kwargs)

        for l in range(gc.L):
            got = np.stack([res["xs"][t][l].numpy() for t in range(T)])
            e = rel_err(got, gc.traj(ci, l))
            worst["x"] = max(worst["x"], e)
            assert e < tol_x, (name, ci, l, "traj", e)
            e = rel_err(pcs[l].get_x().detach().cpu().numpy(), gc.x_final(ci)[l])
            worst["x"] = max(worst["x"], e)
            assert e < tol_x, (name, ci, l, "final", e)
        outs = np.stack([o.detach().cpu().numpy() for o in res["outputs"]])
        e = rel_err(outs, gc.z[f"c{ci}_outputs"])
        worst["x"] = max(worst["x"], e)
        assert e < tol_x, (name, ci, "outputs", e)
        for key in ("energy", "overall", "loss"):
            ref = gc.z[f"c{ci}_{key}"]
            assert len(res[key]) == len(ref), (name, ci, key)
            if len(ref) and np.max(np.abs(ref)) > 0:
                e = rel_err(res[key], ref)
                worst["scalar"] = max(worst["scalar"], e)
                assert e < tol_s, (name, ci, key, e)
        upd_x, upd_p, acc = gc.step_lists(ci)
        if upd_p or trainer._keep_unused_param_grads:
            gW_ref, gb_ref = gc.grads(ci)
            for i, lin in enumerate(lins):
                for ref, p in ((gW_ref[i], lin.weight), (gb_ref[i], lin.bias)):
                    if ref is None or p is None:
                        continue
                    got = p.grad.detach().cpu().numpy()
                    e = float(np.max(np.abs(got - ref)) / max(float(np.max(np.abs(ref))), 1.0))
                    worst["grad"] = max(worst["grad"], e)
                    assert e < tol_g, (name, ci, i, "grad", e)
        W_after, b_after = gc.weights(ci, "after")
        adam_p = tr.get("opt_p", "sgd") == "adam" and bool(upd_p)
        for i, lin in enumerate(lins):
            dW = np.abs(lin.weight.detach().cpu().numpy() - W_after[i])
            if adam_p and w_min_grad > 0:
                # Adam's first step is lr*sign(g): entries whose gradient is ~0 flip with any rounding noise
                gref = gc.grads(ci)[0][i]
                if gref is not None:
                    dW = dW[np.abs(gref) >= w_min_grad * max(float(np.max(np.abs(gref))), 1e-30)]
            e = float(np.max(dW)) if dW.size else 0.0
            worst["w"] = max(worst["w"], e)
            assert e < tol_w, (name, ci, i, "W after", e)
            if b_after[i] is not None and not (adam_p and w_min_grad > 0):
                e = float(np.max(np.abs(lin.bias.detach().cpu().numpy() - b_after[i])))
                worst["w"] = max(worst["w"], e)
                assert e < tol_w, (name, ci, i, "b after", e)
    return worst
