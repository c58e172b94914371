"""SURVEY §8(f) N2 host logic on CPU (kernels replaced by the oracle test double): thinned trajectory recording
(set_trajectory_stride) and the chunked on-device statistics (enable_trajectory_stats) against plain slicing /
torch reductions of the reference-style every-step trajectory; the get_representations mirror against the reference's
formulas (utils/model.py:143-151)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.optim as optim

from oracle_engine import OracleEngine

from montecarlopredictivecoding_b200 import mcpc_utils as mu
from montecarlopredictivecoding_b200 import predictive_coding as pc


def _setup(T=23, B=5, update_p_at="never", seed=0):
    torch.manual_seed(seed)
    cfg = {"input_size": 4, "hidden_size": 8, "hidden2_size": 6, "output_size": 10, "activation_fn": "tanh"}
    model = mu.get_model(cfg, use_cuda=False, sample_x_fn=mu.sample_x_fn_normal)
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.05}, update_p_at=update_p_at,
                      optimizer_p_fn=optim.SGD, optimizer_p_kwargs={"lr": 0.01}, plot_progress_at=[])
    tr._engine = OracleEngine()
    y = torch.randn(B, 10)
    return model, tr, y


def _call(model, tr, y, **kw):
    torch.manual_seed(7)                    # same t=0 latents in every call
    tr.set_noise_seed(99)                   # same Langevin noise in every call
    return tr.train_on_batch(torch.zeros(y.shape[0], 4), loss_fn=mu.fe_fn, loss_fn_kwargs={"_target": y, "_var": 1.0},
                             callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr},
                             is_log_progress=False, is_checking_after_callback_after_t=False, **kw)


@pytest.mark.parametrize("stride,start,update_p_at", [(1, 0, "never"), (3, 0, "never"), (4, 5, "never"), (2, 3, "all"),
                                                        (5, 22, "never"), (7, 30, "never")])
def test_thinned_recording_equals_slices_of_the_full_trajectory(stride, start, update_p_at):
    T = 23
    model, tr, y = _setup(T=T, update_p_at=update_p_at)
    lin_params = [q for m in model if isinstance(m, nn.Linear) for q in (m.weight, m.bias)]   # not the latents (F7)
    w0 = [q.detach().clone() for q in lin_params]
    full = _call(model, tr, y, is_return_xs=True, is_return_outputs=True, is_return_representations=True)
    with torch.no_grad():
        for q, w in zip(lin_params, w0):
            q.copy_(w)
    tr.set_trajectory_stride(stride, start)
    thin = _call(model, tr, y, is_return_xs=True, is_return_outputs=True, is_return_representations=True)
    steps = list(range(start, T, stride))
    assert len(thin["xs"]) == len(steps) == len(thin["outputs"]) == len(thin["representations"])
    assert len(thin["energy"]) == T                      # scalars stay per step
    for r, t in enumerate(steps):
        for l in range(3):
            assert torch.equal(thin["xs"][r][l], full["xs"][t][l])
        assert torch.equal(thin["outputs"][r], full["outputs"][t])
        assert torch.equal(thin["representations"][r], full["representations"][t])
    assert tr.last_trajectories["steps"] == steps


@pytest.mark.parametrize("start,stride,ring_bytes", [(0, 1, 1 << 30), (6, 1, 400), (4, 3, 1), (0, 2, 4 * 5 * 8 * 3)])
def test_trajectory_stats_equal_torch_reductions(start, stride, ring_bytes):
    T = 29
    model, tr, y = _setup(T=T)
    full = _call(model, tr, y, is_return_xs=True)
    ref = [torch.stack([full["xs"][t][l] for t in range(start, T, stride)]) for l in range(3)]
    tr._traj_ring_bytes = ring_bytes        # tiny rings force the chunked fold (bounded memory for T >= 1e4)
    tr.enable_trajectory_stats(start=start, stride=stride)
    res = _call(model, tr, y)               # nothing recorded for the caller
    assert "xs" not in res
    st = tr.trajectory_stats()
    assert st["count"] == ref[0].shape[0]
    for l in range(3):
        assert torch.allclose(st["mean"][l], ref[l].mean(0), rtol=1e-5, atol=1e-6)
        assert torch.allclose(st["var"][l], ref[l].var(0), rtol=1e-4, atol=1e-6)
    # the same statistics when the trajectory IS returned with the same thinning (one ring serves both)
    tr.set_trajectory_stride(stride, start)
    res = _call(model, tr, y, is_return_xs=True)
    st2 = tr.trajectory_stats()
    for l in range(3):
        assert torch.allclose(st2["mean"][l], st["mean"][l], rtol=1e-5, atol=1e-6)
    tr.set_trajectory_stride(stride + 1, start)
    with pytest.raises(ValueError):
        _call(model, tr, y, is_return_xs=True)


def test_get_representations_mirror_matches_the_reference_formulas():
    T_map, mixing, sampling, B = 6, 4, 12, 5
    torch.manual_seed(1)
    cfg = {"input_size": 4, "hidden_size": 8, "hidden2_size": 6, "output_size": 10, "activation_fn": "tanh",
           "loss_fn": mu.fe_fn, "input_var": 1.0, "T_pc": T_map, "optimizer_x_fn_pc": optim.Adam,
           "optimizer_x_kwargs_pc": {"lr": 0.1}, "mixing": mixing, "sampling": sampling,
           "optimizer_x_kwargs_mcpc": {"lr": 0.05}}
    model = mu.get_model(cfg, use_cuda=False, sample_x_fn=mu.sample_x_fn_normal)
    pc_tr = mu.get_pc_trainer(model, cfg, is_mcpc=True)
    mc_tr = mu.get_mcpc_trainer(model, cfg, training=False)
    for t in (pc_tr, mc_tr):
        t._engine = OracleEngine()
    data = torch.randn(2 * B, 10)
    labels = torch.arange(2 * B)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(data, labels), batch_size=B)

    def reference_style(rep_type, n):
        """What utils/model.py:105-151 computes, from every-step host copies."""
        reps, labs = [], []
        indent = 1 if n is None else int(sampling / n)
        for d, lab in loader:
            torch.manual_seed(11)
            mc_tr.set_noise_seed(5)
            pc_tr.train_on_batch(torch.zeros(B, 4), loss_fn=mu.fe_fn, loss_fn_kwargs={"_target": d, "_var": 1.0},
                                 is_log_progress=False, is_return_results_every_t=False)
            r = mc_tr.train_on_batch(torch.zeros(B, 4), loss_fn=mu.fe_fn, loss_fn_kwargs={"_target": d, "_var": 1.0},
                                     callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": mc_tr},
                                     is_log_progress=False, is_sample_x_at_batch_start=False,
                                     is_checking_after_callback_after_t=False, is_return_representations=True)
            temp = torch.stack(r["representations"])
            if rep_type == "expectation":
                reps.append(temp.mean(0))
                labs.append(lab)
            else:
                reps.append(temp[mixing::indent].reshape(-1, temp.shape[2]))
                labs.append(lab.repeat(n if n is not None else sampling))
        return torch.cat(reps), torch.cat(labs)

    class SeededLoader:                       # re-seed before every batch like reference_style does
        def __iter__(self):
            for d, lab in loader:
                torch.manual_seed(11)
                mc_tr.set_noise_seed(5)
                yield d, lab

    for rep_type, n in (("expectation", None), ("full", None), ("full", 4)):
        want_r, want_l = reference_style(rep_type, n)
        ds = mu.get_representations(model, cfg, [pc_tr, mc_tr], SeededLoader(), rep_type=rep_type, n=n)
        got_r, got_l = ds.tensors
        assert got_r.shape == want_r.shape, (rep_type, n, got_r.shape, want_r.shape)
        assert torch.allclose(got_r, want_r, rtol=1e-5, atol=1e-6), (rep_type, n)
        assert torch.equal(got_l, want_l)
    torch.manual_seed(11)
    ds = mu.get_representations(model, cfg, [pc_tr], loader, rep_type="MAP")
    assert ds.tensors[0].shape == (2 * B, 4)
