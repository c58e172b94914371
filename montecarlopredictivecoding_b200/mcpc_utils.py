"""Host-side mirror of the reference's experiment utilities that sit on the hot path
(SURVEY §8 rows a4, a5, a8, a11, a12).  Same names, argument meaning and behaviour as

  * utils/model.py:8-15    sample_x_fn, sample_x_fn_normal, sample_x_fn_cte
  * utils/model.py:17-33   fe_fn, bernoulli_fn, fe_fn_mask, zero_fn, bernoulli_fn_mask
  * utils/model.py:35-44   random_step   (the Langevin noise callback)
  * utils/model.py:47-69   get_model
  * utils/training_evaluation.py:16-70   get_pc_trainer, get_mcpc_trainer, get_mcpc_trainer_one_sample
  * utils/model.py:71-163  get_representations   (trajectory consumer)

so a user of the reference finds the same vocabulary.  The reference's own ``utils`` package also
works unmodified on top of the drop-in ``predictive_coding`` (its ``random_step`` is recognised by
the trainer and folded into the kernel).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from .predictive_coding import PCLayer, PCTrainer


# ---- t=0 initialisers of the latents ---------------------------------------------------------
def sample_x_fn(inputs):
    return inputs["mu"].detach().clone().uniform_(-10.0, 10.0)


def sample_x_fn_normal(inputs):
    return torch.randn_like(inputs["mu"])


def sample_x_fn_cte(inputs):
    return 3 * torch.ones_like(inputs["mu"])


for _fn in (sample_x_fn, sample_x_fn_normal, sample_x_fn_cte):
    _fn.__mcpc_shape_only__ = True      # lets the trainer draw the t=0 latents without a model forward


# ---- sensory-layer losses ------------------------------------------------------------------------
def fe_fn(output, _target, _var):
    return (1 / _var) * 0.5 * (output - _target).pow(2).sum()


def bernoulli_fn(output, _target, _var=None, _reduction="sum"):
    return nn.functional.binary_cross_entropy_with_logits(output, _target, reduction=_reduction)


def _tail(t, perc):
    n = round(t.shape[1] * perc)
    return t[:, -n:]


def fe_fn_mask(output, _target, _var, perc=0.5):
    return (1 / _var) * 0.5 * (_tail(output, perc) - _tail(_target, perc)).pow(2).sum()


def bernoulli_fn_mask(output, _target, _var=None, perc=0.5):
    return nn.functional.binary_cross_entropy_with_logits(_tail(output, perc), _tail(_target, perc), reduction="sum")


def zero_fn(output):
    return torch.tensor(0.0)


# ---- Langevin noise ------------------------------------------------------------------------------
def random_step(t, _pc_trainer, var=2.0):
    """x <- x - lr * n,  n ~ N(0, var/lr0): with SGD this is the sqrt(var*lr)*xi term of the Langevin
    update (var=2 for mathematically correct sampling).  When passed as ``callback_after_t`` the trainer
    recognises it and draws the noise inside the fused kernel; the body below only runs in the
    step-by-step mode (where ``x.grad`` is materialised)."""
    optimizer = _pc_trainer.get_optimizer_x()
    std = float(np.sqrt(var / optimizer.defaults["lr"]))
    for x in _pc_trainer.get_model_xs():
        x.grad.normal_(0.0, std)
    optimizer.step()


random_step.__mcpc_langevin__ = True


# ---- model factory -------------------------------------------------------------------------------
def get_model(config, use_cuda, sample_x_fn=sample_x_fn):
    act = {"relu": nn.ReLU, "tanh": nn.Tanh}[config["activation_fn"]]
    widths = [config["input_size"], config["input_size"], config["hidden_size"], config["hidden2_size"]]
    mods = []
    for d_from, d_to in zip(widths[:-1], widths[1:]):
        mods += [nn.Linear(d_from, d_to), PCLayer(sample_x_fn=sample_x_fn), act()]
    mods.append(nn.Linear(widths[-1], config["output_size"]))
    gen_pc = nn.Sequential(*mods)
    gen_pc.train()
    if use_cuda:
        gen_pc.cuda()
    return gen_pc


# ---- trainer factories ---------------------------------------------------------------------------
def get_pc_trainer(gen_pc, config, is_mcpc=False, training=True):
    kw = dict(T=config["T_pc"], update_x_at="all", optimizer_x_fn=config["optimizer_x_fn_pc"],
              optimizer_x_kwargs=config["optimizer_x_kwargs_pc"], early_stop_condition="False", plot_progress_at=[])
    if is_mcpc:
        kw["update_p_at"] = "never"
    else:
        kw.update(update_p_at="last" if training else "never", optimizer_p_fn=config["optimizer_p_fn"],
                  optimizer_p_kwargs=config["optimizer_p_kwargs"])
    return PCTrainer(gen_pc, **kw)


def _mcpc_p_kwargs(config, training):
    if training:
        return dict(optimizer_p_fn=config["optimizer_p_fn_mcpc"], optimizer_p_kwargs=config["optimizer_p_kwargs_mcpc"])
    return dict(optimizer_p_fn=optim.SGD, optimizer_p_kwargs={"lr": 0.0})


def get_mcpc_trainer(gen_pc, config, training=True):
    return PCTrainer(
        gen_pc, T=config["mixing"] + config["sampling"], update_x_at="all", optimizer_x_fn=optim.SGD,
        optimizer_x_kwargs=config["optimizer_x_kwargs_mcpc"], update_p_at="last" if training else "never",
        accumulate_p_at=[config["mixing"] + i for i in range(config["sampling"])], plot_progress_at=[],
        **_mcpc_p_kwargs(config, training))


def get_mcpc_trainer_one_sample(gen_pc, config, training=True):
    return PCTrainer(
        gen_pc, T=config["K"], update_x_at="all", optimizer_x_fn=optim.SGD,
        optimizer_x_kwargs=config["optimizer_x_kwargs_mcpc"], update_p_at="last" if training else "never",
        plot_progress_at=[], **_mcpc_p_kwargs(config, training))


# ---- trajectory consumers ---------------------------------------------------------------------
def get_representations(gen_pc, config, trainers, loader, rep_type="MAP", use_cuda=False, n=None):
    """utils/model.py:71-163 with the same arguments and return value (a TensorDataset of first-PCLayer
    representations and labels), but the T-step trajectories never leave the GPU:

      * "MAP"          the latent of the first PCLayer after MAP inference (as the reference);
      * "expectation"  mean over ALL T steps of the Langevin chain (the reference's ``temp.mean(0)`` over per-step
                       ``.cpu()`` copies, :143-149) -- here ``enable_trajectory_stats`` folds a bounded device ring into
                       a running mean, nothing is returned per step;
      * "full"         the samples at steps mixing, mixing+indent, ... (the reference records every step and slices
                       ``temp[mixing::indent]``, :150-151) -- here only those steps are recorded
                       (``set_trajectory_stride``) into one device ring."""
    reps, labels = [], []
    input_size = len(gen_pc[0].bias)
    device = gen_pc[0].bias.device

    def map_call(pc_trainer, data, log):
        pseudo_input = torch.zeros(data.shape[0], input_size, device=device)
        pc_trainer.train_on_batch(inputs=pseudo_input, loss_fn=config["loss_fn"],
                                  loss_fn_kwargs={"_target": data, "_var": config["input_var"]}, is_log_progress=log,
                                  is_return_results_every_t=False, is_checking_after_callback_after_t=False)
        return pseudo_input

    if rep_type == "MAP":
        for data, label in loader:
            data, label = data.to(device), label.to(device)
            map_call(trainers[0], data, True)
            reps.append(gen_pc[1].get_x().detach().clone())
            labels.append(label)
    elif len(trainers) == 2:
        assert rep_type in ("full", "expectation")
        pc_trainer, mcpc_trainer = trainers
        indent = 1
        if n is not None:
            indent = int(config["sampling"] / n)
        else:
            n = config["sampling"]
        saved = (mcpc_trainer._traj_stride, mcpc_trainer._traj_start, mcpc_trainer._traj_on_device,
                 mcpc_trainer._traj_stats_cfg)
        try:
            for data, label in loader:
                data, label = data.to(device), label.to(device)
                pseudo_input = map_call(pc_trainer, data, False)
                kw = dict(inputs=pseudo_input, loss_fn=config["loss_fn"],
                          loss_fn_kwargs={"_target": data, "_var": config["input_var"]}, callback_after_t=random_step,
                          callback_after_t_kwargs={"_pc_trainer": mcpc_trainer}, is_log_progress=False,
                          is_return_results_every_t=True, is_checking_after_callback_after_t=False,
                          is_sample_x_at_batch_start=False)
                if rep_type == "expectation":
                    mcpc_trainer.enable_trajectory_stats(start=0, stride=1, layers=[0])
                    mcpc_trainer.train_on_batch(**kw)
                    reps.append(mcpc_trainer.trajectory_stats()["mean"][0].clone())
                    labels.append(label)
                else:
                    mcpc_trainer.disable_trajectory_stats()
                    mcpc_trainer.set_trajectory_stride(indent, start=config["mixing"])
                    mcpc_trainer.set_trajectories_on_device(True)
                    mcpc_trainer.train_on_batch(is_return_representations=True, **kw)
                    ring = mcpc_trainer.last_trajectories["x"][0]              # [n_rec, B, d] on the device
                    reps.append(ring.reshape(-1, ring.shape[2]).clone())
                    labels.append(label.repeat(ring.shape[0]))
        finally:
            (mcpc_trainer._traj_stride, mcpc_trainer._traj_start, mcpc_trainer._traj_on_device,
             mcpc_trainer._traj_stats_cfg) = saved
    else:
        raise NotImplementedError
    from torch.utils.data import TensorDataset
    if not reps:
        return TensorDataset(torch.tensor([]), torch.tensor([]).type(torch.int))
    return TensorDataset(torch.cat(reps, dim=0), torch.cat(labels, dim=0))


# ---- prior sampling and the marginal-likelihood estimate of table_1 --------------------------
def sample_pc(num_samples, model, config, use_cuda=False, is_return_hidden=False):
    """Ancestral samples of the generative model (utils/training_evaluation.py:72-100): every PCLayer adds unit
    Gaussian noise to its prediction; ``is_return_hidden`` returns the pre-sigmoid / pre-noise sensory prediction.
    Same arguments and return value; the noise is drawn on the model's device instead of on the CPU."""
    device = next(model.parameters()).device
    temp = torch.zeros((num_samples, config["input_size"]), device=device)
    with torch.no_grad():
        for layer in model:
            if isinstance(layer, PCLayer):
                temp = temp + torch.randn_like(temp)           # N(mu, I): cholesky(eye) of the reference is the identity
            else:
                temp = layer(temp)
    if is_return_hidden:
        return temp.detach()
    loss_fn = config["loss_fn"]
    if getattr(loss_fn, "__name__", "") == "fe_fn":
        temp = temp + float(config["input_var"]) ** 0.5 * torch.randn_like(temp)
    elif getattr(loss_fn, "__name__", "") == "bernoulli_fn":
        temp = (torch.rand_like(temp) <= temp.sigmoid()).double()
    return temp.detach()


_mll_ws = {}


def bernoulli_marginal_ll(logits, data, clamp_abs=20.0, return_rows=False):
    """``mean_i log mean_s exp(-sum_j BCEWithLogits(clamp(logits[s]), data[i]))`` on the GPU through
    ``mcpc_marginal_ll_bernoulli`` (one tcgen05 GEMM with a streaming min / sum-exp epilogue).
    logits [S, D], data [N, D]: CUDA tensors; returns a 0-d float32 CPU tensor like the reference's ``ml``."""
    import ctypes as C

    from . import _native as N
    if not (logits.is_cuda and data.is_cuda):
        raise RuntimeError("bernoulli_marginal_ll needs CUDA tensors: the B200 build has no CPU path")
    logits = logits.detach().to(torch.float32).contiguous()
    data = data.detach().to(device=logits.device, dtype=torch.float32).contiguous()
    if logits.dim() != 2 or data.dim() != 2 or logits.shape[1] != data.shape[1]:
        raise RuntimeError(f"logits {tuple(logits.shape)} and data {tuple(data.shape)} must be [S, D] and [N, D]")
    S, D = logits.shape
    n = data.shape[0]
    lib = N.load()
    need = C.c_size_t(0)
    N.check(lib.mcpc_marginal_ll_workspace_bytes(n, S, D, C.byref(need)), "mcpc_marginal_ll_workspace_bytes")
    ws = _mll_ws.get(logits.device.index)
    if ws is None or ws.numel() < need.value:
        ws = torch.empty(need.value, dtype=torch.uint8, device=logits.device)
        _mll_ws[logits.device.index] = ws
    ml = torch.zeros(1, dtype=torch.float64, device=logits.device)
    rows = torch.empty(n, dtype=torch.float32, device=logits.device) if return_rows else None
    stream = torch.cuda.current_stream(logits.device).cuda_stream
    with torch.cuda.device(logits.device):
        N.check(lib.mcpc_marginal_ll_bernoulli(logits.data_ptr(), S, data.data_ptr(), n, D, float(clamp_abs), ws.data_ptr(),
                                               ws.numel(), ml.data_ptr(), None if rows is None else rows.data_ptr(),
                                               C.c_void_p(stream)), "mcpc_marginal_ll_bernoulli")
    out = ml.to("cpu", torch.float32).reshape(())
    return (out, rows) if return_rows else out


def get_marginal_likelihood(gen_pc, config, dataloader, use_cuda, n_samples=5000):
    """utils/training_evaluation.py:177-206 for Bernoulli models: S prior samples, every data row scored against all
    of them.  Same arguments and return value; the 4000 x 5000 x 784 elementwise BCE of the reference is one GEMM."""
    if getattr(config["loss_fn"], "__name__", "") != "bernoulli_fn":
        raise NotImplementedError("only the Bernoulli likelihood is implemented (the reference raises for fe_fn too)")
    logits = sample_pc(n_samples, gen_pc, config, use_cuda=use_cuda, is_return_hidden=True)
    dataset = dataloader.dataset
    loader = torch.utils.data.DataLoader(dataset, batch_size=4096)
    data = torch.cat([d.reshape(d.shape[0], -1) for d, _ in loader], dim=0)
    return bernoulli_marginal_ll(logits, data.to(logits.device), clamp_abs=20.0)
