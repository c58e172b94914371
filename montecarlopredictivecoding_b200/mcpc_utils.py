"""Host-side mirror of the reference's experiment utilities that sit on the hot path
(SURVEY §8 rows a4, a5, a8, a11, a12).  Same names, argument meaning and behaviour as

  * utils/model.py:8-15    sample_x_fn, sample_x_fn_normal, sample_x_fn_cte
  * utils/model.py:17-33   fe_fn, bernoulli_fn, fe_fn_mask, zero_fn, bernoulli_fn_mask
  * utils/model.py:35-44   random_step   (the Langevin noise callback)
  * utils/model.py:47-69   get_model
  * utils/training_evaluation.py:16-70   get_pc_trainer, get_mcpc_trainer, get_mcpc_trainer_one_sample

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
