"""Golden STATISTICS recorded from the real reference with its OWN generated noise (BASELINE.json north star: "with
generated noise, posterior means and variances, energy trajectories and table_1 MSE ... must match within stated
statistical tolerance").  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_stats.py

  stat_langevin_mcpc_ml.npz   shipped checkpoint models/mcpc_ml_1 (20-128-128->784 relu, Bernoulli), 64 chains, MAP warm-up
                              (Adam lr 0.1, T=250) then R=32 independent Langevin runs of T=250 (SGD lr 0.03, var 2, stock
                              utils.model.random_step / torch RNG) from the SAME start state: per-run energy / loss
                              trajectories, pooled posterior mean / variance of every first-layer latent over the last 100
                              steps.
  stat_mse_rec.npz            table_1.py get_models_mse -> utils/training_evaluation.py:143-174 get_mse_rec on 128 images
                              sampled from the model itself (sample_pc), checkpoints pc_mse_1 (tanh, [30,256,256]) and
                              mcpc_mse_1 (relu, [10,256,256]): MAP (Adam lr 0.7, T=250, masked BCE) and the MSE on the
                              masked-out half.
"""
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
import torch.optim as optim  # noqa: E402

import predictive_coding as pc  # noqa: E402
from utils import model as rm  # noqa: E402
from utils import training_evaluation as te  # noqa: E402

assert os.path.realpath(pc.__file__).startswith(os.path.realpath(REF))


def load(cfg, ckpt):
    model = rm.get_model(cfg, use_cuda=False)
    sd = torch.load(os.path.join(REF, "models", ckpt), map_location="cpu", weights_only=True)
    model.load_state_dict({k: v for k, v in sd.items() if "_x" not in k}, strict=False)
    model.train()
    return model


def weights(model):
    out = {}
    for i, lin in enumerate(m for m in model if isinstance(m, torch.nn.Linear)):
        out[f"W{i}"] = lin.weight.detach().numpy().copy()
        out[f"b{i}"] = lin.bias.detach().numpy().copy()
    return out


def langevin_stats():
    torch.manual_seed(30)
    np.random.seed(2)
    cfg = {"input_size": 20, "hidden_size": 128, "hidden2_size": 128, "output_size": 784, "activation_fn": "relu",
           "loss_fn": rm.bernoulli_fn, "input_var": None, "T_pc": 250, "optimizer_x_fn_pc": optim.Adam,
           "optimizer_x_kwargs_pc": {"lr": 0.1}, "mixing": 150, "sampling": 100, "optimizer_x_kwargs_mcpc": {"lr": 0.03}}
    model = load(cfg, "mcpc_ml_1")
    B, R, T = 64, 32, 250
    data = te.sample_pc(B, model, cfg).float()               # binary images from the model itself
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    map_tr = te.get_pc_trainer(model, cfg, is_mcpc=True)
    mc_tr = te.get_mcpc_trainer(model, cfg, training=False)
    z = torch.zeros(B, 20)
    map_tr.train_on_batch(inputs=z, loss_fn=rm.bernoulli_fn, loss_fn_kwargs={"_target": data, "_var": None},
                          is_log_progress=False, is_return_results_every_t=False, is_checking_after_callback_after_t=False)
    x_start = [layer.get_x().detach().clone() for layer in pcs]
    energy = np.zeros((R, T))
    loss = np.zeros((R, T))
    s1 = torch.zeros(B, 20, dtype=torch.float64)
    s2 = torch.zeros(B, 20, dtype=torch.float64)
    n = 0
    for r in range(R):
        with torch.no_grad():
            for layer, x0 in zip(pcs, x_start):
                layer.get_x().copy_(x0)
        torch.manual_seed(1000 + r)
        res = mc_tr.train_on_batch(inputs=z, loss_fn=rm.bernoulli_fn, loss_fn_kwargs={"_target": data, "_var": None},
                                   callback_after_t=rm.random_step, callback_after_t_kwargs={"_pc_trainer": mc_tr},
                                   is_sample_x_at_batch_start=False, is_log_progress=False, is_return_results_every_t=True,
                                   is_checking_after_callback_after_t=False, is_return_representations=True)
        energy[r], loss[r] = res["energy"], res["loss"]
        reps = torch.stack(res["representations"][150:]).double()        # [100, B, 20]
        s1 += reps.sum(0)
        s2 += (reps ** 2).sum(0)
        n += reps.shape[0]
        print(f"run {r}: mean energy last 100 = {energy[r, 150:].mean():.2f}, loss = {loss[r, 150:].mean():.2f}")
    mean = s1 / n
    var = (s2 - n * mean ** 2) / (n - 1)
    out = dict(weights(model), data=data.numpy(), energy=energy, loss=loss, post_mean=mean.numpy(), post_var=var.numpy(),
               n_pooled=np.array(n), lr=np.array(0.03), mixing=np.array(150), sampling=np.array(100))
    for l, x0 in enumerate(x_start):
        out[f"x_start{l}"] = x0.numpy()
    np.savez_compressed(os.path.join(HERE, "stat_langevin_mcpc_ml.npz"), **out)


def mse_rec():
    out = {}
    for tag, ckpt, cfg in (
            ("pc", "pc_mse_1", {"input_size": 30, "hidden_size": 256, "hidden2_size": 256, "output_size": 784,
                                "activation_fn": "tanh"}),
            ("mcpc", "mcpc_mse_1", {"input_size": 10, "hidden_size": 256, "hidden2_size": 256, "output_size": 784,
                                    "activation_fn": "relu"})):
        torch.manual_seed(30)
        cfg.update(loss_fn=rm.bernoulli_fn, input_var=None, T_pc=250, optimizer_x_fn_pc=optim.Adam,
                   optimizer_x_kwargs_pc={"lr": 0.7})
        model = load(cfg, ckpt)
        B = 128
        data = te.sample_pc(B, model, cfg).float()
        pcs = [m for m in model if isinstance(m, pc.PCLayer)]
        x0 = [torch.empty(B, d).uniform_(-10.0, 10.0) for d in (cfg["input_size"], 256, 256)]    # utils/model.py:8-9
        for layer, v in zip(pcs, x0):
            layer._sample_x_fn = (lambda inputs, v=v: v.clone())
        loader = [(data, torch.zeros(B))]
        mse = te.get_mse_rec(model, cfg, loader, use_cuda=False)
        print(tag, "MSE on the masked-out half:", float(mse))
        for k, v in weights(model).items():
            out[f"{tag}_{k}"] = v
        out[f"{tag}_data"] = data.numpy()
        out[f"{tag}_mse"] = np.array(float(mse))
        for l, v in enumerate(x0):
            out[f"{tag}_x0_{l}"] = v.numpy()
        out[f"{tag}_x_final0"] = pcs[0].get_x().detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "stat_mse_rec.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["langevin", "mse"]
    if "langevin" in which:
        langevin_stats()
    if "mse" in which:
        mse_rec()
