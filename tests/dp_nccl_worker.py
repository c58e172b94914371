"""Worker of tests/test_gpu_dp_nccl.py: launched with torch.distributed.run on N GPUs (one rank per GPU, NCCL).
Every rank runs its shard of the batch through the REAL kernels; rank 0 also runs the concatenated batch alone on its
GPU (the oracle of a sharded run) and compares: per-step energy / loss, the weights after the p-step, and
every rank's latents against the matching rows of the big-batch run."""
import json
import os
import sys
import warnings

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.optim as optim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from montecarlopredictivecoding_b200 import mcpc_utils as mu  # noqa: E402
from montecarlopredictivecoding_b200 import predictive_coding as pc  # noqa: E402


def build(cfg, dev):
    torch.manual_seed(0)
    return mu.get_model(cfg, use_cuda=False, sample_x_fn=mu.sample_x_fn_normal).to(dev)


def run(model, cfg, x0, y, dp, precision, dev, top):
    mixing, sampling = 5, 7
    tr = pc.PCTrainer(model, T=mixing + sampling, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.03}, update_p_at="last",
                      accumulate_p_at=list(range(mixing, mixing + sampling)), optimizer_p_fn=optim.Adam,
                      optimizer_p_kwargs={"lr": 0.01}, plot_progress_at=[])
    tr.set_precision(precision)
    tr.set_noise_seed(4321)
    if dp:
        tr.set_data_parallel()
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    for layer, v in zip(pcs, x0):
        layer._sample_x_fn = (lambda inputs, v=v: v.clone())
    loss_fn = mu.bernoulli_fn if top == "bernoulli" else mu.fe_fn
    res = tr.train_on_batch(torch.zeros(y.shape[0], cfg["input_size"], device=dev), loss_fn=loss_fn,
                            loss_fn_kwargs={"_target": y, "_var": 1.0}, callback_after_t=mu.random_step,
                            callback_after_t_kwargs={"_pc_trainer": tr}, is_log_progress=False,
                            is_checking_after_callback_after_t=False)
    return res, [layer.get_x().detach().clone() for layer in pcs], tr.last_call_info


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    report = {}
    cases = [
        ("resident_fp32", "fp32", dict(input_size=20, hidden_size=128, hidden2_size=128, output_size=784, activation_fn="relu"), 96, "bernoulli", False),
        ("resident_bf16", "bf16", dict(input_size=20, hidden_size=128, hidden2_size=128, output_size=784, activation_fn="relu"), 1024, "bernoulli", False),
        ("streaming_bf16", "bf16", dict(input_size=64, hidden_size=320, hidden2_size=192, output_size=256, activation_fn="tanh"), 300, "gauss", True),
    ]
    ok = True
    for name, precision, cfg, B_local, top, streaming in cases:
        if streaming:
            os.environ["MCPC_FORCE_STREAMING"] = "1"
        else:
            os.environ.pop("MCPC_FORCE_STREAMING", None)
        dims = (cfg["input_size"], cfg["hidden_size"], cfg["hidden2_size"])
        B = B_local * world
        g = torch.Generator(device="cpu").manual_seed(11)
        x0 = [torch.randn(B, d, generator=g).to(dev) for d in dims]
        y = ((torch.rand(B, cfg["output_size"], generator=g) < 0.5).float() if top == "bernoulli"
             else torch.randn(B, cfg["output_size"], generator=g)).to(dev)
        rows = slice(rank * B_local, (rank + 1) * B_local)
        model = build(cfg, dev)
        res, xs, info = run(model, cfg, [v[rows] for v in x0], y[rows], True, precision, dev, top)
        W_dp = [p.detach().clone() for p in model.parameters() if p.dim() == 2 and p.shape[0] != B_local]
        # every rank's latents against the big-batch run, which every rank repeats locally (cheap at these sizes)
        model_1 = build(cfg, dev)
        res_1, xs_1, _ = run(model_1, cfg, x0, y, False, precision, dev, top)
        W_1 = [p.detach().clone() for p in model_1.parameters() if p.dim() == 2 and p.shape[0] != B]
        tol_x = 1e-5 if precision == "fp32" else 2e-3
        err_x = max(float((a - b[rows]).abs().max() / b.abs().max()) for a, b in zip(xs, xs_1))
        e_dp, e_1 = torch.tensor(res["energy"]), torch.tensor(res_1["energy"])
        l_dp, l_1 = torch.tensor(res["loss"]), torch.tensor(res_1["loss"])
        err_e = float(((e_dp - e_1).abs() / e_1.abs()).max())
        err_l = float(((l_dp - l_1).abs() / l_1.abs()).max())
        # Adam's first step is lr * sign(g): compare the weights where the big-batch update is not a coin flip
        err_w = max(float((a - b).abs().max()) for a, b in zip(W_dp, W_1))
        flips = max(float(((a - b).abs() > 1e-3).float().mean()) for a, b in zip(W_dp, W_1))
        case_ok = err_x < tol_x and err_e < 1e-4 and err_l < 1e-4 and flips < 2e-3
        gathered = [None] * world
        dist.all_gather_object(gathered, dict(rank=rank, err_x=err_x, err_e=err_e, err_l=err_l, err_w=err_w, flips=flips,
                                              ok=case_ok, mode=info.get("mode")))
        report[name] = gathered
        ok = ok and all(gr["ok"] for gr in gathered)
    if rank == 0:
        print("DP_NCCL_REPORT " + json.dumps({"world": world, "ok": ok, "cases": report}))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
