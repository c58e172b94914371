#!/usr/bin/env python
"""Secondary measurements (one JSON line per workload) for the configs of BASELINE.json that are not the
bench.py headline: C1 figure_2 linear model (latency), C3 65,536-chain sampling, C4 deterministic PC (Adam),
C5 wide 4x4096.  Usage: python scripts/bench_configs.py [c1 c3 c4 c5] [--precision bf16|fp32]"""
import json
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.optim as optim  # noqa: E402

from montecarlopredictivecoding_b200 import mcpc_utils as mu  # noqa: E402
from montecarlopredictivecoding_b200 import predictive_coding as pc  # noqa: E402

LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(LOCAL)
DEV = torch.device("cuda", LOCAL)
if WORLD > 1:
    import torch.distributed as dist
    if not dist.is_initialized():          # bench.py imports this module after creating the group itself
        dist.init_process_group("nccl", device_id=DEV)
PEAK_TF = 1663.5
if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")):
    PEAK_TF = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("bf16_tflops", PEAK_TF)


def timed(fn, reps=3, warm=1):
    """Best of `reps` device-timed calls (CUDA events on the launching stream); multi-GPU: barrier on both sides and the
    MAX over ranks of every repetition."""
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        if WORLD > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3
        if WORLD > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=DEV)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        ts.append(dt)
    return min(ts)


def c1(prec):
    model = nn.Sequential(nn.Linear(1, 1), pc.PCLayer(sample_x_fn=mu.sample_x_fn_cte), nn.Linear(1, 1, bias=False)).to(DEV)
    model.train()
    nn.init.constant_(model[0].bias, 0.2)
    nn.init.constant_(model[2].weight, 2.0)
    T = 10000
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.02}, update_p_at="never", plot_progress_at=[])
    tr.set_precision(prec)
    out = {}

    def run():
        out["res"] = tr.train_on_batch(torch.zeros(1, 1, device=DEV), loss_fn=mu.fe_fn,
                                        loss_fn_kwargs={"_target": torch.ones(1, 1, device=DEV), "_var": 1.0},
                                        callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr},
                                        is_log_progress=False, is_return_representations=True)
    s = timed(run)
    reps = torch.stack(out["res"]["representations"][1000:])
    return {"workload": "C1 figure_2 linear-Gaussian, B=1, T=10000, per-step representations", "precision": prec,
            "steps_per_s": T / s, "ms": s * 1e3, "posterior_mean": float(reps.mean()), "posterior_var": float(reps.var()),
            "analytic": [0.44, 0.2]}


def ml_model(act="relu", dims=(20, 128, 128)):
    torch.manual_seed(0)
    cfg = {"input_size": dims[0], "hidden_size": dims[1], "hidden2_size": dims[2], "output_size": 784, "activation_fn": act}
    return mu.get_model(cfg, use_cuda=False).to(DEV)


def c3(prec, B=65536, T=1000, thin=0):
    """BASELINE.json configs[2]: 65,536 chains in total, sharded over the GPUs (B / WORLD chains each, no communication
    at all: update_p_at='never'); `thin` > 0 additionally reads the sensory outputs out every thin-th step into a device
    ring ("outputs read out every k-th step only")."""
    B_total = B
    B = B_total // WORLD
    model = ml_model()
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.1}, update_p_at="never", plot_progress_at=[])
    tr.set_precision(prec)
    if WORLD > 1:
        tr.set_data_parallel(chain_offset=int(os.environ.get("RANK", "0")) * B)      # global chain ids: one Philox stream
    if thin:
        tr.set_trajectory_stride(thin, 0)
        tr.set_trajectories_on_device(True)
    z = torch.zeros(B, 20, device=DEV)
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    x0 = [torch.randn(B, d, device=DEV) for d in (20, 128, 128)]
    for layer, v in zip(pcs, x0):
        layer._sample_x_fn = (lambda inputs, v=v: v.clone())
    first = [True]

    def run():
        tr.train_on_batch(z, loss_fn=mu.zero_fn, callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr},
                          is_sample_x_at_batch_start=first[0], is_log_progress=False,
                          is_return_results_every_t=bool(thin), is_return_outputs=bool(thin),
                          is_checking_after_callback_after_t=False)
        first[0] = False
    s = timed(run, reps=2)
    flops = B_total * T * 4 * (20 * 128 + 128 * 128)
    return {"workload": f"C3 sampling, {B_total} chains over {WORLD} GPU(s), T={T}, zero_fn (sensory Linear is readout only)"
                        + (f", outputs read out every {thin}th step into a device ring" if thin else ""),
            "precision": prec, "n_gpus": WORLD, "chains_per_gpu": B, "scaling": "strong",
            "latent_updates_per_s": B_total * 3 * T / s, "ms": s * 1e3, "us_per_step": s / T * 1e6,
            "algorithmic_tflops": flops / s / 1e12, "frac_of_bf16_peak": flops / s / 1e12 / (PEAK_TF * WORLD)}


def c4(prec, B=1024, T=250):
    model = ml_model("tanh", (25, 128, 128))
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.Adam, optimizer_x_kwargs={"lr": 0.3}, update_p_at="last",
                      optimizer_p_fn=optim.Adam, optimizer_p_kwargs={"lr": 0.01}, plot_progress_at=[])
    tr.set_precision(prec)
    z = torch.zeros(B, 25, device=DEV)
    y = (torch.rand(B, 784, device=DEV) < 0.5).float()

    def run():
        tr.train_on_batch(z, loss_fn=mu.bernoulli_fn_mask, loss_fn_kwargs={"_target": y, "_var": 1.0}, is_log_progress=False,
                          is_return_results_every_t=False)
    s = timed(run)
    return {"workload": f"C4 deterministic PC (pc_ml shape 25-128-128->784 tanh), Adam lr 0.3, T={T}, masked BCE, B={B}",
            "precision": prec, "latent_updates_per_s": B * 3 * T / s, "ms": s * 1e3, "us_per_step": s / T * 1e6,
            "images_per_s": B / s}


def live_cublas_tflops(seconds=1.0, n=8192):
    """cuBLAS bf16 GEMM (torch.matmul, n^3) run back to back for `seconds` on THIS box right now: the same measurement as
    MEASURED_PEAKS.json's sustained figure, taken next to our own number because the achievable rate is set by the 1 kW
    power cap and differs from box to box and minute to minute."""
    a = torch.randn(n, n, device=DEV, dtype=torch.bfloat16)
    b = torch.randn(n, n, device=DEV, dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 0
    e0.record()
    import time
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(20):
            torch.matmul(a, b)
        reps += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12


def c5(prec, B=2048, T=None, width=4096, L=4):
    T = T or int(os.environ.get("MCPC_C5_T", "20"))
    torch.manual_seed(0)
    mods, prev = [], width
    for _ in range(L):
        mods += [nn.Linear(prev, width), pc.PCLayer(sample_x_fn=mu.sample_x_fn_normal), nn.Tanh()]
    mods.append(nn.Linear(width, width))
    model = nn.Sequential(*mods)
    model.train()
    with torch.no_grad():
        for m in model:
            if isinstance(m, nn.Linear):
                m.weight.normal_(0, (1.0 / width) ** 0.5)
                m.bias.zero_()
    model.to(DEV)
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.01}, update_p_at="last",
                      accumulate_p_at=list(range(T)), optimizer_p_fn=optim.SGD, optimizer_p_kwargs={"lr": 1e-4},
                      plot_progress_at=[])
    tr.set_precision(prec)
    if WORLD > 1:
        tr.set_data_parallel()            # chains sharded: B per GPU, one NCCL all-reduce of the 268 MB dW per call
    z = torch.zeros(B, width, device=DEV)
    y = torch.randn(B, width, device=DEV)
    first = [True]

    def run():
        tr.train_on_batch(z, loss_fn=mu.fe_fn, loss_fn_kwargs={"_target": y, "_var": 1.0}, callback_after_t=mu.random_step,
                          callback_after_t_kwargs={"_pc_trainer": tr}, is_sample_x_at_batch_start=first[0],
                          is_log_progress=False, is_return_results_every_t=False)
        first[0] = False
    live = live_cublas_tflops() if os.environ.get("MCPC_C5_LIVE_PEAK", "1") != "0" else None
    s = timed(run, reps=3 if T >= 50 else 2)
    mac = L * width * width          # Linear_0 sees zero inputs; 3 hidden + 1 output contraction of width^2 each ... L total
    flops = B * T * 6 * mac          # fwd + back-projection + dW every step
    sustained = PEAK_TF
    try:
        sustained = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", PEAK_TF)
    except Exception:  # noqa: BLE001
        pass
    return {"workload": f"C5 wide {L}x{width}->{width} tanh Gaussian, B={B} per GPU, T={T}, dW every step"
                        + (f", one NCCL all-reduce of the {L * width * width * 4 / 1e6:.0f} MB dW per call" if WORLD > 1 else ""),
            "precision": prec, "n_gpus": WORLD, "scaling": "weak", "ms_per_call": s * 1e3,
            "ms_per_step": s / T * 1e3, "latent_updates_per_s": WORLD * B * L * T / s, "images_per_s_T100": WORLD * B / (s / T * 100),
            "algorithmic_tflops_per_gpu": flops / s / 1e12, "frac_of_bf16_peak": flops / s / 1e12 / PEAK_TF,
            "frac_of_bf16_peak_sustained": flops / s / 1e12 / sustained,
            "live_cublas_bf16_tflops_sustained": live,
            "frac_of_live_cublas": (flops / s / 1e12 / live) if live else None,
            "peak_note": "per-GPU algorithmic TFLOP/s (6 x MAC x B x T, whole call incl. weight conversion, reductions, "
                         "all-reduce and p-step) over the measured cuBLAS bf16 burst / sustained peaks (MEASURED_PEAKS.json)"}


def n1(prec, N=10000, S=5000, D=784):
    """get_marginal_likelihood of table_1.py (10,000 test images x 5,000 prior samples x 784 pixels)."""
    import time

    import numpy as np

    from oracle import mcpc_oracle as orc
    g = torch.Generator(device="cpu").manual_seed(0)
    logits = (torch.randn(S, D, generator=g) * 4.0).to(DEV)
    data = (torch.rand(N, D, generator=g) < 0.2).float().to(DEV)
    out = {}

    def run():
        out["ml"] = mu.bernoulli_marginal_ll(logits, data)
    s = timed(run, reps=5, warm=2)
    flops = 2.0 * N * S * D
    rows = 40                                              # bounded CPU sample of the same workload, scaled in N
    t0 = time.perf_counter()
    orc.marginal_ll_bernoulli(logits.cpu().numpy(), data[:rows].cpu().numpy(), dtype=np.float32)
    cpu_s = (time.perf_counter() - t0) * N / rows
    return {"workload": f"N1 marginal likelihood, {N} rows x {S} samples x {D} pixels (table_1.py:253)", "precision": "bf16 hi/lo split x3, fp32 accumulate",
            "ms": s * 1e3, "rows_per_s": N / s, "algorithmic_tflops": flops / s / 1e12,
            "tensor_tflops_issued": 3 * flops * (832.0 / 784.0) / s / 1e12,
            "frac_of_bf16_peak_issued": 3 * flops * (832.0 / 784.0) / s / 1e12 / PEAK_TF, "ml": float(out["ml"]),
            "cpu_port_s_scaled": cpu_s, "cpu_sample": f"{rows} rows of the numpy oracle, scaled to {N}"}


if __name__ == "__main__":
    prec = "bf16"
    if "--precision" in sys.argv:
        prec = sys.argv[sys.argv.index("--precision") + 1]
    which = [a for a in sys.argv[1:] if a in ("c1", "c3", "c4", "c5", "n1")] or ["c1", "c3", "c4", "c5"]
    for w in which:
        try:
            out = {"c1": c1, "c3": c3, "c4": c4, "c5": c5, "n1": n1}[w](prec)
            if int(os.environ.get("RANK", "0")) == 0:
                print(json.dumps(out))
        except Exception as exc:  # noqa: BLE001
            print(json.dumps({"workload": w, "error": repr(exc)[:300]}))
        sys.stdout.flush()
    if WORLD > 1:
        dist.destroy_process_group()
