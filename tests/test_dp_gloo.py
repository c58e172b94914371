"""Data-parallel host logic with world_size 2 on CPU (gloo): the batch is sharded over ranks, no
communication during inference, one all-reduce of the flat weight gradient (+ the per-step scalars), the
normalisation uses the GLOBAL batch and the Philox stream is keyed by the global chain id.  The oracle of a
sharded run is the single-process run of the concatenated batch.  Kernels are replaced by the
oracle test double; everything else is the shipped PCTrainer."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn
import torch.optim as optim

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _build(seed=0):
    from montecarlopredictivecoding_b200 import mcpc_utils as mu
    torch.manual_seed(seed)
    cfg = {"input_size": 6, "hidden_size": 16, "hidden2_size": 12, "output_size": 24, "activation_fn": "tanh"}
    return mu.get_model(cfg, use_cuda=False)


def _run(model, x0, y, rows, dp, seed=77):
    from montecarlopredictivecoding_b200 import mcpc_utils as mu
    from montecarlopredictivecoding_b200 import predictive_coding as pc
    from oracle_engine import OracleEngine
    mixing, sampling = 3, 4
    tr = pc.PCTrainer(model, T=mixing + sampling, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.05},
                      update_p_at="last", accumulate_p_at=list(range(mixing, mixing + sampling)),
                      optimizer_p_fn=optim.SGD, optimizer_p_kwargs={"lr": 0.1}, plot_progress_at=[])
    tr._engine = OracleEngine()
    tr.set_noise_seed(seed)
    if dp:
        tr.set_data_parallel()
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    for layer, v in zip(pcs, x0):
        layer._sample_x_fn = (lambda inputs, v=v[rows]: v.clone())
    res = tr.train_on_batch(torch.zeros(len(y[rows]), 6), loss_fn=mu.bernoulli_fn, loss_fn_kwargs={"_target": y[rows]},
                            callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr},
                            is_log_progress=False)
    return res, [p.get_x().detach().clone() for p in pcs]


def _data(B):
    g = torch.Generator().manual_seed(5)
    x0 = [torch.randn(B, d, generator=g) for d in (6, 16, 12)]
    y = (torch.rand(B, 24, generator=g) < 0.5).float()
    return x0, y


def _worker(rank, world, port, B, out_dir):
    import warnings
    warnings.simplefilter("ignore")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _build()
    x0, y = _data(B)
    shard = slice(rank * B // world, (rank + 1) * B // world)
    res, xs = _run(model, x0, y, shard, dp=True)
    lins = [m for m in model if isinstance(m, nn.Linear)]
    torch.save({"W": [l.weight.detach() for l in lins], "b": [l.bias.detach() for l in lins], "xs": xs,
                "energy": res["energy"], "loss": res["loss"]}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("single_collective", ["0", "1"])
def test_two_rank_run_equals_single_process_big_batch(tmp_path, single_collective, monkeypatch):
    """single_collective = 1: the [2, T] scalars ride in the tail of the flat gradient buffer (ONE all-reduce per call);
    0 (default): they are reduced right after the inference launch so that the host gets them early."""
    import warnings
    warnings.simplefilter("ignore")
    monkeypatch.setenv("MCPC_DP_SINGLE_COLLECTIVE", single_collective)
    B, world = 16, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    model = _build()
    x0, y = _data(B)
    res, xs = _run(model, x0, y, slice(0, B), dp=False)
    lins = [m for m in model if isinstance(m, nn.Linear)]
    shards = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(world)]
    for r in range(world):
        for i, lin in enumerate(lins):          # identical parameters on every rank == the single-process update
            assert torch.allclose(shards[r]["W"][i], lin.weight.detach(), atol=2e-6), (r, i)
            assert torch.allclose(shards[r]["b"][i], lin.bias.detach(), atol=2e-6), (r, i)
        assert np.allclose(shards[r]["energy"], res["energy"], rtol=1e-5)     # scalars are global sums
        assert np.allclose(shards[r]["loss"], res["loss"], rtol=1e-5)
    for l in range(3):                          # chains do not depend on the sharding (global chain ids)
        cat = torch.cat([shards[r]["xs"][l] for r in range(world)])
        assert torch.allclose(cat, xs[l], atol=1e-6), l
