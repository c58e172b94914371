"""Unmodified reference call sites on the GPU: the reference's OWN `utils/model.py` /
`utils/training_evaluation.py` (git-ignored copy in baseline/_ref, installed by scripts/install_ref.py) run on top of
THIS repository's drop-in `predictive_coding`, with the NativeEngine and WITHOUT any call to set_precision: the default
('auto') must put them on the fused tcgen05 bf16 path.  The deterministic parts are compared with the same functions on
the reference's own CPU `predictive_coding`.  Runs in a subprocess (both packages are called `predictive_coding`)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

WORKER = r'''
import importlib, json, os, sys, types, warnings
import numpy as np, torch, torch.optim as optim
warnings.simplefilter("ignore")
ROOT, REF = sys.argv[1], sys.argv[2]
os.environ.pop("MCPC_PRECISION", None)          # the product default: auto

def import_utils(pc_first):
    for n in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for name in [m for m in sys.modules if m in ("utils", "predictive_coding") or m.startswith(("utils.", "predictive_coding."))]:
        del sys.modules[name]
    saved = list(sys.path)
    sys.path[:] = [pc_first, REF] + [p for p in saved if p not in (pc_first, REF, ROOT)]
    try:
        return (importlib.import_module("predictive_coding"), importlib.import_module("utils.model"),
                importlib.import_module("utils.training_evaluation"))
    finally:
        sys.path[:] = saved

CFG = {"input_size": 20, "hidden_size": 128, "hidden2_size": 128, "output_size": 784, "activation_fn": "relu", "input_var": None,
       "T_pc": 40, "optimizer_x_fn_pc": optim.Adam, "optimizer_x_kwargs_pc": {"lr": 0.1}, "mixing": 20, "sampling": 30,
       "optimizer_x_kwargs_mcpc": {"lr": 0.03}, "optimizer_p_fn_mcpc": optim.Adam, "optimizer_p_kwargs_mcpc": {"lr": 0.01},
       "optimizer_p_fn": optim.Adam, "optimizer_p_kwargs": {"lr": 0.01}}

def run(pc, mm, te, cuda):
    cfg = dict(CFG); cfg["loss_fn"] = mm.bernoulli_fn
    dev = torch.device("cuda:0" if cuda else "cpu")
    torch.manual_seed(30)
    net = mm.get_model(cfg, use_cuda=False, sample_x_fn=mm.sample_x_fn_cte)
    if cuda:
        net.cuda()
    g = torch.Generator().manual_seed(7)
    data = (torch.rand(64, 784, generator=g) < 0.5).float()
    loader = [(data[:32], torch.zeros(32, dtype=torch.int)), (data[32:], torch.ones(32, dtype=torch.int))]
    out = {}
    pc_tr = te.get_pc_trainer(net, cfg, is_mcpc=True)
    ds = mm.get_representations(net, cfg, [pc_tr], loader, rep_type="MAP", use_cuda=cuda)       # utils/model.py:85-102
    out["reps"] = ds.tensors[0].detach().cpu().numpy().tolist()
    out["mse"] = float(te.get_mse_rec(net, cfg, loader, use_cuda=cuda))                          # training_evaluation.py:143-174
    mc = te.get_mcpc_trainer(net, cfg, training=True)
    res = mc.train_on_batch(inputs=torch.zeros(32, 20, device=dev), loss_fn=cfg["loss_fn"],
                            loss_fn_kwargs={"_target": data[:32].to(dev), "_var": None}, callback_after_t=mm.random_step,
                            callback_after_t_kwargs={"_pc_trainer": mc}, is_sample_x_at_batch_start=False, is_log_progress=False,
                            is_checking_after_callback_after_t=False)
    out["energy0"], out["loss0"], out["n"] = res["energy"][0], res["loss"][0], len(res["energy"])
    out["info"] = {k: (int(v) if isinstance(v, (int, np.integer)) else v) for k, v in getattr(mc, "last_call_info", {}).items()}
    out["map_info"] = {k: (int(v) if isinstance(v, (int, np.integer)) else v) for k, v in getattr(pc_tr, "last_call_info", {}).items()}
    # figure_2.py:29-79 posterior of the linear-Gaussian model, written as the script writes it
    import torch.nn as nn
    lin = nn.Sequential(nn.Linear(1, 1), pc.PCLayer(sample_x_fn=mm.sample_x_fn_cte), nn.Linear(1, 1, bias=False))
    lin.train()
    nn.init.constant_(lin[0].bias, 0.2); nn.init.constant_(lin[2].weight, 2.0)
    if cuda:
        lin.cuda()
    tr = pc.PCTrainer(lin, T=3000 if cuda else 300, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.02}, update_p_at="never",
                      plot_progress_at=[])
    B = 256 if cuda else 4
    r = tr.train_on_batch(torch.zeros(B, 1, device=dev), loss_fn=mm.fe_fn, loss_fn_kwargs={"_target": torch.ones(B, 1, device=dev), "_var": 1.0},
                          callback_after_t=mm.random_step, callback_after_t_kwargs={"_pc_trainer": tr}, is_log_progress=False,
                          is_return_representations=True)
    s = torch.stack(r["representations"][200:])
    out["post_mean"], out["post_var"] = float(s.mean()), float(s.var())
    return out

sys.path.insert(0, os.path.join(ROOT, "tests"))
ref = run(*import_utils(REF), cuda=False)
pc, mm, te = import_utils(ROOT)
assert pc.PCTrainer.__module__.startswith("montecarlopredictivecoding_b200")
from montecarlopredictivecoding_b200 import _native
l0 = _native.load().mcpc_launch_count()
ours = run(pc, mm, te, cuda=True)
ours["launches"] = int(_native.load().mcpc_launch_count() - l0)
print("CALLSITES " + json.dumps({"ref": ref, "ours": ours}))
'''


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "utils", "model.py")),
                    reason="baseline/_ref not installed (python scripts/install_ref.py)")
def test_reference_utils_run_unmodified_on_the_gpu_drop_in(tmp_path):
    import numpy as np
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ)
    env.pop("MCPC_PRECISION", None)
    out = subprocess.run([sys.executable, str(script), ROOT, REF], capture_output=True, text=True, timeout=900, env=env)
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("CALLSITES ")]
    assert line, (out.stdout[-1500:], out.stderr[-3000:])
    d = json.loads(line[-1][len("CALLSITES "):])
    ref, ours = d["ref"], d["ours"]
    from montecarlopredictivecoding_b200 import _native as N
    # the reference's own random_step object is recognised: fused path, bf16 by default, native kernels launched
    assert ours["info"]["mode"] == "fused" and ours["info"]["precision"] == N.PREC_BF16, ours["info"]
    assert ours["map_info"]["precision"] == N.PREC_BF16
    assert ours["launches"] > 0
    assert ours["n"] == ref["n"] == 50
    # deterministic parts against the reference's CPU run (bf16 bound of a 40-step Adam MAP from a constant start)
    a, b = np.array(ours["reps"]), np.array(ref["reps"])
    err = float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    print("MAP representations: max-norm rel err", err, "| mse", ours["mse"], ref["mse"], "| posterior", ours["post_mean"], ours["post_var"])
    assert err < 5e-2
    assert abs(ours["mse"] - ref["mse"]) < 1e-2
    assert abs(ours["energy0"] - ref["energy0"]) < 2e-2 * abs(ref["energy0"])
    assert abs(ours["loss0"] - ref["loss0"]) < 2e-2 * abs(ref["loss0"])
    # figure_2: analytic posterior N(0.44, 0.2) (Euler-Maruyama variance 0.2105 at lr 0.02)
    assert abs(ours["post_mean"] - 0.44) < 0.015 and abs(ours["post_var"] - 0.2105) < 0.015
