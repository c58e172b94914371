"""Drop-in check against the reference's OWN call sites (only runs where /root/reference exists, i.e. in the
build container; skipped on the GPU box).  The reference's unmodified `utils/model.py` and
`utils/training_evaluation.py` are imported on top of THIS repository's `predictive_coding` package: their
trainer factories, `random_step` callback object, `get_representations` and `get_mse_rec` must work as they are.
Kernels are replaced by the oracle test double (no GPU here); results are compared with the same functions
running on the reference's own `predictive_coding`."""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.optim as optim

REF = os.environ.get("MCPC_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "predictive_coding")), reason="reference not mounted")


def _import_utils(pc_first):
    """Import the reference's utils package with either our pc (pc_first=ROOT) or the reference's pc first."""
    for n in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for name in [m for m in sys.modules if m == "utils" or m.startswith("utils.") or m == "predictive_coding"
                 or m.startswith("predictive_coding.")]:
        del sys.modules[name]
    saved = list(sys.path)
    sys.path[:] = [pc_first, REF] + [p for p in saved if p not in (pc_first, REF, ROOT)]
    try:
        pc = importlib.import_module("predictive_coding")
        model = importlib.import_module("utils.model")
        te = importlib.import_module("utils.training_evaluation")
    finally:
        sys.path[:] = saved
    return pc, model, te


CONFIG = {
    "input_size": 6, "hidden_size": 16, "hidden2_size": 12, "output_size": 24, "activation_fn": "relu", "input_var": None,
    "T_pc": 20, "optimizer_x_fn_pc": optim.Adam, "optimizer_x_kwargs_pc": {"lr": 0.1},
    "mixing": 4, "sampling": 6, "optimizer_x_kwargs_mcpc": {"lr": 0.03},
    "optimizer_p_fn_mcpc": optim.Adam, "optimizer_p_kwargs_mcpc": {"lr": 0.01},
    "optimizer_p_fn": optim.Adam, "optimizer_p_kwargs": {"lr": 0.01},
}


def _run(pc, model_mod, te, ours):
    from oracle_engine import OracleEngine
    cfg = dict(CONFIG)
    cfg["loss_fn"] = model_mod.bernoulli_fn
    torch.manual_seed(30)
    net = model_mod.get_model(cfg, use_cuda=False, sample_x_fn=model_mod.sample_x_fn_cte)
    if ours:
        assert pc.PCTrainer.__module__.startswith("montecarlopredictivecoding_b200")
        orig = pc.PCTrainer._get_engine
        pc.PCTrainer._get_engine = lambda self: self.__dict__.setdefault("_engine_obj", OracleEngine())
    try:
        g = torch.Generator().manual_seed(7)
        data = (torch.rand(10, 24, generator=g) < 0.5).float()
        loader = [(data[:5], torch.zeros(5, dtype=torch.int)), (data[5:], torch.ones(5, dtype=torch.int))]
        # MAP representations exactly as utils/model.py:85-102 drives them
        pc_tr = te.get_pc_trainer(net, cfg, is_mcpc=True)
        ds = model_mod.get_representations(net, cfg, [pc_tr], loader, rep_type="MAP")
        reps = ds.tensors[0].detach().clone()
        # masked reconstruction error exactly as utils/training_evaluation.py:143-174 drives it
        mse = float(te.get_mse_rec(net, cfg, loader, use_cuda=False))
        # one MCPC learning call with the reference's own random_step object (utils/model.py:35-44)
        mcpc_tr = te.get_mcpc_trainer(net, cfg, training=True)
        res = mcpc_tr.train_on_batch(inputs=torch.zeros(5, 6), loss_fn=cfg["loss_fn"],
                                     loss_fn_kwargs={"_target": data[:5], "_var": None},
                                     callback_after_t=model_mod.random_step, callback_after_t_kwargs={"_pc_trainer": mcpc_tr},
                                     is_sample_x_at_batch_start=False, is_log_progress=False,
                                     is_checking_after_callback_after_t=False)
        mode = getattr(mcpc_tr, "last_call_info", {}).get("mode")
    finally:
        if ours:
            pc.PCTrainer._get_engine = orig
    return reps, mse, res, mode


def test_reference_utils_run_unmodified_on_the_drop_in():
    import warnings
    warnings.simplefilter("ignore")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    ref = _run(*_import_utils(REF), ours=False)
    ours = _run(*_import_utils(ROOT), ours=True)
    # deterministic parts (MAP with constant init): identical up to fp32 round-off
    assert torch.allclose(ours[0], ref[0], rtol=1e-4, atol=1e-5)
    assert abs(ours[1] - ref[1]) < 1e-6
    # Langevin call: the reference's random_step object is recognised and folded into the fused path; noise streams
    # differ (torch generator vs Philox), so only structure and first-step values are comparable
    assert ours[3] == "fused"
    assert len(ours[2]["energy"]) == len(ref[2]["energy"]) == 10
    assert len(ours[2]["loss"]) == 10
    assert abs(ours[2]["energy"][0] - ref[2]["energy"][0]) <= 1e-5 * abs(ref[2]["energy"][0])
    assert abs(ours[2]["loss"][0] - ref[2]["loss"][0]) <= 1e-5 * abs(ref[2]["loss"][0])
    for m in list(sys.modules):
        if m == "utils" or m.startswith("utils.") or m == "predictive_coding" or m.startswith("predictive_coding."):
            del sys.modules[m]
