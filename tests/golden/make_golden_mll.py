"""Golden vectors for SURVEY §8(f) N1: run the REAL reference get_marginal_likelihood
(utils/training_evaluation.py:177-206) on small synthetic datasets and record the prior samples it drew
(sample_pc is wrapped, not replaced), the data and its result.  Build container only:
``python tests/golden/make_golden_mll.py``; the .npz files are committed."""
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
from torch.utils.data import DataLoader, TensorDataset  # noqa: E402

import predictive_coding as pc  # noqa: E402,F401
from utils import model as ref_model  # noqa: E402
from utils import training_evaluation as te  # noqa: E402

assert os.path.realpath(te.__file__).startswith(os.path.realpath(REF))


def run(name, cfg, n_data, n_samples, binary, checkpoint=None, seed=5):
    torch.manual_seed(seed)
    cfg = dict(cfg)
    cfg["loss_fn"] = ref_model.bernoulli_fn
    model = ref_model.get_model(cfg, use_cuda=False)
    if checkpoint is not None:
        sd = torch.load(os.path.join(REF, "models", checkpoint), map_location="cpu", weights_only=True)
        model.load_state_dict({k: v for k, v in sd.items() if "_x" not in k}, strict=False)
    data = torch.rand(n_data, cfg["output_size"])
    if binary:
        data = (data < 0.3).float()
    loader = DataLoader(TensorDataset(data, torch.zeros(n_data)), batch_size=7)
    captured = {}
    orig = te.sample_pc

    def wrapped(*a, **k):
        out = orig(*a, **k)
        captured["logits"] = out.detach().cpu().clone()
        return out
    te.sample_pc = wrapped
    try:
        ml = te.get_marginal_likelihood(model, cfg, loader, use_cuda=False, n_samples=n_samples)
    finally:
        te.sample_pc = orig
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, logits=captured["logits"].numpy(), data=data.numpy(), ml=np.float64(float(ml)),
                        n_samples=n_samples)
    print(name, float(ml), captured["logits"].shape, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    small = {"input_size": 4, "hidden_size": 8, "hidden2_size": 8, "output_size": 24, "activation_fn": "relu", "input_var": None}
    run("mll_small_soft", small, n_data=37, n_samples=50, binary=False)
    ml_cfg = {"input_size": 20, "hidden_size": 128, "hidden2_size": 128, "output_size": 784, "activation_fn": "relu",
              "input_var": None}
    run("mll_ml_checkpoint_binary", ml_cfg, n_data=33, n_samples=300, binary=True, checkpoint="mcpc_ml_1")
