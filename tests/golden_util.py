"""Helpers shared by the golden-vector tests: load a fixture written by
tests/golden/make_golden.py and describe each recorded ``train_on_batch`` call."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import mcpc_oracle as orc  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
ALL_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith(("mll_", "stat_")))
MLL_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith("mll_"))

ACT = {"identity": orc.ACT_IDENTITY, "relu": orc.ACT_RELU, "tanh": orc.ACT_TANH}


class GoldenCase:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.spec = json.loads(bytes(self.z["spec_json"]).decode())
        self.dims = self.spec["dims"]
        self.L = len(self.dims)
        self.d_out = self.spec.get("d_out", 0)
        self.n_lin = self.L + (1 if self.d_out > 0 else 0)
        self.inputs = self.z["inputs"]
        self.target = self.z["target"]
        self.B = self.spec["B"]
        self.calls = self.spec["calls"]

    def weights(self, call_idx, when="before"):
        """(W list, b list) at the start (``before``) / end (``after``) of call ``call_idx``."""
        if when == "before":
            if call_idx == 0:
                key = "c0_{}{}_before"
            else:
                return self.weights(call_idx - 1, "after")
        else:
            key = f"c{call_idx}_" + "{}{}_after"
        W = [self.z[key.format("W", i)] for i in range(self.n_lin)]
        b = [self.z[key.format("b", i)] if key.format("b", i) in self.z else None for i in range(self.n_lin)]
        return W, b

    def loss_kind(self, call_idx):
        return self.calls[call_idx].get("loss", self.spec.get("loss", "none"))

    def oracle_net(self, call_idx, dtype=np.float32):
        W, b = self.weights(call_idx)
        lk = self.loss_kind(call_idx)
        top = {"none": orc.TOP_NONE, "zero": orc.TOP_ZERO, "gauss": orc.TOP_GAUSS, "gauss_mask": orc.TOP_GAUSS,
               "bernoulli": orc.TOP_BERNOULLI, "bernoulli_mask": orc.TOP_BERNOULLI}[lk]
        ms = 0
        if "mask" in lk:
            ms = self.d_out - round(self.d_out * self.spec.get("perc", 0.5))
        act = [ACT[self.spec["act"]]] * self.L
        if self.d_out == 0:
            act[-1] = orc.ACT_IDENTITY      # a free output PCLayer has no activation after it
        net = orc.OracleNet(W=[w.copy() for w in W], b=[None if v is None else v.copy() for v in b],
                            n_layers=self.L, act=act,
                            energy_scale=self.spec.get("energy_scale", [1.0] * self.L),
                            top=top, top_var=self.spec.get("var", 1.0), mask_start_col=ms)
        return net.cast(dtype)

    def rows(self, ci):
        """Number of chains call ``ci`` ran with (a call may use only the first rows of inputs/target)."""
        return self.calls[ci].get("rows", self.B)

    def x0(self, ci):
        return [self.z[f"c{ci}_traj_x{l}"][0] for l in range(self.L)]

    def traj(self, ci, l):
        return self.z[f"c{ci}_traj_x{l}"]

    def x_final(self, ci):
        return [self.z[f"c{ci}_x{l}_final"] for l in range(self.L)]

    def noise(self, ci):
        if f"c{ci}_noise0" not in self.z:
            return None
        T = self.calls[ci]["trainer"]["T"]
        return [[self.z[f"c{ci}_noise{l}"][t] for l in range(self.L)] for t in range(T)]

    def grads(self, ci):
        gW = [self.z.get(f"c{ci}_gW{i}") if f"c{ci}_gW{i}" in self.z else None for i in range(self.n_lin)]
        gb = [self.z.get(f"c{ci}_gb{i}") if f"c{ci}_gb{i}" in self.z else None for i in range(self.n_lin)]
        return gW, gb

    def steps_run(self, ci):
        """Steps the reference actually ran (< T when its early_stop_condition fired, pc_trainer.py:979-981)."""
        return len(self.z[f"c{ci}_energy"])

    def step_lists(self, ci):
        tr = self.calls[ci]["trainer"]
        T = tr["T"]

        def expand(v):
            if v == "all":
                return list(range(T))
            if v == "last":
                return [T - 1]
            if v == "last_half":
                return list(range(T // 2, T))
            if v == "never":
                return []
            return list(v)
        upd_x, upd_p, acc = (expand(tr.get("update_x_at", "all")), expand(tr.get("update_p_at", "never")),
                             expand(tr.get("accumulate_p_at", "never")))
        n_run = self.steps_run(ci)
        if n_run < T:
            # the stopping step takes a p-update (update_p_at_early_stop=True, pc_trainer.py:853,904): for the oracle,
            # which has no eval()'d condition, that is the schedule truncated at n_run with a p-step on its last step
            upd_x = [t for t in upd_x if t < n_run]
            upd_p = sorted(set([t for t in upd_p if t < n_run] + [n_run - 1]))
            acc = [t for t in acc if t < n_run]
        return upd_x, upd_p, acc


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))
