"""``PCLayer`` -- the latent holder of a predictive-coding network.

Drop-in for the reference class (predictive_coding/pc_layer.py:8-304): same constructor,
same methods, same train/eval behaviour, ``_x`` is a registered ``nn.Parameter`` so it shows
up in ``state_dict()`` exactly like the shipped checkpoints expect.

In the B200 build the T-step loop never calls ``forward`` -- ``PCTrainer`` hands the storage
of ``_x`` to the fused kernel, which updates it in place.  ``forward`` is still a complete
PyTorch implementation because (i) the trainer runs ONE forward at t=0 to apply the user's
arbitrary ``sample_x_fn`` in module order (pc_layer.py:221-233), (ii) eval-mode networks pass
``mu`` through (pc_layer.py:304, used by ``sample_pc``), and (iii) ``energy()`` can be asked
for at any time.
"""
import typing
import warnings

import torch
import torch.nn as nn


def _default_energy(inputs):
    d = inputs["mu"] - inputs["x"]
    return 0.5 * d * d


def _default_sample_x(inputs):
    return inputs["mu"].detach().clone()


class PCLayer(nn.Module):
    """Insert between two layers to make the error between them a local (PC) one."""

    def __init__(
        self,
        energy_fn: typing.Callable = _default_energy,
        sample_x_fn: typing.Callable = _default_sample_x,
        S: torch.Tensor = None,
        M: torch.Tensor = None,
        is_holding_error: bool = False,
        is_keep_energy_per_datapoint: bool = False,
    ):
        super().__init__()
        assert callable(energy_fn)
        assert callable(sample_x_fn)
        assert isinstance(is_holding_error, bool)
        assert isinstance(is_keep_energy_per_datapoint, bool)
        self._energy_fn = energy_fn
        self._sample_x_fn = sample_x_fn
        self._energy = None
        self.set_S(S)
        self.set_M(M)
        self.is_holding_error = is_holding_error
        self.is_keep_energy_per_datapoint = is_keep_energy_per_datapoint
        if is_keep_energy_per_datapoint:
            self._energy_per_datapoint = None
        self._is_sample_x = False
        self._x = None
        # filled by PCTrainer after a fused call: a thunk that recomputes the energy on demand
        self._lazy_energy = None
        self.eval()     # reference: layers start in eval mode (pc_layer.py:104)

    # ---- getters / setters (pc_layer.py:108-133) -------------------------------------------
    def set_M(self, M):
        assert M is None or isinstance(M, torch.Tensor)
        self._M = M

    def set_S(self, S):
        if S is not None:
            assert isinstance(S, torch.Tensor)
            assert S.dim() == 2
        self._S = S

    def get_is_sample_x(self) -> bool:
        return self._is_sample_x

    def set_is_sample_x(self, is_sample_x: bool) -> None:
        assert isinstance(is_sample_x, bool)
        self._is_sample_x = is_sample_x

    def get_x(self) -> nn.Parameter:
        return self._x

    # ---- energy bookkeeping (pc_layer.py:137-159) -------------------------------------------
    def energy(self) -> torch.Tensor:
        if self._energy is None and self._lazy_energy is not None:
            self._lazy_energy()
        return self._energy

    def clear_energy(self):
        self._energy = None
        self._lazy_energy = None

    def energy_per_datapoint(self) -> torch.Tensor:
        assert self.is_keep_energy_per_datapoint
        return self._energy_per_datapoint

    def clear_energy_per_datapoint(self):
        assert self.is_keep_energy_per_datapoint
        self._energy_per_datapoint = None

    # ---- forward (pc_layer.py:161-304) --------------------------------------------------------
    def _needs_resample(self, mu):
        if self._x is None:
            return "The <self._x> has not been initialized yet, run with <pc_layer.set_is_sample_x(True)> first. We will do it for you."
        if mu.device != self._x.device:
            return "The device of <self._x> is not consistent with that of <mu>, run with <pc_layer.set_is_sample_x(True)> first. We will do it for you."
        if mu.size() != self._x.size():
            return ("You have changed the shape of this layer, you should do <pc_layer.set_is_sample_x(True) when changing the shape of this layer. We will do it for you.\n"
                    "This should have been taken care of by <pc_trainer> unless you have set <is_sample_x_at_epoch_start=False> when calling <pc_trainer.train_on_batch()>,\n"
                    "in which case you should be responsible for making sure the batch size stays still.")
        return None

    def forward(self, mu: torch.Tensor, energy_fn_additional_inputs: dict = {}) -> torch.Tensor:
        assert isinstance(mu, torch.Tensor)
        assert isinstance(energy_fn_additional_inputs, dict)
        if not self.training:
            return mu

        if not self._is_sample_x:
            why = self._needs_resample(mu)
            if why is not None:
                warnings.warn(why, category=RuntimeWarning)
                self._is_sample_x = True
        if self._is_sample_x:
            fresh = self._sample_x_fn({"mu": mu, "x": self._x})
            self._x = nn.Parameter(fresh.to(mu.device), True)
            self._is_sample_x = False      # one-shot flag

        x = self._x
        mu_e, x_e = mu, x
        if self._S is not None:
            # pairwise energies between every mu unit and every x unit (linear nets only)
            assert mu.dim() == 2 and x.dim() == 2
            n_mu, n_x = mu.size(1), x.size(1)
            assert self._S.size(0) == n_mu and self._S.size(1) == n_x
            mu_e = mu.unsqueeze(2).expand(-1, -1, n_x)
            x_e = x.unsqueeze(1).expand(-1, n_mu, -1)
        fn_inputs = {"mu": mu_e, "x": x_e}
        fn_inputs.update(energy_fn_additional_inputs)
        energy = self._energy_fn(fn_inputs)
        if self._S is not None:
            energy = energy * self._S.unsqueeze(0)
        elif self._M is not None:
            energy = energy * self._M.unsqueeze(0)
        if self.is_keep_energy_per_datapoint:
            self._energy_per_datapoint = energy.sum(dim=list(range(1, energy.dim())), keepdim=False).unsqueeze(1)
        self._energy = energy.sum()
        self._lazy_energy = None
        if self.is_holding_error:
            self.error = (self._x.data - mu).detach().clone()
        return self._x
