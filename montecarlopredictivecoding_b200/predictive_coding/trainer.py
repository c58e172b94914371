"""``PCTrainer`` -- drop-in for predictive_coding/pc_trainer.py:22-1108 of the reference.

Same constructor, same ``train_on_batch`` signature, same results dict; the T-step loop
(pc_trainer.py:712-983) is executed by the fused sm_100a kernels behind the C ABI
(include/mcpc_b200.h) instead of ~120 ATen launches + 3 host syncs per step.

How a call is executed
----------------------
1. The model is compiled to a ``NetPlan`` (plan.py) -- at call time, because scripts create
   trainers before ``load_state_dict`` / ``add_module`` (figure_5.py:89-93, figure_6.py:85).
2. t=0 sampling (pc_trainer.py:717-724, pc_layer.py:221-233) runs the user's ``sample_x_fn`` in
   ONE ordinary PyTorch forward, layer by layer, like the reference does.
3. *Fused mode* -- no opaque Python has to run between steps (``callback_after_t`` is absent or
   is ``random_step``, no ``callback_after_backward``, ``early_stop_condition == "False"``, fixed
   x-lr): [0, T) is cut into segments at the p-update times and each segment is ONE
   ``mcpc_infer`` launch.  The Langevin noise of ``random_step`` (utils/model.py:35-44) is
   drawn in the kernel.  The operands of the local weight update are saved for the steps
   whose gradients survive the reference's zero_grad rules (pc_trainer.py:852-859) and
   contracted by ``mcpc_weight_grad``; then ``.grad`` is normalised and the user's
   ``optimizer_p`` stepped (pc_trainer.py:904-914).
4. *Step-by-step mode* -- any other callback: one launch per step that materialises
   ``x.grad``; the callbacks and the real ``optimizer_x.step()`` run in Python.
5. Per-step scalars come back as three vectors and are turned into Python lists once.
"""
import os
import typing
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from .. import _native as N
from . import plan as P
from .engine import InferCall, NativeEngine
from .layer import PCLayer


def _slow_down_warning(base, prop, solution):
    # predictive_coding/utils.py:8-16
    warnings.warn("In {}, you have {} enabled, this will slow down training. Set to {} to disable it. ".format(
        base, prop, solution), category=RuntimeWarning)


def _is_shape_only_sampler(fn) -> bool:
    """t=0 initialisers that use nothing but the shape of ``mu`` (tagged, or the reference's own utils/model.py samplers
    after a behavioural probe -- see plan.sampler_is_shape_only)."""
    return P.sampler_is_shape_only(fn)


class PCTrainer(object):
    """Trainer for predictive-coding models built from :class:`PCLayer`."""

    def __init__(
        self,
        model: nn.Module,
        optimizer_x_fn: typing.Callable = optim.SGD,
        optimizer_x_kwargs: dict = {"lr": 0.1},
        manual_optimizer_x_fn: typing.Callable = None,
        x_lr_amplifier: float = 1.0,
        x_lr_discount: float = 1.0,
        loss_x_fn: typing.Callable = None,
        loss_inputs_fn: typing.Callable = None,
        optimizer_p_fn: typing.Callable = optim.Adam,
        optimizer_p_kwargs: dict = {"lr": 0.001},
        manual_optimizer_p_fn: typing.Callable = None,
        T: int = 512,
        update_x_at: typing.Union[str, typing.List[int]] = "all",
        update_p_at: typing.Union[str, typing.List[int]] = "all",
        accumulate_p_at: typing.Union[str, typing.List[int]] = "never",
        energy_coefficient: float = 1.0,
        early_stop_condition: str = "False",
        update_p_at_early_stop: bool = True,
        plot_progress_at: typing.Union[str, typing.List[int]] = "all",
        is_disable_warning_energy_from_different_batch_sizes: bool = False,
    ):
        assert isinstance(model, nn.Module)
        self._model = model

        assert callable(optimizer_x_fn)
        assert isinstance(optimizer_x_kwargs, dict)
        assert manual_optimizer_x_fn is None or callable(manual_optimizer_x_fn)
        self._optimizer_x_fn = optimizer_x_fn
        self._optimizer_x_kwargs = optimizer_x_kwargs
        self._manual_optimizer_x_fn = manual_optimizer_x_fn
        self._optimizer_x = None

        assert isinstance(x_lr_discount, float) and x_lr_discount <= 1.0
        assert isinstance(x_lr_amplifier, float) and x_lr_amplifier >= 1.0
        self._x_lr_discount = x_lr_discount
        self._x_lr_amplifier = x_lr_amplifier

        for fn, label in ((loss_x_fn, "loss_x_fn"), (loss_inputs_fn, "loss_inputs_fn")):
            if fn is not None:
                assert callable(fn)
                assert self.get_is_model_has_pc_layers(), f"<{label}> should only work with models with <PCLayer>. "
        self._loss_x_fn = loss_x_fn
        self._loss_inputs_fn = loss_inputs_fn

        assert callable(optimizer_p_fn)
        assert isinstance(optimizer_p_kwargs, dict)
        assert manual_optimizer_p_fn is None or callable(manual_optimizer_p_fn)
        self._optimizer_p_fn = optimizer_p_fn
        self._optimizer_p_kwargs = optimizer_p_kwargs
        self._manual_optimizer_p_fn = manual_optimizer_p_fn
        self.recreate_optimize_p()

        assert isinstance(T, int) and T > 0
        self._T = T
        if self.get_is_model_has_pc_layers():
            if self._T < self.get_num_pc_layers() + 1:
                warnings.warn(
                    "You should always choose T such that T >= (<pc_trainer.get_num_pc_layers()> + 1), "
                    "as it ensures that the error can be PC-propagated through the network.",
                    category=RuntimeWarning)
            min_t = self.get_least_T()
            if self._T < min_t:
                warnings.warn(
                    f"If you have one pc_layer per layer, T={self._T} is too small. "
                    f"Please use a minimum T of {min_t}, which is just enough to PC-propagate the error through the "
                    "network and have all weigths updated based on these PC-propagated errors. "
                    "In practice, you normally should have T much larger than this minimum T. ",
                    category=RuntimeWarning)

        self._update_x_at = self._preprocess_step_index_list(indices=update_x_at, T=self._T)
        self._update_p_at = self._preprocess_step_index_list(indices=update_p_at, T=self._T)
        self._accumulate_p_at = self._preprocess_step_index_list(indices=accumulate_p_at, T=self._T)

        assert isinstance(energy_coefficient, float)
        self._energy_coefficient = energy_coefficient
        assert isinstance(early_stop_condition, str)
        self._early_stop_condition = early_stop_condition
        assert isinstance(update_p_at_early_stop, bool)
        self._update_p_at_early_stop = update_p_at_early_stop

        if isinstance(plot_progress_at, str):
            assert plot_progress_at in ["all"]
        elif isinstance(plot_progress_at, list):
            for h in plot_progress_at:
                assert isinstance(h, int)
        else:
            raise NotImplementedError
        self._plot_progress_at = plot_progress_at
        self._is_plot_progress = not (isinstance(plot_progress_at, list) and len(plot_progress_at) == 0)
        if self._is_plot_progress:
            self.reset_plot_progress()

        assert isinstance(is_disable_warning_energy_from_different_batch_sizes, bool)
        self.is_disable_warning_energy_from_different_batch_sizes = is_disable_warning_energy_from_different_batch_sizes

        # ---- B200 execution state (not part of the reference API) ----
        self._engine = None                    # NativeEngine, created on first use
        # MCPC_PREC_* of the next call.  Requested mode: env MCPC_PRECISION = auto (default) | bf16 | fp32, or
        # set_precision().  'auto' takes the tcgen05 bf16 kernels whenever they implement the call (the stated bound:
        # tests/test_gpu_bf16_bound.py) and the reference-exact fp32 kernels otherwise (non-zero `inputs`).
        self._precision_req = os.environ.get("MCPC_PRECISION", "auto").strip().lower()
        if self._precision_req not in ("auto", "bf16", "fp32"):
            raise ValueError(f"MCPC_PRECISION={self._precision_req!r}: expected auto, bf16 or fp32")
        self._precision = N.PREC_BF16 if self._precision_req == "bf16" else N.PREC_FP32
        self._seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        self._noise_epoch = 0                  # advances every call so successive calls draw fresh noise
        self._keep_unused_param_grads = False  # see set_keep_unused_param_grads()
        self._dp_group = None                  # torch.distributed group for data-parallel learning
        self._dp_chain_offset = None
        self._adam = None                      # persistent Adam state of the fused x-optimizer
        self._buffers = {}                     # cached device scratch keyed by role
        self._buffer_views = {}                # role -> (shape, shaped view of the buffer)
        self._param_cache = None               # detached Linear weights / biases of the last call
        self._flat_grad = None
        self._zero_inputs_cache = None
        self._save_budget_bytes = int(os.environ.get("MCPC_SAVE_BUDGET_BYTES", str(8 << 30)))
        self._supplied_noise = None            # validation hook: set_supplied_noise()
        self._traj_stride = 1                  # set_trajectory_stride(): record every k-th step ...
        self._traj_start = 0                   # ... from this step on
        self._traj_on_device = False           # set_trajectories_on_device()
        self._traj_stats_cfg = None            # enable_trajectory_stats()
        self._traj_stats = None
        self._traj_ring_bytes = int(os.environ.get("MCPC_TRAJ_RING_BYTES", str(256 << 20)))
        self.last_trajectories = None          # device rings of the last call: {"x": [...], "out": ..., "steps": [...]}
        self._fused_p_optimizer = os.environ.get("MCPC_FUSED_P_STEP", "1") != "0"
        # data-parallel learning calls: 0 (default) = the tiny [2, T] scalar all-reduce is issued right after the inference
        # kernel, so the host gets the results while dW / gradient all-reduce / p-step still run; 1 = the scalars ride in
        # the tail of the gradient buffer (ONE collective per call, but the host then waits for it: measured on 2 GPUs
        # 1.55 ms per C2 call against 1.30 ms)
        self._dp_single_collective = os.environ.get("MCPC_DP_SINGLE_COLLECTIVE", "0") == "1"
        self.last_call_info = {}

    # ======================================================================================
    #  B200-specific knobs
    # ======================================================================================
    def set_precision(self, precision) -> None:
        """'fp32' (reference-exact CUDA-core kernels, 1e-5 parity), 'bf16' (tcgen05 tensor-core kernels: bf16 operands,
        fp32 latents and accumulation; stated bound in tests/test_gpu_bf16_bound.py) or 'auto' (bf16 whenever the bf16
        kernels implement the call, else fp32).  Overrides the MCPC_PRECISION environment variable."""
        table = {"fp32": "fp32", "bf16": "bf16", "auto": "auto", N.PREC_FP32: "fp32", N.PREC_BF16: "bf16"}
        if precision not in table:
            raise ValueError(f"unknown precision {precision!r}")
        self._precision_req = table[precision]
        self._precision = N.PREC_BF16 if self._precision_req == "bf16" else N.PREC_FP32

    def _resolve_precision(self, inputs, netp=None, top=None) -> int:
        """The MCPC_PREC_* this call runs in.  'auto' is bf16: both bf16 modes implement everything the fp32 kernels do
        (non-zero ``inputs`` included); fp32 is the reference-exact validation mode."""
        if self._precision_req == "fp32":
            return N.PREC_FP32
        return N.PREC_BF16

    def set_noise_seed(self, seed: int) -> None:
        self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._noise_epoch = 0

    def set_supplied_noise(self, noise: typing.Optional[torch.Tensor]) -> None:
        """Validation hook: use this [T, B, sum(d_l)] raw gradient-noise tensor (what ``random_step``
        would have written into ``x.grad``) instead of the in-kernel generator for the next call."""
        self._supplied_noise = noise

    def set_keep_unused_param_grads(self, flag: bool) -> None:
        """The reference leaves parameter ``.grad`` behind even when no p-step ever reads it
        (``update_p_at='never'`` trainers).  Off by default: those contractions are skipped."""
        self._keep_unused_param_grads = bool(flag)

    def set_data_parallel(self, group=None, chain_offset: int = None) -> None:
        """Shard the batch over the ranks of ``group``: weight gradients (and the per-step scalars)
        are all-reduced, the normalisation uses the global batch, the Philox stream is keyed by the
        global chain index.  ``group=None`` with an initialised default group uses WORLD."""
        import torch.distributed as dist
        if group is None and dist.is_available() and dist.is_initialized():
            group = dist.group.WORLD
        self._dp_group = group
        self._dp_chain_offset = chain_offset

    # ---- trajectories stay on the device ----------------------------------------------
    def set_trajectory_stride(self, stride: int = 1, start: int = 0) -> None:
        """Record ``outputs`` / ``representations`` / ``xs`` only at steps ``start, start+stride, ...`` (with
        ``is_return_results_every_t=True``); the result lists then hold one entry per RECORDED step.  The reference
        records every step and thins afterwards (``temp[mixing::indent]``, utils/model.py:151; ``xs[mixing::]``,
        figure_5.py:108); ``stride=1, start=0`` (default) is exactly the reference's behaviour."""
        assert isinstance(stride, int) and stride >= 1 and isinstance(start, int) and start >= 0
        self._traj_stride, self._traj_start = stride, start

    def set_trajectories_on_device(self, flag: bool = True) -> None:
        """``representations`` / ``xs`` entries of the results dict stay CUDA tensors (views of one [n_rec, B, d] ring)
        instead of being copied to the host step by step like pc_trainer.py:772-774 does."""
        self._traj_on_device = bool(flag)

    def enable_trajectory_stats(self, start: int = 0, stride: int = 1, layers="all") -> None:
        """Fold the latents of the steps ``start, start+stride, ...`` of every following call into per-element running
        mean / variance ON THE DEVICE (csrc/traj_stats.cu), without returning -- or even keeping -- the trajectory:
        a bounded ring of recorded steps is reduced chunk by chunk.  Read them with :meth:`trajectory_stats`.
        Replaces host-side reductions such as ``temp.mean(0)`` (utils/model.py:149) and the posterior mean / variance
        of figure_2.py:75-79.  ``layers``: 'all' or a list of PCLayer indices (0 = ``get_model_representations()``)."""
        assert isinstance(start, int) and start >= 0 and isinstance(stride, int) and stride >= 1
        self._traj_stats_cfg = {"start": start, "stride": stride, "layers": layers}

    def disable_trajectory_stats(self) -> None:
        self._traj_stats_cfg = None

    def trajectory_stats(self):
        """{"count": n, "mean": [per-layer [B, d] tensor | None], "var": [...]} of the last call (unbiased variance)."""
        st = self._traj_stats
        if st is None:
            return None
        n = st["count"]
        var = [None if m2 is None else (m2 / float(max(n - 1, 1))) for m2 in st["m2"]]
        return {"count": n, "mean": list(st["mean"]), "var": var}

    def _get_engine(self):
        if self._engine is None:
            self._engine = NativeEngine()
        return self._engine

    # ======================================================================================
    #  getters & setters (pc_trainer.py:268-461)
    # ======================================================================================
    def get_T(self) -> int:
        return self._T

    def get_model(self) -> nn.Module:
        return self._model

    def get_optimizer_x(self) -> optim.Optimizer:
        return self._optimizer_x

    def get_optimizer_x_lr(self):
        for group in self._optimizer_x.param_groups:
            return group["lr"]

    def set_optimizer_x(self, optimizer_x: optim.Optimizer) -> None:
        assert isinstance(optimizer_x, optim.Optimizer)
        self._optimizer_x = optimizer_x

    def set_optimizer_x_lr(self, lr: float) -> None:
        for group in self._optimizer_x.param_groups:
            group["lr"] = lr

    def get_optimizer_p(self) -> optim.Optimizer:
        return self._optimizer_p

    def set_optimizer_p(self, optimizer_p: optim.Optimizer) -> None:
        assert isinstance(optimizer_p, optim.Optimizer)
        self._optimizer_p = optimizer_p

    def get_model_pc_layers(self) -> typing.Generator[PCLayer, None, None]:
        for module in self._model.modules():
            if isinstance(module, PCLayer):
                yield module

    def get_named_model_pc_layers(self):
        for name, module in self._model.named_modules():
            if isinstance(module, PCLayer):
                yield name, module

    def get_is_model_has_pc_layers(self) -> bool:
        return next(iter(self.get_model_pc_layers()), None) is not None

    def get_model_pc_layers_training(self) -> list:
        return [layer.training for layer in self.get_model_pc_layers()]

    def get_is_model_training(self):
        """True / False when the model and all its PCLayers agree, else None (pc_trainer.py:312-329)."""
        states = self.get_model_pc_layers_training()
        if self._model.training and all(states):
            return True
        if (not self._model.training) and not any(states):
            return False
        return None

    def get_energies(self, is_per_datapoint: bool = False, named_layers: bool = False):
        energies, batch_sizes = {}, []
        for name, layer in self.get_named_model_pc_layers():
            energy = layer.energy_per_datapoint() if is_per_datapoint else layer.energy()
            if energy is not None:
                energies[name] = energy
                batch_sizes.append(energy.size(0) if is_per_datapoint else energy.size())
        assert len(energies) > 0, "You don't have any pc_layers or none of them is holding energy. "
        if (not self.is_disable_warning_energy_from_different_batch_sizes) and \
                batch_sizes.count(batch_sizes[0]) != len(batch_sizes):
            warnings.warn(
                f"You pc_layers hold energy of different batch_sizes: {batch_sizes}.\n"
                "You can disable this warning by setting is_disable_warning_energy_from_different_batch_sizes in "
                "PCTrainer to True.", category=RuntimeWarning)
        return energies if named_layers else list(energies.values())

    def get_model_xs(self, is_warning_x_not_initialized=True):
        for layer in self.get_model_pc_layers():
            x = layer.get_x()
            if x is not None:
                yield x
            elif is_warning_x_not_initialized:
                warnings.warn(
                    "While you are getting x from all pc layers (calling <pc_trainer.get_model_xs()>), "
                    "some pc layers has not been initialized yet (i.e., has x being None). "
                    "This potentially causes bugs. ", category=RuntimeWarning)

    def get_model_parameters(self):
        """All trainable parameters except the latents (pc_trainer.py:368-382)."""
        xs = {id(x) for x in self.get_model_xs(is_warning_x_not_initialized=False)}
        for param in self._model.parameters():
            if id(param) not in xs:
                yield param

    def get_numparameters(self, is_gen=True):
        params = list(self.get_model_parameters())
        if is_gen:
            return sum(p.numel() for i, p in enumerate(params) if i != 0)
        return sum(p.numel() for p in params)

    def get_weights_norms(self):
        weights_abs, mu_abs = [], []
        for par in self.get_model_parameters():
            if par.dim() == 1:
                mu_abs.append(par.abs().mean())
            elif par.dim() == 2:
                weights_abs.append(par.abs().mean())
        return weights_abs, mu_abs

    def get_model_representations(self):
        # "use for unsupervised learning only" (pc_trainer.py:436-438): the first PCLayer of a Sequential
        return self._model[1].get_x()

    def get_model_xs_copy(self):
        return [x.clone().detach().cpu() for x in self.get_model_xs()]

    def get_num_pc_layers(self) -> int:
        return sum(1 for _ in self.get_model_pc_layers())

    def get_least_T(self) -> int:
        return self.get_num_pc_layers() + 1

    # ======================================================================================
    #  optimizers (pc_trainer.py:465-487)
    # ======================================================================================
    def recreate_optimize_x(self) -> None:
        if self._manual_optimizer_x_fn is None:
            self._optimizer_x = self._optimizer_x_fn(self.get_model_xs(), **self._optimizer_x_kwargs)
        else:
            self._optimizer_x = self._manual_optimizer_x_fn()

    def _reset_optimizer_x(self, pc_layers=None) -> None:
        """``recreate_optimize_x`` without building a new torch object when nothing but its state has to go: the
        latents are the same Parameters as before (no re-sampling), so clearing the state and restoring the
        hyper-parameter defaults leaves the optimizer exactly as a freshly constructed one (pc_trainer.py:749-752)."""
        opt = self._optimizer_x
        if self._manual_optimizer_x_fn is not None or opt is None or len(opt.param_groups) != 1:
            return self.recreate_optimize_x()
        held = opt.param_groups[0]["params"]
        xs = list(self.get_model_xs()) if pc_layers is None else [layer._x for layer in pc_layers]
        if len(held) != len(xs) or any(a is not b for a, b in zip(held, xs)):
            return self.recreate_optimize_x()
        opt.state.clear()
        group = opt.param_groups[0]
        for k, v in opt.defaults.items():
            group[k] = v

    def recreate_optimize_p(self) -> None:
        if self._manual_optimizer_p_fn is None:
            self._optimizer_p = self._optimizer_p_fn(self.get_model_parameters(), **self._optimizer_p_kwargs)
        else:
            self._optimizer_p = self._manual_optimizer_p_fn()

    def reset_plot_progress(self):
        self._h = 0
        self._plot_progress = {"key": [], "h": [], "t": [], "value": []}

    # ======================================================================================
    #  train_on_batch (pc_trainer.py:500-1064)
    # ======================================================================================
    def train_on_batch(
        self,
        inputs: typing.Any,
        loss_fn: typing.Callable = None,
        loss_fn_kwargs: dict = {},
        is_sample_x_at_batch_start: bool = True,
        is_reset_optimizer_x_at_batch_start: bool = True,
        is_reset_optimizer_p_at_batch_start: bool = False,
        is_unwrap_inputs: bool = False,
        is_optimize_inputs: bool = False,
        callback_after_backward: typing.Callable = None,
        callback_after_backward_kwargs: dict = {},
        callback_after_t: typing.Callable = None,
        callback_after_t_kwargs: dict = {},
        is_log_progress: bool = True,
        is_return_results_every_t: bool = True,
        is_checking_after_callback_after_t: bool = True,
        debug: dict = {},
        backward_kwargs: dict = {},
        is_clear_energy_after_use: bool = False,
        is_return_outputs: bool = False,
        is_return_representations: bool = False,
        is_return_xs: bool = False,
        is_return_batchelement_loss: bool = False,
    ):
        """Run T inference steps (and the parameter updates they trigger) on one batch.

        Returns the reference's results dict: ``loss`` / ``energy`` / ``overall`` lists (one entry
        per step, or only the last step when ``is_return_results_every_t=False``; ``loss`` stays
        empty without a ``loss_fn``) plus ``outputs`` / ``representations`` / ``xs`` when asked.
        Entry t describes the state at the START of step t (pc_trainer.py:768-842).
        """
        self.inputs = inputs
        # ---- sanitise exactly like the reference (pc_trainer.py:608-656) ----
        if loss_fn is not None:
            assert callable(loss_fn)
        assert isinstance(loss_fn_kwargs, dict)
        assert isinstance(is_sample_x_at_batch_start, bool)
        assert isinstance(is_reset_optimizer_x_at_batch_start, bool)
        assert isinstance(is_reset_optimizer_p_at_batch_start, bool)
        assert isinstance(is_unwrap_inputs, bool)
        if is_unwrap_inputs:
            assert isinstance(inputs, (tuple, list, dict))
        assert isinstance(is_optimize_inputs, bool)
        if is_optimize_inputs:
            assert self.get_is_model_has_pc_layers(), "<is_optimize_inputs> should only work with models with <PCLayer>. "
            assert not is_unwrap_inputs
        if callback_after_backward is not None:
            assert callable(callback_after_backward)
        assert isinstance(callback_after_backward_kwargs, dict)
        if callback_after_t is not None:
            assert callable(callback_after_t)
        assert isinstance(callback_after_t_kwargs, dict)
        assert isinstance(is_log_progress, bool)
        assert isinstance(is_return_results_every_t, bool)
        assert isinstance(debug, dict)
        assert isinstance(is_return_outputs, bool)
        assert isinstance(is_return_representations, bool)
        assert isinstance(is_return_xs, bool)

        if is_log_progress:
            _slow_down_warning("PCTrainer.train_on_batch", "is_log_progress", "False")
        if self._is_plot_progress:
            raise NotImplementedError(
                "plot_progress needs matplotlib/seaborn and a blocking input() (pc_trainer.py:985-1060); it is "
                "presentation code outside the B200 hot path.  Pass plot_progress_at=[] like every factory in "
                "utils/training_evaluation.py does.")
        if is_return_results_every_t:
            _slow_down_warning("PCTrainer.train_on_batch", "is_return_results_every_t", "False")

        # features of the generic autograd loop that the fused kernels do not implement
        unsupported = []
        if is_unwrap_inputs:
            unsupported.append("is_unwrap_inputs")
        if is_optimize_inputs:
            unsupported.append("is_optimize_inputs")
        if self._loss_x_fn is not None:
            unsupported.append("loss_x_fn")
        if self._loss_inputs_fn is not None:
            unsupported.append("loss_inputs_fn")
        if backward_kwargs:
            unsupported.append("backward_kwargs")
        if is_return_batchelement_loss:
            unsupported.append("is_return_batchelement_loss")
        if unsupported:
            raise NotImplementedError(
                "not implemented by the fused sm_100a path (no script of the reference uses them): "
                + ", ".join(unsupported))
        if not isinstance(inputs, torch.Tensor) or inputs.dim() != 2:
            raise NotImplementedError("inputs must be a [batch, features] tensor")

        netp = P.compile_net(self._model)
        # the plan's PCLayers ARE the model's (flat module list): same check as get_is_model_training() without a walk
        assert self._model.training and all(layer.training for layer in netp.pc_layers), (
            "PCLayer behaves differently in train and eval modes, like Dropout or Batch Normalization. "
            "Thus, call model.eval() before evaluation and model.train() before train. "
            "Make sure your model is in train mode before calling <train_on_batch()>. "
            "It can be done by calling <model.train()>. "
            "Do remember switching your model back to eval mode before evaluating it by calling <model.eval()>. ")
        B = int(inputs.shape[0])
        if inputs.shape[1] != netp.d_in:
            raise RuntimeError(f"inputs has {inputs.shape[1]} features, the first Linear expects {netp.d_in}")
        device = netp.linears[0].weight.device
        top = P.classify_loss(loss_fn, loss_fn_kwargs, B, netp.d_out, device)
        target = loss_fn_kwargs.get(top.target_key) if top.target_key is not None else None
        T = self._T

        # ---- at batch start: sample x (one PyTorch forward), reset optimizers (pc_trainer.py:717-766)
        self._start_of_batch(netp, inputs, is_sample_x_at_batch_start, is_reset_optimizer_x_at_batch_start,
                             is_reset_optimizer_p_at_batch_start)

        self._precision = self._resolve_precision(inputs, netp, top)
        langevin = P.classify_callback_after_t(callback_after_t, callback_after_t_kwargs, self)
        x_opt = self._classify_optimizer_x()
        fused = (
            (callback_after_t is None or langevin is not None)
            and callback_after_backward is None
            and self._early_stop_condition.strip() == "False"
            and self._x_lr_discount == 1.0 and self._x_lr_amplifier == 1.0
            and x_opt is not None
            and not (langevin is not None and x_opt["kind"] != N.OPT_SGD)
        )
        ctx = dict(netp=netp, top=top, target=target, inputs=inputs, B=B, T=T, device=device,
                   every_t=is_return_results_every_t, want_outputs=is_return_outputs,
                   want_reps=is_return_representations, want_xs=is_return_xs)
        if fused:
            rec = self._run_fused(ctx, x_opt, langevin)
        else:
            rec = self._run_stepwise(ctx, loss_fn, callback_after_backward, callback_after_backward_kwargs,
                                     callback_after_t, callback_after_t_kwargs, is_checking_after_callback_after_t)
        self._noise_epoch += 1
        self._install_lazy_energies(netp, inputs)
        if is_clear_energy_after_use:
            for layer in netp.pc_layers:
                layer.clear_energy()
        if is_log_progress:
            self._print_progress(rec, loss_fn is not None)
        return self._build_results(ctx, rec, loss_fn is not None)

    # --------------------------------------------------------------------------------------
    def _start_of_batch(self, netp, inputs, sample_x, reset_x, reset_p):
        need_forward = sample_x
        for layer in netp.pc_layers:
            x = layer.get_x()
            if x is None or x.shape[0] != inputs.shape[0] or x.device != inputs.device:
                need_forward = True      # the layer will warn and sample on its own (pc_layer.py:185-218)
        if sample_x and all(_is_shape_only_sampler(layer._sample_x_fn) for layer in netp.pc_layers):
            # the library samplers (utils/model.py:8-15) only look at the SHAPE of mu: draw the latents layer by layer in
            # module order (same generator calls as the t=0 forward of pc_trainer.py:717-733) without running the model
            lin0 = netp.linears[0].weight
            B = inputs.shape[0]
            with torch.no_grad():
                for l, layer in enumerate(netp.pc_layers):
                    mu_like = torch.empty(B, netp.dims[l], device=lin0.device, dtype=lin0.dtype)
                    fresh = layer._sample_x_fn({"mu": mu_like, "x": layer._x})
                    layer._x = nn.Parameter(fresh.to(lin0.device), True)
                    layer._is_sample_x = False
            need_forward = False
        elif sample_x:
            for layer in netp.pc_layers:
                layer.set_is_sample_x(True)
        if need_forward:
            with torch.no_grad():
                self._model(inputs)
        for l, layer in enumerate(netp.pc_layers):
            x = layer.get_x()
            if x.dim() != 2 or x.shape[1] != netp.dims[l]:
                raise RuntimeError(f"latent of PCLayer {l} has shape {tuple(x.shape)}, expected [B, {netp.dims[l]}]")
            if x.dtype != torch.float32 or not x.is_contiguous():
                layer._x = nn.Parameter(x.detach().to(torch.float32).contiguous(), True)
        if sample_x or self._optimizer_x is None:
            self.recreate_optimize_x()
            self._adam = None
        elif reset_x:
            self._reset_optimizer_x(netp.pc_layers)
            self._adam = None
        if reset_p:
            self.recreate_optimize_p()

    def _classify_optimizer_x(self):
        """Map the torch optimizer the user configured for x onto the kernel's SGD / Adam step."""
        opt = self._optimizer_x
        if len(opt.param_groups) != 1:
            return None
        g = opt.param_groups[0]
        if type(opt) is optim.SGD:
            if g.get("momentum", 0) != 0 or g.get("weight_decay", 0) != 0 or g.get("nesterov", False) or \
                    g.get("dampening", 0) != 0 or g.get("maximize", False):
                return None
            return {"kind": N.OPT_SGD, "lr": float(g["lr"]), "lr0": float(opt.defaults["lr"])}
        if type(opt) is optim.Adam:
            if g.get("weight_decay", 0) != 0 or g.get("amsgrad", False) or g.get("maximize", False):
                return None
            return {"kind": N.OPT_ADAM, "lr": float(g["lr"]), "lr0": float(opt.defaults["lr"]),
                    "betas": tuple(float(b) for b in g["betas"]), "eps": float(g["eps"])}
        return None

    def _buffer(self, role, shape, dtype, device):
        hit = self._buffer_views.get(role)
        if hit is not None and hit[0] == shape and hit[1].dtype == dtype and hit[1].device == device:
            return hit[1]
        n = 1
        for s in shape:
            n *= int(s)
        buf = self._buffers.get(role)
        if buf is None or buf.dtype != dtype or buf.device != device or buf.numel() < n + 1024:
            buf = torch.zeros(n + 1024, dtype=dtype, device=device)     # 1024 elements of slack (mcpc_save_layout)
            self._buffers[role] = buf
            self._buffer_views.pop(role, None)
        hit = self._buffer_views.get(role)
        if hit is None or hit[0] != shape:
            hit = self._buffer_views[role] = (shape, buf[:n].view(*shape))
        return hit[1]

    def _save_layout(self, netp, top):
        eng = self._get_engine()
        if hasattr(eng, "save_layout"):
            return eng.save_layout(netp, top, self._precision)
        return netp.SD + netp.d_out, netp.SD, torch.float32

    def _inputs_or_none(self, inputs):
        """``None`` when inputs are all zero (every script of the reference): Linear_0 then only
        contributes its bias and its weight gradient is exactly zero."""
        key = (inputs.data_ptr(), inputs._version, tuple(inputs.shape))
        if self._zero_inputs_cache is None or self._zero_inputs_cache[0] != key:
            self._zero_inputs_cache = (key, not bool(inputs.any()))
        if self._zero_inputs_cache[1]:
            return None
        return inputs.detach().to(torch.float32).contiguous()

    def _param_tensors(self, netp):
        """Detached views of the Linear weights / biases; rebuilt only when a Parameter object (or its storage) changed."""
        lins = netp.linears
        hit = self._param_cache
        if hit is not None and len(hit[0]) == len(lins):
            ok = True
            for lin, (pw, pb, wp, bp) in zip(lins, hit[0]):
                w, bias = lin.weight, lin.bias
                if w is not pw or bias is not pb or w.data_ptr() != wp or (bias is not None and bias.data_ptr() != bp):
                    ok = False
                    break
            if ok:
                return hit[1], hit[2]
        W = [lin.weight.detach() for lin in lins]
        b = [None if lin.bias is None else lin.bias.detach() for lin in lins]
        self._param_cache = ([(lin.weight, lin.bias, lin.weight.data_ptr(), None if lin.bias is None else lin.bias.data_ptr())
                              for lin in lins], W, b)
        return W, b

    # --------------------------------------------------------------------------------------
    #  parameter-gradient bookkeeping: the reference's zero_grad rules as step-index arithmetic
    # --------------------------------------------------------------------------------------
    def _is_zero_grad_step(self, t):
        acc = self._accumulate_p_at
        return (t in self._update_p_set and t not in self._acc_set) or (len(acc) > 0 and t == acc[0])

    def _refresh_schedule_sets(self):
        """Set views of the three step lists, rebuilt only when a list object was replaced or resized."""
        key = (id(self._update_p_at), len(self._update_p_at), id(self._update_x_at), len(self._update_x_at),
               id(self._accumulate_p_at), len(self._accumulate_p_at))
        if getattr(self, "_sched_key", None) != key:
            self._sched_key = key
            self._update_p_set = set(self._update_p_at)
            self._update_x_set = set(self._update_x_at)
            self._acc_set = set(self._accumulate_p_at)
            self._seg_cache = {}

    def _segments_cached(self, T, split_last):
        """(segments, zero_grad steps inside each segment) for the current schedule."""
        hit = self._seg_cache.get((T, split_last))
        if hit is None:
            segs = self._segments(T, split_last)
            hit = (segs, [[t for t in range(t0, t1) if self._is_zero_grad_step(t)] for (t0, t1) in segs])
            self._seg_cache[(T, split_last)] = hit
        return hit

    def _segments(self, T, split_last):
        """Cut [0, T) where the kernel arguments change: after every p-update step, where the
        ``update_x_at`` membership flips, and (optionally) before the last step."""
        segs = []
        t0 = 0
        for t in range(T):
            end_here = (t in self._update_p_set) or (t == T - 1) or \
                ((t + 1 in self._update_x_set) != (t in self._update_x_set)) or (split_last and t == T - 2)
            if end_here:
                segs.append((t0, t + 1))
                t0 = t + 1
        return segs

    def _ensure_flat_grads(self, netp, zero):
        """All W/b ``.grad`` are views into ONE flat fp32 buffer (a single all-reduce covers them)."""
        flat = self._flat_grad
        if flat is not None and len(flat) == 5 and flat[2] is netp:
            # hit path (it sits in front of the inference launch): same plan, every .grad still IS its view of the buffer
            views = flat[3]
            ok = flat[1].numel() == flat[4] + 4 * self._T
            gW, gb = [], []
            if ok:
                i = 0
                for lin in netp.linears:
                    w, bias = lin.weight, lin.bias
                    if i >= len(views) or w is not views[i][0] or w.grad is not views[i][1]:
                        ok = False
                        break
                    gW.append(views[i][1])
                    i += 1
                    if bias is None:
                        gb.append(None)
                        continue
                    if i >= len(views) or bias is not views[i][0] or bias.grad is not views[i][1]:
                        ok = False
                        break
                    gb.append(views[i][1])
                    i += 1
                ok = ok and i == len(views)
            if ok:
                if zero:
                    flat[1].zero_()
                return flat[1], gW, gb
        params = []
        for lin in netp.linears:
            params.append(lin.weight)
            if lin.bias is not None:
                params.append(lin.bias)
        total = sum(p.numel() for p in params)
        # tail: room for the [2, T] per-step scalars as fp32 (hi, lo) pairs, so that a data-parallel learning call needs
        # ONE all-reduce for the weight gradients and the results dict together (see _p_step)
        tail = 4 * self._T
        self._flat_total = total
        sig = tuple(id(p) for p in params)
        flat = self._flat_grad
        fresh = flat is None or flat[0] != sig or flat[1].numel() != total + tail or flat[1].device != params[0].device
        if fresh:
            buf = torch.zeros(total + tail, dtype=torch.float32, device=params[0].device)
            o = 0
            for p in params:
                n = p.numel()
                view = buf[o:o + n].view_as(p)
                if p.grad is not None and not zero:
                    view.copy_(p.grad)
                p.grad = view
                o += n
        else:
            buf = flat[1]
            o = 0
            for p in params:
                n = p.numel()
                view = buf[o:o + n].view_as(p)
                if p.grad is None:
                    view.zero_()
                    p.grad = view
                elif p.grad.data_ptr() != view.data_ptr():
                    if not zero:
                        view.copy_(p.grad)
                    p.grad = view
                o += n
            if zero:
                buf.zero_()
        self._flat_grad = (sig, buf, netp, [(p_, p_.grad) for p_ in params], total)
        gW = [lin.weight.grad for lin in netp.linears]
        gb = [None if lin.bias is None else lin.bias.grad for lin in netp.linears]
        return buf, gW, gb

    def _global_batch(self, B):
        if self._dp_group is None:
            return B
        import torch.distributed as dist
        return B * dist.get_world_size(self._dp_group)

    def _chain_offset(self, B):
        if self._dp_group is None:
            return 0
        if self._dp_chain_offset is not None:
            return int(self._dp_chain_offset)
        import torch.distributed as dist
        return dist.get_rank(self._dp_group) * B

    def _p_step(self, flat, B, scalars=None):
        """pc_trainer.py:904-914: normalise ``.grad`` in place, then the user's optimizer.  Data-parallel: ONE all-reduce
        (sum) over the flat gradient buffer; when ``scalars`` (the [2, T] fp64 energy / loss sums of this rank) is given
        they ride in the buffer's tail as fp32 (hi, lo) pairs and the reduced values are returned."""
        total = self._flat_total
        reduced = None
        if self._dp_group is not None:
            import torch.distributed as dist
            n_tail = 0
            if scalars is not None and 2 * scalars.numel() <= flat.numel() - total:
                n_tail = 2 * scalars.numel()
                hi = scalars.to(torch.float32)
                lo = (scalars - hi.to(torch.float64)).to(torch.float32)
                flat[total:total + n_tail].view(2, -1).copy_(torch.stack([hi.reshape(-1), lo.reshape(-1)]))
            dist.all_reduce(flat[:total + n_tail], op=dist.ReduceOp.SUM, group=self._dp_group)
            if n_tail:
                pair = flat[total:total + n_tail].view(2, -1).to(torch.float64)
                reduced = (pair[0] + pair[1]).view_as(scalars)
        Bg = self._global_batch(B)
        n_acc = len(self._accumulate_p_at)
        norm = float(n_acc * Bg if n_acc > 0 else Bg)
        if self._fused_p_optimizer and self._fused_p_step(1.0 / norm):
            return reduced
        flat[:total].div_(norm)
        self._optimizer_p.step()
        return reduced

    def _fused_p_step(self, inv_norm) -> bool:
        """normalisation + ``optimizer_p.step()`` as ONE kernel per param group (``mcpc_p_step``), in
        place on the torch optimizer's own state.  Returns False -- the caller then runs the plain torch path -- for
        anything but a stock ``optim.SGD`` / ``optim.Adam`` over contiguous fp32 CUDA parameters."""
        eng = self._get_engine()
        opt = self._optimizer_p
        if not hasattr(eng, "p_step") or type(opt) not in (optim.SGD, optim.Adam):
            return False
        plans = []
        for group in opt.param_groups:
            if group.get("maximize", False) or group.get("differentiable", False) or group.get("fused", False) or \
                    group.get("capturable", False) or group.get("amsgrad", False) or group.get("decoupled_weight_decay", False):
                return False
            params = [q for q in group["params"] if q.grad is not None]
            if not params:
                continue
            if len(params) > N.MAX_PTENSORS:
                return False
            for q in params:
                if not (q.is_cuda and q.dtype == torch.float32 and q.is_contiguous() and q.grad.is_cuda and
                        q.grad.dtype == torch.float32 and q.grad.is_contiguous()) or q.grad.is_sparse:
                    return False
            lr = float(group["lr"])
            wd = float(group.get("weight_decay", 0.0))
            if type(opt) is optim.SGD:
                mom = float(group.get("momentum", 0.0))
                bufs, first = [None] * len(params), 0
                if mom != 0.0:
                    have = ["momentum_buffer" in opt.state[q] and opt.state[q]["momentum_buffer"] is not None for q in params]
                    if any(have) and not all(have):
                        return False
                    first = 0 if all(have) else 1
                plans.append(dict(kind=N.OPT_SGD, params=params, lr=lr, wd=wd, momentum=mom, first=first,
                                  dampening=float(group.get("dampening", 0.0)), nesterov=bool(group.get("nesterov", False))))
            else:
                b1, b2 = (float(v) for v in group["betas"])
                for q in params:
                    st = opt.state[q]
                    if len(st) and (not torch.is_tensor(st.get("step")) or st["step"].is_cuda):
                        return False
                plans.append(dict(kind=N.OPT_ADAM, params=params, lr=lr, wd=wd, beta1=b1, beta2=b2, eps=float(group["eps"])))
        for pl in plans:
            params = pl["params"]
            if pl["kind"] == N.OPT_SGD:
                s1 = [None] * len(params)
                if pl["momentum"] != 0.0:
                    for i, q in enumerate(params):
                        st = opt.state[q]
                        if pl["first"]:
                            st["momentum_buffer"] = torch.empty_like(q, memory_format=torch.preserve_format)
                        s1[i] = st["momentum_buffer"]
                eng.p_step(N.OPT_SGD, params, [q.grad for q in params], s1, [None] * len(params), inv_norm, lr=pl["lr"],
                           weight_decay=pl["wd"], momentum=pl["momentum"], dampening=pl["dampening"],
                           nesterov=pl["nesterov"], first_step=pl["first"])
            else:
                steps = set()
                for q in params:
                    st = opt.state[q]
                    if len(st) == 0:           # torch/optim/adam.py::_init_group
                        st["step"] = torch.tensor(0.0, dtype=torch.float32)
                        st["exp_avg"] = torch.zeros_like(q, memory_format=torch.preserve_format)
                        st["exp_avg_sq"] = torch.zeros_like(q, memory_format=torch.preserve_format)
                    st["step"] += 1
                    steps.add(int(st["step"].item()))
                for k in sorted(steps):       # parameters normally share one step count: one launch
                    sel = [q for q in params if int(opt.state[q]["step"].item()) == k]
                    eng.p_step(N.OPT_ADAM, sel, [q.grad for q in sel], [opt.state[q]["exp_avg"] for q in sel],
                               [opt.state[q]["exp_avg_sq"] for q in sel], inv_norm, lr=pl["lr"], weight_decay=pl["wd"],
                               beta1=pl["beta1"], beta2=pl["beta2"], eps=pl["eps"], step=k)
        return True

    # --------------------------------------------------------------------------------------
    def _run_fused(self, ctx, x_opt, langevin):
        netp, top, B, T, device = ctx["netp"], ctx["top"], ctx["B"], ctx["T"], ctx["device"]
        eng = self._get_engine()
        self._refresh_schedule_sets()
        inputs_dev = self._inputs_or_none(ctx["inputs"])
        target = ctx["target"]
        if target is not None:
            target = target.detach().to(torch.float32).contiguous()
        xs = [layer.get_x().data for layer in netp.pc_layers]
        # one allocation, one D2H at the end; the launches below cover [0, T) and each OVERWRITES its slice (mcpc_b200.h)
        scalars = torch.empty(2, T, dtype=torch.float64, device=device)
        energy, loss = scalars[0], scalars[1]

        want_traj = ctx["want_outputs"] or ctx["want_reps"] or ctx["want_xs"]
        every_t = ctx["every_t"]
        # recorded steps: every step (reference), or start, start+stride, ... (set_trajectory_stride); last step only
        # when is_return_results_every_t=False
        k_rec, s_rec = (self._traj_stride, self._traj_start) if every_t else (1, 0)
        n_rec = (max(0, -(-(T - s_rec) // k_rec)) if every_t else 1)
        traj_x = [None] * netp.L
        traj_out = None
        if want_traj and n_rec > 0:
            for l in range(netp.L):
                need = ctx["want_xs"] or (ctx["want_reps"] and l == 0) or \
                    (ctx["want_outputs"] and netp.d_out == 0 and l == netp.L - 1)
                if need:
                    traj_x[l] = torch.empty(n_rec, B, netp.dims[l], dtype=torch.float32, device=device)
            if ctx["want_outputs"] and netp.d_out > 0:
                traj_out = torch.empty(n_rec, B, netp.d_out, dtype=torch.float32, device=device)
        # on-device statistics (enable_trajectory_stats): either over the rings above (same thinning), or -- when the
        # trajectory itself is not wanted -- over a bounded ring that is folded into the accumulators chunk by chunk
        stats = self._prepare_traj_stats(netp, B, T, device, traj_x, every_t, k_rec, s_rec)
        if stats is not None and stats["own_ring"]:
            k_rec, s_rec = stats["stride"], stats["start"]

        # Adam state of the fused x-optimizer persists while the torch optimizer object does
        adam_m = adam_v = None
        adam_step0 = 0
        if x_opt["kind"] == N.OPT_ADAM:
            if self._adam is None or self._adam["B"] != B:
                self._adam = {"B": B, "step": 0,
                              "m": [torch.zeros_like(x) for x in xs], "v": [torch.zeros_like(x) for x in xs]}
            adam_m, adam_v, adam_step0 = self._adam["m"], self._adam["v"], self._adam["step"]

        noise_mode, noise_scale, noise_all = N.NOISE_NONE, 0.0, None
        if langevin is not None and langevin.var > 0.0:
            noise_scale = float(np.sqrt(langevin.var / x_opt["lr0"]))      # utils/model.py:43
            if self._supplied_noise is not None:
                noise_all = self._supplied_noise.detach().to(device=device, dtype=torch.float32).contiguous()
                if tuple(noise_all.shape) != (T, B, netp.SD):
                    raise RuntimeError(f"supplied noise must be [T={T}, B={B}, {netp.SD}]")
                noise_mode = N.NOISE_SUPPLIED
                self._supplied_noise = None
            else:
                noise_mode = N.NOISE_PHILOX
        seed = (self._seed + 0x9E3779B97F4A7C15 * self._noise_epoch) & 0xFFFFFFFFFFFFFFFF

        W, b = self._param_tensors(netp)
        streaming = hasattr(eng, "infer_mode") and \
            eng.infer_mode(netp, top, B, self._precision) == N.MODE_STREAMING_BF16
        # the native library runs the weight update of the saved steps inside mcpc_infer when it gets the accumulators
        # (resident bf16: on the SMs the inference kernel leaves idle, while it runs); the CPU test double does not
        fused_dw = streaming or (hasattr(eng, "infer_fuses_weight_grad") and
                                 eng.infer_fuses_weight_grad(netp, top, B, self._precision, inputs_dev is not None))
        later_p_updates = sorted(self._update_p_set)
        segs, seg_zero_steps = self._segments_cached(T, split_last=(want_traj and not every_t))
        flat = None
        n_launch = 0
        host_scalars = None
        scalars_reduced = False
        for si, (t0, t1) in enumerate(segs):
            ends_with_p = (t1 - 1) in self._update_p_set
            need_grads = self._keep_unused_param_grads or any(u >= t0 for u in later_p_updates)
            win_begin = None
            flat_ready = False
            zero_steps = []
            if need_grads:
                zero_steps = seg_zero_steps[si]
                win_begin = zero_steps[-1] if zero_steps else t0
                if fused_dw:            # mcpc_infer accumulates dW itself: the (zeroed) buffers must exist up front
                    flat, gW, gb = self._ensure_flat_grads(netp, zero=bool(zero_steps))
                    flat_ready = True
                else:
                    gW = gb = None      # separate weight_grad call: (re)zeroed after the inference launch
            # the window may be cut further so the saved operands fit the scratch budget
            g_w, f_w, s_dtype = self._save_layout(netp, top)
            row_bytes = (4 if s_dtype == torch.float32 else 2) * B * (g_w + f_w)
            max_save = max(1, self._save_budget_bytes // max(row_bytes, 1))
            cuts = [(t0, t1)]
            if need_grads and not streaming and (t1 - win_begin) > max_save:
                cuts = [(t0, win_begin + max_save)] if win_begin + max_save > t0 else []
                s = win_begin + max_save
                while s < t1:
                    cuts.append((s, min(t1, s + max_save)))
                    s += max_save
            recording = (want_traj or stats is not None) and (every_t or stats is not None)
            max_rec = stats["ring_len"] if (stats is not None and stats["own_ring"]) else None
            cuts = self._split_cuts_for_recording(cuts, k_rec, s_rec, max_rec) if recording else [(c0, c1, None) for (c0, c1) in cuts]
            for (c0, c1, r0) in cuts:
                n = c1 - c0
                sb = se = 0
                save_g = save_f = None
                if need_grads and c1 > win_begin:
                    sb, se = max(win_begin, c0) - c0, n
                    if not streaming:
                        save_g = self._buffer("save_g", (se - sb, B, g_w), s_dtype, device)
                        save_f = self._buffer("save_f", (se - sb, B, f_w), s_dtype, device)
                # trajectory slices of this cut: r0 = index of its first recorded step (None: nothing recorded here)
                n_r = 0
                tx_cut = [None] * netp.L
                to_cut = None
                if want_traj and not every_t:
                    if c1 == T:                                        # last step only (split_last made it its own cut)
                        tx_cut, to_cut = list(traj_x), traj_out
                elif r0 is not None:
                    n_r = -(-n // k_rec)
                    if stats is not None and stats["own_ring"]:
                        tx_cut = [None if ring is None else ring[:n_r] for ring in stats["rings"]]
                    else:
                        tx_cut = [None if tx is None else tx[r0:r0 + n_r] for tx in traj_x]
                        to_cut = None if traj_out is None else traj_out[r0:r0 + n_r]
                call = InferCall(
                    plan=netp, top=top, energy_coefficient=self._energy_coefficient, B=B, W=W, b=b, x=xs,
                    inputs=inputs_dev, target=target, energy=energy[c0:c1], loss=loss[c0:c1], n_steps=n, t_begin=c0,
                    optimizer=x_opt["kind"], update_x=(c0 in self._update_x_set), lr=x_opt["lr"],
                    betas=x_opt.get("betas", (0.9, 0.999)), adam_eps=x_opt.get("eps", 1e-8),
                    adam_step0=adam_step0, adam_m=adam_m, adam_v=adam_v,
                    noise_mode=noise_mode, noise=None if noise_all is None else noise_all[c0:c1],
                    noise_scale=noise_scale, seed=seed, chain_offset=self._chain_offset(B),
                    traj_x=tx_cut, traj_out=to_cut,
                    traj_every=k_rec if every_t or stats is not None else 1,
                    save_g=save_g, save_f=save_f, save_begin=sb, save_end=se,
                    precision=self._precision,
                    gW=gW if (fused_dw and need_grads and se > sb) else None,
                    gb=gb if (fused_dw and need_grads and se > sb) else None)
                eng.infer(call)
                n_launch += 1
                if stats is not None and n_r > 0:
                    self._fold_traj_stats(eng, stats, tx_cut, n_r)
                if c1 == T and not (ends_with_p and self._dp_group is not None and self._dp_single_collective):
                    # the per-step scalars are final here: start their read-back on a side stream now, so that the
                    # host gets them while the weight-gradient / optimizer_p kernels of this call are still running
                    # (data-parallel learning calls: they travel with the gradient all-reduce instead, see below)
                    host_scalars = self._start_scalar_readback(scalars)
                if x_opt["kind"] == N.OPT_ADAM and (c0 in self._update_x_set):
                    adam_step0 += n
                    self._adam["step"] = adam_step0
                if need_grads and not flat_ready:
                    flat, gW, gb = self._ensure_flat_grads(netp, zero=bool(zero_steps))
                    flat_ready = True
                if save_g is not None and not fused_dw:
                    eng.weight_grad(netp, top, self._energy_coefficient, B, se - sb, save_g, save_f, inputs_dev,
                                    gW, gb, self._precision)
                    n_launch += 1
            if ends_with_p:
                last = (t1 == T) and self._dp_group is not None and self._dp_single_collective
                reduced = self._p_step(flat, B, scalars=scalars if last else None)
                if last:
                    if reduced is not None:
                        scalars, scalars_reduced = reduced, True
                    host_scalars = self._start_scalar_readback(scalars, reduced=scalars_reduced)
        self.last_call_info = {"mode": "fused", "launches": n_launch, "segments": len(segs),
                               "noise": noise_mode, "precision": self._precision}
        if stats is not None:
            self._traj_stats = {"count": stats["count"], "mean": stats["mean"], "m2": stats["m2"]}
        steps = list(range(s_rec, T, k_rec)) if every_t else [T - 1]
        self.last_trajectories = {"x": traj_x, "out": traj_out, "steps": steps} if want_traj else None
        return {"energy": energy, "loss": loss, "scalars": scalars, "scalars_reduced": scalars_reduced,
                "host_scalars": host_scalars, "traj_x": traj_x, "traj_out": traj_out, "n_rec": n_rec}

    # --------------------------------------------------------------------------------------
    #  SURVEY 8(f) N2 helpers: thinned recording and on-device statistics
    # --------------------------------------------------------------------------------------
    @staticmethod
    def _split_cuts_for_recording(cuts, k, s0, max_rec):
        """Split [c0, c1) launches so that every launch that records starts ON a recorded step (the kernels record
        local steps 0, k, 2k, ...) and records at most ``max_rec`` steps.  Returns (c0, c1, r0) with r0 the global
        record index of the launch's first recorded step, or None when the launch records nothing."""
        out = []
        for (c0, c1) in cuts:
            first = s0 if c0 <= s0 else s0 + -(-(c0 - s0) // k) * k        # first recorded step >= c0
            if first >= c1:
                out.append((c0, c1, None))
                continue
            if first > c0:
                out.append((c0, first, None))
            t = first
            while t < c1:
                e = c1 if max_rec is None else min(c1, t + max_rec * k)
                out.append((t, e, (t - s0) // k))
                t = e
        return out

    def _prepare_traj_stats(self, netp, B, T, device, traj_x, every_t, k_rec, s_rec):
        cfg = self._traj_stats_cfg
        self._traj_stats = None
        if cfg is None:
            return None
        if not every_t and any(t is not None for t in traj_x):
            raise ValueError("trajectory statistics cannot be combined with last-step-only trajectories "
                             "(is_return_results_every_t=False together with is_return_xs / representations)")
        layers = list(range(netp.L)) if cfg["layers"] == "all" else [int(l) for l in cfg["layers"]]
        shared = every_t and any(traj_x[l] is not None for l in layers)
        if shared:
            if (cfg["stride"], cfg["start"]) != (k_rec, s_rec) or any(traj_x[l] is None for l in layers):
                raise ValueError("enable_trajectory_stats(start, stride) must equal set_trajectory_stride(stride, start) "
                                 "when the same layers are also returned as trajectories (the kernels record once)")
        st = {"own_ring": not shared, "start": cfg["start"], "stride": cfg["stride"], "layers": layers, "count": 0,
              "mean": [None] * netp.L, "m2": [None] * netp.L, "rings": [None] * netp.L, "ring_len": 0}
        n_total = max(0, -(-(T - cfg["start"]) // cfg["stride"]))
        if n_total == 0:
            return None
        for l in layers:
            st["mean"][l] = torch.zeros(B, netp.dims[l], dtype=torch.float32, device=device)
            st["m2"][l] = torch.zeros(B, netp.dims[l], dtype=torch.float32, device=device)
        if not shared:
            widest = max(netp.dims[l] for l in layers)
            st["ring_len"] = int(max(1, min(n_total, self._traj_ring_bytes // max(1, 4 * B * widest))))
            for l in layers:
                st["rings"][l] = self._buffer(f"traj_ring{l}", (st["ring_len"], B, netp.dims[l]), torch.float32, device)
        return st

    def _fold_traj_stats(self, eng, st, tx_cut, n_r):
        for l in st["layers"]:
            ring = tx_cut[l]
            if ring is None:
                continue
            eng.traj_stats(ring, n_r, st["count"], st["mean"][l], st["m2"][l])
        st["count"] += n_r

    # --------------------------------------------------------------------------------------
    def _run_stepwise(self, ctx, loss_fn, cb_bwd, cb_bwd_kwargs, cb_t, cb_t_kwargs, check_after_cb):
        """One launch per step with ``x.grad`` materialised; callbacks, the torch ``optimizer_x`` and
        the dynamic x-lr rule run in Python between launches (pc_trainer.py:845-926)."""
        netp, top, B, T, device = ctx["netp"], ctx["top"], ctx["B"], ctx["T"], ctx["device"]
        eng = self._get_engine()
        self._update_p_set = set(self._update_p_at)
        self._update_x_set = set(self._update_x_at)
        self._acc_set = set(self._accumulate_p_at)
        inputs_dev = self._inputs_or_none(ctx["inputs"])
        target = ctx["target"]
        if target is not None:
            target = target.detach().to(torch.float32).contiguous()
        energy = torch.zeros(T, dtype=torch.float64, device=device)
        loss = torch.zeros(T, dtype=torch.float64, device=device)
        every_t = ctx["every_t"]
        n_rec = T if every_t else 1
        traj_x = [None] * netp.L
        traj_out = None
        for l in range(netp.L):
            if ctx["want_xs"] or (ctx["want_reps"] and l == 0) or \
                    (ctx["want_outputs"] and netp.d_out == 0 and l == netp.L - 1):
                traj_x[l] = torch.empty(n_rec, B, netp.dims[l], dtype=torch.float32, device=device)
        if ctx["want_outputs"] and netp.d_out > 0:
            traj_out = torch.empty(n_rec, B, netp.d_out, dtype=torch.float32, device=device)
        is_dynamic_x_lr = (self._x_lr_discount < 1.0) or (self._x_lr_amplifier > 1.0)
        overalls = []
        params = [layer.get_x() for layer in netp.pc_layers]
        for p_ in params:
            if p_.grad is None:
                p_.grad = torch.zeros_like(p_.data)
        n_launch = 0
        flat = None
        t_done = T
        for t in range(T):
            W, b = self._param_tensors(netp)
            xs = [p_.data for p_ in params]
            need_grads = self._keep_unused_param_grads or any(u >= t for u in self._update_p_set) or \
                (self._early_stop_condition.strip() != "False" and self._update_p_at_early_stop)
            save_g = save_f = None
            if need_grads:
                g_w, f_w, s_dtype = self._save_layout(netp, top)
                save_g = self._buffer("save_g", (1, B, g_w), s_dtype, device)
                save_f = self._buffer("save_f", (1, B, f_w), s_dtype, device)
            rec = every_t or t == T - 1
            ri = t if every_t else 0
            call = InferCall(
                plan=netp, top=top, energy_coefficient=self._energy_coefficient, B=B, W=W, b=b, x=xs,
                inputs=inputs_dev, target=target, energy=energy[t:t + 1], loss=loss[t:t + 1], n_steps=1, t_begin=t,
                optimizer=N.OPT_SGD, update_x=False, lr=0.0, x_grad=[p_.grad for p_ in params],
                traj_x=[None if (tx is None or not rec) else tx[ri:ri + 1] for tx in traj_x],
                traj_out=None if (traj_out is None or not rec) else traj_out[ri:ri + 1],
                save_g=save_g, save_f=save_f, save_begin=0, save_end=1 if need_grads else 0,
                precision=self._precision)
            eng.infer(call)
            n_launch += 1
            # early stop is decided from this step's readouts BEFORE the parameter gradients of the step are
            # accumulated: it takes part in the zero_grad rule (pc_trainer.py:844-859)
            early_stop = False
            if self._early_stop_condition.strip() != "False" or is_dynamic_x_lr:
                e_t, l_t = float(energy[t]), float(loss[t])
                overall = (l_t if loss_fn is not None else 0.0) + e_t * self._energy_coefficient
                overalls.append(overall)
                early_stop = bool(eval(self._early_stop_condition, {}, {
                    "t": t, "overall": overall, "loss": l_t if loss_fn is not None else None, "energy": e_t,
                    "overalls": overalls, "self": self}))
            if need_grads:
                zero = self._is_zero_grad_step(t) or \
                    (early_stop and self._update_p_at_early_stop and t not in self._acc_set)
                flat, gW, gb = self._ensure_flat_grads(netp, zero=zero)
                eng.weight_grad(netp, top, self._energy_coefficient, B, 1, save_g, save_f, inputs_dev, gW, gb,
                                self._precision)
                n_launch += 1
            if cb_bwd is not None:
                cb_bwd(t, **cb_bwd_kwargs)
            if t in self._update_x_set:
                self._optimizer_x.step()
                if is_dynamic_x_lr and len(overalls) >= 2:
                    factor = self._x_lr_discount if not (overalls[-1] < overalls[-2]) else self._x_lr_amplifier
                    if factor != 1.0:
                        for g in self._optimizer_x.param_groups:
                            g["lr"] = g["lr"] * factor
            if (t in self._update_p_set) or (early_stop and self._update_p_at_early_stop):
                self._p_step(flat, B)
            if cb_t is not None:
                cb_t(t, **cb_t_kwargs)
                if check_after_cb:
                    _slow_down_warning("PCTrainer.train_on_batch", "is_checking_after_callback_after_t", "False")
                    if not (self.get_is_model_training() == True):  # noqa: E712
                        raise RuntimeError(
                            "If you do <model.eval()> in <callback_after_t()>, you need to put model back to train "
                            "mode when leaving <callback_after_t()>. ")
            if early_stop:
                t_done = t + 1
                if not every_t:
                    warnings.warn("early stop hit before the last step: trajectories of the final step were not "
                                  "recorded", category=RuntimeWarning)
                break
        self.last_call_info = {"mode": "stepwise", "launches": n_launch, "segments": t_done,
                               "noise": N.NOISE_NONE, "precision": self._precision}
        return {"energy": energy[:t_done], "loss": loss[:t_done], "traj_x": traj_x, "traj_out": traj_out,
                "n_rec": min(n_rec, t_done)}

    # --------------------------------------------------------------------------------------
    def _install_lazy_energies(self, netp, inputs):
        """``PCLayer.energy()`` after a fused call: recomputed by one PyTorch forward when asked."""
        model = self._model

        def recompute():
            for layer in netp.pc_layers:
                layer._lazy_energy = None
            if self.get_is_model_training() == True:  # noqa: E712
                with torch.no_grad():
                    model(inputs)

        for layer in netp.pc_layers:
            layer._energy = None
            layer._lazy_energy = recompute

    def _start_scalar_readback(self, scalars, reduced=False):
        """Asynchronous device->host copy of the [2, T] energy / loss scalars on a side stream (CUDA tensors only).
        Returns (pinned host tensor, completion event) or None; ``_build_results`` waits for the event only."""
        if not scalars.is_cuda:
            return None
        vec = scalars if reduced else self._reduce_scalars(scalars)   # data-parallel: all-reduce on the main stream first
        dev = vec.device
        side = getattr(self, "_copy_stream", None)
        if side is None or side.device != dev:
            side = self._copy_stream = torch.cuda.Stream(device=dev)
        host = getattr(self, "_host_scalars", None)
        if host is None or host.shape != vec.shape:
            host = self._host_scalars = torch.empty(vec.shape, dtype=torch.float64).pin_memory()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            side.wait_event(ready)
            host.copy_(vec, non_blocking=True)
            vec.record_stream(side)
            done = torch.cuda.Event()
            done.record(side)
        return host, done

    def _reduce_scalars(self, vec):
        if self._dp_group is not None:
            import torch.distributed as dist
            vec = vec.clone()
            dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=self._dp_group)
        return vec

    def _build_results(self, ctx, rec, has_loss):
        every_t = ctx["every_t"]
        netp = ctx["netp"]
        pending = rec.get("host_scalars")
        if pending is not None:
            pending[1].synchronize()                          # the ONE host wait of the call: the scalars only
            host = pending[0].numpy().copy()
        else:
            both = rec.get("scalars")
            stacked = both if both is not None else torch.stack([rec["energy"], rec["loss"]])
            if not rec.get("scalars_reduced", False):
                stacked = self._reduce_scalars(stacked)
            host = stacked.to("cpu", torch.float64).numpy()   # the ONE device->host sync of the call
        e, l = host[0], host[1]
        sel = slice(None) if every_t else slice(len(e) - 1, len(e))
        # round to fp32 like `.item()` of the reference's fp32 scalars (pc_trainer.py:780,794,836)
        e32 = e.astype(np.float32)
        l32 = l.astype(np.float32)
        o32 = ((l32 if has_loss else 0.0) + e32 * np.float32(self._energy_coefficient)).astype(np.float32)
        results = {
            "loss": l32[sel].astype(np.float64).tolist() if has_loss else [],      # Python floats, like .item()
            "energy": e32[sel].astype(np.float64).tolist(),
            "overall": o32[sel].astype(np.float64).tolist(),
        }
        n_rec = rec["n_rec"]
        if ctx["want_outputs"]:
            src = rec["traj_out"] if netp.d_out > 0 else rec["traj_x"][netp.L - 1]
            results["outputs"] = list(src[:n_rec].unbind(0)) if n_rec > 0 else []
        # the reference returns host copies (pc_trainer.py:772-774); set_trajectories_on_device(True) keeps the rings on
        # the GPU (ONE bulk copy per layer otherwise -- never one per step)
        keep = (lambda t: t) if self._traj_on_device else (lambda t: t.cpu())
        if ctx["want_reps"]:
            results["representations"] = list(keep(rec["traj_x"][0][:n_rec]).unbind(0)) if n_rec > 0 else []
        if ctx["want_xs"]:
            per_layer = [keep(rec["traj_x"][l][:n_rec]) for l in range(netp.L)] if n_rec > 0 else []
            results["xs"] = [[per_layer[l][t] for l in range(netp.L)] for t in range(n_rec)]
        return results

    def _print_progress(self, rec, has_loss):
        e = float(rec["energy"][-1])
        l = float(rec["loss"][-1])
        msg = "|"
        if has_loss:
            msg += " l: {:.3e} |".format(l)
        msg += " e: {:.3e} |".format(e)
        msg += " o: {:.3e} |".format((l if has_loss else 0.0) + e * self._energy_coefficient)
        print(f"{msg} T={len(rec['energy'])} [{self.last_call_info.get('mode')}]")

    # ======================================================================================
    #  private (pc_trainer.py:1068-1108)
    # ======================================================================================
    def _preprocess_step_index_list(self, indices, T: int) -> typing.List[int]:
        assert isinstance(indices, (str, list))
        assert isinstance(T, int) and T > 0
        if isinstance(indices, str):
            table = {"all": lambda: list(range(T)), "last": lambda: [T - 1],
                     "last_half": lambda: list(range(T // 2, T)), "never": lambda: []}
            if indices not in table:
                raise NotImplementedError
            return table[indices]()
        for t in indices:
            assert isinstance(t, int)
            assert 0 <= t < T
        return indices
