"""Wall-clock breakdown of the host side of one C2 MCPC learning call (where do the ~0.4 ms outside the kernels go?)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); warnings.simplefilter('ignore')
import torch, torch.optim as optim
from montecarlopredictivecoding_b200 import mcpc_utils as mu
from montecarlopredictivecoding_b200.predictive_coding import trainer as TR, plan as P, engine as E
dev = torch.device('cuda:0'); torch.manual_seed(0)
CFG = dict(input_size=20, hidden_size=128, hidden2_size=128, output_size=784, activation_fn="relu")
model = mu.get_model(CFG, use_cuda=False).to(dev)
config = {"mixing": 50, "sampling": 100, "optimizer_x_kwargs_mcpc": {"lr": 0.03}, "optimizer_p_fn_mcpc": optim.Adam, "optimizer_p_kwargs_mcpc": {"lr": 0.01}}
tr = mu.get_mcpc_trainer(model, config, training=True); tr.set_precision('bf16')
B = 1024; y = (torch.rand(B, 784, device=dev) < 0.5).float(); z = torch.zeros(B, 20, device=dev)
acc = {}
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        t = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
    setattr(obj, name, g)
eng = tr._get_engine()
for o, n in [(P, "compile_net"), (P, "classify_loss"), (P, "classify_callback_after_t"), (tr, "_start_of_batch"), (tr, "_run_fused"),
             (eng, "infer"), (eng, "weight_grad"), (tr, "_p_step"), (tr, "_build_results"), (tr, "_ensure_flat_grads"),
             (tr, "_install_lazy_energies"), (tr, "_segments"), (tr, "_param_tensors"), (tr, "_save_layout")]:
    if hasattr(o, n):
        wrap(o, n)
if os.environ.get("MODE") == "map":
    CFG2 = dict(input_size=25, hidden_size=128, hidden2_size=128, output_size=784, activation_fn="tanh")
    model = mu.get_model(CFG2, use_cuda=False).to(dev)
    tr = mu.get_pc_trainer(model, {"T_pc": 250, "optimizer_x_fn_pc": optim.Adam, "optimizer_x_kwargs_pc": {"lr": 0.3}}, is_mcpc=True)
    tr.set_precision('bf16')
    z = torch.zeros(B, 25, device=dev)
    eng = tr._get_engine()
    for o, n in [(tr, "_start_of_batch"), (tr, "_run_fused"), (eng, "infer"), (eng, "weight_grad"), (tr, "_build_results"),
                 (tr, "_ensure_flat_grads"), (tr, "recreate_optimize_x")]:
        wrap(o, n)


def call():
    if os.environ.get("MODE") == "map":
        return tr.train_on_batch(inputs=z, loss_fn=mu.bernoulli_fn_mask, loss_fn_kwargs={"_target": y, "_var": 1.0}, is_log_progress=False,
                                 is_return_results_every_t=False, is_checking_after_callback_after_t=False)
    return tr.train_on_batch(inputs=z, loss_fn=mu.bernoulli_fn, loss_fn_kwargs={"_target": y, "_var": 1.0}, callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr}, is_sample_x_at_batch_start=False, is_log_progress=False, is_checking_after_callback_after_t=False)
for _ in range(5): call()
torch.cuda.synchronize(); acc.clear()
N = 100
t0 = time.perf_counter()
for _ in range(N): call()
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / N * 1e3
print(f"wall {tot:.3f} ms/call")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print(f"  {k:28s} {v / N * 1e3:.3f} ms")
