"""Stamps around the host functions of a real (GPU-bound) C2 call loop: entry/exit times relative to call entry."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); warnings.simplefilter('ignore')
import numpy as np, torch, torch.optim as optim
from montecarlopredictivecoding_b200 import mcpc_utils as mu
from montecarlopredictivecoding_b200.predictive_coding import plan as P, trainer as TR, engine as E
dev = torch.device('cuda:0')
CFG = dict(input_size=20, hidden_size=128, hidden2_size=128, output_size=784, activation_fn="relu")
model = mu.get_model(CFG, use_cuda=False).to(dev)
config = {"mixing": 50, "sampling": 100, "optimizer_x_kwargs_mcpc": {"lr": 0.03}, "optimizer_p_fn_mcpc": optim.Adam, "optimizer_p_kwargs_mcpc": {"lr": 0.01}}
tr = mu.get_mcpc_trainer(model, config, training=True); tr.set_precision('bf16')
B = 1024; y = (torch.rand(B, 784, device=dev) < 0.5).float(); z = torch.zeros(B, 20, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
kw = {"_target": y, "_var": 1.0}; ckw = {"_pc_trainer": tr}
eng = tr._get_engine()
log = []
pc = time.perf_counter
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        t = pc()
        try:
            return f(*a, **k)
        finally:
            log.append((name, t, pc()))
    setattr(obj, name, g)
for o, n in [(tr, "get_is_model_training"), (P, "compile_net"), (P, "classify_loss"), (tr, "_start_of_batch"), (P, "classify_callback_after_t"),
             (tr, "_classify_optimizer_x"), (tr, "_run_fused"), (tr, "_inputs_or_none"), (tr, "_prepare_traj_stats"), (tr, "_param_tensors"),
             (eng, "infer_mode"), (tr, "_segments_cached"), (tr, "_save_layout"), (tr, "_buffer"), (eng, "infer"), (tr, "_start_scalar_readback"),
             (tr, "_ensure_flat_grads"), (eng, "weight_grad"), (tr, "_p_step"), (tr, "_install_lazy_energies"), (tr, "_build_results")]:
    wrap(o, n)
def call():
    return tr.train_on_batch(inputs=z, loss_fn=mu.bernoulli_fn, loss_fn_kwargs=kw, callback_after_t=mu.random_step, callback_after_t_kwargs=ckw,
                             is_sample_x_at_batch_start=False, is_log_progress=False, is_checking_after_callback_after_t=False)
for _ in range(5): call()
torch.cuda.synchronize()
runs = []
for i in range(60):
    if not os.environ.get("NOFLUSH"):
        flush.fill_(i & 0xFF)
    log.clear(); h0 = pc(); call(); h1 = pc()
    runs.append([(n, (a - h0) * 1e6, (b - h0) * 1e6) for n, a, b in log] + [("return", (h1 - h0) * 1e6, (h1 - h0) * 1e6)])
torch.cuda.synchronize()
names = [r[0] for r in runs[0]]
print(f"{'function':28s} {'enter':>8s} {'exit':>8s} {'dur':>7s}   (median us after call entry; wrappers nest)")
for j, n in enumerate(names):
    a = np.median([r[j][1] for r in runs]); b = np.median([r[j][2] for r in runs])
    print(f"{n:28s} {a:8.1f} {b:8.1f} {b - a:7.1f}")
