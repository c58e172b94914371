"""cProfile of the host side of one C2 MCPC learning call (where do the ~0.4 ms outside the kernels go?)."""
import cProfile, os, pstats, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); warnings.simplefilter('ignore')
import torch, torch.optim as optim
from montecarlopredictivecoding_b200 import mcpc_utils as mu
dev = torch.device('cuda:0'); torch.manual_seed(0)
CFG = dict(input_size=20, hidden_size=128, hidden2_size=128, output_size=784, activation_fn="relu")
model = mu.get_model(CFG, use_cuda=False).to(dev)
config = {"mixing": 50, "sampling": 100, "optimizer_x_kwargs_mcpc": {"lr": 0.03}, "optimizer_p_fn_mcpc": optim.Adam, "optimizer_p_kwargs_mcpc": {"lr": 0.01}}
tr = mu.get_mcpc_trainer(model, config, training=True); tr.set_precision('bf16')
B = 1024; y = (torch.rand(B, 784, device=dev) < 0.5).float(); z = torch.zeros(B, 20, device=dev)
def call():
    return tr.train_on_batch(inputs=z, loss_fn=mu.bernoulli_fn, loss_fn_kwargs={"_target": y, "_var": 1.0}, callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr}, is_sample_x_at_batch_start=False, is_log_progress=False, is_checking_after_callback_after_t=False)
for _ in range(5): call()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): call()
torch.cuda.synchronize()
print("wall ms/call", (time.perf_counter() - t0) / 50 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(50): call()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
