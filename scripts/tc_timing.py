"""Cycle trace of CTA 0 of the resident bf16 kernel on the C2 workload (MCPC_TC_TIMING=<first step> prints 8 steps)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); warnings.simplefilter('ignore')
import torch, torch.optim as optim
from montecarlopredictivecoding_b200 import mcpc_utils as mu
dev = torch.device('cuda:0'); torch.manual_seed(0)
CFG = dict(input_size=20, hidden_size=128, hidden2_size=128, output_size=784, activation_fn="relu")
model = mu.get_model(CFG, use_cuda=False).to(dev)
MIX, SAMP = int(os.environ.get("MIX", 5)), int(os.environ.get("SAMP", 10))
config = {"mixing": MIX, "sampling": SAMP, "optimizer_x_kwargs_mcpc": {"lr": 0.03}, "optimizer_p_fn_mcpc": optim.Adam, "optimizer_p_kwargs_mcpc": {"lr": 0.01}}
tr = mu.get_mcpc_trainer(model, config, training=True); tr.set_precision('bf16')
B = 1024; y = (torch.rand(B, 784, device=dev) < 0.5).float(); z = torch.zeros(B, 20, device=dev)
if os.environ.get("MODE") == "map":      # deterministic PC / MAP inference (Adam on x, tanh, masked BCE): config C4
    CFG2 = dict(input_size=25, hidden_size=128, hidden2_size=128, output_size=784, activation_fn="tanh")
    model = mu.get_model(CFG2, use_cuda=False).to(dev)
    tr = mu.get_pc_trainer(model, {"T_pc": 250, "optimizer_x_fn_pc": optim.Adam, "optimizer_x_kwargs_pc": {"lr": 0.3}}, is_mcpc=True)
    tr.set_precision('bf16')
    z = torch.zeros(B, 25, device=dev)
    for i in range(2):
        tr.train_on_batch(inputs=z, loss_fn=mu.bernoulli_fn_mask, loss_fn_kwargs={"_target": y, "_var": 1.0}, is_log_progress=False,
                          is_return_results_every_t=False, is_checking_after_callback_after_t=False)
    torch.cuda.synchronize()
    sys.exit(0)
for i in range(2):
    tr.train_on_batch(inputs=z, loss_fn=mu.bernoulli_fn, loss_fn_kwargs={"_target": y, "_var": 1.0}, callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr}, is_sample_x_at_batch_start=(i == 0), is_log_progress=False, is_checking_after_callback_after_t=False)
torch.cuda.synchronize()
