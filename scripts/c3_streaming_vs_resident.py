import sys, os, json
sys.path.insert(0, '/root/repo/scripts'); sys.path.insert(0, '/root/repo')
import bench_configs as bc
import torch
bc_ml = bc.ml_model
bc.ml_model = lambda act="relu", dims=(32, 128, 128): bc_ml(act, (32, 128, 128))
import torch.optim as optim
from montecarlopredictivecoding_b200 import mcpc_utils as mu, predictive_coding as pc
def c3(prec, B, T):
    model = bc.ml_model()
    tr = pc.PCTrainer(model, T=T, optimizer_x_fn=optim.SGD, optimizer_x_kwargs={"lr": 0.1}, update_p_at="never", plot_progress_at=[])
    tr.set_precision(prec)
    z = torch.zeros(B, 32, device=bc.DEV)
    pcs = [m for m in model if isinstance(m, pc.PCLayer)]
    x0 = [torch.randn(B, d, device=bc.DEV) for d in (32, 128, 128)]
    for layer, v in zip(pcs, x0):
        layer._sample_x_fn = (lambda inputs, v=v: v.clone())
    first = [True]
    def run():
        tr.train_on_batch(z, loss_fn=mu.zero_fn, callback_after_t=mu.random_step, callback_after_t_kwargs={"_pc_trainer": tr},
                          is_sample_x_at_batch_start=first[0], is_log_progress=False, is_return_results_every_t=False)
        first[0] = False
    s = bc.timed(run, reps=2)
    return {"B": B, "T": T, "us_per_step": s / T * 1e6, "lu_per_s": B * 3 * T / s}
for B in (1024, 8192, 65536):
    print(os.environ.get("MCPC_FORCE_STREAMING"), json.dumps(c3("bf16", B, 200)))
