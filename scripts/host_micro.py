"""timeit of the host-side pieces that run BEFORE the inference kernel of a C2 MCPC call is launched (GPU box)."""
import os, sys, time, warnings, timeit
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); warnings.simplefilter('ignore')
import ctypes as C
import torch, torch.optim as optim
from montecarlopredictivecoding_b200 import mcpc_utils as mu, _native as N
from montecarlopredictivecoding_b200.predictive_coding import plan as P, trainer as TR, engine as E
dev = torch.device('cuda:0')
CFG = dict(input_size=20, hidden_size=128, hidden2_size=128, output_size=784, activation_fn="relu")
model = mu.get_model(CFG, use_cuda=False).to(dev)
config = {"mixing": 50, "sampling": 100, "optimizer_x_kwargs_mcpc": {"lr": 0.03}, "optimizer_p_fn_mcpc": optim.Adam, "optimizer_p_kwargs_mcpc": {"lr": 0.01}}
tr = mu.get_mcpc_trainer(model, config, training=True); tr.set_precision('bf16')
B = 1024; y = (torch.rand(B, 784, device=dev) < 0.5).float(); z = torch.zeros(B, 20, device=dev)
kw = {"_target": y, "_var": 1.0}; ckw = {"_pc_trainer": tr}
def call():
    return tr.train_on_batch(inputs=z, loss_fn=mu.bernoulli_fn, loss_fn_kwargs=kw, callback_after_t=mu.random_step, callback_after_t_kwargs=ckw,
                             is_sample_x_at_batch_start=False, is_log_progress=False, is_checking_after_callback_after_t=False)
for _ in range(5): call()
torch.cuda.synchronize()
netp = P.compile_net(model); top = P.classify_loss(mu.bernoulli_fn, kw, B, 784, dev); eng = tr._get_engine()
def T(name, f, n=2000):
    f(); torch.cuda.synchronize()
    t = timeit.timeit(f, number=n) / n * 1e6; torch.cuda.synchronize(); print(f"{name:40s} {t:7.2f} us", flush=True)
T("get_is_model_training", tr.get_is_model_training)
T("compile_net", lambda: P.compile_net(model))
T("classify_loss", lambda: P.classify_loss(mu.bernoulli_fn, kw, B, 784, dev))
T("_start_of_batch", lambda: tr._start_of_batch(netp, z, False, True, False))
T("_reset_optimizer_x", tr._reset_optimizer_x)
T("classify_callback", lambda: P.classify_callback_after_t(mu.random_step, ckw, tr))
T("_classify_optimizer_x", tr._classify_optimizer_x)
T("_slow_down_warning", lambda: TR._slow_down_warning("a", "b", "c"))
T("_inputs_or_none", lambda: tr._inputs_or_none(z))
T("target prep", lambda: y.detach().to(torch.float32).contiguous())
T("zeros(2,T) f64 cuda", lambda: torch.zeros(2, 150, dtype=torch.float64, device=dev), 500)
T("empty(2,T) f64 cuda", lambda: torch.empty(2, 150, dtype=torch.float64, device=dev), 500)
T("_param_tensors", lambda: tr._param_tensors(netp))
T("infer_mode", lambda: eng.infer_mode(netp, top, B, N.PREC_BF16))
T("_env_key", E._env_key)
T("_net_key", lambda: E._net_key(netp, top, 1.0))
T("net_struct", lambda: E.net_struct(netp, top, 1.0))
T("_save_layout", lambda: tr._save_layout(netp, top))
T("_buffer x2", lambda: (tr._buffer("save_g", (100, B, 1060), torch.bfloat16, dev), tr._buffer("save_f", (100, B, 276), torch.bfloat16, dev)))
T("_segments_cached", lambda: tr._segments_cached(150, False))
T("current_stream", lambda: torch.cuda.current_stream(dev).cuda_stream)
T("_OnDevice", lambda: E._OnDevice(dev).__enter__())
T("McpcIO()+McpcOpts()", lambda: (N.McpcIO(), N.McpcOpts()))
T("_ptr x16", lambda: [E._ptr(y, "t") for _ in range(16)])
T("event create+record", lambda: torch.cuda.Event(enable_timing=True).record(), 500)
io = N.McpcIO()
def fill():
    for i in range(16): io.W[i % 8] = 12345
T("16 ctypes array stores", fill)
lib = eng._lib
T("mcpc_launch_count (ctypes call)", lib.mcpc_launch_count)
# a whole call and the inference launch alone
calls = []
orig = eng.infer
def rec(c): calls.append(c); orig(c)
eng.infer = rec; call(); eng.infer = orig
c0 = calls[0]
T("eng.infer (python + C + launches)", lambda: eng.infer(c0), 200)
T("whole call (GPU-bound)", call, 100)
