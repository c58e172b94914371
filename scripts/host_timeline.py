"""Host / device timeline of a C2 MCPC learning call as bench.py issues it (flush, e0, call, e1): where inside the
event-timed region is the GPU idle?  Prints host stamps (us after call entry) and device intervals (us)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); warnings.simplefilter('ignore')
import numpy as np, torch, torch.optim as optim
from montecarlopredictivecoding_b200 import mcpc_utils as mu
dev = torch.device('cuda:0'); torch.manual_seed(0)
CFG = dict(input_size=20, hidden_size=128, hidden2_size=128, output_size=784, activation_fn="relu")
model = mu.get_model(CFG, use_cuda=False).to(dev)
config = {"mixing": 50, "sampling": 100, "optimizer_x_kwargs_mcpc": {"lr": 0.03}, "optimizer_p_fn_mcpc": optim.Adam, "optimizer_p_kwargs_mcpc": {"lr": 0.01}}
tr = mu.get_mcpc_trainer(model, config, training=True); tr.set_precision('bf16')
B = 1024; y = (torch.rand(B, 784, device=dev) < 0.5).float(); z = torch.zeros(B, 20, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
eng = tr._get_engine()
st = {}
orig_infer, orig_wg, orig_p, orig_br = eng.infer, eng.weight_grad, tr._p_step, tr._build_results
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
def infer(c):
    st["h_infer_in"] = time.perf_counter(); st["d_infer0"] = ev(); orig_infer(c); st["d_infer1"] = ev(); st["h_infer_out"] = time.perf_counter()
def wg(*a, **k):
    orig_wg(*a, **k); st["d_wg1"] = ev(); st["h_wg_out"] = time.perf_counter()
def pstep(*a, **k):
    r = orig_p(*a, **k); st["d_p1"] = ev(); st["h_p_out"] = time.perf_counter(); return r
def br(*a, **k):
    st["h_br_in"] = time.perf_counter(); r = orig_br(*a, **k); st["h_br_out"] = time.perf_counter(); return r
if not os.environ.get("PLAIN"):
    eng.infer, eng.weight_grad, tr._p_step, tr._build_results = infer, wg, pstep, br
kw = {"_target": y, "_var": 1.0}; ckw = {"_pc_trainer": tr}
def call():
    return tr.train_on_batch(inputs=z, loss_fn=mu.bernoulli_fn, loss_fn_kwargs=kw, callback_after_t=mu.random_step, callback_after_t_kwargs=ckw,
                             is_sample_x_at_batch_start=False, is_log_progress=False, is_checking_after_callback_after_t=False)
for _ in range(5): call()
torch.cuda.synchronize()
rows = []
for i in range(60):
    flush.fill_(i & 0xFF)
    e0 = ev(); h0 = time.perf_counter()
    call()
    h1 = time.perf_counter(); e1 = ev()
    rows.append((dict(st), e0, e1, h0, h1)); st.clear()
torch.cuda.synchronize()
tot = [a.elapsed_time(b) * 1e3 for _, a, b, _, _ in rows]
print(f"event-timed call: median {np.median(tot):.1f} us  mean {np.mean(tot):.1f} us; host wall per call {np.median([(h1 - h0) * 1e6 for *_, h0, h1 in rows]):.1f} us")
if not os.environ.get("PLAIN"):
    def med(f): return float(np.median([f(r) for r in rows]))
    print("host stamps after entry (us):")
    for k in ("h_infer_in", "h_infer_out", "h_wg_out", "h_p_out", "h_br_in", "h_br_out"):
        if k in rows[0][0]:
            print(f"  {k:12s} {med(lambda r: (r[0][k] - r[3]) * 1e6):8.1f}")
    print(f"  return       {med(lambda r: (r[4] - r[3]) * 1e6):8.1f}")
    print("device intervals (us):")
    print(f"  e0 -> infer start (GPU idle before the kernel) {med(lambda r: r[1].elapsed_time(r[0]['d_infer0']) * 1e3):8.1f}")
    print(f"  infer (packs + kernel)                         {med(lambda r: r[0]['d_infer0'].elapsed_time(r[0]['d_infer1']) * 1e3):8.1f}")
    if "d_wg1" in rows[0][0]:
        print(f"  infer end -> weight_grad end                   {med(lambda r: r[0]['d_infer1'].elapsed_time(r[0]['d_wg1']) * 1e3):8.1f}")
        print(f"  weight_grad end -> p_step end                  {med(lambda r: r[0]['d_wg1'].elapsed_time(r[0]['d_p1']) * 1e3):8.1f}")
    else:
        print(f"  infer (+ overlapped dW) end -> p_step end      {med(lambda r: r[0]['d_infer1'].elapsed_time(r[0]['d_p1']) * 1e3):8.1f}")
    print(f"  p_step end -> e1                               {med(lambda r: r[0]['d_p1'].elapsed_time(r[2]) * 1e3):8.1f}")
if os.environ.get("CPROF"):
    import cProfile, pstats
    eng.infer, eng.weight_grad, tr._p_step, tr._build_results = orig_infer, orig_wg, orig_p, orig_br
    pr = cProfile.Profile(); pr.enable()
    for _ in range(200): call()
    pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(40)
