/*
 * mcpc_b200.h -- C ABI of the B200-native MCPC hot path (libmcpc_b200.so).
 *
 * The reference (gaspardol/MonteCarloPredictiveCoding) is pure Python: it has no FFI of its
 * own.  Its boundary for this path is the Python class API
 *     predictive_coding/__init__.py:1-2   (PCLayer, PCTrainer)
 * and the work the entry points below replace is the body of the T-step loop of
 *     predictive_coding/pc_trainer.py:712-983 (PCTrainer.train_on_batch).
 * The drop-in Python classes (montecarlopredictivecoding_b200/predictive_coding) bind these
 * symbols through ctypes; INTEGRATION.md shows the binding a maintainer of the reference
 * would add.
 *
 * Conventions: plain pointers and sizes only; every pointer in McpcIO / McpcGradIO is a
 * DEVICE pointer owned by the caller (no ownership transfer, nothing is allocated or freed
 * here); every call is asynchronous on `stream` (a cudaStream_t passed as void*) and never
 * synchronises the host; return value 0 = success, negative = error (mcpc_last_error()
 * gives the text, thread-local).  All matrices are row-major fp32 unless stated.
 */
#ifndef MCPC_B200_H_
#define MCPC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCPC_ABI_VERSION 1
#define MCPC_MAX_LAYERS 8

/* activation applied to the latent x_l before the next Linear (utils/model.py:57,60,63) */
enum { MCPC_ACT_IDENTITY = 0, MCPC_ACT_RELU = 1, MCPC_ACT_TANH = 2 };
/* what sits on top of the last PCLayer:
 *   NONE      no loss_fn (the output Linear, if any, is readout only)         pc_trainer.py:777-782
 *   ZERO      utils/model.py:27-28 zero_fn: loss value 0, no gradient
 *   GAUSS     utils/model.py:17-18,24-25 fe_fn / fe_fn_mask: (1/var)*0.5*sum (o-y)^2
 *   BERNOULLI utils/model.py:20-22,31-33 bernoulli_fn[_mask]: sum BCEWithLogits(o, y)          */
enum { MCPC_TOP_NONE = 0, MCPC_TOP_ZERO = 1, MCPC_TOP_GAUSS = 2, MCPC_TOP_BERNOULLI = 3 };
/* optimizer on the latents: optim.SGD (no momentum) / optim.Adam  pc_trainer.py:465-475,877 */
enum { MCPC_OPT_SGD = 0, MCPC_OPT_ADAM = 1 };
/* Langevin noise source (utils/model.py:35-44 random_step) */
enum { MCPC_NOISE_NONE = 0, MCPC_NOISE_SUPPLIED = 1, MCPC_NOISE_PHILOX = 2 };
/* arithmetic of the contractions: FP32 = CUDA-core fp32 FMA (reference-exact mode);
 * BF16 = tcgen05 kind::f16 with bf16 operands and fp32 accumulation in TMEM               */
enum { MCPC_PREC_FP32 = 0, MCPC_PREC_BF16 = 1 };

/* how mcpc_infer executes (net, B, precision) -- see mcpc_infer_mode:
 *   RESIDENT_*  one persistent launch for all steps, latents on chip; the weight update is a separate
 *               mcpc_weight_grad over operands saved in save_g / save_f
 *   STREAMING   networks too wide for one SM: per-step grouped GEMM kernels; the weight update is accumulated
 *               directly into McpcIO.gW / gb                                                              */
enum { MCPC_MODE_RESIDENT_FP32 = 0, MCPC_MODE_RESIDENT_BF16 = 1, MCPC_MODE_STREAMING_BF16 = 2 };

enum {
  MCPC_OK = 0,
  MCPC_ERR_INVALID = -1,      /* bad argument (NULL, negative size, unknown enum)            */
  MCPC_ERR_UNSUPPORTED = -2,  /* shape / feature outside what the kernels implement          */
  MCPC_ERR_WORKSPACE = -3,    /* workspace too small                                         */
  MCPC_ERR_CUDA = -4          /* CUDA runtime error at launch (text in mcpc_last_error)      */
};

/* Chain  inputs -> Linear_0 -> PC_0 -> act_0 -> Linear_1 -> PC_1 -> ... [-> Linear_out]
 * (utils/model.py:54-65, figure_2.py:40-44, figure_3.py:50-55). */
typedef struct McpcNet {
  int32_t n_layers;                       /* L = number of PCLayers (1..MCPC_MAX_LAYERS)     */
  int32_t d_in;                           /* width of `inputs`                               */
  int32_t dims[MCPC_MAX_LAYERS];          /* d_l, width of latent x_l                        */
  int32_t d_out;                          /* width of the output Linear, 0 if there is none  */
  int32_t act[MCPC_MAX_LAYERS];           /* MCPC_ACT_* applied to x_l                       */
  float energy_scale[MCPC_MAX_LAYERS];    /* c_l of energy_fn = c_l*0.5*(mu-x)^2  pc_layer.py:17-18, figure_3.py:47-48 */
  float energy_coefficient;               /* overall = loss + energy*coef  pc_trainer.py:825-827 */
  int32_t top;                            /* MCPC_TOP_*                                      */
  float top_inv_var;                      /* 1/_var of the Gaussian top                      */
  int32_t mask_start_col;                 /* loss covers output columns >= this (the *_mask fns) */
} McpcNet;

typedef struct McpcIO {
  const float* W[MCPC_MAX_LAYERS + 1];    /* W[l]: [d_l, d_{l-1}] (nn.Linear layout); W[L]: [d_out, d_{L-1}] */
  const float* b[MCPC_MAX_LAYERS + 1];    /* bias or NULL (bias=False)                       */
  float* x[MCPC_MAX_LAYERS];              /* latents [B, d_l], updated in place (PCLayer._x, pc_layer.py:230) */
  const float* inputs;                    /* [B, d_in]; NULL means all zeros (every script passes zeros) */
  const float* target;                    /* [B, d_out] `_target`; may be NULL for TOP_NONE/ZERO */
  const float* noise;                     /* NOISE_SUPPLIED: [n_steps, B, sum(d_l)] raw gradient noise as random_step
                                             writes it into x.grad (std sqrt(var/lr0)); the step is x -= lr*noise */
  float* adam_m[MCPC_MAX_LAYERS];         /* OPT_ADAM: exp_avg / exp_avg_sq state [B, d_l]    */
  float* adam_v[MCPC_MAX_LAYERS];
  float* x_grad[MCPC_MAX_LAYERS];         /* optional: d overall / d x_l of the LAST step of this call */
  double* energy;                         /* [n_steps] sum_l E_l at the START of each step  pc_trainer.py:785-795 */
  double* loss;                           /* [n_steps] loss at the start of each step (0 when TOP_NONE) :777-780 */
  float* traj_x[MCPC_MAX_LAYERS];         /* optional [n_rec, B, d_l]: x_l at the start of step t_k = k*traj_every */
  float* traj_out;                        /* optional [n_rec, B, d_out]: outputs of the same steps (:769-770) */
  float* gW[MCPC_MAX_LAYERS + 1];         /* optional weight-gradient accumulators: the update of the steps [save_begin,    */
  float* gb[MCPC_MAX_LAYERS + 1];         /*   save_end) is ADDED here by mcpc_infer itself.  MODE_STREAMING: the only way     */
                                          /*   (no save_g/save_f).  RESIDENT modes: save_g/save_f must be given too (scratch) */
                                          /*   and no mcpc_weight_grad call is needed; RESIDENT_BF16 runs the update on the   */
                                          /*   SMs the inference kernel leaves idle, WHILE it runs (pc_trainer.py:862)        */
  void* save_g;                           /* optional [n_save, B, g_width]: d overall / d mu_l and d loss / d out      */
  void* save_f;                           /* optional [n_save, B, f_width]: act_l(x_l); operands of mcpc_weight_grad;
                                             widths / element type from mcpc_save_layout                           */
} McpcIO;

typedef struct McpcOpts {
  double lr;                /* current param-group lr of optimizer_x                                 */
  double adam_beta1, adam_beta2, adam_eps;
  double noise_scale;       /* PHILOX: sqrt(var / lr0), the std random_step gives x.grad             */
  uint64_t seed;            /* PHILOX key                                                            */
  uint64_t chain_offset;    /* global index of local row 0 (multi-GPU shards share one stream)       */
  int32_t n_steps;          /* steps in this call                                                    */
  int32_t t_begin;          /* absolute index of the first step (Philox counter)                     */
  int32_t optimizer;        /* MCPC_OPT_*                                                            */
  int32_t update_x;         /* 0: latents only receive noise (t not in update_x_at, pc_trainer.py:874) */
  int32_t adam_step0;       /* Adam steps already taken (bias correction continues from here)        */
  int32_t noise_mode;       /* MCPC_NOISE_*                                                          */
  int32_t traj_every;       /* record trajectories every k-th step (>=1) when traj pointers are set  */
  int32_t save_begin;       /* steps [save_begin, save_end) (relative to this call) are written to   */
  int32_t save_end;         /*   save_g / save_f, slot = step - save_begin                           */
  int32_t precision;        /* MCPC_PREC_*                                                           */
} McpcOpts;

/* Operands of the local weight update  gW_l += G_l^T act(x_{l-1}),  gb_l += colsum(G_l)
 * summed over n_save steps and B rows (autograd's dW of pc_trainer.py:862). */
typedef struct McpcGradIO {
  const void* save_g;                     /* as written by mcpc_infer                        */
  const void* save_f;
  const float* inputs;                    /* [B, d_in] or NULL (zeros => gW_0 receives nothing) */
  float* gW[MCPC_MAX_LAYERS + 1];         /* accumulators, ADDED to (never zeroed here); NULL skips */
  float* gb[MCPC_MAX_LAYERS + 1];
  void* scratch;                          /* device scratch, needed only for MCPC_PREC_BF16 with non-NULL inputs:      */
  size_t scratch_bytes;                   /* B * dims[0] floats (sum over the saved steps of d overall / d mu_0)       */
} McpcGradIO;

int mcpc_version(void);
const char* mcpc_last_error(void);
/* Number of CUDA kernels this library has launched in the calling process (monotonic). */
uint64_t mcpc_launch_count(void);

/* Bytes of scratch mcpc_infer needs for (net, B, n_steps). */
int mcpc_workspace_bytes(const McpcNet* net, int32_t B, int32_t n_steps, int32_t precision, size_t* out_bytes);

/* Row layout of the operands mcpc_infer saves for mcpc_weight_grad: save_g is [n_save, B, *g_width],
 * save_f is [n_save, B, *f_width], elements of *elem_bytes bytes (fp32: unpadded concatenation; bf16: every
 * layer's block padded to 8 columns).  Allocate 1024 elements of slack behind each buffer. */
int mcpc_save_layout(const McpcNet* net, int32_t precision, int32_t* g_width, int32_t* f_width, int32_t* elem_bytes);

int mcpc_infer_mode(const McpcNet* net, int32_t B, int32_t precision, int32_t* mode);

/* n_steps fused steps of: forward, energy/loss readout, latent gradient, x-step, Langevin noise
 * (pc_trainer.py:733-918 with utils/model.py:35-44 folded in). */
int mcpc_infer(const McpcNet* net, const McpcIO* io, const McpcOpts* opts, int32_t B,
               void* workspace, size_t workspace_bytes, void* stream);

/* *out = 1 when passing McpcIO.gW/gb to mcpc_infer is the faster way to get the weight update for this (net, B, precision):
 * MODE_STREAMING (the only way there), or MODE_RESIDENT_BF16 with MCPC_TC_DW_OVERLAP=1 in the environment when the
 * weight-gradient kernel fits the SMs the inference kernel leaves idle and runs NEXT to it (B <= ~1100 chains, zero
 * inputs; opt-in: it pays only for callers that do not read results between calls).  0: mcpc_infer would
 * run the update after the inference kernel -- a caller that wants the per-step scalars early calls mcpc_weight_grad itself. */
int mcpc_infer_fuses_weight_grad(const McpcNet* net, int32_t B, int32_t precision, int32_t has_inputs, int32_t* out);

/* Local weight update from the operands saved by mcpc_infer. */
int mcpc_weight_grad(const McpcNet* net, const McpcGradIO* io, int32_t B, int32_t n_save,
                     int32_t precision, void* stream);

/* The raw gradient noise NOISE_PHILOX applies, materialised as [n_steps, B, sum(d_l)] fp32
 * (validation: feed it to the reference / oracle as recorded noise). */
int mcpc_fill_noise(uint64_t seed, int32_t t_begin, int32_t n_steps, uint64_t chain_offset, int32_t B,
                    int32_t n_units, float noise_scale, float* out, void* stream);

/* SURVEY 8(f) N1 -- importance-sampling estimate of the marginal log-likelihood of a Bernoulli generative model
 * (utils/training_evaluation.py:177-206, get_marginal_likelihood):
 *   losses[i][s] = sum_j BCEWithLogits(clamp(logits[s][j], +-clamp_abs), data[i][j])
 *   row_ll[i]    = log(mean_s exp(-(losses[i][s] - m_i))) - m_i,  m_i = min_s losses[i][s]      (:203-205)
 *   *ml_out      = mean_i row_ll[i]
 * logits [S, D] are the pre-sigmoid outputs of S prior samples (sample_pc(..., is_return_hidden=True), :72-100),
 * data [N, D] targets in [0, 1]; both fp32 row-major device pointers.  ml_out: one device double; row_ll: optional
 * device [N] fp32.  One tcgen05 GEMM (bf16 hi/lo split operands, fp32 accumulation) with a streaming min / sum-exp
 * epilogue; clamp_abs <= 0 disables the clamp. */
int mcpc_marginal_ll_workspace_bytes(int32_t N, int32_t S, int32_t D, size_t* bytes);
int mcpc_marginal_ll_bernoulli(const float* logits, int32_t S, const float* data, int32_t N, int32_t D, float clamp_abs,
                               void* workspace, size_t workspace_bytes, double* ml_out, float* row_ll, void* stream);

/* SURVEY 8(f) N2 -- running per-element mean / variance of a recorded trajectory, on the device.  Replaces the host-side
 * reductions over per-step `.cpu()` copies of the reference (pc_trainer.py:772-774 feeding utils/model.py:143-149
 * `temp.mean(0)` and figure_2.py:75-79).  traj: [n_rec, n_elems] fp32 ring as written through McpcIO.traj_x / traj_out;
 * mean, m2: [n_elems] fp32 running accumulators (Welford: m2 = sum of squared deviations); count_before = samples
 * already folded in (0: the accumulators are initialised here).  After the call they hold count_before + n_rec samples;
 * the unbiased variance is m2 / (count - 1). */
int mcpc_traj_stats_update(const float* traj, int32_t n_rec, uint64_t n_elems, uint64_t count_before, float* mean, float* m2,
                           void* stream);

/* SURVEY 8(f) N3 -- the parameter update of pc_trainer.py:904-914 in one launch over all W / b tensors:
 *   grad *= inv_norm (= 1 / (len(accumulate_p_at) * batch), written back like the reference's `param.grad = param.grad / ...`),
 *   then optim.SGD (momentum / dampening / nesterov / weight_decay) or optim.Adam (betas, eps, L2 weight_decay) exactly as
 *   torch's single-tensor implementations compute them, IN PLACE on the optimizer's own state tensors:
 *   state1 = momentum_buffer (SGD, NULL without momentum) or exp_avg (Adam); state2 = exp_avg_sq (Adam).
 *   first_step: SGD momentum buffers are initialised with the gradient (torch clones it on the first step);
 *   step: 1-based index of this Adam step (bias corrections). */
#define MCPC_MAX_PTENSORS (2 * (MCPC_MAX_LAYERS + 1))
typedef struct McpcPStep {
  int32_t kind;                           /* MCPC_OPT_SGD | MCPC_OPT_ADAM */
  int32_t n_tensors;
  float* param[MCPC_MAX_PTENSORS];
  float* grad[MCPC_MAX_PTENSORS];
  float* state1[MCPC_MAX_PTENSORS];
  float* state2[MCPC_MAX_PTENSORS];
  uint64_t numel[MCPC_MAX_PTENSORS];
  double inv_norm;
  double lr, weight_decay, momentum, dampening, beta1, beta2, eps;
  int32_t nesterov, first_step, step;
} McpcPStep;
int mcpc_p_step(const McpcPStep* step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCPC_B200_H_ */
