// Shared device/host declarations for libmcpc_b200.so.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "mcpc_b200.h"

namespace mcpc {

constexpr int kMaxL = MCPC_MAX_LAYERS;

// Net description with the derived offsets every kernel needs.
struct NetDev {
  int L, d_in, d_out;
  int SD;               // sum_l d_l        (width of the concatenated latent / F rows)
  int NG;               // SD + d_out       (width of the concatenated G rows)
  int dims[kMaxL];
  int off[kMaxL + 1];   // unit offset of layer l inside a concatenated row; off[L] = SD
  int act[kMaxL];
  float c[kMaxL];       // energy scale (readout)
  float gc[kMaxL];      // energy_coefficient * c_l (gradients)
  int top;
  float inv_var;
  int mask_start;
  bool top_has_grad;    // GAUSS / BERNOULLI with an output Linear
};

void set_error(const char* fmt, ...);
void count_launch(int n = 1);   // feeds mcpc_launch_count()
int check_net(const McpcNet* net, NetDev* out);   // validates + fills NetDev; returns MCPC_* code

#define MCPC_CUDA_CHECK(expr)                                                          \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::mcpc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MCPC_ERR_CUDA;                                                            \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ float act_apply(int kind, float x) {
  if (kind == MCPC_ACT_RELU) return fmaxf(x, 0.0f);
  if (kind == MCPC_ACT_TANH) return tanhf(x);
  return x;
}
// derivative from the pre-activation x and the activation value a = act(x)
__device__ __forceinline__ float act_deriv(int kind, float x, float a) {
  if (kind == MCPC_ACT_RELU) return x > 0.0f ? 1.0f : 0.0f;
  if (kind == MCPC_ACT_TANH) return 1.0f - a * a;
  return 1.0f;
}

// Launchers implemented in the .cu files (host side, return MCPC_* codes).
struct InferArgs;   // infer_rows.cu
int infer_rows_workspace(const NetDev& nd, int B, int n_steps, size_t* bytes);
int launch_infer_rows(const NetDev& nd, const McpcIO* io, const McpcOpts* o, int B, void* ws, size_t ws_bytes,
                      cudaStream_t stream);
int launch_weight_grad_fp32(const NetDev& nd, const McpcGradIO* io, int B, int n_save, cudaStream_t stream);
int launch_wgrad_tn_fp32(const float* A, int lda, const float* Bm, int ldb, float* C, int M, int N, int rows,
                         cudaStream_t stream);
int launch_fill_noise(uint64_t seed, int t_begin, int n_steps, uint64_t chain_offset, int B, int n_units,
                      float noise_scale, float* out, cudaStream_t stream);

int launch_reduce_partials(const float* partials, int n_steps, int n_tiles, double* energy, double* loss,
                           cudaStream_t stream);
int infer_tc_workspace(const NetDev& nd, int B, int n_steps, size_t* bytes);
int launch_infer_tc(const NetDev& nd, const McpcIO* io, const McpcOpts* o, int B, void* ws, size_t ws_bytes,
                    cudaStream_t stream);
void save_layout_bf16(const NetDev& nd, int* g_off, int* g_width, int* f_off, int* f_width);
int launch_weight_grad_tc(const NetDev& nd, const McpcGradIO* io, int B, int n_save, cudaStream_t stream);
int weight_grad_tc_tiles(const NetDev& nd, const McpcGradIO* io);
bool infer_tc_overlaps_weight_grad(const NetDev& nd, int B, bool has_inputs);
int launch_weight_grad_tc_overlapped(const NetDev& nd, const McpcGradIO* io, int B, int n_save, const unsigned* ready,
                                     unsigned ready_target, int max_ctas, cudaStream_t stream);
bool infer_tc_fits(const NetDev& nd, int B);
int infer_wide_workspace(const NetDev& nd, int B, int n_steps, size_t* bytes);
int launch_infer_wide(const NetDev& nd, const McpcIO* io, const McpcOpts* o, int B, void* ws, size_t ws_bytes,
                      cudaStream_t stream);
int marginal_ll_workspace(int N, int S, int D, size_t* bytes);
int launch_marginal_ll(const float* logits, int S, const float* data, int N, int D, float clamp_abs, void* ws,
                       size_t ws_bytes, double* ml_out, float* row_ll, cudaStream_t stream);
int launch_traj_stats(const float* traj, int n_rec, size_t n_elems, double count_before, float* mean, float* m2,
                      cudaStream_t stream);
int launch_p_step(const McpcPStep* s, cudaStream_t stream);
int launch_tma_probe(const float* A, const float* B, int N, int a_mn, int b_mn, float* D, void* ws, cudaStream_t stream);
int launch_umma_probe(const float* Wt, const float* Bx, const float* G, int Kin, int N, float* D1, float* D2, void* ws,
                      cudaStream_t stream);

}  // namespace mcpc
