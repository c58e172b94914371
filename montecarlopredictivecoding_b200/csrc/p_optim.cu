// mcpc_p_step -- SURVEY §8(f) N3: the parameter update of pc_trainer.py:904-914 as ONE kernel over all W / b tensors:
//     param.grad = param.grad / (len(accumulate_p_at) * batch_size)       (:905-913)
//     optimizer_p.step()                                                   (:914; optim.SGD or optim.Adam)
// The reference (and round 1 of this build) runs a div per parameter plus torch's multi-tensor optimizer kernels
// (read g, write g, read p / state, write p / state: 4-6 launches and ~2x the bytes).  This kernel reads g, p and the
// state once, writes g (normalised, as the reference leaves it), p and the state once.  It works IN PLACE on the
// state tensors of the user's torch optimizer (momentum_buffer / exp_avg / exp_avg_sq), so the optimizer object stays
// consistent for state_dict(), later torch steps and learning-rate schedulers.  HBM-bound: 5 (SGD) / 7 (Adam) floats of
// traffic per parameter.  Arithmetic mirrors torch/optim/sgd.py::_single_tensor_sgd and adam.py::_single_tensor_adam.
#include "mcpc_common.cuh"

namespace mcpc {
namespace {

struct PStepDev {
  float* param[MCPC_MAX_PTENSORS];
  float* grad[MCPC_MAX_PTENSORS];
  float* s1[MCPC_MAX_PTENSORS];
  float* s2[MCPC_MAX_PTENSORS];
  unsigned long long numel[MCPC_MAX_PTENSORS];
  int kind, nesterov, first_step;
  float inv_norm, lr, wd, momentum, one_minus_damp, beta1, beta2, eps;
  float step_size, inv_sqrt_bc2;            // Adam: lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t)
};

__global__ void p_step_kernel(const __grid_constant__ PStepDev a) {
  const int ti = blockIdx.y;
  float* __restrict__ p = a.param[ti];
  float* __restrict__ g = a.grad[ti];
  float* __restrict__ s1 = a.s1[ti];
  float* __restrict__ s2 = a.s2[ti];
  const size_t n = a.numel[ti];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gn = g[i] * a.inv_norm;
    g[i] = gn;                                              // the reference leaves the divided gradient in .grad
    float w = p[i];
    float d = (a.wd != 0.0f) ? fmaf(a.wd, w, gn) : gn;      // L2 weight decay (both optimizers)
    if (a.kind == MCPC_OPT_SGD) {
      if (s1 != nullptr) {                                  // momentum
        float buf = a.first_step ? d : fmaf(a.momentum, s1[i], a.one_minus_damp * d);
        s1[i] = buf;
        d = a.nesterov ? fmaf(a.momentum, buf, d) : buf;
      }
      w = fmaf(-a.lr, d, w);
    } else {
      float m = s1[i], v = s2[i];
      m = fmaf(1.0f - a.beta1, d - m, m);                   // exp_avg.lerp_(grad, 1 - beta1)
      v = fmaf((1.0f - a.beta2) * d, d, a.beta2 * v);       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      s1[i] = m;
      s2[i] = v;
      const float denom = fmaf(sqrtf(v), a.inv_sqrt_bc2, a.eps);
      w = fmaf(-a.step_size, m / denom, w);
    }
    p[i] = w;
  }
}

}  // namespace

int launch_p_step(const McpcPStep* s, cudaStream_t stream) {
  PStepDev a{};
  size_t biggest = 0;
  for (int i = 0; i < s->n_tensors; ++i) {
    a.param[i] = s->param[i];
    a.grad[i] = s->grad[i];
    a.s1[i] = s->state1[i];
    a.s2[i] = s->state2[i];
    a.numel[i] = s->numel[i];
    if (s->numel[i] > biggest) biggest = s->numel[i];
  }
  a.kind = s->kind;
  a.nesterov = s->nesterov;
  a.first_step = s->first_step;
  a.inv_norm = (float)s->inv_norm;
  a.lr = (float)s->lr;
  a.wd = (float)s->weight_decay;
  a.momentum = (float)s->momentum;
  a.one_minus_damp = (float)(1.0 - s->dampening);
  a.beta1 = (float)s->beta1;
  a.beta2 = (float)s->beta2;
  a.eps = (float)s->eps;
  if (s->kind == MCPC_OPT_ADAM) {
    const double bc1 = 1.0 - pow(s->beta1, (double)s->step), bc2 = 1.0 - pow(s->beta2, (double)s->step);
    a.step_size = (float)(s->lr / bc1);
    a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  }
  const size_t want = (biggest + 1023) / 1024;             // 4 elements per thread in flight
  dim3 grid((unsigned)(want < 592 ? (want > 0 ? want : 1) : 592), (unsigned)s->n_tensors);
  p_step_kernel<<<grid, 256, 0, stream>>>(a);
  MCPC_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return MCPC_OK;
}

}  // namespace mcpc
