// mcpc_traj_stats_update -- SURVEY §8(f) N2: on-device running mean / variance of recorded latent trajectories.
//
// The reference ships every step's latents to the host (`.clone().detach().cpu()`, pc_trainer.py:772-774), concatenates
// Python lists of tensors and reduces them there (utils/model.py:143-149 `temp.mean(0)`; figure_2.py:75-79 posterior
// mean / variance of 9,000 recorded samples).  Here the trajectory stays in a device ring [n_rec][n_elems] written by
// the inference kernels (McpcIO.traj_x with McpcOpts.traj_every thinning) and is folded into running per-element
// (mean, M2) accumulators -- Welford's update, one thread per element, every read a contiguous row segment: the kernel
// reads each recorded value exactly once (HBM-bound, 4 bytes per sample) and a T >= 10^4 run needs only a bounded ring.
#include "mcpc_common.cuh"

namespace mcpc {
namespace {

__global__ void traj_stats_kernel(const float* __restrict__ traj, int n_rec, size_t n_elems, double count_before,
                                  float* __restrict__ mean, float* __restrict__ m2) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n_elems; e += (size_t)gridDim.x * blockDim.x) {
    float mu = count_before > 0.0 ? mean[e] : 0.0f;
    float s2 = count_before > 0.0 ? m2[e] : 0.0f;
    float n = (float)count_before;
    const float* src = traj + e;
    int r = 0;
    for (; r + 4 <= n_rec; r += 4) {                       // 4 independent loads in flight per thread
      const float v0 = src[(size_t)r * n_elems], v1 = src[(size_t)(r + 1) * n_elems];
      const float v2 = src[(size_t)(r + 2) * n_elems], v3 = src[(size_t)(r + 3) * n_elems];
      float d;
      n += 1.0f; d = v0 - mu; mu += __fdividef(d, n); s2 = fmaf(d, v0 - mu, s2);
      n += 1.0f; d = v1 - mu; mu += __fdividef(d, n); s2 = fmaf(d, v1 - mu, s2);
      n += 1.0f; d = v2 - mu; mu += __fdividef(d, n); s2 = fmaf(d, v2 - mu, s2);
      n += 1.0f; d = v3 - mu; mu += __fdividef(d, n); s2 = fmaf(d, v3 - mu, s2);
    }
    for (; r < n_rec; ++r) {
      const float v = src[(size_t)r * n_elems];
      n += 1.0f;
      const float d = v - mu;
      mu += __fdividef(d, n);
      s2 = fmaf(d, v - mu, s2);
    }
    mean[e] = mu;
    m2[e] = s2;
  }
}

}  // namespace

int launch_traj_stats(const float* traj, int n_rec, size_t n_elems, double count_before, float* mean, float* m2,
                      cudaStream_t stream) {
  const size_t want = (n_elems + 255) / 256;
  const int blocks = (int)(want < (size_t)148 * 16 ? want : (size_t)148 * 16);
  traj_stats_kernel<<<blocks > 0 ? blocks : 1, 256, 0, stream>>>(traj, n_rec, n_elems, count_before, mean, m2);
  MCPC_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return MCPC_OK;
}

}  // namespace mcpc
