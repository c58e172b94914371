// mcpc_weight_grad, MCPC_PREC_FP32: the local (Hebbian) weight update
//     gW_l += G_l^T act(x_{l-1}),   gb_l += colsum(G_l)
// summed over the saved steps and the batch -- what autograd's dW contractions of
// predictive_coding/pc_trainer.py:862 leave in `.grad`.  G and act(x)
// were written by mcpc_infer for the steps of the accumulate window, so the whole window is
// ONE reduction of length n_save*B per layer (split-K over the grid, fp32 FMA, fp32 atomics
// for the cross-CTA sum).
#include "mcpc_common.cuh"

namespace mcpc {

namespace {

constexpr int kTM = 64, kTN = 64, kTK = 16;

// C[M][N] += sum_r A[r][m] * Bm[r % b_mod][n]; A has leading dim lda, Bm ldb.  bias[m] += sum_r A[r][m].
__global__ void __launch_bounds__(256) wgrad_tn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm,
                                                       int ldb, int b_mod, float* __restrict__ C, float* __restrict__ bias,
                                                       int M, int N, int rows, int rows_per_slab) {
  __shared__ float As[kTK][kTM + 4];
  __shared__ float Bs[kTK][kTN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, 4x4 outputs each
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
  const int r_begin = blockIdx.z * rows_per_slab;
  const int r_end = min(rows, r_begin + rows_per_slab);
  const bool do_bias = (bias != nullptr) && (blockIdx.x == 0);
  const bool do_w = (C != nullptr);
  float acc[4][4] = {};
  float bsum = 0.0f;
  for (int r0 = r_begin; r0 < r_end; r0 += kTK) {
    // stage a [kTK][64] slab of each operand (rows are the reduction index; columns contiguous)
    for (int i = tid; i < kTK * kTM; i += 256) {
      const int kk = i / kTM, c = i % kTM;
      const int r = r0 + kk;
      As[kk][c] = (r < r_end && m0 + c < M) ? A[(size_t)r * lda + m0 + c] : 0.0f;
    }
    if (do_w) {
      for (int i = tid; i < kTK * kTN; i += 256) {
        const int kk = i / kTN, c = i % kTN;
        const int r = r0 + kk;
        Bs[kk][c] = (r < r_end && n0 + c < N) ? Bm[(size_t)(r % b_mod) * ldb + n0 + c] : 0.0f;
      }
    }
    __syncthreads();
    if (do_w) {
#pragma unroll
      for (int kk = 0; kk < kTK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    if (do_bias && tid < kTM) {
#pragma unroll
      for (int kk = 0; kk < kTK; ++kk) bsum += As[kk][tid];
    }
    __syncthreads();
  }
  if (do_w) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
        if (m < M && n < N) atomicAdd(&C[(size_t)m * N + n], acc[i][j]);
      }
  }
  if (do_bias && tid < kTM && m0 + tid < M) atomicAdd(&bias[m0 + tid], bsum);
}

}  // namespace

// C[M][N] += A^T Bm over `rows` fp32 rows (one launch; used by the bf16 path for Linear_0 with non-zero inputs)
int launch_wgrad_tn_fp32(const float* A, int lda, const float* Bm, int ldb, float* C, int M, int N, int rows,
                         cudaStream_t stream) {
  const int gx = (N + kTN - 1) / kTN, gy = (M + kTM - 1) / kTM;
  int slabs = (148 * 4 + gx * gy - 1) / (gx * gy);
  int rows_per_slab = (rows + slabs - 1) / slabs;
  rows_per_slab = ((rows_per_slab + kTK - 1) / kTK) * kTK;
  if (rows_per_slab < 4 * kTK) rows_per_slab = 4 * kTK;
  slabs = (rows + rows_per_slab - 1) / rows_per_slab;
  wgrad_tn_kernel<<<dim3(gx, gy, slabs), 256, 0, stream>>>(A, lda, Bm, ldb, rows, C, nullptr, M, N, rows, rows_per_slab);
  MCPC_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return MCPC_OK;
}

int launch_weight_grad_fp32(const NetDev& nd, const McpcGradIO* io, int B, int n_save, cudaStream_t stream) {
  const float* G = reinterpret_cast<const float*>(io->save_g);
  const float* F = reinterpret_cast<const float*>(io->save_f);
  const int rows = n_save * B;
  const int n_lin = nd.L + (nd.d_out > 0 ? 1 : 0);
  for (int l = 0; l < n_lin; ++l) {
    const bool is_out = (l == nd.L);
    if (is_out && !nd.top_has_grad) continue;                 // readout-only Linear: its gradient is identically 0
    const int M = is_out ? nd.d_out : nd.dims[l];
    const int N = (l == 0) ? nd.d_in : nd.dims[l - 1];
    const float* A = G + (is_out ? nd.SD : nd.off[l]);
    const float* Bm = nullptr;
    int ldb = 0, b_mod = rows;
    float* C = io->gW[l];
    if (l == 0) {
      if (io->inputs == nullptr) C = nullptr;                 // zero inputs: Linear_0's weight gradient is exactly 0
      Bm = io->inputs; ldb = nd.d_in; b_mod = B;
    } else {
      Bm = F + nd.off[l - 1]; ldb = nd.SD;
    }
    float* bias = io->gb[l];
    if (C == nullptr && bias == nullptr) continue;
    const int gx = C != nullptr ? (N + kTN - 1) / kTN : 1, gy = (M + kTM - 1) / kTM;
    int slabs = (148 * 4 + gx * gy - 1) / (gx * gy);
    int rows_per_slab = (rows + slabs - 1) / slabs;
    rows_per_slab = ((rows_per_slab + kTK - 1) / kTK) * kTK;
    if (rows_per_slab < 4 * kTK) rows_per_slab = 4 * kTK;
    slabs = (rows + rows_per_slab - 1) / rows_per_slab;
    dim3 grid(gx, gy, slabs);
    wgrad_tn_kernel<<<grid, 256, 0, stream>>>(A, nd.NG, Bm, ldb, b_mod, C, bias, M, N, rows, rows_per_slab);
    MCPC_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  return MCPC_OK;
}

}  // namespace mcpc
