// mcpc_marginal_ll_bernoulli: the importance-sampling estimate of the marginal log-likelihood of a Bernoulli
// generative model (reference utils/training_evaluation.py:177-206, get_marginal_likelihood).
//
// The reference materialises, on the CPU, losses[i, s] = sum_j BCEWithLogits(o[s, j], y[i, j]) for every data row i
// and every prior sample s (4000 x 5000 x 784 elementwise terms per DataLoader batch) and then reduces
//     m_i = min_s losses[i, s],   p_i = mean_s exp(-(losses[i, s] - m_i)),   ml = mean_i (log p_i - m_i).
// BCE-with-logits is linear in the target:  losses[i, s] = sp_s - (Y O^T)[i, s]  with  sp_s = sum_j softplus(o[s, j]),
// so the whole thing is ONE GEMM with a row-wise streaming min / sum-exp epilogue:
//   * operands: bf16 hi/lo splits concatenated along K -- A' = [Yhi | Yhi | Ylo], B' = [Ohi | Olo | Ohi] -- so that a
//     single K-major tcgen05 GEMM accumulates Yhi.Ohi + Yhi.Olo + Ylo.Ohi in fp32 (relative error ~2^-16 per
//     product, below the fp32 summation error of the reference); TMA + SWIZZLE_128B, 4-stage ring, as infer_wide.cu;
//   * persistent CTAs over 128 (data rows) x 256 (samples) tiles, two TMEM accumulators so the epilogue of tile i
//     overlaps the mainloop of tile i+1;
//   * epilogue: thread = data row; two passes over its 128 accumulator columns (min, then sum of exp) -> one
//     (m, s) partial per (row, half tile); mll_combine_kernel merges the partials of a row and averages in fp64.
#include <cfloat>

#include "mcpc_common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace mcpc {
namespace {

using namespace umma;

constexpr int kS = 4;                       // pipeline stages
constexpr int kBK = 64;                     // K elements per stage (128 bytes: one SWIZZLE_128B row)
constexpr int kBM = 128, kBN = 256;
constexpr uint32_t kABytes = kBM * kBK * 2, kBBytes = kBN * kBK * 2, kStage = kABytes + kBBytes;

struct MllParams {
  int N, S, K3;                             // data rows, samples, concatenated K extent (3 * pad64(D))
  int n_mt, n_nt;                           // tiles along rows / samples
  const float* sp;                          // [n_nt * 256] softplus sums, +inf past S
  float2* partials;                         // [n_mt * 128][n_nt * 2] (min, sum of exp) per row and half tile
};

struct MllPipe {
  uint64_t full[kS], empty[kS], acc_full[2], acc_empty[2];
};

__device__ __forceinline__ bool elect1m() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

// fp32 [R x D] -> bf16 [R x 3*Kp]: hi/lo split in the block order `lo_block` asks for (1: [hi|lo|hi] for the logits,
// 2: [hi|hi|lo] for the data); logits are clamped to +-clamp_abs first (training_evaluation.py:180) and their softplus
// row sums written to sp.  One warp per row.
__global__ void mll_prep_kernel(const float* __restrict__ X, int R, int D, int Kp, int lo_block, float clamp_abs,
                                __nv_bfloat16* __restrict__ out, float* __restrict__ sp) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const float* x = X + (size_t)warp * D;
  __nv_bfloat16* o = out + (size_t)warp * 3 * Kp;
  float acc = 0.0f;
  for (int j = lane; j < Kp; j += 32) {
    float v = 0.0f;
    if (j < D) {
      v = x[j];
      if (clamp_abs > 0.0f) v = fminf(fmaxf(v, -clamp_abs), clamp_abs);
      if (sp != nullptr) acc += fmaxf(v, 0.0f) + log1pf(expf(-fabsf(v)));     // softplus, overflow-free
    }
    const __nv_bfloat16 hi = __float2bfloat16(v);
    const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
    o[j] = hi;
    o[Kp + j] = (lo_block == 1) ? lo : hi;
    o[2 * Kp + j] = (lo_block == 2) ? lo : hi;
  }
  if (sp != nullptr) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) sp[warp] = acc;
  }
}

__global__ void mll_fill_kernel(float* p, int begin, int end, float v) {
  const int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < end) p[i] = v;
}

__global__ void __launch_bounds__(320, 1) mll_gemm_kernel(const __grid_constant__ MllParams p,
                                                          const __grid_constant__ CUtensorMap mapA,
                                                          const __grid_constant__ CUtensorMap mapB) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ MllPipe pipe;
  __shared__ uint32_t tmem_s;
  __shared__ float s_sp[2][kBN];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = p.n_mt * p.n_nt;
  const int n_stage = p.K3 / kBK;

  if (tid == 0) {
    for (int s = 0; s < kS; ++s) {
      mbar_init(&pipe.full[s], 1);
      mbar_init(&pipe.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&pipe.acc_full[b], 1);
      mbar_init(&pipe.acc_empty[b], 256);
    }
    fence_mbar_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1) tmem_alloc(&tmem_s, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_s;
  const uint32_t smem_base = smem_u32(smem);

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      uint32_t issued = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.n_nt) * kBM, n0 = (tile % p.n_nt) * kBN;
        for (int s = 0; s < n_stage; ++s, ++issued) {
          const uint32_t slot = issued % kS;
          mbar_wait(&pipe.empty[slot], ((issued / kS) & 1u) ^ 1u);
          uint8_t* sa = smem + slot * kStage;
          mbar_expect_tx(&pipe.full[slot], kStage);
          tma_load_2d(sa, &mapA, s * kBK, m0, &pipe.full[slot]);
          tma_load_2d(sa + kABytes, &mapB, s * kBK, n0, &pipe.full[slot]);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: both operands K-major, SWIZZLE_128B (tma.cuh) ----------------
    const uint32_t id = idesc_bf16(kBM, kBN, false, false);
    uint32_t sc = 0, gi = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++gi) {
      const uint32_t ab = gi & 1u;
      mbar_wait(&pipe.acc_empty[ab], ((gi >> 1) & 1u) ^ 1u);
      fence_after_sync();
      for (int s = 0; s < n_stage; ++s, ++sc) {
        const uint32_t slot = sc % kS;
        mbar_wait(&pipe.full[slot], (sc / kS) & 1u);
        fence_after_sync();
        const uint64_t ad0 = smem_desc_sw128(smem_base + slot * kStage, 16u, 1024u);
        const uint64_t bd0 = smem_desc_sw128(smem_base + slot * kStage + kABytes, 16u, 1024u);
        if (elect1m()) {
#pragma unroll
          for (int ks = 0; ks < kBK / 16; ++ks)            // 16 K elements = 32 bytes further into every row
            mma_bf16_ss(tmem + ab * kBN, ad0 + (uint64_t)(ks * 2), bd0 + (uint64_t)(ks * 2), id, s > 0 || ks > 0);
          mma_commit(&pipe.empty[slot]);
          if (s == n_stage - 1) mma_commit(&pipe.acc_full[ab]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------- epilogue warps 2..9: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 ----------------
    const int q = warp & 3, ew = warp - 2, half = ew >> 2;
    const int etid = tid - 64;                             // 0..255
    const int c_begin = half * (kBN / 2);
    const int row_in_tile = q * 32 + lane;
    uint32_t gi = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++gi) {
      const int mt = tile / p.n_nt, nt = tile % p.n_nt;
      const uint32_t ab = gi & 1u;
      s_sp[ab][etid] = __ldg(p.sp + nt * kBN + etid);      // this buffer was last read two tiles ago (bar below)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&pipe.acc_full[ab], (gi >> 1) & 1u);
      fence_after_sync();
      const uint32_t acc = tmem + ((uint32_t)(q * 32) << 16) + ab * kBN + c_begin;
      const float* sp = s_sp[ab] + c_begin;
      // pass 1: the smallest loss of this row among the 128 samples of the half tile
      float m = FLT_MAX;
      for (int c = 0; c < kBN / 2; c += 16) {
        float v[16];
        tmem_ld16(acc + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) m = fminf(m, sp[c + i] - v[i]);
      }
      // pass 2: sum of exp(-(loss - m)); padded samples carry sp = +inf and add exactly 0
      float s = 0.0f;
      for (int c = 0; c < kBN / 2; c += 16) {
        float v[16];
        tmem_ld16(acc + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) s += __expf(m - (sp[c + i] - v[i]));
      }
      fence_before_sync();
      mbar_arrive(&pipe.acc_empty[ab]);
      p.partials[(size_t)(mt * kBM + row_in_tile) * (p.n_nt * 2) + nt * 2 + half] = make_float2(m, s);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// one thread per data row: merge its (min, sum-exp) partials, row_ll = log(mean_s exp(-loss)) ; block sums in fp64
__global__ void mll_combine_kernel(const float2* __restrict__ partials, int N, int S, int n_part, float* __restrict__ row_ll,
                                   double* __restrict__ ml_sum) {
  __shared__ double red[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double val = 0.0;
  if (i < N) {
    const float2* pr = partials + (size_t)i * n_part;
    float m = FLT_MAX;
    for (int k = 0; k < n_part; ++k) m = fminf(m, pr[k].x);
    float s = 0.0f;
    for (int k = 0; k < n_part; ++k) s += pr[k].y * expf(m - pr[k].x);
    const float ll = logf(s / (float)S) - m;               // training_evaluation.py:203-205
    if (row_ll != nullptr) row_ll[i] = ll;
    val = (double)ll;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    atomicAdd(ml_sum, t / (double)N);
  }
}

inline int pad_to(int v, int a) { return (v + a - 1) / a * a; }

struct MllLayout {
  int Kp, n_mt, n_nt;
  size_t off_a, off_b, off_sp, off_part, total;
};

MllLayout mll_layout(int N, int S, int D) {
  MllLayout L{};
  L.Kp = pad_to(D, kBK);
  L.n_mt = (N + kBM - 1) / kBM;
  L.n_nt = (S + kBN - 1) / kBN;
  size_t o = 0;
  L.off_a = o;
  o += (size_t)N * 3 * L.Kp * 2;
  o = (o + 1023) & ~(size_t)1023;
  L.off_b = o;
  o += (size_t)S * 3 * L.Kp * 2;
  o = (o + 1023) & ~(size_t)1023;
  L.off_sp = o;
  o += (size_t)L.n_nt * kBN * 4;
  o = (o + 1023) & ~(size_t)1023;
  L.off_part = o;
  o += (size_t)L.n_mt * kBM * L.n_nt * 2 * sizeof(float2);
  L.total = o + 1024;
  return L;
}

}  // namespace

int marginal_ll_workspace(int N, int S, int D, size_t* bytes) {
  if (N < 1 || S < 1 || D < 1) {
    set_error("marginal likelihood: N, S and D must be positive (got %d, %d, %d)", N, S, D);
    return MCPC_ERR_INVALID;
  }
  *bytes = mll_layout(N, S, D).total;
  return MCPC_OK;
}

int launch_marginal_ll(const float* logits, int S, const float* data, int N, int D, float clamp_abs, void* ws,
                       size_t ws_bytes, double* ml_out, float* row_ll, cudaStream_t stream) {
  size_t need = 0;
  int rc = marginal_ll_workspace(N, S, D, &need);
  if (rc != MCPC_OK) return rc;
  if (logits == nullptr || data == nullptr || ml_out == nullptr) {
    set_error("marginal likelihood: logits, data and ml_out must not be NULL");
    return MCPC_ERR_INVALID;
  }
  if (ws == nullptr || ws_bytes < need) {
    set_error("workspace too small: %zu B given, %zu B needed", ws_bytes, need);
    return MCPC_ERR_WORKSPACE;
  }
  const MllLayout L = mll_layout(N, S, D);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(base + L.off_a);
  __nv_bfloat16* Bm = reinterpret_cast<__nv_bfloat16*>(base + L.off_b);
  float* sp = reinterpret_cast<float*>(base + L.off_sp);
  float2* partials = reinterpret_cast<float2*>(base + L.off_part);

  mll_prep_kernel<<<(N * 32 + 255) / 256, 256, 0, stream>>>(data, N, D, L.Kp, 2, 0.0f, A, nullptr);
  mll_prep_kernel<<<(S * 32 + 255) / 256, 256, 0, stream>>>(logits, S, D, L.Kp, 1, clamp_abs, Bm, sp);
  const int sp_end = L.n_nt * kBN;
  if (sp_end > S) mll_fill_kernel<<<(sp_end - S + 255) / 256, 256, 0, stream>>>(sp, S, sp_end, INFINITY);
  MCPC_CUDA_CHECK(cudaMemsetAsync(ml_out, 0, sizeof(double), stream));
  count_launch(3);

  MllParams p{};
  p.N = N;
  p.S = S;
  p.K3 = 3 * L.Kp;
  p.n_mt = L.n_mt;
  p.n_nt = L.n_nt;
  p.sp = sp;
  p.partials = partials;
  CUtensorMap mapA, mapB;
  rc = make_tmap_bf16(&mapA, A, (uint64_t)p.K3, (uint64_t)N, (uint64_t)p.K3, kBK, kBM);
  if (rc != MCPC_OK) return rc;
  rc = make_tmap_bf16(&mapB, Bm, (uint64_t)p.K3, (uint64_t)S, (uint64_t)p.K3, kBK, kBN);
  if (rc != MCPC_OK) return rc;
  const size_t smem = (size_t)kS * kStage + 1024;
  MCPC_CUDA_CHECK(cudaFuncSetAttribute(mll_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int n_tiles = p.n_mt * p.n_nt;
  mll_gemm_kernel<<<n_tiles < n_sm ? n_tiles : n_sm, 320, smem, stream>>>(p, mapA, mapB);
  MCPC_CUDA_CHECK(cudaGetLastError());
  mll_combine_kernel<<<(N + 255) / 256, 256, 0, stream>>>(partials, N, S, p.n_nt * 2, row_ll, ml_out);
  MCPC_CUDA_CHECK(cudaGetLastError());
  count_launch(2);
  return MCPC_OK;
}

}  // namespace mcpc
