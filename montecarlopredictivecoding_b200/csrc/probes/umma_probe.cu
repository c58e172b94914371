// mcpc_debug_umma: a one-CTA known-answer test of the tensor-core primitives the fused bf16 kernel is
// built from (umma.cuh): canonical no-swizzle operand layouts, K-major and MN-major (transposed) reads
// of ONE weight tile, bulk async copy + mbarrier, TMEM alloc / tcgen05.ld.  tests/test_gpu_umma.py
// compares both products with torch.
//     D1[128][N] = Wt[128][Kin] . Bx[N][Kin]^T          (prediction-style: A = tile, K-major)
//     D2[m][N]   = sum_j Wt[j][m] . G[N][j]             (back-projection-style: A = tile^T, MN-major), m < Kin
#include <cstdlib>

#include "mcpc_common.cuh"
#include "umma.cuh"

namespace mcpc {
namespace {

using namespace umma;

__global__ void pack_tile_kernel(const float* __restrict__ W, int ld, int rows, int cols, int Kp, __nv_bfloat16* __restrict__ out) {
  // out: 128 x Kp bf16 in canonical K-major order (LBO = 128 B, SBO = Kp/8*128 B), zero padded
  const uint32_t sbo = (uint32_t)(Kp / 8) * 128u;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 128 * Kp; i += gridDim.x * blockDim.x) {
    const int r = i / Kp, k = i % Kp;
    const float v = (r < rows && k < cols) ? W[(size_t)r * ld + k] : 0.0f;
    out[kmajor_off(r, k, 128u, sbo) / 2] = __float2bfloat16(v);
  }
}

__global__ void __launch_bounds__(160) umma_probe_kernel(const __nv_bfloat16* __restrict__ packed, const float* __restrict__ Bx,
                                                         const float* __restrict__ G, int Kin, int N, float* __restrict__ D1,
                                                         float* __restrict__ D2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_w, bar_m1, bar_m2;
  __shared__ uint32_t tmem_base_s;
  const uint32_t S = (uint32_t)(Kin / 8) * 128u;           // SBO of the weight tile
  uint8_t* tile = smem;                                     // 128 x Kin bf16
  uint8_t* b1 = tile + 128 * Kin * 2 + 4096;                // [N x Kin] K-major (slack: MN-major reads overrun the tile)
  uint8_t* b2 = b1 + N * Kin * 2;                           // [N x 128] K-major
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_m1, 1);
    mbar_init(&bar_m2, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(&tmem_base_s, 64);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    mbar_expect_tx(&bar_w, (uint32_t)(128 * Kin * 2));
    bulk_g2s(tile, packed, (uint32_t)(128 * Kin * 2), &bar_w);
  }
  for (int i = tid; i < N * Kin; i += blockDim.x) {
    const int n = i / Kin, k = i % Kin;
    *reinterpret_cast<__nv_bfloat16*>(b1 + kmajor_off(n, k, 128u, S)) = __float2bfloat16(Bx[i]);
  }
  for (int i = tid; i < N * 128; i += blockDim.x) {
    const int n = i / 128, k = i % 128;
    *reinterpret_cast<__nv_bfloat16*>(b2 + kmajor_off(n, k, 128u, 2048u)) = __float2bfloat16(G[i]);
  }
  fence_async_smem();
  __syncthreads();

  if (warp == 4 && lane == 0) {
    mbar_wait(&bar_w, 0);
    fence_after_sync();
    const uint32_t id_a = idesc_bf16(128, N, false, false);
    for (int ks = 0; ks < Kin / 16; ++ks)
      mma_bf16_ss(tmem, smem_desc(smem_u32(tile) + ks * 256, 128u, S), smem_desc(smem_u32(b1) + ks * 256, 128u, S), id_a, ks > 0);
    mma_commit(&bar_m1);
    const uint32_t id_b = idesc_bf16(128, N, true, false);
    for (int ks = 0; ks < 8; ++ks)      // K' = 128 output units, 16 per instruction = 2 row groups of the tile
      mma_bf16_ss(tmem + N, smem_desc(smem_u32(tile) + ks * 2 * S, /*LBO (K' dir)*/ S, /*SBO (M' dir)*/ 128u),
                  smem_desc(smem_u32(b2) + ks * 256, 128u, 2048u), id_b, ks > 0);
    mma_commit(&bar_m2);
  }
  if (warp < 4) {
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    float v[16];
    mbar_wait(&bar_m1, 0);
    fence_after_sync();
    for (int c = 0; c < N; c += 16) {
      tmem_ld16(lane_addr + c, v);
      for (int i = 0; i < 16; ++i) D1[(size_t)row * N + c + i] = v[i];
    }
    mbar_wait(&bar_m2, 0);
    fence_after_sync();
    for (int c = 0; c < N; c += 16) {
      tmem_ld16(lane_addr + N + c, v);
      for (int i = 0; i < 16; ++i) D2[(size_t)row * N + c + i] = v[i];
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 64);
}

// Cost model of small tcgen05.mma instructions (debug, MCPC_UMMA_TIMING=1): cycles from first issue to the
// commit barrier for n_mma instructions of shape M x N x 16 spread round-robin over n_acc accumulators.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

// variant: 0 = issued inside `if (lane == 0)`, 1 = converged warp + elect.sync, 2 = as 1 with SWIZZLE_128B
// descriptors, 3 = as 1 with the A operand in TMEM, 4 = as 1 with an MN-major A operand in the no-swizzle canonical
// layout (the resident kernel's back-projection), 5 = MN-major A with SWIZZLE_128B, 6 = as 1, unrolled by 8
__global__ void __launch_bounds__(160) umma_timing_kernel(int M, int N, int n_mma, int n_acc, int variant, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 4) tmem_alloc(&tmem_base_s, 512);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 4) {
    const uint32_t id = idesc_bf16(M, N, variant >= 4, false);
    uint64_t ad = smem_desc(smem_u32(smem), 128u, 2048u);
    uint32_t a_step = (variant == 2) ? 2u : 16u;
    if (variant == 4) { ad = smem_desc(smem_u32(smem), 2048u, 128u); a_step = 256u; }
    if (variant == 5) { ad = smem_desc(smem_u32(smem), 8192u, 1024u) | ((uint64_t)2 << 61); a_step = 128u; }
    uint64_t bd = smem_desc(smem_u32(smem) + 32768, 128u, 256u);
    if (variant == 2) {
      ad = smem_desc(smem_u32(smem), 16u, 1024u) | ((uint64_t)2 << 61);
      bd = smem_desc(smem_u32(smem) + 32768, 16u, 1024u) | ((uint64_t)2 << 61);
    }
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = 0, t1 = 0, t2 = 0;
      if (variant == 0) {
        if (lane == 0) {
          t0 = clock64();
          for (int i = 0; i < n_mma; ++i)
            mma_bf16_ss(tmem + (uint32_t)((i % n_acc) * N), ad + (uint64_t)((i & 7) * 16), bd, id, i >= n_acc);
          mma_commit(&bar);
          t1 = clock64();
        }
      } else {
        t0 = clock64();
        if (elect_one()) {
          if (variant == 6) {
            // descriptors precomputed: fully unrolled groups of 8 so that every MMA reads its own uniform registers
            for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                mma_bf16_ss(tmem, ad + (uint64_t)(j * 16), bd + (uint64_t)(j * 16), id, (i + j) > 0);
            }
          } else if (variant == 3) {
            for (int i = 0; i < n_mma; ++i)
              mma_bf16_ts(tmem + (uint32_t)((i % n_acc) * N), tmem + 256 + (uint32_t)((i & 7) * 8), bd, id, i >= n_acc);
          } else {
            for (int i = 0; i < n_mma; ++i)
              mma_bf16_ss(tmem + (uint32_t)((i % n_acc) * N), ad + (uint64_t)((i & 7) * a_step), bd, id, i >= n_acc);
          }
          mma_commit(&bar);
        }
        __syncwarp();
        t1 = clock64();
      }
      mbar_wait(&bar, rep & 1);
      t2 = clock64();
      if (rep == 2 && lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

}  // namespace

int launch_umma_timing(cudaStream_t stream) {
  long long* d = nullptr;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(umma_timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int Ms[2] = {128, 64};
  const int Ns[5] = {16, 32, 64, 128, 256};
  for (int variant = 0; variant < 7; ++variant)
    for (int mi = 0; mi < 2; ++mi)
      for (int ni = 0; ni < 5; ni += 2)
        for (int n_acc = 1; n_acc <= 2; ++n_acc) {
          const int M = Ms[mi], N = Ns[ni];
          if (n_acc * N > 256 || (variant >= 3 && M == 64)) continue;
          umma_timing_kernel<<<1, 160, 64 * 1024, stream>>>(M, N, 32, n_acc, variant, d);
          long long h[2];
          cudaStreamSynchronize(stream);
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          fprintf(stderr, "[umma timing] variant=%d M=%d N=%d n_acc=%d: issue %lld cyc, complete %lld cyc for 32 MMAs (%.1f cyc/MMA)\n",
                  variant, M, N, n_acc, h[0], h[1], h[1] / 32.0);
        }
  cudaFree(d);
  return MCPC_OK;
}

int launch_umma_probe(const float* Wt, const float* Bx, const float* G, int Kin, int N, float* D1, float* D2, void* ws,
                      cudaStream_t stream) {
#ifdef MCPC_DEBUG_BUILD
  if (getenv("MCPC_UMMA_TIMING") != nullptr) return launch_umma_timing(stream);      // tcgen05.mma issue-cost table (experiment)
#endif
  if (Kin % 16 != 0 || Kin < 16 || Kin > 256 || N % 16 != 0 || N < 16 || N > 32) {
    set_error("umma probe: Kin must be a multiple of 16 in [16,256], N in {16,32}");
    return MCPC_ERR_INVALID;
  }
  __nv_bfloat16* packed = reinterpret_cast<__nv_bfloat16*>(ws);
  pack_tile_kernel<<<32, 256, 0, stream>>>(Wt, Kin, 128, Kin, Kin, packed);
  const size_t smem = (size_t)128 * Kin * 2 + 4096 + (size_t)N * Kin * 2 + (size_t)N * 128 * 2 + 1024;
  MCPC_CUDA_CHECK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 160, smem, stream>>>(packed, Bx, G, Kin, N, D1, D2);
  MCPC_CUDA_CHECK(cudaGetLastError());
  count_launch(2);
  return MCPC_OK;
}

}  // namespace mcpc
