// mcpc_debug_umma: a one-CTA known-answer test of the tensor-core primitives the fused bf16 kernel is
// built from (umma.cuh): canonical no-swizzle operand layouts, K-major and MN-major (transposed) reads
// of ONE weight tile, bulk async copy + mbarrier, TMEM alloc / tcgen05.ld.  tests/test_gpu_umma.py
// compares both products with torch.
//     D1[128][N] = Wt[128][Kin] . Bx[N][Kin]^T          (prediction-style: A = tile, K-major)
//     D2[m][N]   = sum_j Wt[j][m] . G[N][j]             (back-projection-style: A = tile^T, MN-major), m < Kin
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

}  // namespace

int launch_umma_probe(const float* Wt, const float* Bx, const float* G, int Kin, int N, float* D1, float* D2, void* ws,
                      cudaStream_t stream) {
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
