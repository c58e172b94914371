// mcpc_debug_tma: known-answer test of the TMA (cp.async.bulk.tensor, SASS UTMALDG) + SWIZZLE_128B operand path
// the streaming kernels use.  One CTA loads a 128 x 64 A tile and an N x 64 B tile (K = 64, bf16) with tensor
// maps and multiplies them four ways: both operands K-major or MN-major ("transposed storage").
//   a_mn = 0: A stored [128][64] (k contiguous);   a_mn = 1: A stored [64][128] (m contiguous)
//   b_mn = 0: B stored [N][64];                    b_mn = 1: B stored [64][N]
//   D [128][N] = A . B^T  in every case.
#include <cuda.h>

#include "mcpc_common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace mcpc {
namespace {

using namespace umma;

__global__ void __launch_bounds__(160) tma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                        int N, int a_mn, int b_mn, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_ld, bar_mma;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sa = smem;                      // 16 KB: A stage
  uint8_t* sb = smem + 16384;              // up to 32 KB: B stage
  if (tid == 0) {
    mbar_init(&bar_ld, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(&tmem_s, 256);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_s;
  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(&bar_ld, (uint32_t)(128 * 64 * 2 + N * 64 * 2));
      // K-major: one box [64 k x rows]; MN-major: boxes of [64 mn x 64 k], 8 KB each, one per 64 units
      if (!a_mn) tma_load_2d(sa, &tmA, 0, 0, &bar_ld);
      else
        for (int j = 0; j < 2; ++j) tma_load_2d(sa + j * 8192, &tmA, j * 64, 0, &bar_ld);
      if (!b_mn) tma_load_2d(sb, &tmB, 0, 0, &bar_ld);
      else
        for (int j = 0; j < N / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, j * 64, 0, &bar_ld);
    }
    __syncwarp();
    mbar_wait(&bar_ld, 0);
    fence_after_sync();
    if (lane == 0) {
      const uint32_t id = idesc_bf16(128, N, a_mn != 0, b_mn != 0);
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = a_mn ? smem_desc_sw128(smem_u32(sa) + ks * 2048, 8192u, 1024u) : smem_desc_sw128(smem_u32(sa) + ks * 32, 16u, 1024u);
        const uint64_t bd = b_mn ? smem_desc_sw128(smem_u32(sb) + ks * 2048, 8192u, 1024u) : smem_desc_sw128(smem_u32(sb) + ks * 32, 16u, 1024u);
        mma_bf16_ss(tmem, ad, bd, id, ks > 0);
      }
      mma_commit(&bar_mma);
    }
    __syncwarp();
  }
  if (warp < 4) {
    mbar_wait(&bar_mma, 0);
    fence_after_sync();
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < N; c += 16) {
      float v[16];
      tmem_ld16(lane_addr + c, v);
      for (int i = 0; i < 16; ++i) D[(size_t)row * N + c + i] = v[i];
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = __float2bfloat16(src[i]);
}

}  // namespace

// A: fp32 [128][64] (a_mn=0) or [64][128] (a_mn=1); B: fp32 [N][64] or [64][N]; ws: >= (128*64 + N*64)*2 bytes
int launch_tma_probe(const float* A, const float* B, int N, int a_mn, int b_mn, float* D, void* ws, cudaStream_t stream) {
  if (N % 64 != 0 || N < 64 || N > 256) {
    set_error("tma probe: N must be 64, 128, 192 or 256");
    return MCPC_ERR_INVALID;
  }
  __nv_bfloat16* Ab = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* Bb = Ab + 128 * 64;
  f32_to_bf16_kernel<<<16, 256, 0, stream>>>(A, Ab, 128 * 64);
  f32_to_bf16_kernel<<<16, 256, 0, stream>>>(B, Bb, N * 64);
  CUtensorMap tmA, tmB;
  int rc;
  // tensor map = (inner extent, outer extent, row pitch in elements, box inner, box outer)
  rc = a_mn ? make_tmap_bf16(&tmA, Ab, 128, 64, 128, 64, 64) : make_tmap_bf16(&tmA, Ab, 64, 128, 64, 64, 128);
  if (rc != MCPC_OK) return rc;
  rc = b_mn ? make_tmap_bf16(&tmB, Bb, N, 64, N, 64, 64) : make_tmap_bf16(&tmB, Bb, 64, N, 64, 64, N);
  if (rc != MCPC_OK) return rc;
  const size_t smem = 16384 + 32768 + 1024;
  MCPC_CUDA_CHECK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tma_probe_kernel<<<1, 160, smem, stream>>>(tmA, tmB, N, a_mn, b_mn, D);
  MCPC_CUDA_CHECK(cudaGetLastError());
  count_launch(3);
  return MCPC_OK;
}

}  // namespace mcpc
