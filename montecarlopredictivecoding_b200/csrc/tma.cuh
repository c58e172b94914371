// TMA (cp.async.bulk.tensor, SASS UTMALDG) helpers: tensor-map construction through the driver entry point
// (no libcuda link dependency) and the SWIZZLE_128B shared-memory descriptors that match what TMA writes.
//
// SWIZZLE_128B layouts (bf16): shared-memory rows are 128 bytes (64 elements), 8 rows form a 1024-byte atom, the
// 16-byte chunk index of every row is XOR-ed with (row % 8).  TMA produces exactly this from a row-major box of
// 64 elements x R rows, conflict-free on both sides and with one instruction per box.
//   K-major operand [R x 64k]:    rows = M/N index.  desc: LBO field 1 (ignored), SBO = 1024 (next 8 rows);
//                                 the 16-element K step of one MMA advances the start address by 32 bytes.
//   MN-major operand [64k x 64mn]: rows = K index, one 8 KB block per 64 M/N units.  desc: LBO = 8192 (next block
//                                 of 64 units), SBO = 1024 (next 8 k-rows); one MMA (K=16) advances by 2048 bytes.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "mcpc_common.cuh"
#include "umma.cuh"

namespace mcpc {

typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TmapEncodeFn tmap_encode_fn() {
  static TmapEncodeFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TmapEncodeFn>(p);
  }
  return fn;
}

// Row-major bf16 matrix [outer][inner] with `pitch` elements between rows; box = box_inner x box_outer elements
// (box_inner * 2 bytes must be 128 for SWIZZLE_128B); out-of-bounds elements are filled with zeros.
inline int make_tmap_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch,
                          uint32_t box_inner, uint32_t box_outer) {
  TmapEncodeFn fn = tmap_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return MCPC_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {inner, outer};
  const cuuint64_t gstride[1] = {pitch * 2};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for inner=%llu outer=%llu pitch=%llu box=%ux%u", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch, box_inner, box_outer);
    return MCPC_ERR_CUDA;
  }
  return MCPC_OK;
}

// Row-major fp32 matrix, no swizzle: only used for L2 prefetches (cp.async.bulk.prefetch.tensor), never as an operand.
inline int make_tmap_f32(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch,
                         uint32_t box_inner, uint32_t box_outer) {
  TmapEncodeFn fn = tmap_encode_fn();
  if (fn == nullptr) return MCPC_ERR_CUDA;
  const cuuint64_t gdim[2] = {inner, outer};
  const cuuint64_t gstride[1] = {pitch * 4};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MCPC_OK : MCPC_ERR_CUDA;
}

// The same row-major matrix [k rows][n columns] viewed as 3-D (64 columns, k, column block): ONE box of
// 64 x box_k x box_blocks lands as `box_blocks` consecutive MN-major operand blocks of 8 KB (n must be a
// multiple of 64 so that no block straddles the end of the matrix).
inline int make_tmap_bf16_mn3(CUtensorMap* tm, const void* base, uint64_t n_cols, uint64_t k_rows, uint64_t pitch,
                              uint32_t box_k, uint32_t box_blocks) {
  TmapEncodeFn fn = tmap_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return MCPC_ERR_CUDA;
  }
  const cuuint64_t gdim[3] = {64, k_rows, n_cols / 64};
  const cuuint64_t gstride[2] = {pitch * 2, 128};
  const cuuint32_t box[3] = {64, box_k, box_blocks};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3-D) failed (%d) for cols=%llu rows=%llu pitch=%llu", (int)r,
              (unsigned long long)n_cols, (unsigned long long)k_rows, (unsigned long long)pitch);
    return MCPC_ERR_CUDA;
  }
  return MCPC_OK;
}

namespace umma {

// shared-memory matrix descriptor with layout type SWIZZLE_128B (= 2 in bits [61,64))
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return smem_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)2 << 61);
}

// one TMA box: element coordinates (c_inner, c_outer) of its first element; completion counted on `bar`
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tm, int c_inner, int c_outer, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :: "r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst_smem, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :: "r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// the same boxes, fetched into L2 only (no shared-memory destination, no mbarrier): hides the DRAM latency of
// first-touch operands behind stages that are still far away
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* tm, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               :: "l"(reinterpret_cast<uint64_t>(tm)), "r"(c_inner), "r"(c_outer) : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* tm, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               :: "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

}  // namespace umma
}  // namespace mcpc
