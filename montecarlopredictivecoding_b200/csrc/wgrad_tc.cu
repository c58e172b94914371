// mcpc_weight_grad, MCPC_PREC_BF16: the local weight update on the tensor cores.
//
//     gW_l [d_l x d_{l-1}] += G_l^T F_{l-1},   gb_l += colsum(G_l)        (autograd dW of pc_trainer.py:862)
//
// G (d overall / d mu) and F (act(x)) were saved by infer_tc_kernel as bf16 row-major [n_save*B, width] with
// every layer's block padded to 8 columns.  The reduction index (rows = saved steps x chains, ~1e5 for the
// mcpc_ml call) is the MMA K dimension and BOTH operands are "MN-major": rows of 8 units (16 B) are the
// core-matrix rows, so a row-major slab is scattered into the canonical no-swizzle layout with plain 16-byte
// cp.async (LDGSTS) -- no register staging, no tensor map.
//
// Grid = (output tile, K slab).  Output tile = 128 output units x <=128 input units (+ a constant block of
// ones appended to the B operand, which makes column N of the accumulator the bias gradient for free).
// Warps 0-3: cp.async producers (4-stage ring, 64 rows per stage), later the epilogue; warp 4: MMA issuer.
// Partial tiles are added to global memory with coalesced fp32 reductions (via a shared-memory transpose).
#include "mcpc_common.cuh"
#include "umma.cuh"

namespace mcpc {
namespace {

using namespace umma;

constexpr int kStages = 4;
constexpr int kStageRows = 64;
constexpr int kMaxWTiles = 64;

struct WTile {
  int lin;        // Linear index (0..L)
  int m0;         // first output unit
  int n0;         // first input unit
  int n_real;     // input units in this tile (multiple of 16, 0 for a bias-only tile)
  int with_bias;  // column n_real of the accumulator is the bias gradient
  int g_col;      // first column of the A slab in save_g
  int f_col;      // first column of the B slab in save_f
};

struct WgradParams {
  WTile tiles[kMaxWTiles];
  const __nv_bfloat16* G;
  const __nv_bfloat16* F;
  int g_pitch, f_pitch;           // row pitches in elements
  float* gW[kMaxL + 1];
  float* gb[kMaxL + 1];
  int d_out_units[kMaxL + 1];     // rows of gW_l
  int d_in_units[kMaxL + 1];      // cols of gW_l
  int rows, rows_per_slab;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ bool elect_one_w() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(160, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[kStages], empty[kStages], done;
  __shared__ uint32_t tmem_base_s;

  const WTile& T = p.tiles[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tot = T.n_real + 16;                       // + the ones block
  const uint32_t lbo_a = 16 * 128, lbo_b = (uint32_t)(n_tot / 8) * 128u;     // k-group strides; mn-group stride is 128
  const uint32_t a_bytes = (kStageRows / 8) * lbo_a, b_bytes = (kStageRows / 8) * lbo_b;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int r_begin = blockIdx.y * p.rows_per_slab;
  const int r_end = min(p.rows, r_begin + p.rows_per_slab);
  const int n_stage = (r_end > r_begin) ? (r_end - r_begin + kStageRows - 1) / kStageRows : 0;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 128);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(&tmem_base_s, 256);
  // the ones block of every stage's B operand: element (k, unit n_real) = 1, the other 15 columns 0
  for (int i = tid; i < kStages * kStageRows * 2; i += blockDim.x) {
    const int s = i / (kStageRows * 2), r = (i / 2) % kStageRows, g = i & 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (g == 0) v.x = 0x00003F80u;                         // bf16 1.0 in the first of the 8 units
    *reinterpret_cast<uint4*>(smem + s * stage_bytes + a_bytes + (r >> 3) * lbo_b + (T.n_real / 8 + g) * 128 + (r & 7) * 16) = v;
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (n_stage == 0) {
    if (warp == 4) tmem_dealloc(tmem, 256);
    return;
  }

  if (warp < 4) {
    // ---------------- producers: 64 rows x (128 + n_real) units per stage ----------------
    const int r_in = tid >> 1, half = tid & 1;
    const uint32_t smem_base = smem_u32(smem);
    const int nb = T.n_real / 8;
    constexpr int D = kStages - 1;                          // stages in flight
    for (int s = 0; s < n_stage + D; ++s) {
      if (s < n_stage) {
        const int slot = s % kStages;
        mbar_wait(&empty[slot], ((s / kStages) & 1) ^ 1);
        const int row = r_begin + s * kStageRows + r_in;
        const uint32_t ok = (row < r_end) ? 16u : 0u;       // rows past the slab are zero-filled
        const size_t rr = (size_t)min(row, p.rows - 1);
        const uint32_t dst_row = smem_base + slot * stage_bytes + (r_in >> 3) * lbo_a + (r_in & 7) * 16;
        const __nv_bfloat16* ga = p.G + rr * p.g_pitch + T.g_col + half * 64;
#pragma unroll
        for (int c = 0; c < 8; ++c) cp_async16(dst_row + (half * 8 + c) * 128, ga + c * 8, ok);
        const uint32_t dst_rowb = smem_base + slot * stage_bytes + a_bytes + (r_in >> 3) * lbo_b + (r_in & 7) * 16;
        const __nv_bfloat16* fb = p.F + rr * p.f_pitch + T.f_col;
        for (int c = half; c < nb; c += 2) cp_async16(dst_rowb + c * 128, fb + c * 8, ok);
      }
      cp_async_commit();
      if (s >= D) {
        cp_async_wait<D>();                                 // the group of stage s-D has landed
        fence_async_smem();
        mbar_arrive(&full[(s - D) % kStages]);
      }
    }
  } else {
    // ---------------- MMA issuer ----------------
    const uint32_t id = idesc_bf16(128, n_tot, true, true);
    const uint32_t smem_base = smem_u32(smem);
    for (int s = 0; s < n_stage; ++s) {
      const int slot = s % kStages;
      mbar_wait(&full[slot], (s / kStages) & 1);
      fence_after_sync();
      const uint64_t ad0 = smem_desc(smem_base + slot * stage_bytes, lbo_a, 128u);
      const uint64_t bd0 = smem_desc(smem_base + slot * stage_bytes + a_bytes, lbo_b, 128u);
      if (elect_one_w()) {
#pragma unroll
        for (int ks = 0; ks < kStageRows / 16; ++ks)
          mma_bf16_ss(tmem, ad0 + (uint64_t)(ks * ((2 * lbo_a) >> 4)), bd0 + (uint64_t)(ks * ((2 * lbo_b) >> 4)), id, s > 0 || ks > 0);
        mma_commit(&empty[slot]);
        if (s == n_stage - 1) mma_commit(&done);
      }
      __syncwarp();
    }
  }

  // ---------------- epilogue: TMEM -> smem transpose -> coalesced global reductions ----------------
  if (warp < 4) {
    mbar_wait(&done, 0);
    fence_after_sync();
    float* tr = reinterpret_cast<float*>(smem);               // [128][n_tot + 1] fp32, the ring is idle now
    const int pitch = n_tot + 1;
    const int m = warp * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < n_tot; c += 16) {
      float v[16];
      tmem_ld16(lane_addr + c, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) tr[m * pitch + c + i] = v[i];
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int d_out_l = p.d_out_units[T.lin], d_in_l = p.d_in_units[T.lin];
    float* gW = p.gW[T.lin];
    float* gb = p.gb[T.lin];
    for (int r = warp; r < 128; r += 4) {
      const int mo = T.m0 + r;
      if (mo >= d_out_l) break;
      if (gW != nullptr)
        for (int n = lane; n < T.n_real; n += 32)
          if (T.n0 + n < d_in_l) atomicAdd(gW + (size_t)mo * d_in_l + T.n0 + n, tr[r * pitch + n]);
      if (T.with_bias && gb != nullptr && lane == 0) atomicAdd(gb + mo, tr[r * pitch + T.n_real]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

inline int pad8i(int v) { return (v + 7) & ~7; }
inline int pad16i(int v) { return (v + 15) & ~15; }

}  // namespace

// Column layout of the bf16 saved operands: every layer's block starts at a multiple of 8 columns (16 B).
void save_layout_bf16(const NetDev& nd, int* g_off, int* g_width, int* f_off, int* f_width) {
  int o = 0;
  for (int l = 0; l < nd.L; ++l) {
    g_off[l] = o;
    f_off[l] = o;
    o += pad8i(nd.dims[l]);
  }
  *f_width = o;
  g_off[nd.L] = o;
  *g_width = o + pad8i(nd.d_out);
}

int launch_weight_grad_tc(const NetDev& nd, const McpcGradIO* io, int B, int n_save, cudaStream_t stream) {
  if (io->inputs != nullptr) {
    set_error("bf16 weight-grad: non-zero inputs are not implemented; use MCPC_PREC_FP32");
    return MCPC_ERR_UNSUPPORTED;
  }
  WgradParams p{};
  int g_off[kMaxL + 1], f_off[kMaxL + 1], gw = 0, fw = 0;
  save_layout_bf16(nd, g_off, &gw, f_off, &fw);
  p.G = reinterpret_cast<const __nv_bfloat16*>(io->save_g);
  p.F = reinterpret_cast<const __nv_bfloat16*>(io->save_f);
  p.g_pitch = gw;
  p.f_pitch = fw;
  p.rows = n_save * B;
  int nt = 0;
  const int n_lin = nd.L + (nd.d_out > 0 ? 1 : 0);
  for (int l = 0; l < n_lin; ++l) {
    const bool is_out = (l == nd.L);
    if (is_out && !nd.top_has_grad) continue;
    const int d_o = is_out ? nd.d_out : nd.dims[l];
    const int d_i = (l == 0) ? 0 : nd.dims[l - 1];             // zero inputs: Linear_0 has no weight gradient
    p.gW[l] = (l == 0) ? nullptr : io->gW[l];
    p.gb[l] = io->gb[l];
    p.d_out_units[l] = d_o;
    p.d_in_units[l] = d_i;
    if (p.gW[l] == nullptr && p.gb[l] == nullptr) continue;
    const int n_ntiles = (d_i == 0 || p.gW[l] == nullptr) ? 1 : (d_i + 127) / 128;
    for (int m0 = 0; m0 < d_o; m0 += 128)
      for (int j = 0; j < n_ntiles; ++j) {
        if (nt >= kMaxWTiles) {
          set_error("bf16 weight-grad: more than %d output tiles", kMaxWTiles);
          return MCPC_ERR_UNSUPPORTED;
        }
        WTile& T = p.tiles[nt++];
        T.lin = l;
        T.m0 = m0;
        T.n0 = j * 128;
        T.n_real = (d_i == 0 || p.gW[l] == nullptr) ? 0 : pad16i(d_i - j * 128 < 128 ? d_i - j * 128 : 128);
        T.with_bias = (j == 0) ? 1 : 0;
        T.g_col = g_off[l] + m0;
        T.f_col = (l == 0) ? 0 : f_off[l - 1] + j * 128;
      }
  }
  if (nt == 0) return MCPC_OK;
  int slabs = (148 + nt - 1) / nt;
  int rps = (p.rows + slabs - 1) / slabs;
  rps = ((rps + kStageRows - 1) / kStageRows) * kStageRows;
  slabs = (p.rows + rps - 1) / rps;
  p.rows_per_slab = rps;
  // widest stage: A 16 KB + B (128+16)/8 * 128 * 8 = 18 KB
  const size_t smem = (size_t)kStages * ((kStageRows / 8) * 16 * 128 + (kStageRows / 8) * (144 / 8) * 128) + 1024;
  const size_t tr_bytes = (size_t)128 * 145 * 4 + 1024;
  const size_t dyn = smem > tr_bytes ? smem : tr_bytes;
  MCPC_CUDA_CHECK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  wgrad_tc_kernel<<<dim3(nt, slabs), 160, dyn, stream>>>(p);
  MCPC_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return MCPC_OK;
}

}  // namespace mcpc
