// mcpc_weight_grad, MCPC_PREC_BF16: the local weight update on the tensor cores.
//
//     gW_l [d_l x d_{l-1}] += G_l^T F_{l-1},   gb_l += colsum(G_l)        (autograd dW of pc_trainer.py:862)
//
// G (d overall / d mu) and F (act(x)) were saved by infer_tc_kernel as bf16 row-major [n_save*B, width] with
// every layer's block padded to 8 columns.  The reduction index (rows = saved steps x chains, ~1e5 for the
// mcpc_ml call) is the MMA K dimension, so BOTH operands are "MN-major": a 3-D tensor map (64 columns, rows,
// column block) over the row-major slab lets ONE TMA box land as consecutive SWIZZLE_128B MN-major operand
// blocks of 64 units x 64 rows (tma.cuh).  (The first version scattered the slab with 16-byte cp.async: ~2,200
// LDGSTS per stage, producer-bound at 0.27 ms per mcpc_ml call.)
//
// Grid = (output tile, K slab).  Output tile = 128 output units x <=128 input units; a constant block whose first
// column is ones follows the B operand's two blocks, which makes column 128 of the accumulator the bias gradient
// for free.  Warp 0: TMA producer (one elected lane, 4-stage mbarrier ring, 64 rows per stage); warp 4: MMA issuer;
// warps 0-3: epilogue.  Partial tiles are added to global memory with fp32 atomics from a shared-memory transpose.
#include "mcpc_common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace mcpc {
namespace {

using namespace umma;

constexpr int kStages = 4;
constexpr int kStageRows = 64;
constexpr int kMaxWTiles = 64;
constexpr uint32_t kBlk = 8192;       // one MN-major operand block: 64 rows (K) x 64 units, SWIZZLE_128B

struct WTile {
  int lin;        // Linear index (0..L)
  int m0;         // first output unit
  int n0;         // first input unit
  int n_real;     // input units in this tile (multiple of 16, 0 for a bias-only tile)
  int with_bias;  // the column after the F blocks of the accumulator is the bias gradient
  int nblk;       // 64-unit blocks of F the tile loads (0, 1 or 2)
};

struct WgradParams {
  WTile tiles[kMaxWTiles];
  CUtensorMap mapG[kMaxL + 1];    // G block of Linear l:   (64 cols, rows, column blocks) from its first column
  CUtensorMap mapF[kMaxL + 1];    // F block of layer l-1 (input of Linear l), same view
  float* gW[kMaxL + 1];
  float* gb[kMaxL + 1];
  int d_out_units[kMaxL + 1];     // rows of gW_l
  int d_in_units[kMaxL + 1];      // cols of gW_l
  int rows, rows_per_slab;
  // Overlapped mode (launched next to infer_tc_kernel on the SMs it leaves idle): CTA y takes the 64-row stages y, y + slabs,
  // ... so that every CTA follows the producer, and waits until ready[step] == ready_target before it loads rows of a step
  const unsigned* ready;
  unsigned ready_target;
  int rows_per_step;              // B
};

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
  // the F box always carries two 64-unit blocks (a second block the tile does not need is junk or zero fill and is
  // never written back), so the ones block sits at a fixed place and the accumulator is 144 columns wide
  const int f_blocks = (T.nblk > 0) ? 2 : 0;
  const int n_tot = f_blocks * 64 + 16;                  // F blocks + the first 16 columns of the ones block
  const uint32_t a_bytes = 2 * kBlk, b_bytes = (uint32_t)f_blocks * kBlk;
  const uint32_t stage_bytes = a_bytes + 2 * kBlk + kBlk;          // fixed stride: A | up to 2 F blocks | ones block
  int r_begin, r_stride, n_stage;
  if (p.ready != nullptr) {
    const int total = (p.rows + kStageRows - 1) / kStageRows;
    r_begin = (int)blockIdx.y * kStageRows;
    r_stride = (int)gridDim.y * kStageRows;
    n_stage = ((int)blockIdx.y < total) ? (total - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y : 0;
  } else {
    r_begin = blockIdx.y * p.rows_per_slab;
    r_stride = kStageRows;
    const int r_end = min(p.rows, r_begin + p.rows_per_slab);
    n_stage = (r_end > r_begin) ? (r_end - r_begin + kStageRows - 1) / kStageRows : 0;
  }

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&done, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.mapG[T.lin]);
    if (T.nblk > 0) tma_prefetch_desc(&p.mapF[T.lin]);
  }
  if (warp == 4) tmem_alloc(&tmem_base_s, 256);
  // the ones block right after the F blocks of every stage: element (k, column 0) = 1, everything else 0.  In the
  // SWIZZLE_128B layout row k keeps its 16-byte chunk c at position c ^ (k % 8).
  for (int i = tid; i < kStages * kStageRows * 8; i += blockDim.x) {
    const int s = i / (kStageRows * 8), k = (i / 8) % kStageRows, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c == (k & 7)) v.x = 0x00003F80u;                   // chunk 0 of row k: bf16 1.0 in its first element
    *reinterpret_cast<uint4*>(smem + s * stage_bytes + a_bytes + b_bytes + k * 128 + c * 16) = v;
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

  if (warp == 0) {
    // ---------------- TMA producer: 64 rows x (128 + 64*nblk) units per stage, two instructions ----------------
    if (elect_one_w()) {
      int steps_ready = 0;                                 // saved steps known to be complete (overlapped mode)
      for (int s = 0; s < n_stage; ++s) {
        const int slot = s % kStages;
        mbar_wait(&empty[slot], ((s / kStages) & 1) ^ 1);
        uint8_t* st = smem + slot * stage_bytes;
        const int row = r_begin + s * r_stride;            // rows past the end of the buffer are zero-filled by TMA
        if (p.ready != nullptr) {
          const int step_hi = min(row + kStageRows - 1, p.rows - 1) / p.rows_per_step;
          if (step_hi >= steps_ready) {
            // every producer CTA finishes its steps in order, so the newest step's flag covers the older ones
            unsigned v;
            unsigned long long t0 = 0, t1;
            for (;;) {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.ready + step_hi) : "memory");
              if (v >= p.ready_target) break;
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
              if (t0 == 0) t0 = t1;
              if (t1 - t0 > 4000000000ull) __trap();      // 4 s without progress: the producer is gone -- fail, never hang
              __nanosleep(200);
            }
            steps_ready = step_hi + 1;
            asm volatile("fence.proxy.async;" ::: "memory");   // the TMA (async proxy) reads below come after the acquire
          }
        }
        mbar_expect_tx(&full[slot], a_bytes + b_bytes);
        tma_load_3d(st, &p.mapG[T.lin], 0, row, T.m0 / 64, &full[slot]);
        if (f_blocks > 0) tma_load_3d(st + a_bytes, &p.mapF[T.lin], 0, row, T.n0 / 64, &full[slot]);
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // ---------------- MMA issuer ----------------
    const uint32_t id = idesc_bf16(128, n_tot, true, true);
    const uint32_t smem_base = smem_u32(smem);
    for (int s = 0; s < n_stage; ++s) {
      const int slot = s % kStages;
      mbar_wait(&full[slot], (s / kStages) & 1);
      fence_after_sync();
      const uint64_t ad0 = smem_desc_sw128(smem_base + slot * stage_bytes, kBlk, 1024u);
      const uint64_t bd0 = smem_desc_sw128(smem_base + slot * stage_bytes + a_bytes, kBlk, 1024u);
      if (elect_one_w()) {
#pragma unroll
        for (int ks = 0; ks < kStageRows / 16; ++ks)      // 16 rows (K) = 2048 bytes further into every block
          mma_bf16_ss(tmem, ad0 + (uint64_t)(ks * 128), bd0 + (uint64_t)(ks * 128), id, s > 0 || ks > 0);
        mma_commit(&empty[slot]);
        if (s == n_stage - 1) mma_commit(&done);
      }
      __syncwarp();
    }
  }

  // ---------------- epilogue: TMEM -> smem transpose -> global reductions ----------------
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
    const int bias_col = f_blocks * 64;
    for (int r = warp; r < 128; r += 4) {
      const int mo = T.m0 + r;
      if (mo >= d_out_l) break;
      if (gW != nullptr)
        for (int n = lane; n < T.n_real; n += 32)
          if (T.n0 + n < d_in_l) atomicAdd(gW + (size_t)mo * d_in_l + T.n0 + n, tr[r * pitch + n]);
      if (T.with_bias && gb != nullptr && lane == 0) atomicAdd(gb + mo, tr[r * pitch + bias_col]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

// GS[c][u] = sum over the saved steps of the bf16 G_0 operand (d overall / d mu_0) of chain c: with non-zero inputs
// gW_0 = sum_steps G_0^T inputs = GS^T inputs, and the inputs do not change during the call.
__global__ void sum_slots_kernel(const __nv_bfloat16* __restrict__ G, int n_save, int B, int gw, int d0,
                                 float* __restrict__ GS) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * d0) return;
  const int c = (int)(idx / d0), u = (int)(idx % d0);
  float s = 0.0f;
  for (int slot = 0; slot < n_save; ++slot) s += __bfloat162float(G[((size_t)slot * B + c) * gw + u]);
  GS[idx] = s;
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
  return launch_weight_grad_tc_overlapped(nd, io, B, n_save, nullptr, 0, 0, stream);
}

// Number of output tiles (= CTAs per K slab) of the tensor-core weight-gradient kernel for this network
int weight_grad_tc_tiles(const NetDev& nd, const McpcGradIO* io) {
  int nt = 0;
  const int n_lin = nd.L + (nd.d_out > 0 ? 1 : 0);
  for (int l = 0; l < n_lin; ++l) {
    if (l == nd.L && !nd.top_has_grad) continue;
    const int d_o = (l == nd.L) ? nd.d_out : nd.dims[l];
    const int d_i = (l == 0) ? 0 : nd.dims[l - 1];
    const bool has_w = (l > 0) && io->gW[l] != nullptr;
    if (!has_w && io->gb[l] == nullptr) continue;
    nt += ((d_o + 127) / 128) * ((d_i == 0 || !has_w) ? 1 : (d_i + 127) / 128);
  }
  return nt;
}

// ready != nullptr: overlapped mode -- the kernel is launched while infer_tc_kernel is still writing save_g / save_f and
// follows it step by step (ready[step] reaches ready_target when a step's rows are complete); at most max_ctas CTAs so
// that it fits the SMs the inference kernel leaves idle.
int launch_weight_grad_tc_overlapped(const NetDev& nd, const McpcGradIO* io, int B, int n_save, const unsigned* ready,
                                     unsigned ready_target, int max_ctas, cudaStream_t stream) {
  if (ready != nullptr && io->inputs != nullptr) {
    set_error("internal: the overlapped weight-gradient launch does not take non-zero inputs");
    return MCPC_ERR_INVALID;
  }
  if (io->inputs != nullptr && io->gW[0] != nullptr) {
    // Linear_0 with non-zero inputs: the inputs are the same rows for every saved step, so its weight gradient is
    // (sum over steps of G_0)^T inputs -- a [B]-row contraction with exact fp32 inputs (the bias gradient comes from
    // the tensor-core kernel below like for zero inputs)
    const size_t need = (size_t)B * nd.dims[0] * sizeof(float);
    if (io->scratch == nullptr || io->scratch_bytes < need) {
      set_error("bf16 weight-grad with non-zero inputs needs McpcGradIO.scratch of %zu bytes (%zu given)", need,
                io->scratch == nullptr ? (size_t)0 : io->scratch_bytes);
      return MCPC_ERR_WORKSPACE;
    }
    int g_off0[kMaxL + 1], f_off0[kMaxL + 1], gw0 = 0, fw0 = 0;
    save_layout_bf16(nd, g_off0, &gw0, f_off0, &fw0);
    float* GS = reinterpret_cast<float*>(io->scratch);
    const size_t n = (size_t)B * nd.dims[0];
    sum_slots_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(io->save_g), n_save, B, gw0, nd.dims[0], GS);
    MCPC_CUDA_CHECK(cudaGetLastError());
    count_launch();
    const int rc0 = launch_wgrad_tn_fp32(GS, nd.dims[0], io->inputs, nd.d_in, io->gW[0], nd.dims[0], nd.d_in, B, stream);
    if (rc0 != MCPC_OK) return rc0;
  }
  WgradParams p{};
  int g_off[kMaxL + 1], f_off[kMaxL + 1], gw = 0, fw = 0;
  save_layout_bf16(nd, g_off, &gw, f_off, &fw);
  const __nv_bfloat16* G = reinterpret_cast<const __nv_bfloat16*>(io->save_g);
  const __nv_bfloat16* F = reinterpret_cast<const __nv_bfloat16*>(io->save_f);
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
        T.nblk = (T.n_real + 63) / 64;
      }
    // tensor maps from the first column of the layer's block; the last 64-column block may run past the block
    // (neighbouring columns, or the start of the next row / the buffer's slack after the last row): those
    // accumulator rows / columns are never written back
    int rc = make_tmap_bf16_mn3(&p.mapG[l], G + g_off[l], (uint64_t)((gw - g_off[l] + 63) / 64) * 64, (uint64_t)p.rows,
                                (uint64_t)gw, kStageRows, 2);
    if (rc != MCPC_OK) return rc;
    if (l > 0 && p.gW[l] != nullptr) {
      rc = make_tmap_bf16_mn3(&p.mapF[l], F + f_off[l - 1], (uint64_t)((fw - f_off[l - 1] + 63) / 64) * 64, (uint64_t)p.rows,
                              (uint64_t)fw, kStageRows, 2);
      if (rc != MCPC_OK) return rc;
    }
  }
  if (nt == 0) return MCPC_OK;
  // K slabs: as many as keep the whole grid in ONE wave (nt * slabs <= SMs; one CTA per SM by shared memory) --
  // rounding up instead put 153 CTAs on 148 SMs and doubled the kernel time
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  int slabs = (ready != nullptr ? max_ctas : n_sm) / nt;
  if (slabs < 1) slabs = 1;
  p.ready = ready;
  p.ready_target = ready_target;
  p.rows_per_step = B;
  int rps = (p.rows + slabs - 1) / slabs;
  rps = ((rps + kStageRows - 1) / kStageRows) * kStageRows;
  slabs = (p.rows + rps - 1) / rps;
  p.rows_per_slab = rps;
  // stage: A 16 KB + up to 2 F blocks + the ones block = 40 KB; the epilogue's transpose reuses the ring
  const size_t smem = (size_t)kStages * 5 * kBlk + 1024;
  const size_t tr_bytes = (size_t)128 * 145 * 4 + 1024;
  const size_t dyn = smem > tr_bytes ? smem : tr_bytes;
  MCPC_CUDA_CHECK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  wgrad_tc_kernel<<<dim3(nt, slabs), 160, dyn, stream>>>(p);
  MCPC_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return MCPC_OK;
}

}  // namespace mcpc
