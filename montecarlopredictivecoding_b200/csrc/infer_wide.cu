// mcpc_infer, MCPC_PREC_BF16, networks too wide to stay on chip (SURVEY config C5: 4 x 4096): the
// "streaming" path.  Latents x (fp32), their activations (bf16) and the error signals live in HBM / L2;
// every Langevin step is three grouped tcgen05 GEMM kernels with fused epilogues:
//
//   wide_predict_kernel   for every Linear l:  mu = act(x_{l-1}) W_l^T + b  ->  eps = x_l - mu, energy, loss,
//                         G_l = d overall / d mu_l (bf16 operand copy + fp32 own-layer term), e_out
//   wide_wgrad_kernel     (steps of the accumulate window)  gW_l += G_l^T act(x_{l-1}),  gb_l += colsum G_l
//   wide_update_kernel    for every PCLayer l: bp = G_{l+1} W_{l+1};  grad = -G_l + act'(x_l) * bp;
//                         x <- SGD | Adam step; x <- x - lr * noise (Philox);  act(x) re-emitted as bf16
//
// All three share one persistent mainloop: 128 x 256 output tile, K in stages of 64, operands brought in by TMA
// (cp.async.bulk.tensor with SWIZZLE_128B tensor maps over the row-major global matrices; K-major or MN-major
// as the operand's storage order dictates -- no transposed copies of anything), 4-stage mbarrier ring, two
// 256-column accumulators in TMEM so the epilogue of tile i overlaps the mainloop of tile i+1.  Warp roles:
// warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue (TMEM lane quarter = warp % 4, column half = (warp-2)/4;
// sub-blocks are transposed through shared memory so that global accesses are contiguous, see "coalesced epilogues").
// (A first version scattered operands with 16-byte cp.async: it was bound by the L1TEX wavefront rate -- 8 cache
// lines per instruction, ~2000 cycles to issue one stage -- see DESIGN.md.)
// Reference semantics: predictive_coding/pc_trainer.py:733-918 + utils/model.py:35-44.
#include <cstdlib>

#include "mcpc_common.cuh"
#include "philox.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace mcpc {
namespace {

using namespace umma;

constexpr int kWS = 4;                      // pipeline stages
constexpr int kBK = 64;                     // K per stage
constexpr int kBN = 256;                    // output-tile width (N of the MMA)
constexpr uint32_t kABytes = 128 * kBK * 2;             // A operand stage: 128 (M) x 64 (K) bf16
constexpr uint32_t kBBytes = kBN * kBK * 2;             // B operand stage: 256 (N) x 64 (K) bf16

struct WideParams {
  NetDev net;
  float* x[kMaxL];
  float* m[kMaxL];
  float* v[kMaxL];
  float* xgrad[kMaxL];
  float* traj_x[kMaxL];
  float* traj_out;
  const float* b[kMaxL + 1];
  const __nv_bfloat16* Wb[kMaxL + 1];     // bf16 copies of the weights, row-major [d_l][d_{l-1}]
  __nv_bfloat16* act;                     // [B][a_pitch]: act(x_l) at column poff[l]
  __nv_bfloat16* Gb;                      // [B][g_pitch]: G_l at column poff[l], e_out at poff[L]
  float* G32;                             // [B][SD]: fp32 G_l (own-layer gradient term) at column off[l]
  const float* target;
  const float* noise;
  float* gW[kMaxL + 1];
  float* gb[kMaxL + 1];
  float* partials;                        // [n_steps][n_part][2]
  int poff[kMaxL + 1];
  int a_pitch, g_pitch;
  int B, mt;                              // chains, chain tiles of 128
  int tP_first[kMaxL + 2];                // predict tiles: prefix over Linear 1..L (index lin)
  int tU_first[kMaxL + 1];                // update tiles: prefix over layers 0..L-1
  int tW_first[kMaxL + 2];                // wgrad tiles: prefix over Linear 0..L
  int n_part;                             // partial slots per step (predict tiles + update tiles of layer 0)
  int optimizer, update_x;
  float lr, adam_eps, one_minus_b1, one_minus_b2, beta2f;
  int noise_mode;
  float noise_scale;
  uint64_t seed, chain_offset;
  int mn3;                                // MN-major operands come in through 3-D tensor maps (all widths % 64 == 0)
  int l2_prefetch;                        // stages ahead of the ring whose boxes are prefetched into L2 (0 = off)
  int skip_epilogue;                      // debug: MCPC_WIDE_SKIP_EPI=1
  long long* dbg;                         // MCPC_WIDE_TIMING=1: per-role cycle counters of CTA 0 (debug)
};

struct StepArgs {
  int ts, t_abs, rec, do_traj, last;
  int acc;                                // this step is inside the weight-gradient window: bias gradients accumulate
  float step_size, inv_bc2_sqrt;          // Adam bias corrections of this step
};

__device__ __forceinline__ bool elect1() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float tanh_fast_w(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float act_w(int kind, float x) {
  return kind == MCPC_ACT_RELU ? fmaxf(x, 0.0f) : (kind == MCPC_ACT_TANH ? tanh_fast_w(x) : x);
}
__device__ __forceinline__ float dact_w(int kind, float x, float a) {
  return kind == MCPC_ACT_RELU ? (x > 0.0f ? 1.0f : 0.0f) : (kind == MCPC_ACT_TANH ? fmaf(-a, a, 1.0f) : 1.0f);
}
__device__ __forceinline__ float warp_sum_w(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Tensor maps over the row-major bf16 matrices, one per layer block so that out-of-range K / M / N are zero-filled:
//   *_k : box 64 (inner = contraction index) x 128|256 rows  -> K-major operand stage
//   *_mn: box 64 (inner = M/N index) x 64 rows (contraction) -> MN-major operand stage, one box per 64 units
struct WideMaps {
  CUtensorMap act_k[kMaxL], act_mn[kMaxL];
  CUtensorMap gb_k[kMaxL + 1], gb_mn[kMaxL + 1];
  CUtensorMap w_k[kMaxL + 1], w_mn[kMaxL + 1];
};

struct Pipe {
  uint64_t full[kWS], empty[kWS], acc_full[2], acc_empty[2];
};

enum { KIND_PREDICT = 0, KIND_UPDATE = 1, KIND_WGRAD = 2 };

// Epilogue warps per CTA (8 = two per sub-partition).  Measured on C5: 16 warps for the update kernel made it slower
// (500k vs 437k cycles per CTA: its epilogue is bound by the 32-lines-per-instruction global access pattern and by
// Philox latency chains, and a lane = unit version of it executed 2x the instructions); the predict and wgrad
// epilogues are memory-shaped and, transposed through shared memory, hide behind the mainloop.
__host__ __device__ constexpr int epi_warps(int kind) { return kind >= 0 ? 8 : 8; }

struct TileDesc {
  int idx;          // Linear index (predict / wgrad) or layer index (update)
  int m0, n0;
  int k_ext;        // 0: no contraction for this tile
  const CUtensorMap* mapA;
  const CUtensorMap* mapB;
};

template <int KIND>
__device__ __forceinline__ int n_tiles_of(const WideParams& p) {
  return KIND == KIND_PREDICT ? p.tP_first[p.net.L + 1] : (KIND == KIND_UPDATE ? p.tU_first[p.net.L] : p.tW_first[p.net.L + 1]);
}

template <int KIND>
__device__ __forceinline__ TileDesc decode_tile(const WideParams& p, const WideMaps& mp, int tile) {
  const NetDev& nd = p.net;
  TileDesc t{};
  if (KIND == KIND_PREDICT) {
    int lin = 0;
    while (tile >= p.tP_first[lin + 1]) ++lin;
    const int d_o = (lin == nd.L) ? nd.d_out : nd.dims[lin];
    const int d_i = (lin == 0) ? 0 : nd.dims[lin - 1];            // Linear_0 sees zero inputs: mu_0 = b_0
    const int ntn = (d_o + kBN - 1) / kBN, local = tile - p.tP_first[lin];
    t.idx = lin; t.m0 = (local / ntn) * 128; t.n0 = (local % ntn) * kBN; t.k_ext = d_i;
    if (d_i > 0) {
      t.mapA = &mp.act_k[lin - 1];                                // act(x_{l-1}) [B x d_i], K-major
      t.mapB = &mp.w_k[lin];                                      // W_l [d_o x d_i], K-major
    }
  } else if (KIND == KIND_UPDATE) {
    int l = 0;
    while (tile >= p.tU_first[l + 1]) ++l;
    const int dl = nd.dims[l];
    const int ntn = (dl + kBN - 1) / kBN, local = tile - p.tU_first[l];
    const bool has_above = (l + 1 < nd.L) || nd.top_has_grad;
    const int d_up = (l + 1 < nd.L) ? nd.dims[l + 1] : nd.d_out;
    t.idx = l; t.m0 = (local / ntn) * 128; t.n0 = (local % ntn) * kBN; t.k_ext = has_above ? d_up : 0;
    if (has_above) {
      t.mapA = &mp.gb_k[l + 1];                                   // G_{l+1} [B x d_up], K-major
      t.mapB = &mp.w_mn[l + 1];                                   // W_{l+1} [d_up x d_l] as B[k][n]: MN-major
    }
  } else {
    int lin = 1;
    while (tile >= p.tW_first[lin + 1]) ++lin;
    const int d_i = nd.dims[lin - 1];
    const int ntn = (d_i + kBN - 1) / kBN, local = tile - p.tW_first[lin];
    t.idx = lin; t.m0 = (local / ntn) * 128; t.n0 = (local % ntn) * kBN; t.k_ext = p.B;
    t.mapA = &mp.gb_mn[lin];                                      // G_l as A[k=chain][m]: MN-major
    t.mapB = &mp.act_mn[lin - 1];                                 // act(x_{l-1}) as B[k=chain][n]: MN-major
  }
  return t;
}

// ---- coalesced epilogues -------------------------------------------------------------------------------------
// tcgen05.ld hands every lane one ROW of the accumulator; reading / writing global memory in that shape touches 32
// cache lines per instruction (the L1 wavefront rate, not HBM, bounded the first epilogues: the three kernels ran at
// 0.98 ms per C5 step against 0.57 ms with the epilogues switched off).  Each epilogue warp therefore transposes its
// 32 x 32 sub-blocks through a private padded shared-memory tile: afterwards lane = COLUMN, registers = rows, and
// every global access of a warp is one contiguous 128-byte row segment.
constexpr int kTP = 33;                                   // padded pitch of the transpose tile (floats)
constexpr uint32_t kTransBytes = 8 * 32 * kTP * 4;        // 8 epilogue warps

__device__ __forceinline__ void acc_block_to_columns(uint32_t acc_addr, bool has_acc, float* tb, int lane, float (&col)[32]) {
  if (!has_acc) {
#pragma unroll
    for (int r = 0; r < 32; ++r) col[r] = 0.0f;
    return;
  }
  float v[16];
  tmem_ld16(acc_addr, v);
#pragma unroll
  for (int i = 0; i < 16; ++i) tb[lane * kTP + i] = v[i];
  tmem_ld16(acc_addr + 16, v);
#pragma unroll
  for (int i = 0; i < 16; ++i) tb[lane * kTP + 16 + i] = v[i];
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 32; ++r) col[r] = tb[r * kTP + lane];
  __syncwarp();                                           // the tile is rewritten by the next sub-block
}

// errors of the units a tile predicts, lane = unit: eps / energy / own-layer G (hidden Linears) or loss / dLoss (output)
__device__ __forceinline__ void epilogue_predict_t(const WideParams& p, const StepArgs& st, const TileDesc& t, uint32_t acc_q,
                                                   int q, int c_begin, float* tb, int lane, float& e_part, float& l_part,
                                                   float (&gsum)[4]) {
  const NetDev& nd = p.net;
  const int lin = t.idx;
  const bool is_out = (lin == nd.L);
  const int d_o = is_out ? nd.d_out : nd.dims[lin];
  const int row0 = t.m0 + q * 32;
  const int n_rows = min(32, p.B - row0);
  const float ce = is_out ? 0.0f : 0.5f * nd.c[lin], gc = is_out ? 0.0f : nd.gc[lin];
  const bool bern = nd.top == MCPC_TOP_BERNOULLI;
#pragma unroll
  for (int sb = 0; sb < 4; ++sb) gsum[sb] = 0.0f;
#pragma unroll
  for (int sb = 0; sb < 4; ++sb) {
    const int n0 = t.n0 + c_begin + sb * 32;
    if (n0 >= d_o) break;                                               // uniform over the warp
    float d[32];
    acc_block_to_columns(acc_q + c_begin + sb * 32, t.k_ext > 0, tb, lane, d);
    const int n = n0 + lane;
    if (n >= d_o || n_rows <= 0) continue;
    const float bias = (p.b[lin] != nullptr) ? __ldg(p.b[lin] + n) : 0.0f;
    __nv_bfloat16* gbp = p.Gb + (size_t)row0 * p.g_pitch + p.poff[lin] + n;
    if (!is_out) {
      const float* xp = p.x[lin] + (size_t)row0 * d_o + n;
      float xv[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) xv[r] = (r < n_rows) ? xp[(size_t)r * d_o] : 0.0f;
      float* g32 = p.G32 + (size_t)row0 * nd.SD + nd.off[lin] + n;
#pragma unroll
      for (int r = 0; r < 32; ++r)
        if (r < n_rows) {
          const float eps = xv[r] - (d[r] + bias);
          e_part = fmaf(ce * eps, eps, e_part);
          const float g = -gc * eps;
          g32[(size_t)r * nd.SD] = g;
          const __nv_bfloat16 gb16 = __float2bfloat16(g);
          gbp[(size_t)r * p.g_pitch] = gb16;
          gsum[sb] += __bfloat162float(gb16);                           // the bias gradient sums the operand the dW GEMM sees
        }
    } else {
      const bool use_y = nd.top >= MCPC_TOP_GAUSS;
      const bool on = use_y && (n >= nd.mask_start);
      const float* yp = p.target + (size_t)row0 * d_o + n;
      float yv[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) yv[r] = (use_y && r < n_rows) ? yp[(size_t)r * d_o] : 0.0f;
      float* to = (st.do_traj && p.traj_out != nullptr) ? p.traj_out + ((size_t)st.rec * p.B + row0) * d_o + n : nullptr;
#pragma unroll
      for (int r = 0; r < 32; ++r)
        if (r < n_rows) {
          const float o = d[r] + bias;
          float lv, e;
          if (bern) {
            const float z = __expf(-fabsf(o));
            lv = fmaxf(o, 0.0f) - o * yv[r] + __logf(1.0f + z);
            e = __fdividef(o >= 0.0f ? 1.0f : z, 1.0f + z) - yv[r];
          } else {
            const float dd = o - yv[r];
            lv = 0.5f * nd.inv_var * dd * dd;
            e = dd * nd.inv_var;
          }
          l_part += on ? lv : 0.0f;
          const __nv_bfloat16 eb16 = __float2bfloat16(on ? e : 0.0f);
          gbp[(size_t)r * p.g_pitch] = eb16;
          gsum[sb] += __bfloat162float(eb16);
          if (to != nullptr) to[(size_t)r * d_o] = o;
        }
    }
  }
}

// Second shape, for the ALU-heavy update epilogue: lane = (row quad rq2 = lane / 8, column quad cq = lane % 8) and the
// lane owns rows 4*(rq2 + 4i) + j (i < 2, j < 4) x columns 4*cq .. 4*cq+3 of the 32 x 32 sub-block.  One warp
// instruction then covers 4 rows x 128 contiguous bytes with 16-byte accesses (4 cache lines instead of 32, a quarter
// of the instructions of the lane = column shape) and the 4 rows of a quad are 4 consecutive chains of one unit:
// exactly one Philox counter, no shuffles.
constexpr int kTQ = 32;                                   // unpadded 32 x 32 tile; the 16-byte chunk q of row r sits at q ^ (r % 8)
__device__ __forceinline__ void acc_block_to_quads(uint32_t acc_addr, bool has_acc, float* tb, int lane, float (&v)[2][4][4]) {
  if (!has_acc) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) v[i][j][c] = 0.0f;
    return;
  }
  float a[16];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    tmem_ld16(acc_addr + h * 16, a);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      *reinterpret_cast<float4*>(tb + lane * kTQ + 4 * ((h * 4 + k) ^ (lane & 7))) = make_float4(a[4 * k], a[4 * k + 1], a[4 * k + 2], a[4 * k + 3]);
  }
  __syncwarp();
  const int rq2 = lane >> 3, cq = lane & 7;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int R = 4 * (rq2 + 4 * i) + j;
      const float4 t4 = *reinterpret_cast<const float4*>(tb + R * kTQ + 4 * (cq ^ (R & 7)));
      v[i][j][0] = t4.x; v[i][j][1] = t4.y; v[i][j][2] = t4.z; v[i][j][3] = t4.w;
    }
  __syncwarp();
}

// latent update of layer t.idx: x <- x - lr*grad (SGD | Adam), x <- x - lr*noise, act(x) re-emitted in bf16
// SPEC = 1 folds the Langevin call's modes into constants (SGD, in-kernel Philox with aligned chain quads, no
// trajectories, no x.grad read-out): the epilogue is bound by instruction issue / fetch, dead branches cost.
template <int SPEC>
__device__ __forceinline__ void epilogue_update_q(const WideParams& p, const StepArgs& st, const TileDesc& t, uint32_t acc_q,
                                                  int q, int c_begin, float* tb, int lane) {
  const NetDev& nd = p.net;
  const int l = t.idx;
  const int dl = nd.dims[l];
  const int row0 = t.m0 + q * 32;
  const int kind = nd.act[l];
  const bool adam = (SPEC == 1) ? false : (p.optimizer == MCPC_OPT_ADAM);
  const bool quad_rng = (SPEC == 1) ? true : (((p.chain_offset + (uint64_t)row0) & 3) == 0);   // a row quad = one Philox counter
  const int noise_kind = (SPEC == 1) ? (int)MCPC_NOISE_PHILOX : p.noise_mode;
  const bool do_traj = (SPEC == 1) ? false : (st.do_traj != 0);
  const bool want_xgrad = (SPEC == 1) ? false : (st.last != 0);
  const bool upd_x = (SPEC == 1) ? true : (p.update_x != 0);
  const int rq2 = lane >> 3, cq = lane & 7;
  for (int sb = 0; sb < 4; ++sb) {
    const int n0 = t.n0 + c_begin + sb * 32;
    if (n0 >= dl) break;                                                // uniform over the warp (widths % 16 == 0)
    float bp[2][4][4];
    acc_block_to_quads(acc_q + c_begin + sb * 32, t.k_ext > 0, tb, lane, bp);
    const int n = n0 + 4 * cq;
    if (n >= dl) continue;
    const int gu = nd.off[l] + n;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int rbase = row0 + 4 * (rq2 + 4 * i);                       // first of the 4 consecutive chains
      if (rbase >= p.B) continue;
      float xv[4][4], g[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = rbase + j < p.B;
        const float4 x4 = ok ? *reinterpret_cast<const float4*>(p.x[l] + (size_t)(rbase + j) * dl + n) : make_float4(0, 0, 0, 0);
        const float4 g4 = ok ? *reinterpret_cast<const float4*>(p.G32 + (size_t)(rbase + j) * nd.SD + gu) : make_float4(0, 0, 0, 0);
        xv[j][0] = x4.x; xv[j][1] = x4.y; xv[j][2] = x4.z; xv[j][3] = x4.w;
        g[j][0] = g4.x; g[j][1] = g4.y; g[j][2] = g4.z; g[j][3] = g4.w;
      }
      if (do_traj && p.traj_x[l] != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (rbase + j < p.B)
            *reinterpret_cast<float4*>(p.traj_x[l] + ((size_t)st.rec * p.B + rbase + j) * dl + n) =
                make_float4(xv[j][0], xv[j][1], xv[j][2], xv[j][3]);
      }
      float nz[4][4];                                                   // [chain j][unit c]
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) nz[j][c] = 0.0f;
      if (noise_kind == MCPC_NOISE_PHILOX) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (quad_rng) {
            float q4[4];
            langevin_normals4(p.seed, (uint32_t)(gu + c), (uint32_t)st.t_abs, (p.chain_offset + (uint64_t)rbase) >> 2, q4);
#pragma unroll
            for (int j = 0; j < 4; ++j) nz[j][c] = p.noise_scale * q4[j];
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t chain = p.chain_offset + (uint64_t)(rbase + j);
              float q4[4];
              langevin_normals4(p.seed, (uint32_t)(gu + c), (uint32_t)st.t_abs, chain >> 2, q4);
              const int kc = (int)(chain & 3);
              nz[j][c] = p.noise_scale * (kc == 0 ? q4[0] : (kc == 1 ? q4[1] : (kc == 2 ? q4[2] : q4[3])));
            }
          }
        }
      } else if (noise_kind == MCPC_NOISE_SUPPLIED) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (rbase + j < p.B) {
            const float4 z4 = *reinterpret_cast<const float4*>(p.noise + ((size_t)st.ts * p.B + rbase + j) * nd.SD + gu);
            nz[j][0] = z4.x; nz[j][1] = z4.y; nz[j][2] = z4.z; nz[j][3] = z4.w;
          }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (rbase + j >= p.B) continue;
        const size_t xo = (size_t)(rbase + j) * dl + n;
        float mv[4] = {0, 0, 0, 0}, vv[4] = {0, 0, 0, 0};
        if (adam && upd_x) {
          const float4 m4 = *reinterpret_cast<const float4*>(p.m[l] + xo), v4 = *reinterpret_cast<const float4*>(p.v[l] + xo);
          mv[0] = m4.x; mv[1] = m4.y; mv[2] = m4.z; mv[3] = m4.w;
          vv[0] = v4.x; vv[1] = v4.y; vv[2] = v4.z; vv[3] = v4.w;
        }
        float gradv[4], a_new[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float x = xv[j][c];
          const float a = act_w(kind, x);
          const float grad = fmaf(dact_w(kind, x, a), bp[i][j][c], -g[j][c]);
          gradv[c] = grad;
          if (upd_x) {
            if (!adam) {
              x = fmaf(-p.lr, grad, x);
            } else {
              mv[c] = fmaf(p.one_minus_b1, grad - mv[c], mv[c]);
              vv[c] = fmaf(p.one_minus_b2 * grad, grad, vv[c] * p.beta2f);
              x = fmaf(-st.step_size, __fdividef(mv[c], fmaf(sqrtf(vv[c]), st.inv_bc2_sqrt, p.adam_eps)), x);
            }
          }
          x = fmaf(-p.lr, nz[j][c], x);
          xv[j][c] = x;
          a_new[c] = act_w(kind, x);
        }
        if (want_xgrad && p.xgrad[l] != nullptr)
          *reinterpret_cast<float4*>(p.xgrad[l] + xo) = make_float4(gradv[0], gradv[1], gradv[2], gradv[3]);
        if (adam && upd_x) {
          *reinterpret_cast<float4*>(p.m[l] + xo) = make_float4(mv[0], mv[1], mv[2], mv[3]);
          *reinterpret_cast<float4*>(p.v[l] + xo) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        }
        *reinterpret_cast<float4*>(p.x[l] + xo) = make_float4(xv[j][0], xv[j][1], xv[j][2], xv[j][3]);
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(a_new[0], a_new[1]), h1 = __floats2bfloat162_rn(a_new[2], a_new[3]);
        *reinterpret_cast<uint2*>(p.act + (size_t)(rbase + j) * p.a_pitch + p.poff[l] + n) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      }
    }
  }
}

// gW tile += accumulator (exactly one CTA owns each tile: plain read-modify-write), lane = input unit
__device__ __forceinline__ void epilogue_wgrad_t(const WideParams& p, const TileDesc& t, uint32_t acc_q, int q, int c_begin,
                                                 float* tb, int lane) {
  const NetDev& nd = p.net;
  const int lin = t.idx;
  const int d_o = (lin == nd.L) ? nd.d_out : nd.dims[lin], d_i = nd.dims[lin - 1];
  float* gW = p.gW[lin];
  const int mo0 = t.m0 + q * 32;
  for (int sb = 0; sb < 4; ++sb) {
    const int n0 = t.n0 + c_begin + sb * 32;
    if (n0 >= d_i) break;                                 // uniform over the warp
    float col[32];
    acc_block_to_columns(acc_q + c_begin + sb * 32, true, tb, lane, col);
    const int n = n0 + lane;
    if (gW == nullptr || n >= d_i) continue;
    float* dst = gW + (size_t)mo0 * d_i + n;
    float cur[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) cur[r] = (mo0 + r < d_o) ? dst[(size_t)r * d_i] : 0.0f;
#pragma unroll
    for (int r = 0; r < 32; ++r)
      if (mo0 + r < d_o) dst[(size_t)r * d_i] = cur[r] + col[r];
  }
}

// Persistent grouped GEMM: each CTA walks tiles blockIdx.x, +gridDim.x, ...  Warps 0-3 cp.async producers (4-stage ring),
// warp 4 MMA issuer, warps 5-8 epilogue; two 256-column accumulators in TMEM so the epilogue of tile i overlaps the
// mainloop of tile i+1.
template <int KIND, int SPEC>
__global__ void __launch_bounds__(64 + 32 * epi_warps(KIND), 1) wide_kernel(const __grid_constant__ WideParams p, const __grid_constant__ StepArgs st,
                                                      const __grid_constant__ WideMaps mp) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ Pipe pipe;
  __shared__ uint32_t tmem_s;
  __shared__ float s_red[8][2];
  constexpr bool A_MN = (KIND == KIND_WGRAD), B_MN = (KIND != KIND_PREDICT);
  constexpr uint32_t stage_bytes = kABytes + kBBytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = n_tiles_of<KIND>(p);

  if (tid == 0) {
    for (int s = 0; s < kWS; ++s) {
      mbar_init(&pipe.full[s], 1);
      mbar_init(&pipe.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&pipe.acc_full[b], 1);
      mbar_init(&pipe.acc_empty[b], 32 * epi_warps(KIND));
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_s, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_s;
  const uint32_t smem_base = smem_u32(smem);

  if (warp == 0) {
    // ---------------- TMA producer: one elected lane, one mbarrier transaction per 48 KB stage ----------------
    if (lane == 0) {
      uint32_t issued = 0;
      long long c_empty = 0, c_issue = 0;
      const bool prof = p.dbg != nullptr && blockIdx.x == 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileDesc t = decode_tile<KIND>(p, mp, tile);
        const int n_stage = (t.k_ext + kBK - 1) / kBK;
        for (int s = 0; s < n_stage; ++s, ++issued) {
          const uint32_t slot = issued % kWS;
          const long long t0 = prof ? clock64() : 0;
          if (p.l2_prefetch > 0 && s + p.l2_prefetch < n_stage) {
            // operands of a stage far beyond the shared-memory ring: into L2 now (first touches come from DRAM)
            const int kp = (s + p.l2_prefetch) * kBK;
            if (!A_MN) tma_prefetch_l2_2d(t.mapA, kp, t.m0);
            else if (p.mn3) tma_prefetch_l2_3d(t.mapA, 0, kp, t.m0 / 64);
            if (!B_MN) tma_prefetch_l2_2d(t.mapB, kp, t.n0);
            else if (p.mn3) tma_prefetch_l2_3d(t.mapB, 0, kp, t.n0 / 64);
          }
          mbar_wait(&pipe.empty[slot], ((issued / kWS) & 1u) ^ 1u);
          const long long t1 = prof ? clock64() : 0;
          uint8_t* sa = smem + slot * stage_bytes;
          uint8_t* sb = sa + kABytes;
          mbar_expect_tx(&pipe.full[slot], stage_bytes);
          const int k0 = s * kBK;
          if (!A_MN) tma_load_2d(sa, t.mapA, k0, t.m0, &pipe.full[slot]);
          else if (p.mn3) tma_load_3d(sa, t.mapA, 0, k0, t.m0 / 64, &pipe.full[slot]);
          else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d(sa + j * 8192, t.mapA, t.m0 + j * 64, k0, &pipe.full[slot]);
          }
          if (!B_MN) tma_load_2d(sb, t.mapB, k0, t.n0, &pipe.full[slot]);
          else if (p.mn3) tma_load_3d(sb, t.mapB, 0, k0, t.n0 / 64, &pipe.full[slot]);
          else {
#pragma unroll
            for (int j = 0; j < kBN / 64; ++j) tma_load_2d(sb + j * 8192, t.mapB, t.n0 + j * 64, k0, &pipe.full[slot]);
          }
          if (prof) { c_empty += t1 - t0; c_issue += clock64() - t1; }
        }
      }
      if (prof) { p.dbg[KIND * 8 + 0] = c_empty; p.dbg[KIND * 8 + 1] = c_issue; p.dbg[KIND * 8 + 2] = 0; p.dbg[KIND * 8 + 3] = issued; }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    const uint32_t id = idesc_bf16(128, kBN, A_MN, B_MN);
    // SWIZZLE_128B operand descriptors (tma.cuh): K-major LBO field 1 / SBO 1024, 32 B per K step;
    // MN-major LBO 8192 (next 64 units) / SBO 1024 (next 8 k-rows), 2048 B per K step
    constexpr uint32_t lbo_a = A_MN ? 8192u : 16u, lbo_b = B_MN ? 8192u : 16u;
    constexpr uint32_t adv_a = A_MN ? (2048u >> 4) : (32u >> 4), adv_b = B_MN ? (2048u >> 4) : (32u >> 4);
    uint32_t sc = 0, gi = 0;
    long long c_acc = 0, c_full = 0;
    const bool prof = p.dbg != nullptr && blockIdx.x == 0 && lane == 0;
    const long long k0 = clock64();
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const TileDesc t = decode_tile<KIND>(p, mp, tile);
      const int n_stage = (t.k_ext + kBK - 1) / kBK;
      if (n_stage == 0) continue;
      const uint32_t ab = gi & 1u;
      const long long a0 = clock64();
      mbar_wait(&pipe.acc_empty[ab], ((gi >> 1) & 1u) ^ 1u);
      c_acc += clock64() - a0;
      fence_after_sync();
      for (int s = 0; s < n_stage; ++s, ++sc) {
        const uint32_t slot = sc % kWS;
        const long long f0 = clock64();
        mbar_wait(&pipe.full[slot], (sc / kWS) & 1u);
        c_full += clock64() - f0;
        fence_after_sync();
        const uint64_t ad0 = smem_desc_sw128(smem_base + slot * stage_bytes, lbo_a, 1024u);
        const uint64_t bd0 = smem_desc_sw128(smem_base + slot * stage_bytes + kABytes, lbo_b, 1024u);
        if (elect1()) {
#pragma unroll
          for (int ks = 0; ks < kBK / 16; ++ks)
            mma_bf16_ss(tmem + ab * kBN, ad0 + (uint64_t)(ks * adv_a), bd0 + (uint64_t)(ks * adv_b), id, s > 0 || ks > 0);
          mma_commit(&pipe.empty[slot]);
          if (s == n_stage - 1) mma_commit(&pipe.acc_full[ab]);
        }
        __syncwarp();
      }
      ++gi;
    }
    if (prof) { p.dbg[KIND * 8 + 4] = c_acc; p.dbg[KIND * 8 + 5] = c_full; p.dbg[KIND * 8 + 6] = clock64() - k0; }
  } else {
    // ---------------- epilogue warps 2..9: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 ----------------
    const int q = warp & 3;
    const int ew = warp - 2;                               // 0 .. epi_warps-1
    constexpr int kColsPerWarp = kBN / (epi_warps(KIND) / 4);
    const int c_begin = (ew >> 2) * kColsPerWarp;
    float* trans = reinterpret_cast<float*>(smem + kWS * stage_bytes);     // 8 private transpose tiles after the ring
    uint32_t gi = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const TileDesc t = decode_tile<KIND>(p, mp, tile);
      const bool has_gemm = t.k_ext > 0;
      const uint32_t ab = gi & 1u;
      if (has_gemm) {
        mbar_wait(&pipe.acc_full[ab], (gi >> 1) & 1u);
        fence_after_sync();
      }
      const uint32_t acc_addr = tmem + ((uint32_t)(q * 32) << 16) + ab * kBN;
      if (p.skip_epilogue) {
        // debug (MCPC_WIDE_SKIP_EPI=1, results are garbage): mainloop-only rate of the three kernels
      } else if (KIND == KIND_PREDICT) {
        float e_part = 0.0f, l_part = 0.0f, gsum[4];
        float* my_tile = trans + ew * 32 * kTP;
        epilogue_predict_t(p, st, t, acc_addr, q, c_begin, my_tile, lane, e_part, l_part, gsum);
        e_part = warp_sum_w(e_part);
        l_part = warp_sum_w(l_part);
        if (lane == 0) { s_red[ew][0] = e_part; s_red[ew][1] = l_part; }
        // gb_l += column sums of G: every warp leaves its 128 partial sums (rows of its lane quarter) in its private
        // tile, the quarter-0 warp of each column half adds the four up and issues ONE atomic per column and tile
        if (st.acc) {
#pragma unroll
          for (int sb = 0; sb < 4; ++sb) my_tile[sb * 32 + lane] = gsum[sb];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ew == 0 && lane < 2) {
          float sum = 0.0f;
#pragma unroll
          for (int w = 0; w < 8; ++w) sum += s_red[w][lane];
          p.partials[((size_t)st.ts * p.n_part + tile) * 2 + lane] = sum;
        }
        if (st.acc && q == 0) {
          const NetDev& nd = p.net;
          const int lin = t.idx;
          const int d_o = (lin == nd.L) ? nd.d_out : nd.dims[lin];
          const bool live = (lin < nd.L) || nd.top_has_grad;
          if (live && p.gb[lin] != nullptr) {
#pragma unroll
            for (int sb = 0; sb < 4; ++sb) {
              const int n = t.n0 + c_begin + sb * 32 + lane;
              float tot = 0.0f;
#pragma unroll
              for (int w = 0; w < 4; ++w) tot += trans[((ew & 4) + w) * 32 * kTP + sb * 32 + lane];
              if (n < d_o) atomicAdd(p.gb[lin] + n, tot);
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      } else if (KIND == KIND_UPDATE) {
        epilogue_update_q<SPEC>(p, st, t, acc_addr, q, c_begin, trans + ew * 32 * kTQ, lane);
      } else {
        epilogue_wgrad_t(p, t, acc_addr, q, c_begin, trans + ew * 32 * kTP, lane);
      }
      if (has_gemm) {
        fence_before_sync();
        mbar_arrive(&pipe.acc_empty[ab]);
        ++gi;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

__global__ void to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

__global__ void init_act_kernel(WideParams p) {
  const NetDev& nd = p.net;
  const size_t total = (size_t)p.B * nd.SD;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / nd.SD), u = (int)(i % nd.SD);
    int l = 0;
    while (u >= nd.off[l + 1]) ++l;
    const int k = u - nd.off[l];
    p.act[(size_t)row * p.a_pitch + p.poff[l] + k] = __float2bfloat16(act_w(nd.act[l], p.x[l][(size_t)row * nd.dims[l] + k]));
  }
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct WideLayout {
  size_t wb_off[kMaxL + 1], act_off, gb_off, g32_off, part_off, total;
  int poff[kMaxL + 1], a_pitch, g_pitch, n_part;
};

int wide_layout(const NetDev& nd, int B, int n_steps, WideLayout* lay) {
  for (int l = 0; l < nd.L; ++l)
    if (nd.dims[l] % 16 != 0) {
      set_error("bf16 streaming path: layer widths must be multiples of 16 (got %d)", nd.dims[l]);
      return MCPC_ERR_UNSUPPORTED;
    }
  if (nd.d_out % 16 != 0) {
    set_error("bf16 streaming path: output width must be a multiple of 16 (got %d)", nd.d_out);
    return MCPC_ERR_UNSUPPORTED;
  }
  int f_off[kMaxL + 1];
  save_layout_bf16(nd, lay->poff, &lay->g_pitch, f_off, &lay->a_pitch);
  size_t o = 0;
  const int n_lin = nd.L + (nd.d_out > 0 ? 1 : 0);
  for (int l = 1; l < n_lin; ++l) {
    lay->wb_off[l] = o;
    o += align256((size_t)(l == nd.L ? nd.d_out : nd.dims[l]) * nd.dims[l - 1] * 2);
  }
  lay->act_off = o;
  o += align256((size_t)B * lay->a_pitch * 2 + 4096);
  lay->gb_off = o;
  o += align256((size_t)B * lay->g_pitch * 2 + 4096);
  lay->g32_off = o;
  o += align256((size_t)B * nd.SD * 4);
  const int mt = (B + 127) / 128;
  int n_part = 0;
  for (int l = 0; l < n_lin; ++l) n_part += mt * (((l == nd.L ? nd.d_out : nd.dims[l]) + kBN - 1) / kBN);
  lay->n_part = n_part;
  lay->part_off = o;
  o += align256((size_t)n_steps * n_part * 2 * sizeof(float));
  lay->total = o + 512;
  return MCPC_OK;
}

}  // namespace

int infer_wide_workspace(const NetDev& nd, int B, int n_steps, size_t* bytes) {
  WideLayout lay;
  int rc = wide_layout(nd, B, n_steps, &lay);
  if (rc != MCPC_OK) return rc;
  *bytes = lay.total;
  return MCPC_OK;
}

int launch_infer_wide(const NetDev& nd, const McpcIO* io, const McpcOpts* o, int B, void* ws, size_t ws_bytes,
                      cudaStream_t stream) {
  if (io->inputs != nullptr) {
    set_error("bf16 streaming path: non-zero `inputs` are not implemented; use MCPC_PREC_FP32");
    return MCPC_ERR_UNSUPPORTED;
  }
  if (io->save_g != nullptr) {
    set_error("bf16 streaming path accumulates the weight update itself (McpcIO.gW/gb); save_g/save_f are not used");
    return MCPC_ERR_INVALID;
  }
  WideLayout lay;
  int rc = wide_layout(nd, B, o->n_steps, &lay);
  if (rc != MCPC_OK) return rc;
  if (ws == nullptr || ws_bytes < lay.total) {
    set_error("workspace too small: %zu B given, %zu B needed", ws_bytes, lay.total);
    return MCPC_ERR_WORKSPACE;
  }
  uint8_t* wsb = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  WideParams p{};
  p.net = nd;
  p.B = B;
  p.mt = (B + 127) / 128;
  const int n_lin = nd.L + (nd.d_out > 0 ? 1 : 0);
  for (int l = 0; l <= nd.L; ++l) {
    p.poff[l] = lay.poff[l];
    p.b[l] = io->b[l];
    p.gW[l] = io->gW[l];
    p.gb[l] = io->gb[l];
  }
  p.a_pitch = lay.a_pitch;
  p.g_pitch = lay.g_pitch;
  p.act = reinterpret_cast<__nv_bfloat16*>(wsb + lay.act_off);
  p.Gb = reinterpret_cast<__nv_bfloat16*>(wsb + lay.gb_off);
  p.G32 = reinterpret_cast<float*>(wsb + lay.g32_off);
  p.partials = reinterpret_cast<float*>(wsb + lay.part_off);
  p.n_part = lay.n_part;
  for (int l = 1; l < n_lin; ++l) {
    __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(wsb + lay.wb_off[l]);
    const size_t n = (size_t)(l == nd.L ? nd.d_out : nd.dims[l]) * nd.dims[l - 1];
    to_bf16_kernel<<<(int)((n + 1023) / 1024 < 1184 ? (n + 1023) / 1024 : 1184), 256, 0, stream>>>(io->W[l], wb, n);
    count_launch();
    p.Wb[l] = wb;
  }
  for (int l = 0; l < nd.L; ++l) {
    p.x[l] = io->x[l];
    p.m[l] = io->adam_m[l];
    p.v[l] = io->adam_v[l];
    p.xgrad[l] = io->x_grad[l];
    p.traj_x[l] = io->traj_x[l];
  }
  p.traj_out = io->traj_out;
  p.target = io->target;
  p.noise = io->noise;
  p.optimizer = o->optimizer;
  p.update_x = o->update_x;
  p.lr = (float)o->lr;
  p.one_minus_b1 = (float)(1.0 - o->adam_beta1);
  p.one_minus_b2 = (float)(1.0 - o->adam_beta2);
  p.beta2f = (float)o->adam_beta2;
  p.adam_eps = (float)o->adam_eps;
  p.noise_mode = o->noise_mode;
  p.noise_scale = (float)o->noise_scale;
  p.seed = o->seed;
  p.chain_offset = o->chain_offset;
  // tile tables (128 chains / output units x kBN columns per tile)
  int t = 0;
  for (int l = 0; l <= nd.L; ++l) {
    p.tP_first[l] = t;
    if (l < nd.L || nd.d_out > 0) t += p.mt * (((l == nd.L ? nd.d_out : nd.dims[l]) + kBN - 1) / kBN);
  }
  p.tP_first[nd.L + 1] = t;
  const int n_predict = t;
  t = 0;
  for (int l = 0; l < nd.L; ++l) {
    p.tU_first[l] = t;
    t += p.mt * ((nd.dims[l] + kBN - 1) / kBN);
  }
  p.tU_first[nd.L] = t;
  const int n_update = t;
  t = 0;
  bool any_grad = false;
  p.tW_first[0] = p.tW_first[1] = 0;
  for (int l = 1; l <= nd.L; ++l) {
    p.tW_first[l] = t;
    const bool is_out = (l == nd.L);
    if (is_out && (nd.d_out == 0 || !nd.top_has_grad)) continue;
    if (p.gW[l] == nullptr) continue;
    const int d_o = is_out ? nd.d_out : nd.dims[l];
    t += ((d_o + 127) / 128) * ((nd.dims[l - 1] + kBN - 1) / kBN);
  }
  p.tW_first[nd.L + 1] = t;
  const int n_wgrad = t;
  for (int l = 0; l <= nd.L; ++l) any_grad = any_grad || p.gW[l] != nullptr || p.gb[l] != nullptr;
  if (p.n_part != n_predict) {
    set_error("internal: partial-slot count mismatch (%d vs %d)", p.n_part, n_predict);
    return MCPC_ERR_INVALID;
  }

  bool any_traj = io->traj_out != nullptr;
  for (int l = 0; l < nd.L; ++l) any_traj = any_traj || io->traj_x[l] != nullptr;
  const int traj_every = any_traj ? (o->traj_every > 0 ? o->traj_every : 1) : 0;

  // tensor maps: one per layer block (base offset = the block's first column) so TMA zero-fills past its extent
  WideMaps mp;
  bool mn3 = (nd.d_out % 64 == 0);
  for (int l = 0; l < nd.L; ++l) mn3 = mn3 && (nd.dims[l] % 64 == 0);
  p.mn3 = mn3 ? 1 : 0;
  p.l2_prefetch = 0;     // measured on C5: 0.98 ms/step without, 1.09-1.14 with 4/8/16 stages of L2 prefetch (extra TMA work, no gain)
  if (const char* env = getenv("MCPC_WIDE_L2PF")) p.l2_prefetch = atoi(env);
  p.skip_epilogue = getenv("MCPC_WIDE_SKIP_EPI") != nullptr ? 1 : 0;
  for (int l = 0; l < nd.L; ++l) {
    rc = make_tmap_bf16(&mp.act_k[l], p.act + p.poff[l], nd.dims[l], B, p.a_pitch, 64, 128);
    if (rc == MCPC_OK)
      rc = mn3 ? make_tmap_bf16_mn3(&mp.act_mn[l], p.act + p.poff[l], nd.dims[l], B, p.a_pitch, 64, kBN / 64)   // wgrad B operand
               : make_tmap_bf16(&mp.act_mn[l], p.act + p.poff[l], nd.dims[l], B, p.a_pitch, 64, 64);
    if (rc != MCPC_OK) return rc;
  }
  for (int l = 0; l < n_lin; ++l) {
    const int d_o = (l == nd.L) ? nd.d_out : nd.dims[l];
    rc = make_tmap_bf16(&mp.gb_k[l], p.Gb + p.poff[l], d_o, B, p.g_pitch, 64, 128);
    if (rc == MCPC_OK)
      rc = mn3 ? make_tmap_bf16_mn3(&mp.gb_mn[l], p.Gb + p.poff[l], d_o, B, p.g_pitch, 64, 2)                   // wgrad A operand
               : make_tmap_bf16(&mp.gb_mn[l], p.Gb + p.poff[l], d_o, B, p.g_pitch, 64, 64);
    if (rc == MCPC_OK && l >= 1) rc = make_tmap_bf16(&mp.w_k[l], p.Wb[l], nd.dims[l - 1], d_o, nd.dims[l - 1], 64, 256);
    if (rc == MCPC_OK && l >= 1)
      rc = mn3 ? make_tmap_bf16_mn3(&mp.w_mn[l], p.Wb[l], nd.dims[l - 1], d_o, nd.dims[l - 1], 64, kBN / 64)       // update B operand
               : make_tmap_bf16(&mp.w_mn[l], p.Wb[l], nd.dims[l - 1], d_o, nd.dims[l - 1], 64, 64);
    if (rc != MCPC_OK) return rc;
  }
  const size_t smem_g = (size_t)kWS * (kABytes + kBBytes) + kTransBytes + 1024;
  MCPC_CUDA_CHECK(cudaFuncSetAttribute(wide_kernel<KIND_PREDICT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
  MCPC_CUDA_CHECK(cudaFuncSetAttribute(wide_kernel<KIND_UPDATE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
  MCPC_CUDA_CHECK(cudaFuncSetAttribute(wide_kernel<KIND_UPDATE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
  MCPC_CUDA_CHECK(cudaFuncSetAttribute(wide_kernel<KIND_WGRAD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
  int n_sm = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (const char* env = getenv("MCPC_WIDE_CTAS")) {        // testing hook: few persistent CTAs => many tiles per CTA
      const int v = atoi(env);
      if (v >= 1 && v <= n_sm) n_sm = v;
    }
  }

  const bool timing = getenv("MCPC_WIDE_TIMING") != nullptr;      // debug only: allocates + synchronises
  if (timing) {
    cudaMalloc(&p.dbg, 32 * sizeof(long long));
    cudaMemsetAsync(p.dbg, 0, 32 * sizeof(long long), stream);
  }
  init_act_kernel<<<1184, 256, 0, stream>>>(p);
  count_launch();
  double b1p = pow(o->adam_beta1, (double)o->adam_step0), b2p = pow(o->adam_beta2, (double)o->adam_step0);
  bool spec_update = !any_traj && o->update_x && o->optimizer == MCPC_OPT_SGD && o->noise_mode == MCPC_NOISE_PHILOX &&
                     (o->chain_offset & 3) == 0;                       // what wide_kernel<KIND_UPDATE, 1> assumes
  for (int l = 0; l < nd.L; ++l) spec_update = spec_update && io->x_grad[l] == nullptr;
  if (getenv("MCPC_TC_NOSPEC") != nullptr) spec_update = false;       // testing hook: the generic instantiation
  for (int ts = 0; ts < o->n_steps; ++ts) {
    StepArgs st{};
    st.ts = ts;
    st.t_abs = o->t_begin + ts;
    st.do_traj = (traj_every > 0 && ts % traj_every == 0) ? 1 : 0;
    st.rec = st.do_traj ? ts / traj_every : 0;
    st.last = (ts == o->n_steps - 1) ? 1 : 0;
    if (o->optimizer == MCPC_OPT_ADAM && o->update_x) {
      b1p *= o->adam_beta1;
      b2p *= o->adam_beta2;
      st.step_size = (float)(o->lr / (1.0 - b1p));
      st.inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - b2p));
    }
    const bool acc = any_grad && ts >= o->save_begin && ts < o->save_end;
    st.acc = acc ? 1 : 0;                  // the predict epilogue adds the bias gradients (column sums of G) on these steps
    if (n_predict > 0) {
      wide_kernel<KIND_PREDICT, 0><<<n_predict < n_sm ? n_predict : n_sm, 320, smem_g, stream>>>(p, st, mp);
      count_launch();
    }
    // the weight update reads G (this step's errors) and act(x) of the state BEFORE the update: it runs between the two
    if (acc) {
      if (n_wgrad > 0) {
        wide_kernel<KIND_WGRAD, 0><<<n_wgrad < n_sm ? n_wgrad : n_sm, 320, smem_g, stream>>>(p, st, mp);
        count_launch();
      }
    }
    if (spec_update)
      wide_kernel<KIND_UPDATE, 1><<<n_update < n_sm ? n_update : n_sm, 64 + 32 * epi_warps(KIND_UPDATE), smem_g, stream>>>(p, st, mp);
    else
      wide_kernel<KIND_UPDATE, 0><<<n_update < n_sm ? n_update : n_sm, 64 + 32 * epi_warps(KIND_UPDATE), smem_g, stream>>>(p, st, mp);
    count_launch();
  }
  MCPC_CUDA_CHECK(cudaGetLastError());
  if (timing) {
    long long h[32];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    const char* names[3] = {"predict", "update", "wgrad"};
    for (int k = 0; k < 3; ++k)
      fprintf(stderr, "[wide timing] %s (CTA 0, last launch): producer wait-empty %lld, issue %lld, (%lld) cyc over %lld stages; "
                      "mma wait-acc %lld, wait-full %lld, total %lld cyc\n", names[k], h[k * 8], h[k * 8 + 1], h[k * 8 + 2], h[k * 8 + 3],
              h[k * 8 + 4], h[k * 8 + 5], h[k * 8 + 6]);
  }
  if (io->energy != nullptr || io->loss != nullptr) return launch_reduce_partials(p.partials, o->n_steps, p.n_part, io->energy, io->loss, stream);
  return MCPC_OK;
}

}  // namespace mcpc
