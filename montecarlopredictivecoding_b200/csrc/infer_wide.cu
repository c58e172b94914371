// mcpc_infer, MCPC_PREC_BF16, networks too wide to stay on chip (SURVEY config C5: 4 x 4096): the
// "streaming" path.  Latents x (fp32), their activations (bf16) and the error signals (bf16) live in HBM / L2;
// every Langevin step is two grouped tcgen05 GEMM kernels with fused epilogues, plus a third one every few steps:
//
//   PREDICT   for every Linear l:  mu = act(x_{l-1}) W_l^T + b  ->  eps = x_l - mu, energy, loss,
//             G_l = d overall / d mu_l (bf16: GEMM operand AND own-layer gradient term of the update), e_out, bias gradients
//   UPDATE    for every PCLayer l: bp = G_{l+1} W_{l+1};  grad = -G_l + act'(x_l) * bp;
//             x <- SGD | Adam step; x <- x - lr * noise (Philox);  act(x) re-emitted as bf16
//   WGRAD     gW_l += sum over the last s accumulate steps of G_l^T act(x_{l-1}): the bf16 operands of s steps stay
//             in a ring of s slots, so that the contraction index is (step, chain) and the fp32 gW tiles are read and
//             written once per s steps instead of once per step (537 MB of HBM traffic per C5 step otherwise).
//
// Orientation: UNITS on the M axis (TMEM lanes), CHAINS on the N axis (TMEM columns).  tcgen05.ld hands every thread
// one accumulator row, i.e. one unit and a run of chains: the 32 lanes of a warp then touch 32 CONSECUTIVE units of
// one chain -- every global access of the epilogues is a contiguous 128-byte row segment without any transpose
// (the first version had chains on M and needed a shared-memory transpose per 32 x 32 block), and the four chains of
// a Philox counter are four registers of one thread.
//
// All three kinds share one persistent mainloop: (128*CG) x 256 output tile, K in stages of 64, operands brought in
// by TMA (cp.async.bulk.tensor with SWIZZLE_128B tensor maps over the row-major global matrices; K-major or MN-major
// as the operand's storage order dictates -- no transposed copies of anything), mbarrier ring, two 256-column
// accumulators in TMEM so the epilogue of tile i overlaps the mainloop of tile i+1.  CG = 2 runs the tile on a CTA
// PAIR (cluster of 2, tcgen05.mma.cta_group::2, M = 256): each CTA stages its 128 units of A and HALF of the chains of
// B, so a stage is 32 KB per CTA instead of 48 KB and every byte of B is fetched into shared memory once per pair
// instead of once per CTA.  Warp roles: warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warps 2-9 epilogue
// (TMEM lane quarter = warp % 4, chain half = (warp - 2) / 4).
//
// Epilogue inputs (latents, own-layer term, targets) are staged through per-warp shared-memory rings by cp.async, several
// 16-chain chunks ahead and across the tile boundary: with registers alone a warp keeps 32 requests in flight, and at
// ~2 us of DRAM latency under load Little's law then makes a tile's epilogue as long as the mainloop it should hide
// behind (measured: 27-34 k cycles per tile before, 10-17 k after, mainloop 33 k).
//
// Measured on C5 (profiles/r02_*): the mainloop issues at the full tensor rate (128 cycles per 256x256x16 MMA)
// whenever operands are there; the step is bound by the 1 kW power cap (SM clock 1.05-1.3 GHz under this load, as
// for cuBLAS in a long loop), so what remains is energy per step, not issue slots.
// Reference semantics: predictive_coding/pc_trainer.py:733-918 + utils/model.py:35-44.
#include <cstdlib>

#include "mcpc_common.cuh"
#include "philox.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace mcpc {
namespace {

using namespace umma;

constexpr int kBK = 64;                     // K per stage (128 bytes: one SWIZZLE_128B row)
constexpr int kTM = 128;                    // units per CTA tile (M of one CTA)
constexpr int kTN = 256;                    // chains (PREDICT / UPDATE) or output units (WGRAD) per tile (N of the MMA)
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxStages = 6;
// operand ring depth: the UPDATE kernel trades one stage for deeper staging of its epilogue inputs (below)
#ifndef MCPC_UPD_NS
#define MCPC_UPD_NS 4
#endif
#ifndef MCPC_UPD_STG
#define MCPC_UPD_STG 12288
#endif
__host__ __device__ constexpr int n_stages(int cg, int kind) { return cg == 2 ? (kind == 1 ? MCPC_UPD_NS : 5) : 4; }
// per epilogue warp: shared-memory staging of the epilogue's global INPUTS (latents, own-layer term, targets), filled by
// cp.async several chunks ahead.  Registers could only keep one chunk (32 requests per warp) in flight, and with ~2 us
// of DRAM latency under load Little's law then caps the eight epilogue warps of an SM at ~16 GB/s -- the tile
// epilogues took as long as the mainloop they are supposed to hide behind.
// UPDATE stages 3 KB per chunk (fp32 latents + the bf16 own-layer term), PREDICT 2 KB (latents or targets): 4 chunks deep
// with CTA pairs, 1-2 chunks deep on single CTAs (whose operand stages are 48 KB).
__host__ __device__ constexpr uint32_t kUpdSlot = 3072u, kPredSlot = 2048u;
__host__ __device__ constexpr uint32_t stg_bytes(int cg, int kind) { return cg == 2 ? (kind == 1 ? (uint32_t)MCPC_UPD_STG : 8192u) : 4096u; }
__host__ __device__ constexpr uint32_t a_bytes() { return kTM * kBK * 2; }
__host__ __device__ constexpr uint32_t b_bytes(int cg) { return (kTN / cg) * kBK * 2; }
__host__ __device__ constexpr uint32_t stage_bytes(int cg) { return a_bytes() + b_bytes(cg); }

struct WideParams {
  NetDev net;
  float* x[kMaxL];
  float* m[kMaxL];
  float* v[kMaxL];
  float* xgrad[kMaxL];
  float* traj_x[kMaxL];
  float* traj_out;
  const float* b[kMaxL + 1];
  // One block per layer, each a dense row-major matrix of its own (pitch = width rounded up to 8): a tile's 256 chains
  // x 4096 units are then ONE contiguous 2 MB (bf16) region -- one page -- instead of 256 row segments spread over a
  // [chain][all layers] row of 32-40 KB.
  __nv_bfloat16* act_l[kMaxL];            // [S][Bpad][apitch[l]]: act(x_l), S ring slots
  __nv_bfloat16* gb_l[kMaxL + 1];         // [S][Bpad][gpitch[l]]: G_l (Linear l), e_out for l = L
  int apitch[kMaxL], gpitch[kMaxL + 1];
  int d_in_eff;                           // width of non-zero `inputs` (they enter Linear_0 as a bf16 operand block replicated
                                          // in every ring slot); 0: zero inputs, Linear_0 is bias-only
  const float* target;
  const float* noise;
  float* gW[kMaxL + 1];
  float* gb[kMaxL + 1];
  float* partials;                        // [n_steps][n_part][2]
  int B, Bpad;                            // chains; rows of one ring slot (B rounded up to 64)
  int tP_first[kMaxL + 2];                // predict tiles: prefix over Linear 0..L (index lin)
  int tU_first[kMaxL + 1];                // update tiles: prefix over layers 0..L-1
  int tW_first[kMaxL + 2];                // wgrad tiles: prefix over Linear 0..L
  int n_part;                             // partial slots per step (8 per predict CTA tile)
  int optimizer, update_x;
  float lr, adam_eps, one_minus_b1, one_minus_b2, beta2f;
  int noise_mode;
  float noise_scale;
  uint64_t seed, chain_offset;
  int mn3;                                // MN-major operands come in through 3-D tensor maps (all widths % 64 == 0)
  int pf_x, pf_t;                         // the fp32 prefetch maps exist (16-byte aligned bases and row strides)
  int cs;                                 // streaming cache hints on the fp32 epilogue streams
  int pdl;                                // launched with programmatic stream serialization
  int wpf;                                // stages ahead of the ring at which the WEIGHT operand is prefetched into L2 (0 = off)
  int skip_epilogue;                      // debug build only
  long long* dbg_buf;                     // debug build only: per-tile timeline of CTA 0 (MCPC_WIDE_TIMING)
};

struct StepArgs {
  int ts, t_abs, rec, do_traj, last;
  int acc;                                // this step is inside the weight-gradient window: bias gradients accumulate
  int slot, slot_next;                    // ring slot this step's act / G live in; slot the update writes act(x) to
  int k_rows;                             // WGRAD: contraction extent = (slots in use) * Bpad
  float step_size, inv_bc2_sqrt;          // Adam bias corrections of this step
};

// Tensor maps over the row-major bf16 matrices, one per layer block so that out-of-range K / M / N are zero-filled:
//   *_k : box 64 (inner = contraction index) x rows           -> K-major operand stage
//   *_mn: box 64 (inner = M/N index) x 64 rows (contraction)  -> MN-major operand stage, one 8 KB block per 64 units
struct WideMaps {
  CUtensorMap act_k[kMaxL], act_mn[kMaxL];
  CUtensorMap in_k, in_mn;                 // the inputs block (non-zero `inputs` only): B operand of PREDICT / A operand of WGRAD, Linear 0
  CUtensorMap gb_k[kMaxL + 1], gb_mn[kMaxL + 1];
  CUtensorMap w_k[kMaxL + 1];              // W_l   [d_l x d_{l-1}] bf16, K-major A operand of PREDICT
  CUtensorMap w_kt[kMaxL + 1];             // W_l^T [d_{l-1} x d_l] bf16 (transposed copy made per call), K-major A operand of UPDATE
  // fp32 views used only for L2 prefetches of the epilogue inputs (box = 128 units x 256 chains): the latents, the
  // own-layer term and the targets are first touched by latency-bound epilogue loads otherwise
  CUtensorMap x32[kMaxL], tgt;
};

struct Pipe {
  uint64_t full[kMaxStages], empty[kMaxStages], acc_full[2], acc_empty[2];
};

enum { KIND_PREDICT = 0, KIND_UPDATE = 1, KIND_WGRAD = 2 };

// Experiment hook of the DEBUG build (-DMCPC_DEBUG_BUILD, env MCPC_WIDE_EPI_MODE): 1 = epilogues skipped entirely (mainloop-only
// rate), 2 = epilogues without their global stores, 3 = without their global loads.  Results are garbage in all three; the
// release build compiles the constant 0.
#ifdef MCPC_DEBUG_BUILD
#define MCPC_EPI_MODE(p) ((p).skip_epilogue)
#else
#define MCPC_EPI_MODE(p) 0
#endif

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
__device__ __forceinline__ float warp_sum_w(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Streaming (evict-first) accesses for the fp32 epilogue streams -- latents, own-layer term, targets: 134 MB each per C5
// step, touched once per kernel -- so that they do not push the bf16 operand panels (reused by 8-16 tiles) out of L2.
__device__ __forceinline__ void st_stream(float* p, float v, bool cs) {
  if (cs) __stcs(p, v);
  else *p = v;
}

__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
  return v;
}

// ---- CTA-pair (cta_group::2) primitives ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t* dst_smem, uint32_t ncols) {      // one full warp (of each CTA)
  if (CG == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else {
    tmem_alloc(dst_smem, ncols);
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr, uint32_t ncols) {
  if (CG == 2)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
  else
    tmem_dealloc(taddr, ncols);
}
template <int CG>
__device__ __forceinline__ void mma_bf16_cg(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  if (CG == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
  } else {
    mma_bf16_ss(tmem_d, adesc, bdesc, idesc, accumulate);
  }
}
// arrive on the barrier at this shared-memory offset (of BOTH CTAs of the pair for CG = 2) once every MMA issued so far
// has completed
template <int CG>
__device__ __forceinline__ void mma_commit_cg(uint64_t* mbar) {
  if (CG == 2) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(mbar)), "h"((uint16_t)3) : "memory");
  } else {
    mma_commit(mbar);
  }
}
// TMA box loads; `bar` is a shared::cluster address (for CG = 2: the LEADER's full barrier, which counts the bytes of
// both CTAs)
template <int CG>
__device__ __forceinline__ void tma2d(void* dst_smem, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  if (CG == 2) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
  } else {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tma3d(void* dst_smem, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  if (CG == 2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
  } else {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
}

// ---- tiles ---------------------------------------------------------------------------------------------------
struct TileDesc {
  int idx;          // Linear index (predict / wgrad) or layer index (update)
  int m0;           // first unit (M) of THIS CTA's 128 rows of the tile
  int n0;           // first chain / output unit (N) of the tile (all 256 columns)
  int nb0;          // first N index this CTA stages into shared memory (its 256 / CG columns of B)
  int ni, ntn;      // index of the tile along N inside its group and the number of N tiles (tiles sharing the A panel)
  int k_ext;        // contraction extent; 0: no GEMM for this tile
  int k_base;       // first row (chain axis) of the K-major B operand: ring slot * Bpad (PREDICT / UPDATE), 0 for WGRAD
  const CUtensorMap* mapA;
  const CUtensorMap* mapB;
};

template <int KIND>
__device__ __forceinline__ int n_tiles_of(const WideParams& p) {
  return KIND == KIND_PREDICT ? p.tP_first[p.net.L + 1] : (KIND == KIND_UPDATE ? p.tU_first[p.net.L] : p.tW_first[p.net.L + 1]);
}

// Tiles are numbered per CTA pair (CG = 2) / CTA (CG = 1); inside a group the chain (N) tile index runs fastest so that
// the CTAs working at the same time share the weight rows of a few unit tiles and all the chains.
template <int KIND, int CG>
__device__ __forceinline__ TileDesc decode_tile(const WideParams& p, const StepArgs& st, const WideMaps& mp, int tile, int rank) {
  const NetDev& nd = p.net;
  TileDesc t{};
  if (KIND == KIND_PREDICT) {
    int lin = 0;
    while (tile >= p.tP_first[lin + 1]) ++lin;
    const int d_i = (lin == 0) ? p.d_in_eff : nd.dims[lin - 1];   // zero inputs: Linear_0 has no GEMM, mu_0 = b_0
    const int ntn = (p.B + kTN - 1) / kTN, local = tile - p.tP_first[lin];
    t.idx = lin; t.m0 = (local / ntn) * (kTM * CG) + rank * kTM; t.n0 = (local % ntn) * kTN; t.k_ext = d_i;
    t.ni = local % ntn; t.ntn = ntn;
    t.k_base = st.slot * p.Bpad;
    if (d_i > 0) {
      t.mapA = &mp.w_k[lin];                                      // W_l [d_o x d_i], K-major
      t.mapB = (lin == 0) ? &mp.in_k : &mp.act_k[lin - 1];        // act(x_{l-1}) (or the inputs) [chains x d_i], K-major
    }
  } else if (KIND == KIND_UPDATE) {
    int l = 0;
    while (tile >= p.tU_first[l + 1]) ++l;
    const int ntn = (p.B + kTN - 1) / kTN, local = tile - p.tU_first[l];
    const bool has_above = (l + 1 < nd.L) || nd.top_has_grad;
    const int d_up = (l + 1 < nd.L) ? nd.dims[l + 1] : nd.d_out;
    t.idx = l; t.m0 = (local / ntn) * (kTM * CG) + rank * kTM; t.n0 = (local % ntn) * kTN; t.k_ext = has_above ? d_up : 0;
    t.ni = local % ntn; t.ntn = ntn;
    t.k_base = st.slot * p.Bpad;
    if (has_above) {
      t.mapA = &mp.w_kt[l + 1];                                   // W_{l+1}^T [d_l x d_up], K-major (like PREDICT's operands)
      t.mapB = &mp.gb_k[l + 1];                                   // G_{l+1} [chains x d_up], K-major
    }
  } else {
    int lin = 0;                                                  // Linear 0 owns tiles only with non-zero inputs
    while (tile >= p.tW_first[lin + 1]) ++lin;
    const int d_o = (lin == nd.L) ? nd.d_out : nd.dims[lin];
    const int ntn = (d_o + kTN - 1) / kTN, local = tile - p.tW_first[lin];
    t.idx = lin; t.m0 = (local / ntn) * (kTM * CG) + rank * kTM; t.n0 = (local % ntn) * kTN; t.k_ext = st.k_rows;
    t.ni = local % ntn; t.ntn = ntn;
    t.k_base = 0;
    t.mapA = (lin == 0) ? &mp.in_mn : &mp.act_mn[lin - 1];        // act(x_{l-1}) (or the inputs) as A[m = input unit][k = chain]
    t.mapB = &mp.gb_mn[lin];                                      // G_l as B[n = output unit][k = chain]: MN-major
  }
  t.nb0 = t.n0 + rank * (kTN / CG);
  return t;
}

// ---- epilogues: lane = unit (accumulator row), registers = chains ---------------------------------------------
struct EpiPos {
  int q, h, lane, ew;     // TMEM lane quarter, column half, lane, epilogue warp index
};

// errors of the units a tile predicts: eps / energy / own-layer G (hidden Linears) or loss / dLoss (output).
// One chunk = 16 chains of one unit per lane; GUARD = false: every (lane, chain) of the chunk exists (steady state, no
// predicates); GUARD = true: edge chunks.  All global offsets are 32-bit element indices from warp-uniform bases.
struct PredCtx {
  const float* in;             // hidden: p.x[lin]; output: p.target
  __nv_bfloat16* gb;           // G ring slot of this step
  float* traj;                 // output: trajectory record of this step or nullptr
  uint32_t d_o, g_pitch;       // pitches of in and gb
  uint32_t io, bo;             // element offsets of (chain 0 of the tile half, this lane's unit) in in / gb
  float bias, ce, gc, inv_var;
  int mode;                    // output: 0 no target (TOP_NONE / ZERO), 1 Gaussian, 2 Bernoulli
  bool on;                     // output: unit inside the loss mask
  int dbg;                     // MCPC_EPI_MODE
  bool cs;
  // tile geometry of this warp
  int lin, c_base, n_ok, n_warp, lane;
  bool is_out, lanes_full, want_in, u_ok;
  uint32_t stg;                // shared-memory address of the warp's staging buffer
};

// stage chunk `cc` (16 chains of this lane's unit) of the predict input into ring slot `slot` of the warp's buffer
__device__ __forceinline__ void pred_stage(const PredCtx& c, int cc, int slot) {
  if (c.want_in && cc < c.n_warp) {
    const uint32_t dst = c.stg + (uint32_t)slot * kPredSlot + (uint32_t)c.lane * 4u;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (cc + j < c.n_ok) cp_async4(dst + j * 128, c.in + (c.io + (uint32_t)(cc + j) * c.d_o));
  }
  cp_async_commit();
}
__device__ __forceinline__ void pred_fetch(const PredCtx& c, int slot, float (&v)[16]) {
  const uint32_t src = c.stg + (uint32_t)slot * kPredSlot + (uint32_t)c.lane * 4u;
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = c.want_in ? lds_f32(src + j * 128) : 0.0f;
}

template <bool GUARD>
__device__ __forceinline__ void pred_chunk_hidden(const PredCtx& c, int cc, int n_ok, uint32_t acc_addr, bool has_acc,
                                                  const float (&xv)[16], float& e_part, float& gsum) {
  float d[16];
  if (has_acc) {
    tmem_ld16(acc_addr, d);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) d[j] = 0.0f;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const bool ok = !GUARD || (cc + j < n_ok);
    const float eps = ok ? xv[j] - (d[j] + c.bias) : 0.0f;                // (unstaged slots hold stale bytes: never used)
    e_part = fmaf(c.ce * eps, eps, e_part);
    const float g = -c.gc * eps;
    const __nv_bfloat16 gb16 = __float2bfloat16(g);
    gsum += __bfloat162float(gb16);                                   // the bias gradient sums the operand the dW GEMM sees
    if (ok && c.dbg != 2) {
      c.gb[c.bo + (uint32_t)(cc + j) * c.g_pitch] = gb16;
    }
  }
}

template <bool GUARD>
__device__ __forceinline__ void pred_chunk_out(const PredCtx& c, int cc, int n_ok, uint32_t acc_addr, bool has_acc,
                                               const float (&yv)[16], float& l_part, float& gsum) {
  float d[16];
  if (has_acc) {
    tmem_ld16(acc_addr, d);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) d[j] = 0.0f;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const bool ok = !GUARD || (cc + j < n_ok);
    const float o = d[j] + c.bias;
    float lv = 0.0f, e = 0.0f;
    const float yj = (GUARD && !ok) ? 0.0f : yv[j];
    if (c.mode == 2) {
      const float z = __expf(-fabsf(o));
      lv = fmaxf(o, 0.0f) - o * yj + __logf(1.0f + z);
      e = __fdividef(o >= 0.0f ? 1.0f : z, 1.0f + z) - yj;
    } else if (c.mode == 1) {
      const float dd = o - yj;
      lv = 0.5f * c.inv_var * dd * dd;
      e = dd * c.inv_var;
    }
    const bool live = ok && c.on;
    l_part += live ? lv : 0.0f;
    const __nv_bfloat16 eb16 = __float2bfloat16(live ? e : 0.0f);
    gsum += __bfloat162float(eb16);
    if (ok && c.dbg != 2) {
      c.gb[c.bo + (uint32_t)(cc + j) * c.g_pitch] = eb16;
      if (c.traj != nullptr) st_stream(c.traj + (c.io + (uint32_t)(cc + j) * c.d_o), o, c.cs);
    }
  }
}

// DEPTH = chunks the staging ring holds (stg_bytes / 2 KB)
__device__ __forceinline__ void predict_ctx(const WideParams& p, const StepArgs& st, const TileDesc& t, const EpiPos& ep, uint32_t stg,
                                            PredCtx& c) {
  const NetDev& nd = p.net;
  const int lin = t.idx;
  const bool is_out = (lin == nd.L);
  const int d_o = is_out ? nd.d_out : nd.dims[lin];
  const int u = t.m0 + ep.q * 32 + ep.lane;
  const bool u_ok = u < d_o;
  const int c_base = t.n0 + ep.h * (kTN / 2);
  c.lin = lin; c.is_out = is_out; c.c_base = c_base; c.lane = ep.lane; c.u_ok = u_ok; c.stg = stg;
  c.d_o = (uint32_t)d_o; c.g_pitch = (uint32_t)p.gpitch[lin];
  c.bias = (u_ok && p.b[lin] != nullptr) ? __ldg(p.b[lin] + u) : 0.0f;
  c.ce = is_out ? 0.0f : 0.5f * nd.c[lin];
  c.gc = is_out ? 0.0f : nd.gc[lin];
  c.inv_var = nd.inv_var;
  c.mode = nd.top == MCPC_TOP_BERNOULLI ? 2 : (nd.top == MCPC_TOP_GAUSS ? 1 : 0);
  c.on = c.mode != 0 && u >= nd.mask_start;
  c.dbg = MCPC_EPI_MODE(p);
  c.cs = p.cs != 0;
  c.in = is_out ? p.target : p.x[lin];
  c.gb = p.gb_l[lin] + ((size_t)st.slot * p.Bpad) * p.gpitch[lin];
  c.traj = (is_out && st.do_traj && p.traj_out != nullptr) ? p.traj_out + (size_t)st.rec * p.B * d_o : nullptr;
  c.io = (uint32_t)c_base * c.d_o + (uint32_t)u;
  c.bo = (uint32_t)c_base * c.g_pitch + (uint32_t)u;
  c.want_in = (!is_out || c.mode != 0) && c.dbg != 3;                    // the output tile reads the target only under a loss
  c.n_ok = u_ok ? (p.B - c_base) : 0;                                    // chunk-relative chains cc < n_ok are this lane's
  c.n_warp = min(kTN / 2, p.B - c_base);                                 // chains of this tile half that exist (uniform)
  c.lanes_full = (t.m0 + ep.q * 32 + 32 <= d_o);                         // uniform: every lane of the warp has a unit
}

template <int DEPTH>
__device__ __forceinline__ void predict_prestage(const PredCtx& c) {
#pragma unroll
  for (int k = 0; k < DEPTH; ++k) pred_stage(c, 16 * k, k);
}

template <int DEPTH>
__device__ __forceinline__ void epilogue_predict(const WideParams& p, const StepArgs& st, const PredCtx& c, uint32_t acc, bool has_acc,
                                                 int slot_id) {
  const NetDev& nd = p.net;
  float e_part = 0.0f, l_part = 0.0f, gsum = 0.0f;
  int ci = 0;
#pragma unroll 1
  for (int cc = 0; cc < c.n_warp; cc += 16, ++ci) {
    cp_async_wait<DEPTH - 1>();                                          // chunk ci has landed (its group is the oldest)
    const int slot = ci % DEPTH;
    float v[16];
    pred_fetch(c, slot, v);
    pred_stage(c, cc + 16 * DEPTH, slot);                                // refill the slot just read
    const bool full = c.lanes_full && cc + 16 <= c.n_warp;
    if (!c.is_out) {
      if (full) pred_chunk_hidden<false>(c, cc, c.n_ok, acc + cc, has_acc, v, e_part, gsum);
      else pred_chunk_hidden<true>(c, cc, c.n_ok, acc + cc, has_acc, v, e_part, gsum);
    } else {
      if (full) pred_chunk_out<false>(c, cc, c.n_ok, acc + cc, has_acc, v, l_part, gsum);
      else pred_chunk_out<true>(c, cc, c.n_ok, acc + cc, has_acc, v, l_part, gsum);
    }
  }
  cp_async_wait<0>();
  e_part = warp_sum_w(e_part);
  l_part = warp_sum_w(l_part);
  if (c.lane == 0) {
    float* dst = p.partials + ((size_t)st.ts * p.n_part + slot_id) * 2;
    dst[0] = e_part;
    dst[1] = l_part;
  }
  // gb_l += column sums of G over this warp's chains (lane = unit: the sum is already in a register)
  if (st.acc && c.u_ok && p.gb[c.lin] != nullptr && ((c.lin < nd.L) || nd.top_has_grad))
    atomicAdd(p.gb[c.lin] + (c.io - (uint32_t)c.c_base * c.d_o), gsum);
}

// latent update of layer t.idx: x <- x - lr*grad (SGD | Adam), x <- x - lr*noise, act(x) re-emitted in bf16.
// SPEC = 1 folds the Langevin call's modes into constants (SGD, in-kernel Philox with aligned chain quads, no
// trajectories, no x.grad read-out) and ACT is the layer's activation as a compile-time constant (ACT < 0: read from
// the net at run time).  Only loads and stores are predicated: the arithmetic runs unconditionally on every lane.
template <int ACT>
__device__ __forceinline__ float act_t(int kind, float x) {
  if (ACT == MCPC_ACT_RELU) return fmaxf(x, 0.0f);
  if (ACT == MCPC_ACT_TANH) return tanh_fast_w(x);
  if (ACT == MCPC_ACT_IDENTITY) return x;
  const float th = tanh_fast_w(x), rl = fmaxf(x, 0.0f);                  // run-time kind: branch-free selects
  return kind == MCPC_ACT_TANH ? th : (kind == MCPC_ACT_RELU ? rl : x);
}
template <int ACT>
__device__ __forceinline__ float dact_t(int kind, float x, float a) {
  if (ACT == MCPC_ACT_RELU) return x > 0.0f ? 1.0f : 0.0f;
  if (ACT == MCPC_ACT_TANH) return fmaf(-a, a, 1.0f);
  if (ACT == MCPC_ACT_IDENTITY) return 1.0f;
  return kind == MCPC_ACT_TANH ? fmaf(-a, a, 1.0f) : (kind == MCPC_ACT_RELU ? (x > 0.0f ? 1.0f : 0.0f) : 1.0f);
}

// One chunk = 16 chains of one unit per lane.  GUARD = false: every (lane, chain) of the chunk exists (the steady state:
// no predicates at all); GUARD = true: edge chunks, loads and stores predicated by `n_ok` (chains c < n_ok are valid).
// All global offsets are 32-bit element indices from warp-uniform base pointers (wide_layout checks the ranges).
struct UpdCtx {
  float* x;                    // p.x[l]
  const __nv_bfloat16* gown;   // G_l block of this step's ring slot: the own-layer gradient term (bf16 operand copy)
  __nv_bfloat16* act;          // act ring slot the update writes
  uint32_t dl, a_pitch, g_pitch;
  uint32_t xo, ao, bo;         // element offsets of (chain 0 of the tile half, this lane's unit) in x / act / gown
  uint32_t gu;                 // global unit index (Philox counter word 0)
  float nlr, nscale;
  int dbg;                     // MCPC_EPI_MODE
  bool cs;
  // tile geometry of this warp
  int l, kind, c_base, n_ok, n_warp, lane;
  bool lanes_full;
  uint32_t stg;                // shared-memory address of the warp's staging buffer
};

// stage chunk `cc` (latents and own-layer term of 16 chains of this lane's unit) into ring slot `slot`: fp32 x at bytes
// [0, 2048), bf16 G_l at [2048, 3072).  cp.async moves at least 4 bytes, so the even lanes fetch their unit's and their
// right neighbour's bf16 value (the block pitch is padded to 8 units: the pair always exists in memory).
__device__ __forceinline__ void upd_stage(const UpdCtx& c, int cc, int slot) {
  if (cc < c.n_warp && c.dbg != 3) {
    const uint32_t dst = c.stg + (uint32_t)slot * kUpdSlot + (uint32_t)c.lane * 4u;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (cc + j < c.n_ok) cp_async4(dst + j * 128, c.x + (c.xo + (uint32_t)(cc + j) * c.dl));
    const int n_pair = max(c.n_ok, __shfl_down_sync(0xffffffffu, c.n_ok, 1));
    if ((c.lane & 1) == 0) {
      const uint32_t dstb = c.stg + (uint32_t)slot * kUpdSlot + 2048u + (uint32_t)c.lane * 2u;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (cc + j < n_pair) cp_async4(dstb + j * 64, c.gown + (c.bo + (uint32_t)(cc + j) * c.g_pitch));
    }
  }
  cp_async_commit();
}
__device__ __forceinline__ void upd_fetch(const UpdCtx& c, int slot, float (&xv)[16], float (&gv)[16]) {
  const uint32_t src = c.stg + (uint32_t)slot * kUpdSlot + (uint32_t)c.lane * 4u;
  const uint32_t srcb = c.stg + (uint32_t)slot * kUpdSlot + 2048u + (uint32_t)c.lane * 2u;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    xv[j] = lds_f32(src + j * 128);
    uint16_t h;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(srcb + j * 64) : "memory");
    gv[j] = __uint_as_float((uint32_t)h << 16);
  }
}

template <int SPEC, int ACT, bool GUARD>
__device__ __forceinline__ void upd_chunk(const WideParams& p, const StepArgs& st, const UpdCtx& c, int l, int kind, int c_abs, int cc,
                                          int n_ok, uint32_t acc_addr, bool has_acc, const float (&xv)[16], const float (&gv)[16]) {
  const bool adam = (SPEC == 1) ? false : (p.optimizer == MCPC_OPT_ADAM);
  const bool quad_rng = (SPEC == 1) ? true : ((p.chain_offset & 3) == 0);       // 4 aligned chains = one Philox counter
  const int noise_kind = (SPEC == 1) ? (int)MCPC_NOISE_PHILOX : p.noise_mode;
  const bool do_traj = (SPEC == 1) ? false : (st.do_traj != 0 && p.traj_x[l] != nullptr);
  const bool want_xgrad = (SPEC == 1) ? false : (st.last != 0 && p.xgrad[l] != nullptr);
  const bool upd_x = (SPEC == 1) ? true : (p.update_x != 0);
  float nz[16];
  if (noise_kind == MCPC_NOISE_PHILOX) {
    if (quad_rng) {
#pragma unroll
      for (int qd = 0; qd < 4; ++qd)
        langevin_normals4(p.seed, c.gu, (uint32_t)st.t_abs, (p.chain_offset + (uint64_t)(c_abs + 4 * qd)) >> 2, &nz[4 * qd]);
    } else {
#pragma unroll 1
      for (int j = 0; j < 16; ++j) {
        const uint64_t chain = p.chain_offset + (uint64_t)(c_abs + j);
        float q4[4];
        langevin_normals4(p.seed, c.gu, (uint32_t)st.t_abs, chain >> 2, q4);
        const int kc = (int)(chain & 3);
        nz[j] = kc == 0 ? q4[0] : (kc == 1 ? q4[1] : (kc == 2 ? q4[2] : q4[3]));
      }
    }
  } else if (noise_kind == MCPC_NOISE_SUPPLIED) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      nz[j] = (!GUARD || cc + j < n_ok) ? __ldg(p.noise + ((size_t)st.ts * p.B + c_abs + j) * p.net.SD + c.gu) : 0.0f;
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) nz[j] = 0.0f;
  }
  float bp[16];
  if (has_acc) {
    tmem_ld16(acc_addr, bp);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) bp[j] = 0.0f;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const bool ok = !GUARD || (cc + j < n_ok);
    const uint32_t xo = c.xo + (uint32_t)(cc + j) * c.dl;
    float x = (GUARD && !ok) ? 0.0f : xv[j];                              // (unstaged slots hold stale bytes)
    if (SPEC == 0 && do_traj && ok) p.traj_x[l][((size_t)st.rec * p.B + c_abs + j) * c.dl + (c.gu - (uint32_t)p.net.off[l])] = x;
    const float a = act_t<ACT>(kind, x);
    const float grad = fmaf(dact_t<ACT>(kind, x, a), bp[j], -((GUARD && !ok) ? 0.0f : gv[j]));
    if (SPEC == 0 && want_xgrad && ok) p.xgrad[l][xo] = grad;
    if (SPEC == 1) {
      x = fmaf(c.nlr, grad, x);
    } else if (upd_x) {
      if (!adam) {
        x = fmaf(c.nlr, grad, x);
      } else if (ok) {
        float mv = p.m[l][xo], vv = p.v[l][xo];
        mv = fmaf(p.one_minus_b1, grad - mv, mv);
        vv = fmaf(p.one_minus_b2 * grad, grad, vv * p.beta2f);
        x = fmaf(-st.step_size, __fdividef(mv, fmaf(sqrtf(vv), st.inv_bc2_sqrt, p.adam_eps)), x);
        p.m[l][xo] = mv;
        p.v[l][xo] = vv;
      }
    }
    x = fmaf(c.nscale, nz[j], x);
    const __nv_bfloat16 a16 = __float2bfloat16(act_t<ACT>(kind, x));
    if (ok && c.dbg != 2) {
      st_stream(c.x + xo, x, c.cs);
      c.act[c.ao + (uint32_t)(cc + j) * c.a_pitch] = a16;
    }
  }
}

__device__ __forceinline__ void update_ctx(const WideParams& p, const StepArgs& st, const TileDesc& t, const EpiPos& ep, uint32_t stg,
                                           UpdCtx& c) {
  const NetDev& nd = p.net;
  const int l = t.idx;
  const int dl = nd.dims[l];
  const int u = t.m0 + ep.q * 32 + ep.lane;
  const bool u_ok = u < dl;
  const int c_base = t.n0 + ep.h * (kTN / 2);
  c.l = l; c.kind = nd.act[l]; c.c_base = c_base; c.lane = ep.lane; c.stg = stg;
  c.x = p.x[l];
  c.gown = p.gb_l[l] + ((size_t)st.slot * p.Bpad) * p.gpitch[l];
  c.act = p.act_l[l] + ((size_t)st.slot_next * p.Bpad) * p.apitch[l];
  c.dl = (uint32_t)dl; c.a_pitch = (uint32_t)p.apitch[l]; c.g_pitch = (uint32_t)p.gpitch[l];
  c.xo = (uint32_t)c_base * c.dl + (uint32_t)u;
  c.ao = (uint32_t)c_base * c.a_pitch + (uint32_t)u;
  c.bo = (uint32_t)c_base * c.g_pitch + (uint32_t)u;
  c.gu = (uint32_t)(nd.off[l] + u);
  c.nlr = -p.lr;
  c.nscale = c.nlr * p.noise_scale;                                      // x <- x - lr * (noise_scale * xi)
  c.dbg = MCPC_EPI_MODE(p);
  c.cs = p.cs != 0;
  c.n_ok = u_ok ? (p.B - c_base) : 0;                                    // chunk-relative chains cc < n_ok are this lane's
  c.n_warp = min(kTN / 2, p.B - c_base);                                 // chains of this tile half that exist (uniform)
  c.lanes_full = (t.m0 + ep.q * 32 + 32 <= dl);                          // uniform: every lane of the warp has a unit
}

template <int DEPTH>
__device__ __forceinline__ void update_prestage(const UpdCtx& c) {
#pragma unroll
  for (int k = 0; k < DEPTH; ++k) upd_stage(c, 16 * k, k);
}

template <int SPEC, int ACT, int DEPTH>
__device__ __forceinline__ void epilogue_update(const WideParams& p, const StepArgs& st, const UpdCtx& c, uint32_t acc, bool has_acc) {
  int ci = 0;
#pragma unroll 1
  for (int cc = 0; cc < c.n_warp; cc += 16, ++ci) {
    cp_async_wait<DEPTH - 1>();                                          // chunk ci has landed (its group is the oldest)
    const int slot = ci % DEPTH;
    float xv[16], gv[16];
    upd_fetch(c, slot, xv, gv);
    upd_stage(c, cc + 16 * DEPTH, slot);                                 // refill the slot just read
    if (c.lanes_full && cc + 16 <= c.n_warp)
      upd_chunk<SPEC, ACT, false>(p, st, c, c.l, c.kind, c.c_base + cc, cc, c.n_ok, acc + cc, has_acc, xv, gv);
    else
      upd_chunk<SPEC, ACT, true>(p, st, c, c.l, c.kind, c.c_base + cc, cc, c.n_ok, acc + cc, has_acc, xv, gv);
  }
  cp_async_wait<0>();
}

// gW tile += accumulator (exactly one CTA owns each tile: plain read-modify-write); lane = INPUT unit, so for a fixed
// output unit the 32 lanes touch 32 consecutive floats of one gW row
__device__ __forceinline__ void epilogue_wgrad(const WideParams& p, const TileDesc& t, uint32_t acc, const EpiPos& ep) {
  const NetDev& nd = p.net;
  const int lin = t.idx;
  const int d_o = (lin == nd.L) ? nd.d_out : nd.dims[lin], d_i = (lin == 0) ? p.d_in_eff : nd.dims[lin - 1];
  const int mi = t.m0 + ep.q * 32 + ep.lane;
  const bool m_ok = mi < d_i && p.gW[lin] != nullptr;
  float* gW = p.gW[lin] + mi;
  const int n_base = t.n0 + ep.h * (kTN / 2);
#pragma unroll 1
  for (int ch = 0; ch < kTN / 32; ++ch) {
    const int n0 = n_base + ch * 16;
    if (n0 >= d_o) break;                                 // uniform over the warp
    float cur[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) cur[j] = (m_ok && n0 + j < d_o) ? gW[(size_t)(n0 + j) * d_i] : 0.0f;
    float col[16];
    tmem_ld16(acc + ch * 16, col);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (m_ok && n0 + j < d_o) gW[(size_t)(n0 + j) * d_i] = cur[j] + col[j];
  }
}

// Persistent grouped GEMM: each CTA (pair) walks tiles first, +stride, ...
template <int KIND, int SPEC, int CG>
__global__ void __launch_bounds__(kThreads, 1) wide_kernel(const __grid_constant__ WideParams p, const __grid_constant__ StepArgs st,
                                                           const __grid_constant__ WideMaps mp) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ Pipe pipe;
  __shared__ uint32_t tmem_s;
  constexpr bool A_MN = (KIND == KIND_WGRAD), B_MN = (KIND == KIND_WGRAD);
  constexpr int NS = n_stages(CG, KIND);
  constexpr uint32_t kA = a_bytes(), kB = b_bytes(CG), kStage = stage_bytes(CG);
  constexpr int kBRows = kTN / CG;                       // N indices of B this CTA stages
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = n_tiles_of<KIND>(p);
  const int rank = (CG == 2) ? (int)cluster_ctarank() : 0;
  const int first_tile = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_stride = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;

#ifdef MCPC_DEBUG_BUILD
  if (p.dbg_buf != nullptr && blockIdx.x == 0 && tid == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.dbg_buf[KIND * 256 + 250] = (long long)ns;
  }
#endif
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&pipe.full[s], 1);
      mbar_init(&pipe.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&pipe.acc_full[b], 1);
      mbar_init(&pipe.acc_empty[b], CG * kEpiWarps);
    }
    fence_mbar_init();
  }
  // Programmatic dependent launch: the next kernel of the step may become resident on this SM as soon as this CTA is gone
  // (its barrier init / TMEM allocation then overlap the other SMs' last tiles instead of a full grid drain + launch);
  // it touches no global memory before its own griddepcontrol.wait below.
  if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 1) tmem_alloc_cg<CG>(&tmem_s, 512);
  fence_before_sync();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_s;
  const uint32_t smem_base = smem_u32(smem);
  if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");      // everything the previous kernels wrote is visible from here on

  if (warp == 0) {
    // ---------------- TMA producer: one elected lane, one mbarrier transaction per stage ----------------
    if (lane == 0) {
      uint32_t issued = 0;
      for (int tile = first_tile; tile < n_tiles; tile += tile_stride) {
        const TileDesc t = decode_tile<KIND, CG>(p, st, mp, tile, rank);
        const int n_stage = (t.k_ext + kBK - 1) / kBK;
        // what this tile's epilogue will read, into L2 now: the epilogue runs one mainloop later
        if (KIND == KIND_PREDICT) {
          if (t.idx < p.net.L) {
            if (p.pf_x) tma_prefetch_l2_2d(&mp.x32[t.idx], t.m0, t.n0);
          } else if (p.pf_t) {
            tma_prefetch_l2_2d(&mp.tgt, t.m0, t.n0);
          }
        } else if (KIND == KIND_UPDATE) {
          if (p.pf_x) tma_prefetch_l2_2d(&mp.x32[t.idx], t.m0, t.n0);
        }
        // The weight panel of a tile (A operand of PREDICT / UPDATE) is cold in L2: the tiles that share it -- the ntn chain
        // tiles of one unit tile -- run in lock step on different SMs, so every one of them waits for the SAME DRAM fetch,
        // whose latency under the epilogues' traffic exceeds what the shared-memory ring covers.  Each of the sharing tiles
        // therefore prefetches every ntn-th stage of the panel into L2, p.wpf stages ahead (across the tile boundary too).
        const bool do_wpf = (KIND != KIND_WGRAD) && p.wpf > 0;
        TileDesc tn{};
        int n_stage_next = 0;
        if (do_wpf && tile + tile_stride < n_tiles) {
          tn = decode_tile<KIND, CG>(p, st, mp, tile + tile_stride, rank);
          n_stage_next = (tn.k_ext + kBK - 1) / kBK;
        }
        for (int s = 0; s < n_stage; ++s, ++issued) {
          if (do_wpf) {
            int sp = s + p.wpf;
            const TileDesc* tp = &t;
            if (sp >= n_stage) { sp -= n_stage; tp = (sp < n_stage_next) ? &tn : nullptr; }
            if (tp != nullptr && (sp % tp->ntn) == tp->ni) {
              if (!A_MN) tma_prefetch_l2_2d(tp->mapA, sp * kBK, tp->m0);
              else if (p.mn3) tma_prefetch_l2_3d(tp->mapA, 0, sp * kBK, tp->m0 / 64);
            }
          }
          const uint32_t slot = issued % NS;
          mbar_wait(&pipe.empty[slot], ((issued / NS) & 1u) ^ 1u);
          uint8_t* sa = smem + slot * kStage;
          uint8_t* sb = sa + kA;
          uint32_t bar = smem_u32(&pipe.full[slot]);
          if (CG == 2) {
            if (rank == 0) mbar_expect_tx(&pipe.full[slot], 2 * kStage);
            bar = mapa_rank(bar, 0);
          } else {
            mbar_expect_tx(&pipe.full[slot], kStage);
          }
          const int k0 = s * kBK;
          if (!A_MN) tma2d<CG>(sa, t.mapA, k0, t.m0, bar);
          else if (p.mn3) tma3d<CG>(sa, t.mapA, 0, k0, t.m0 / 64, bar);
          else {
#pragma unroll
            for (int j = 0; j < kTM / 64; ++j) tma2d<CG>(sa + j * 8192, t.mapA, t.m0 + j * 64, k0, bar);
          }
          if (!B_MN) tma2d<CG>(sb, t.mapB, k0, t.k_base + t.nb0, bar);
          else if (p.mn3) tma3d<CG>(sb, t.mapB, 0, k0, t.nb0 / 64, bar);
          else {
#pragma unroll
            for (int j = 0; j < kBRows / 64; ++j) tma2d<CG>(sb + j * 8192, t.mapB, t.nb0 + j * 64, k0, bar);
          }
        }
      }
      // tail: every commit that releases one of this CTA's slots has landed before the CTA may exit (for CG = 2 they are
      // remote arrivals from the leader's tensor core)
      for (uint32_t k = issued; k < issued + NS; ++k) mbar_wait(&pipe.empty[k % NS], ((k / NS) & 1u) ^ 1u);
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (leader CTA of a pair only) ----------------
    if (rank == 0) {
      const uint32_t id = idesc_bf16(kTM * CG, kTN, A_MN, B_MN);
      // SWIZZLE_128B operand descriptors (tma.cuh): K-major LBO field 1 / SBO 1024, 32 B per K step;
      // MN-major LBO 8192 (next 64 units) / SBO 1024 (next 8 k-rows), 2048 B per K step
      constexpr uint32_t lbo_a = A_MN ? 8192u : 16u, lbo_b = B_MN ? 8192u : 16u;
      constexpr uint32_t adv_a = A_MN ? (2048u >> 4) : (32u >> 4), adv_b = B_MN ? (2048u >> 4) : (32u >> 4);
      uint32_t sc = 0, gi = 0;
      for (int tile = first_tile; tile < n_tiles; tile += tile_stride) {
        const TileDesc t = decode_tile<KIND, CG>(p, st, mp, tile, rank);
        const int n_stage = (t.k_ext + kBK - 1) / kBK;
        if (n_stage == 0) continue;
        const uint32_t ab = gi & 1u;
#ifdef MCPC_DEBUG_BUILD
        const bool tl = p.dbg_buf != nullptr && blockIdx.x == 0 && lane == 0 && gi < 16;
        if (tl) {
          p.dbg_buf[KIND * 256 + gi * 8 + 0] = clock64();
          unsigned long long ns;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
          p.dbg_buf[KIND * 256 + gi * 8 + 7] = (long long)ns;
        }
#endif
        mbar_wait(&pipe.acc_empty[ab], ((gi >> 1) & 1u) ^ 1u);
        fence_after_sync();
#ifdef MCPC_DEBUG_BUILD
        if (tl) p.dbg_buf[KIND * 256 + gi * 8 + 1] = clock64();
#endif
        for (int s = 0; s < n_stage; ++s, ++sc) {
          const uint32_t slot = sc % NS;
          mbar_wait(&pipe.full[slot], (sc / NS) & 1u);
          fence_after_sync();
          const uint64_t ad0 = smem_desc_sw128(smem_base + slot * kStage, lbo_a, 1024u);
          const uint64_t bd0 = smem_desc_sw128(smem_base + slot * kStage + kA, lbo_b, 1024u);
          if (elect1()) {
#pragma unroll
            for (int ks = 0; ks < kBK / 16; ++ks)
              mma_bf16_cg<CG>(tmem + ab * kTN, ad0 + (uint64_t)(ks * adv_a), bd0 + (uint64_t)(ks * adv_b), id, s > 0 || ks > 0);
            mma_commit_cg<CG>(&pipe.empty[slot]);
            if (s == n_stage - 1) mma_commit_cg<CG>(&pipe.acc_full[ab]);
          }
          __syncwarp();
        }
#ifdef MCPC_DEBUG_BUILD
        if (tl) p.dbg_buf[KIND * 256 + gi * 8 + 2] = clock64();      // last MMA of the tile ISSUED
#endif
        ++gi;
      }
    }
  } else {
    // ---------------- epilogue warps 2..9: TMEM lane quarter = warp % 4, chain half = (warp - 2) / 4 ----------------
    EpiPos ep;
    ep.q = warp & 3;
    ep.ew = warp - 2;
    ep.h = ep.ew >> 2;
    ep.lane = lane;
    constexpr int kDepthU = (int)(stg_bytes(CG, KIND_UPDATE) / kUpdSlot), kDepthP = (int)(stg_bytes(CG, KIND_PREDICT) / kPredSlot);
    const uint32_t stg = smem_base + NS * kStage + (uint32_t)ep.ew * stg_bytes(CG, KIND);
    const uint32_t acc_empty_addr = (CG == 2) ? mapa_rank(smem_u32(&pipe.acc_empty[0]), 0) : smem_u32(&pipe.acc_empty[0]);
    const bool run_epi = MCPC_EPI_MODE(p) != 1;            // debug mode 1: mainloop-only rate, results are garbage
    uint32_t gi = 0;
    int tile = first_tile;
    TileDesc t{};
    PredCtx pc;
    UpdCtx uc;
    // the inputs of a tile's first chunks are staged BEFORE the wait for its accumulator: they fly during the mainloop
    auto begin_tile = [&]() {
      t = decode_tile<KIND, CG>(p, st, mp, tile, rank);
      if (!run_epi) return;
      if (KIND == KIND_PREDICT) {
        predict_ctx(p, st, t, ep, stg, pc);
        predict_prestage<kDepthP>(pc);
      } else if (KIND == KIND_UPDATE) {
        update_ctx(p, st, t, ep, stg, uc);
        update_prestage<kDepthU>(uc);
      }
    };
    if (tile < n_tiles) begin_tile();
    while (tile < n_tiles) {
      const bool has_gemm = t.k_ext > 0;
      const uint32_t ab = gi & 1u;
#ifdef MCPC_DEBUG_BUILD
      const bool tl = p.dbg_buf != nullptr && blockIdx.x == 0 && lane == 0 && ep.ew == 0 && gi < 16 && has_gemm;
      if (tl) p.dbg_buf[KIND * 256 + gi * 8 + 3] = clock64();
#endif
      if (has_gemm) {
        mbar_wait(&pipe.acc_full[ab], (gi >> 1) & 1u);
        fence_after_sync();
      }
#ifdef MCPC_DEBUG_BUILD
      if (tl) p.dbg_buf[KIND * 256 + gi * 8 + 4] = clock64();
#endif
      const uint32_t acc = tmem + ((uint32_t)(ep.q * 32) << 16) + ab * kTN + ep.h * (kTN / 2);
      if (!run_epi) {
      } else if (KIND == KIND_PREDICT) {
        epilogue_predict<kDepthP>(p, st, pc, acc, has_gemm, (tile * CG + rank) * kEpiWarps + ep.ew);
      } else if (KIND == KIND_UPDATE) {
        if (SPEC == 1) {
          if (uc.kind == MCPC_ACT_TANH) epilogue_update<1, MCPC_ACT_TANH, kDepthU>(p, st, uc, acc, has_gemm);
          else if (uc.kind == MCPC_ACT_RELU) epilogue_update<1, MCPC_ACT_RELU, kDepthU>(p, st, uc, acc, has_gemm);
          else epilogue_update<1, MCPC_ACT_IDENTITY, kDepthU>(p, st, uc, acc, has_gemm);
        } else {
          epilogue_update<0, -1, kDepthU>(p, st, uc, acc, has_gemm);
        }
      } else {
        epilogue_wgrad(p, t, acc, ep);
      }
      if (has_gemm) {
#ifdef MCPC_DEBUG_BUILD
        if (tl) p.dbg_buf[KIND * 256 + gi * 8 + 5] = clock64();
#endif
        fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(acc_empty_addr + ab * 8);
          else mbar_arrive(&pipe.acc_empty[ab]);
        }
#ifdef MCPC_DEBUG_BUILD
        if (tl) p.dbg_buf[KIND * 256 + gi * 8 + 6] = clock64();
#endif
        ++gi;
      }
      tile += tile_stride;
      if (tile < n_tiles) begin_tile();
    }
  }
  fence_before_sync();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) tmem_dealloc_cg<CG>(tmem, 512);
#ifdef MCPC_DEBUG_BUILD
  if (p.dbg_buf != nullptr && blockIdx.x == 0 && tid == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.dbg_buf[KIND * 256 + 251] = (long long)ns;
  }
#endif
}

// fp32 [rows][cols] -> bf16 [rows][pitch] (pitch = cols rounded up to 8: TMA needs 16-byte row strides); pad columns zero
__global__ void to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int cols, int pitch) {
  const size_t total = (size_t)rows * pitch;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / pitch), c = (int)(i % pitch);
    dst[i] = __float2bfloat16(c < cols ? src[(size_t)r * cols + c] : 0.0f);
  }
}

// fp32 [rows][cols] -> bf16 TRANSPOSED [cols][pitch_t] (pitch_t = rows rounded up to 8), 32 x 32 tiles through shared memory
__global__ void to_bf16_transposed_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int cols,
                                          int pitch_t) {
  __shared__ float tile[32][33];
  const int tiles_c = (cols + 31) / 32, tiles_r = (pitch_t + 31) / 32;
  for (int t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const int r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int r = r0 + i, c = c0 + threadIdx.x;
      tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int c = c0 + i, r = r0 + threadIdx.x;
      if (c < cols && r < pitch_t) dst[(size_t)c * pitch_t + r] = __float2bfloat16(tile[threadIdx.x][i]);
    }
    __syncthreads();
  }
}

// act(x) of the initial latents into ring slot 0
__global__ void init_act_kernel(WideParams p) {
  const NetDev& nd = p.net;
  const size_t total = (size_t)p.B * nd.SD;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / nd.SD), u = (int)(i % nd.SD);
    int l = 0;
    while (u >= nd.off[l + 1]) ++l;
    const int k = u - nd.off[l];
    p.act_l[l][(size_t)row * p.apitch[l] + k] = __float2bfloat16(act_w(nd.act[l], p.x[l][(size_t)row * nd.dims[l] + k]));
  }
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
inline int pad8w(int v) { return (v + 7) & ~7; }

struct WideLayout {
  size_t wb_off[kMaxL + 1], wbt_off[kMaxL + 1], act_off[kMaxL], gb_off[kMaxL + 1], in_off, part_off, total;
  int apitch[kMaxL], gpitch[kMaxL + 1], in_pitch, n_part, Bpad, S, cg;
};

struct WideKnobs {
  int cg, slots, ctas, nospec, epi_pf, cs, wpf, pdl;
#ifdef MCPC_DEBUG_BUILD
  int skip_epi;
#endif
};

// Test hooks (documented in DESIGN.md): MCPC_WIDE_CG = 1 | 2 (CTA pairs off / on), MCPC_WIDE_SLOTS (steps per weight-gradient
// launch), MCPC_WIDE_CTAS (few persistent CTAs => many tiles per CTA), MCPC_TC_NOSPEC (generic update instantiation).
WideKnobs wide_knobs() {
  WideKnobs k{};
  k.cg = 2;
  if (const char* env = getenv("MCPC_WIDE_CG")) k.cg = (atoi(env) == 1) ? 1 : 2;
  k.slots = 0;
  if (const char* env = getenv("MCPC_WIDE_SLOTS")) k.slots = atoi(env);
  k.ctas = 0;
  if (const char* env = getenv("MCPC_WIDE_CTAS")) k.ctas = atoi(env);
  k.nospec = getenv("MCPC_TC_NOSPEC") != nullptr ? 1 : 0;
  k.cs = 0;          // measured on C5: no effect (0.724 / 0.710 ms per step without / with)
  if (const char* env = getenv("MCPC_WIDE_CS")) k.cs = atoi(env) != 0 ? 1 : 0;
  k.pdl = 0;         // programmatic dependent launch: measured on C5 (T=100) 0.669 ms per step without, 0.687 with
  if (const char* env = getenv("MCPC_WIDE_PDL")) k.pdl = atoi(env) != 0 ? 1 : 0;
  k.wpf = 0;         // measured on C5: 0.719 ms per step without the weight-panel prefetch, 0.753-0.756 with 6 / 12 / 24 stages
  if (const char* env = getenv("MCPC_WIDE_WPF")) k.wpf = atoi(env);
  k.epi_pf = 0;      // measured on C5: 0.734 ms/step with the L2 prefetch of the epilogue inputs, 0.700 without
  if (const char* env = getenv("MCPC_WIDE_EPIPF")) k.epi_pf = atoi(env) != 0 ? 1 : 0;
#ifdef MCPC_DEBUG_BUILD
  k.skip_epi = 0;
  if (const char* env = getenv("MCPC_WIDE_EPI_MODE")) k.skip_epi = atoi(env);
#endif
  return k;
}

int wide_layout(const NetDev& nd, int B, int n_steps, WideLayout* lay) {
  const WideKnobs kn = wide_knobs();
  const int n_lin = nd.L + (nd.d_out > 0 ? 1 : 0);
  lay->cg = kn.cg;
  lay->Bpad = (B + 63) & ~63;
  size_t slot_bytes = 0;                                 // one ring slot of all bf16 operand blocks
  int widest = 8;
  for (int l = 0; l < nd.L; ++l) {
    lay->apitch[l] = pad8w(nd.dims[l]);
    lay->gpitch[l] = pad8w(nd.dims[l]);
    slot_bytes += (size_t)lay->Bpad * (lay->apitch[l] + lay->gpitch[l]) * 2;
    if (lay->apitch[l] > widest) widest = lay->apitch[l];
  }
  lay->gpitch[nd.L] = pad8w(nd.d_out > 0 ? nd.d_out : 8);
  if (nd.d_out > 0) slot_bytes += (size_t)lay->Bpad * lay->gpitch[nd.L] * 2;
  if (lay->gpitch[nd.L] > widest) widest = lay->gpitch[nd.L];
  {
    // the epilogues index global memory with 32-bit element offsets
    if ((size_t)(lay->Bpad + kTN) * (size_t)widest >= ((size_t)1 << 31)) {
      set_error("bf16 streaming path: B * layer width = %zu elements exceeds the 2^31 the kernels index; shard the batch",
                (size_t)B * (size_t)widest);
      return MCPC_ERR_UNSUPPORTED;
    }
  }
  size_t o = 0;
  lay->in_pitch = pad8w(nd.d_in > 0 ? nd.d_in : 8);
  lay->wb_off[0] = o;                                     // W_0 (non-zero inputs only; no transposed copy: nothing flows into them)
  o += align256((size_t)nd.dims[0] * lay->in_pitch * 2);
  lay->wbt_off[0] = 0;
  for (int l = 1; l < n_lin; ++l) {
    const int d_o = (l == nd.L) ? nd.d_out : nd.dims[l];
    lay->wb_off[l] = o;
    o += align256((size_t)d_o * pad8w(nd.dims[l - 1]) * 2);
    lay->wbt_off[l] = o;                                  // transposed copy: the back-projection reads W^T K-major
    o += align256((size_t)nd.dims[l - 1] * pad8w(d_o) * 2);
  }
  // ring of S slots of the bf16 operands: S steps of the accumulate window are contracted by ONE weight-gradient launch.
  // Default 4: with C5's 2048 chains both operands of a layer (134 MB) stay L2-resident during the launch.
  int S = 4;
  const size_t budget = (size_t)4 << 30;
  while (S > 1 && (size_t)S * slot_bytes > budget) --S;
  if (kn.slots >= 1 && kn.slots <= 64) S = kn.slots;
  if (S > n_steps) S = n_steps;
  lay->S = S;
  for (int l = 0; l < nd.L; ++l) {
    lay->act_off[l] = o;
    o += align256((size_t)S * lay->Bpad * lay->apitch[l] * 2 + 65536);
  }
  for (int l = 0; l < n_lin; ++l) {
    lay->gb_off[l] = o;
    o += align256((size_t)S * lay->Bpad * lay->gpitch[l] * 2 + 65536);
  }
  lay->in_off = o;                                        // bf16 copy of the inputs in every ring slot (sized whether or not
  o += align256((size_t)S * lay->Bpad * lay->in_pitch * 2 + 65536);   // the call has inputs: the workspace depends on the net only)
  const int ntn = (B + kTN - 1) / kTN;
  int n_part = 0;
  for (int l = 0; l < n_lin; ++l) {
    const int d_o = (l == nd.L) ? nd.d_out : nd.dims[l];
    n_part += ((d_o + kTM * lay->cg - 1) / (kTM * lay->cg)) * ntn * lay->cg * kEpiWarps;
  }
  lay->n_part = n_part;
  lay->part_off = o;
  o += align256((size_t)n_steps * n_part * 2 * sizeof(float));
  lay->total = o + 512;
  return MCPC_OK;
}

template <int KIND, int SPEC, int CG>
int launch_wide(const WideParams& p, const StepArgs& st, const WideMaps& mp, int n_tiles, int max_ctas, size_t smem_bytes,
                cudaStream_t stream) {
  if (n_tiles <= 0) return MCPC_OK;
  cudaLaunchConfig_t cfg{};
  int grid = n_tiles * CG < max_ctas ? n_tiles * CG : max_ctas;
  if (CG == 2) grid &= ~1;
  if (grid < CG) grid = CG;
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 2 : 1;
  MCPC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, wide_kernel<KIND, SPEC, CG>, p, st, mp));
  count_launch();
  return MCPC_OK;
}

template <int CG>
int run_wide(const NetDev& nd, const McpcIO* io, const McpcOpts* o, WideParams& p, const WideLayout& lay, const WideKnobs& kn,
             const __nv_bfloat16* const* Wb, const __nv_bfloat16* const* WbT, __nv_bfloat16* in_l, cudaStream_t stream) {
  const int B = p.B;
  const int n_lin = nd.L + (nd.d_out > 0 ? 1 : 0);
  // tile tables: (128 * CG) units x 256 chains per (pair) tile
  const int ntn = (B + kTN - 1) / kTN;
  int t = 0;
  for (int l = 0; l <= nd.L; ++l) {
    p.tP_first[l] = t;
    if (l < nd.L || nd.d_out > 0) t += (((l == nd.L ? nd.d_out : nd.dims[l]) + kTM * CG - 1) / (kTM * CG)) * ntn;
  }
  p.tP_first[nd.L + 1] = t;
  const int n_predict = t;
  t = 0;
  for (int l = 0; l < nd.L; ++l) {
    p.tU_first[l] = t;
    t += ((nd.dims[l] + kTM * CG - 1) / (kTM * CG)) * ntn;
  }
  p.tU_first[nd.L] = t;
  const int n_update = t;
  t = 0;
  bool any_grad = false;
  for (int l = 0; l <= nd.L; ++l) {
    p.tW_first[l] = t;
    const bool is_out = (l == nd.L);
    if (is_out && (nd.d_out == 0 || !nd.top_has_grad)) continue;
    if (p.gW[l] == nullptr) continue;
    const int d_o = is_out ? nd.d_out : nd.dims[l], d_i = (l == 0) ? p.d_in_eff : nd.dims[l - 1];
    if (d_i == 0) continue;                                                             // zero inputs: gW_0 receives nothing
    t += ((d_i + kTM * CG - 1) / (kTM * CG)) * ((d_o + kTN - 1) / kTN);                // M = input units, N = output units
  }
  p.tW_first[nd.L + 1] = t;
  const int n_wgrad = t;
  for (int l = 0; l <= nd.L; ++l) any_grad = any_grad || p.gW[l] != nullptr || p.gb[l] != nullptr;
  if (p.n_part != n_predict * CG * kEpiWarps) {
    set_error("internal: partial-slot count mismatch (%d vs %d)", p.n_part, n_predict * CG * kEpiWarps);
    return MCPC_ERR_INVALID;
  }

  bool any_traj = io->traj_out != nullptr;
  for (int l = 0; l < nd.L; ++l) any_traj = any_traj || io->traj_x[l] != nullptr;
  const int traj_every = any_traj ? (o->traj_every > 0 ? o->traj_every : 1) : 0;

  // tensor maps: one per layer block (base offset = the block's first column) so TMA zero-fills past its extent;
  // the row (chain) axis spans all S ring slots
  WideMaps mp;
  bool mn3 = (nd.d_out % 64 == 0);
  for (int l = 0; l < nd.L; ++l) mn3 = mn3 && (nd.dims[l] % 64 == 0);
  if (p.d_in_eff > 0) mn3 = mn3 && (p.d_in_eff % 64 == 0);
  p.mn3 = mn3 ? 1 : 0;
  const uint64_t rows = (uint64_t)lay.S * lay.Bpad;
  constexpr int kBRows = kTN / CG;
  int rc = MCPC_OK;
  for (int l = 0; l < nd.L; ++l) {
    rc = make_tmap_bf16(&mp.act_k[l], p.act_l[l], nd.dims[l], rows, p.apitch[l], 64, kBRows);                  // predict B
    if (rc == MCPC_OK)
      rc = mn3 ? make_tmap_bf16_mn3(&mp.act_mn[l], p.act_l[l], nd.dims[l], rows, p.apitch[l], 64, kTM / 64)       // wgrad A
               : make_tmap_bf16(&mp.act_mn[l], p.act_l[l], nd.dims[l], rows, p.apitch[l], 64, 64);
    if (rc != MCPC_OK) return rc;
  }
  if (p.d_in_eff > 0) {
    rc = make_tmap_bf16(&mp.in_k, in_l, p.d_in_eff, rows, lay.in_pitch, 64, kBRows);                            // predict B, Linear 0
    if (rc == MCPC_OK)
      rc = mn3 ? make_tmap_bf16_mn3(&mp.in_mn, in_l, p.d_in_eff, rows, lay.in_pitch, 64, kTM / 64)                // wgrad A, Linear 0
               : make_tmap_bf16(&mp.in_mn, in_l, p.d_in_eff, rows, lay.in_pitch, 64, 64);
    if (rc == MCPC_OK) rc = make_tmap_bf16(&mp.w_k[0], Wb[0], p.d_in_eff, nd.dims[0], lay.in_pitch, 64, kTM);    // predict A
    if (rc != MCPC_OK) return rc;
  }
  for (int l = 0; l < n_lin; ++l) {
    const int d_o = (l == nd.L) ? nd.d_out : nd.dims[l];
    rc = make_tmap_bf16(&mp.gb_k[l], p.gb_l[l], d_o, rows, p.gpitch[l], 64, kBRows);                            // update B
    if (rc == MCPC_OK)
      rc = mn3 ? make_tmap_bf16_mn3(&mp.gb_mn[l], p.gb_l[l], d_o, rows, p.gpitch[l], 64, kBRows / 64)             // wgrad B
               : make_tmap_bf16(&mp.gb_mn[l], p.gb_l[l], d_o, rows, p.gpitch[l], 64, 64);
    if (rc == MCPC_OK && l >= 1) {
      const int d_i = nd.dims[l - 1], wp = pad8w(d_i);
      rc = make_tmap_bf16(&mp.w_k[l], Wb[l], d_i, d_o, wp, 64, kTM);                                             // predict A
      if (rc == MCPC_OK) rc = make_tmap_bf16(&mp.w_kt[l], WbT[l], d_o, d_i, pad8w(d_o), 64, kTM);                // update A
    }
    if (rc != MCPC_OK) return rc;
  }
  // fp32 prefetch views (optional: skipped when a base or a row stride is not 16-byte aligned)
  p.pf_x = p.pf_t = 0;
  if (kn.epi_pf) {
    bool ok = true;
    for (int l = 0; l < nd.L && ok; ++l)
      ok = (nd.dims[l] % 4 == 0) && (reinterpret_cast<uintptr_t>(p.x[l]) % 16 == 0) &&
           make_tmap_f32(&mp.x32[l], p.x[l], nd.dims[l], B, nd.dims[l], kTM, kTN) == MCPC_OK;
    p.pf_x = ok ? 1 : 0;
    p.pf_t = (p.target != nullptr && nd.top >= MCPC_TOP_GAUSS && nd.d_out % 4 == 0 &&
              reinterpret_cast<uintptr_t>(p.target) % 16 == 0 &&
              make_tmap_f32(&mp.tgt, p.target, nd.d_out, B, nd.d_out, kTM, kTN) == MCPC_OK) ? 1 : 0;
  }
  auto smem_of = [](int kind) {
    return (size_t)n_stages(CG, kind) * stage_bytes(CG) + (size_t)kEpiWarps * stg_bytes(CG, kind) + 1024;
  };
  const size_t smem_p = smem_of(KIND_PREDICT), smem_u = smem_of(KIND_UPDATE), smem_w = smem_of(KIND_WGRAD);
  MCPC_CUDA_CHECK(cudaFuncSetAttribute(wide_kernel<KIND_PREDICT, 0, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
  MCPC_CUDA_CHECK(cudaFuncSetAttribute(wide_kernel<KIND_UPDATE, 0, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u));
  MCPC_CUDA_CHECK(cudaFuncSetAttribute(wide_kernel<KIND_UPDATE, 1, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u));
  MCPC_CUDA_CHECK(cudaFuncSetAttribute(wide_kernel<KIND_WGRAD, 0, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
  int n_sm = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (CG == 2) {
      // persistent pairs: never more CTAs than clusters that can be resident at once (a GPC with an odd number of free
      // SMs leaves one unpaired)
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(n_sm & ~1));
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = smem_u;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int n_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&n_clusters, wide_kernel<KIND_UPDATE, 0, CG>, &cfg) == cudaSuccess && n_clusters > 0 &&
          2 * n_clusters < n_sm)
        n_sm = 2 * n_clusters;
    }
    if (kn.ctas >= 1 && kn.ctas <= n_sm) n_sm = kn.ctas;          // testing hook: few persistent CTAs => many tiles per CTA
    if (CG == 2 && n_sm < 2) n_sm = 2;
  }

  // rows B..Bpad of every ring slot enter the weight-gradient contraction: they must be zero in both operands
  if (lay.Bpad > B) {
    for (int s = 0; s < lay.S; ++s) {
      for (int l = 0; l < nd.L; ++l)
        MCPC_CUDA_CHECK(cudaMemsetAsync(p.act_l[l] + ((size_t)s * lay.Bpad + B) * p.apitch[l], 0,
                                        (size_t)(lay.Bpad - B) * p.apitch[l] * 2, stream));
      for (int l = 0; l < n_lin; ++l)
        MCPC_CUDA_CHECK(cudaMemsetAsync(p.gb_l[l] + ((size_t)s * lay.Bpad + B) * p.gpitch[l], 0,
                                        (size_t)(lay.Bpad - B) * p.gpitch[l] * 2, stream));
    }
  }
#ifdef MCPC_DEBUG_BUILD
  const bool timing = getenv("MCPC_WIDE_TIMING") != nullptr;      // debug only: allocates + synchronises
  if (timing) {
    cudaMalloc(&p.dbg_buf, 3 * 256 * sizeof(long long));
    cudaMemsetAsync(p.dbg_buf, 0, 3 * 256 * sizeof(long long), stream);
  }
#endif
  init_act_kernel<<<1184, 256, 0, stream>>>(p);
  count_launch();
  double b1p = pow(o->adam_beta1, (double)o->adam_step0), b2p = pow(o->adam_beta2, (double)o->adam_step0);
  bool spec_update = !any_traj && o->update_x && o->optimizer == MCPC_OPT_SGD && o->noise_mode == MCPC_NOISE_PHILOX &&
                     (o->chain_offset & 3) == 0;                       // what wide_kernel<KIND_UPDATE, 1> assumes
  for (int l = 0; l < nd.L; ++l) spec_update = spec_update && io->x_grad[l] == nullptr;
  if (kn.nospec) spec_update = false;                                 // testing hook: the generic instantiation
  int used = 0;                                                       // accumulate steps waiting in ring slots [0, used)
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
    st.slot = used;
    rc = launch_wide<KIND_PREDICT, 0, CG>(p, st, mp, n_predict, n_sm, smem_p, stream);
    if (rc != MCPC_OK) return rc;
    // the weight update reads G (this step's errors) and act(x) of the state BEFORE the update: it runs between the two,
    // once the ring is full or the window ends
    if (acc) {
      ++used;
      const bool window_ends = (ts + 1 >= o->save_end) || (ts + 1 >= o->n_steps);
      if (used == lay.S || window_ends) {
        st.k_rows = used * lay.Bpad;
        rc = launch_wide<KIND_WGRAD, 0, CG>(p, st, mp, n_wgrad, n_sm, smem_w, stream);
        if (rc != MCPC_OK) return rc;
        used = 0;
      }
    }
    st.slot_next = used;
    if (spec_update) rc = launch_wide<KIND_UPDATE, 1, CG>(p, st, mp, n_update, n_sm, smem_u, stream);
    else rc = launch_wide<KIND_UPDATE, 0, CG>(p, st, mp, n_update, n_sm, smem_u, stream);
    if (rc != MCPC_OK) return rc;
  }
  MCPC_CUDA_CHECK(cudaGetLastError());
#ifdef MCPC_DEBUG_BUILD
  if (timing) {
    static long long h[3 * 256];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, p.dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg_buf);
    const char* names[3] = {"predict", "update", "wgrad"};
    fprintf(stderr, "[wide timeline] last step: predict exit -> %s entry %lld ns; predict entry -> update exit %lld ns\n",
            h[2 * 256 + 250] > h[0 * 256 + 251] && h[2 * 256 + 250] < h[1 * 256 + 250] ? "wgrad" : "update",
            (h[2 * 256 + 250] > h[0 * 256 + 251] && h[2 * 256 + 250] < h[1 * 256 + 250] ? h[2 * 256 + 250] : h[1 * 256 + 250]) - h[0 * 256 + 251],
            h[1 * 256 + 251] - h[0 * 256 + 250]);
    for (int k = 0; k < 3; ++k) {
      const long long t0 = h[k * 256 + 0];
      if (t0 == 0) continue;
      fprintf(stderr, "[wide timeline] %s, CTA 0, last launch (cycles since the MMA warp's first wait)\n", names[k]);
      for (int g = 0; g < 16 && h[k * 256 + g * 8 + 1] != 0; ++g) {
        const long long* r = h + k * 256 + g * 8;
        fprintf(stderr, "  tile %2d: mma wait-acc %7lld..%7lld issue-done %7lld | epi wait %7lld..%7lld done %7lld arrive %7lld\n", g,
                r[0] - t0, r[1] - t0, r[2] - t0, r[3] - t0, r[4] - t0, r[5] - t0, r[6] - t0);
      }
      fprintf(stderr, "  CTA 0 alive %lld ns (entry -> exit)\n", h[k * 256 + 251] - h[k * 256 + 250]);
      int last = 0;
      while (last + 1 < 16 && h[k * 256 + (last + 1) * 8 + 1] != 0) ++last;
      if (last > 0) {
        const double cyc = (double)(h[k * 256 + last * 8 + 0] - h[k * 256 + 0]);
        const double ns = (double)(h[k * 256 + last * 8 + 7] - h[k * 256 + 7]);
        fprintf(stderr, "  SM clock over tiles 0..%d: %.0f cycles / %.0f ns = %.3f GHz\n", last, cyc, ns, cyc / ns);
      }
    }
  }
#endif
  if (io->energy != nullptr || io->loss != nullptr) return launch_reduce_partials(p.partials, o->n_steps, p.n_part, io->energy, io->loss, stream);
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
  if (io->save_g != nullptr) {
    set_error("bf16 streaming path accumulates the weight update itself (McpcIO.gW/gb); save_g/save_f are not used");
    return MCPC_ERR_INVALID;
  }
  const WideKnobs kn = wide_knobs();
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
  p.Bpad = lay.Bpad;
  const int n_lin = nd.L + (nd.d_out > 0 ? 1 : 0);
  for (int l = 0; l <= nd.L; ++l) {
    p.b[l] = io->b[l];
    p.gW[l] = io->gW[l];
    p.gb[l] = io->gb[l];
    p.gpitch[l] = lay.gpitch[l];
    p.gb_l[l] = (l < nd.L || nd.d_out > 0) ? reinterpret_cast<__nv_bfloat16*>(wsb + lay.gb_off[l]) : nullptr;
  }
  for (int l = 0; l < nd.L; ++l) {
    p.apitch[l] = lay.apitch[l];
    p.act_l[l] = reinterpret_cast<__nv_bfloat16*>(wsb + lay.act_off[l]);
  }
  p.partials = reinterpret_cast<float*>(wsb + lay.part_off);
  p.n_part = lay.n_part;
  const __nv_bfloat16* Wb[kMaxL + 1] = {};
  const __nv_bfloat16* WbT[kMaxL + 1] = {};
  __nv_bfloat16* in_l = nullptr;
  p.d_in_eff = 0;
  if (io->inputs != nullptr && nd.d_in > 0) {
    // non-zero inputs: Linear_0 becomes a GEMM like every other Linear -- its B operand is a bf16 copy of the inputs,
    // replicated in every ring slot (the weight-gradient launch contracts over the slots), rows B..Bpad zero
    p.d_in_eff = nd.d_in;
    in_l = reinterpret_cast<__nv_bfloat16*>(wsb + lay.in_off);
    __nv_bfloat16* wb0 = reinterpret_cast<__nv_bfloat16*>(wsb + lay.wb_off[0]);
    const size_t n0 = (size_t)nd.dims[0] * lay.in_pitch;
    to_bf16_kernel<<<(int)((n0 + 1023) / 1024 < 1184 ? (n0 + 1023) / 1024 : 1184), 256, 0, stream>>>(io->W[0], wb0, nd.dims[0], nd.d_in,
                                                                                                  lay.in_pitch);
    count_launch();
    Wb[0] = wb0;
    const size_t ni = (size_t)B * lay.in_pitch;
    for (int s = 0; s < lay.S; ++s) {
      __nv_bfloat16* dst = in_l + (size_t)s * lay.Bpad * lay.in_pitch;
      to_bf16_kernel<<<(int)((ni + 1023) / 1024 < 1184 ? (ni + 1023) / 1024 : 1184), 256, 0, stream>>>(io->inputs, dst, B, nd.d_in, lay.in_pitch);
      count_launch();
      if (lay.Bpad > B)
        MCPC_CUDA_CHECK(cudaMemsetAsync(dst + (size_t)B * lay.in_pitch, 0, (size_t)(lay.Bpad - B) * lay.in_pitch * 2, stream));
    }
  }
  for (int l = 1; l < n_lin; ++l) {
    __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(wsb + lay.wb_off[l]);
    __nv_bfloat16* wbt = reinterpret_cast<__nv_bfloat16*>(wsb + lay.wbt_off[l]);
    const int rows = (l == nd.L) ? nd.d_out : nd.dims[l], cols = nd.dims[l - 1];
    const size_t n = (size_t)rows * pad8w(cols);
    to_bf16_kernel<<<(int)((n + 1023) / 1024 < 1184 ? (n + 1023) / 1024 : 1184), 256, 0, stream>>>(io->W[l], wb, rows, cols, pad8w(cols));
    const int n_t = ((cols + 31) / 32) * ((pad8w(rows) + 31) / 32);
    to_bf16_transposed_kernel<<<n_t < 2368 ? n_t : 2368, dim3(32, 8), 0, stream>>>(io->W[l], wbt, rows, cols, pad8w(rows));
    count_launch(2);
    Wb[l] = wb;
    WbT[l] = wbt;
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
  p.cs = kn.cs;
  p.wpf = kn.wpf < 0 ? 0 : kn.wpf;
  p.pdl = kn.pdl;
#ifdef MCPC_DEBUG_BUILD
  p.skip_epilogue = kn.skip_epi;
#endif
  return lay.cg == 2 ? run_wide<2>(nd, io, o, p, lay, kn, Wb, WbT, in_l, stream)
                     : run_wide<1>(nd, io, o, p, lay, kn, Wb, WbT, in_l, stream);
}

}  // namespace mcpc
