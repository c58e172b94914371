// mcpc_infer, MCPC_PREC_BF16: persistent tensor-core Langevin / PC inference kernel for sm_100a.
//
// One CTA owns NR chains (batch rows) for ALL n_steps steps.  The contractions run on the 5th-gen tensor
// cores in "swapped" orientation -- units on the 128-lane M axis, chains on the N axis -- so a small
// batch tile still fills the MMA:
//     phase A   mu_l^T  [128 units x NR] = W_l tile [128 x K]   . act(x_{l-1})^T      (A K-major)
//     phase B   bp_l^T  [128 units x NR] += (W_{l+1} tile)^T    . G_{l+1}^T           (A MN-major view of the SAME tile)
// (the chain-side B operands of both phases are MN-major, see st_chains_bf16)
// Weight tiles (bf16, canonical no-swizzle layout, packed once per launch by pack_weights_kernel) are
// resident in shared memory as far as they fit; the rest streams through a 2-slot ring with bulk async
// copies (UBLKCP) every step.  The latents never leave the SM: the fp32 master copy of x, the fp32
// own-layer error and all accumulators live in TMEM (lane = unit, column = chain), the bf16 operand
// copies in shared memory.
//
// Warp roles are listed at the kernel (8 tile-epilogue warps, 8 update warps, 2 MMA issuers, 1 loader, 1 signal).  Per step and
// per weight tile t:  MMA_A(t) -> tile epilogue(t): eps, energy / loss, G (bf16 -> smem)  -> MMA_B(t); then the update
// epilogue applies  x <- x - lr*grad (SGD | Adam)  and  x <- x - lr*noise (Philox)  and re-emits act(x).
// All hand-offs are mbarriers; the tensor pipe never waits on a __syncthreads.
//
// Reference semantics: predictive_coding/pc_trainer.py:733-918 + utils/model.py:35-44 (see infer_rows.cu
// for the fp32 restatement this kernel is validated against).
#include <algorithm>
#include <cstdlib>

#include "mcpc_common.cuh"
#include "philox.cuh"
#include "umma.cuh"

namespace mcpc {
namespace {

using namespace umma;

constexpr int kMaxTiles = 32;         // weight tiles (128 output units each) per network
constexpr int kMaxHT = 8;             // hidden unit tiles (128 latent units each)
constexpr uint32_t kSmemBudget = 225 * 1024;

struct Tile {
  int lin;          // Linear index: 1..L-1 hidden, L = output Linear
  int out_tile;     // which block of 128 output units
  int Kp;           // in-features padded to 16 (phase-A K extent)
  int sbo;          // (Kp/8)*128: byte stride between 8-row groups
  int bytes;        // shared-memory footprint = bytes copied: 128*Kp*2, or only the valid 16-row groups of the partial last
                    // tile of the output Linear (plan_tc)
  int smem_off;     // byte offset of the tile (resident) or of its ring slot (streamed)
  int slot;         // -1 resident, else ring slot
  int h_out;        // hidden unit-tile index of the units this tile predicts (-1 for output tiles)
  size_t gsrc;      // byte offset inside the packed-weights workspace
};

struct TcParams {
  NetDev net;
  int n_hid_tiles, n_out_tiles;
  Tile tiles[kMaxTiles];
  int y_tmem;                   // targets of the output tiles are kept in TMEM (they fit beside the accumulators)
  int adam_tmem;                // Adam's m and v of the latents live in TMEM for the whole call (they fit as well)
  int HT;                       // hidden unit tiles in total
  int h_layer[kMaxHT];          // layer of hidden unit tile h
  int h_index[kMaxHT];          // its index inside the layer
  int h_off[kMaxL + 1];         // first hidden unit tile of layer l
  int ut[kMaxL];                // unit tiles per layer
  int act_off[kMaxL];           // byte offset of layer l's bf16 activation operand [NR x Kp_act[l]]
  int act_kp[kMaxL];            // pad16(d_l)
  int gbuf_off[4];              // byte offsets of the bf16 G operand buffers [NR x 128] (2, or 4 with four tile sub-groups)
  int n_resident_bytes;
  const uint8_t* packed;        // packed bf16 weight tiles
  const float* b[kMaxL + 1];
  float* x[kMaxL];
  float* m[kMaxL];
  float* v[kMaxL];
  float* xgrad[kMaxL];
  float* traj_x[kMaxL];
  float* traj_out;
  int nz_off;                   // pure sampling (instantiation 3): shared-memory offset of the two noise buffers that group T
                                // fills one step ahead for group U ([2][HT][RV][128] fp32), or -1: group U draws its own noise
  int tab_off;                  // byte offset of the MMA issuers' per-tile tables (kTabBytes at the end of the allocation)
  unsigned* ready;              // [n_save] per saved step: += 1 per CTA once its rows of save_g / save_f are written (the
                                // concurrent weight-gradient kernel consumes them while this kernel runs), or nullptr
  const float* mu0;             // fp32 [B, dims[0]]: W_0 inputs + b_0 per chain (non-zero `inputs`), else nullptr
  __nv_bfloat16* save_g;        // bf16 [n_save, B, sg_pitch], layer blocks padded to 8 columns (wgrad_tc.cu)
  __nv_bfloat16* save_f;        // bf16 [n_save, B, sf_pitch]
  int sg_off[kMaxL + 1];
  int sg_pitch, sf_pitch;
  const float* target;
  const float* noise;
  float* partials;
  const float* adam_tab;        // [n_steps][2]: lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t) (adam_table_kernel)
  int B, n_ctas, n_steps, t_begin;
  int optimizer, update_x;
  float lr, adam_eps, one_minus_b1, one_minus_b2, beta2f;
  double lr_d, beta1, beta2, b1_pow0, b2_pow0;
  int noise_mode;
  float noise_scale;
  uint64_t seed, chain_offset;
  int traj_every, save_begin, save_end;
  long long* dbg;               // optional timing trace of CTA 0 (MCPC_TC_TIMING=<first step>), 8 steps x 64 slots
  int dbg_t0;
};

struct Barriers {
  uint64_t w_res;            // resident tiles landed
  uint64_t w_full[2], w_empty[2];
  uint64_t dA_full[8], dA_empty[8];   // prediction accumulators: 4 (8 with four tile sub-groups) in flight with 16-column MMAs, 2 with 32
  uint64_t g_full[4], g_empty[4];
  uint64_t acts_ready[kMaxL];   // act(x_l) / x_l of the coming step are in place (group U -> MMA warp, group T)
  uint64_t bp_ready[kMaxL];     // back-projection into layer l complete (MMA warp -> group U)
  uint64_t g_ready[kMaxL];      // own-layer error of layer l stored in TMEM (group T -> group U)
  uint64_t nz_full[2][kMaxL];   // pure sampling: the noise of layer l for step s is in shared memory buffer s & 1 (group T -> U);
                                // one barrier per buffer: group T is at most one step ahead, so a barrier never completes twice
                                // before group U has waited on it (a single barrier per layer could, and U would wait forever)
  uint64_t out_read;            // read-out only output Linear (no loss gradient), on the steps that record outputs: every output
                                // tile's prediction has READ act(x_{L-1}) (MMA warp -> group U, which overwrites it next)
};

// Per-tile constants of the two MMA issuers (shared memory, built once per launch)
struct __align__(16) MmaA {
  uint64_t ad0, bd0;      // descriptors of the weight tile (K-major) and of the activation operand of its input layer
  int meta;               // in_layer [0,4) | slot + 1 [4,6) | has back-projection [6] | first tile of the step that reads
                          // this layer's activations [7] | K steps of 16 [8,16)
  int pad[3];
};
struct __align__(16) MmaB {
  uint64_t ad0;           // descriptor of the MN-major view of the weight tile (first 128 input units)
  uint32_t a_step;        // descriptor step per 16 output units
  int meta;               // has back-projection [0] | last tile of its Linear [1] | slot + 1 [2,4) | in_layer [4,8) |
                          // K steps (valid output units / 16) [8,12) | unit tiles of the input layer [12,16) | first [16,20) |
                          // first tile of its Linear [20]
};

// The chain-side operands (act(x_l) for the predictions, G for the back-projections: N = chains, K = units) are kept
// MN-major: element (unit u, chain c) at (u/8)*(NR*16) + (c/8)*128 + (u%8)*16 + (c%8)*2, so the chains a thread holds for
// its unit are 8 or 16 CONTIGUOUS bytes and go out as one vector store; a warp covers whole 128-byte rows.  (K-major they
// were 2-byte stores 16 bytes apart: 9 shared-memory wavefronts per store instruction, a quarter of the shared-memory
// data pipe that the tensor core's operand reads need -- ncu r02b.)  Descriptor: LBO = NR*16 (next 8 units), SBO = 128.
template <int N>
__device__ __forceinline__ void st_chains_bf16(uint8_t* dst, const float (&v)[N]) {
  uint32_t w[N / 2];
#pragma unroll
  for (int j = 0; j < N / 2; ++j) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<const uint32_t*>(&h);
  }
  if constexpr (N == 4) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
  } else {
#pragma unroll
    for (int j = 0; j < N / 8; ++j) *reinterpret_cast<uint4*>(dst + j * 128) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
  }
}
// byte offset of (unit u, chain c) inside such an operand
__device__ __forceinline__ uint32_t chain_op_off(int NR, int u, int c) {
  return (uint32_t)(u >> 3) * (uint32_t)(NR * 16) + (uint32_t)(c >> 3) * 128u + (uint32_t)(u & 7) * 16u + (uint32_t)(c & 7) * 2u;
}

constexpr uint32_t kTabBytes = kMaxTiles * (sizeof(MmaA) + sizeof(MmaB));

// One arrival per WARP on the epilogue groups' barriers: every lane has fenced its own writes, __syncwarp orders them
// before lane 0's releasing arrive.
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) umma::mbar_arrive(bar);
}

__device__ __forceinline__ float warp_sum_tc(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// compiled in only for the TRACE instantiation (MCPC_TC_TIMING): the stamps cost ~60 instructions per tile, and the
// epilogue warps are bound by instruction fetch / issue
#define TC_STAMP(cond, ts, idx) do { if constexpr (TRACE) { if (blockIdx.x == 0 && (cond) && (unsigned)((ts) - p.dbg_t0) < 8u) p.dbg[((ts) - p.dbg_t0) * 64 + (idx)] = clock64(); } } while (0)

// fp32 W [rows x cols] (nn.Linear layout) -> bf16 tiles of 128 output units in canonical K-major order
// (every Linear of the network in ONE launch: blockIdx.y picks the Linear -- the call is latency-bound)
struct PackJob {
  const float* W;
  uint8_t* out;
  int rows, cols, Kp, n_tiles;
};
struct PackJobs {
  PackJob job[kMaxL + 1];
};
__global__ void pack_weights_kernel(const PackJobs jobs) {
  const PackJob& J = jobs.job[blockIdx.y];
  const float* __restrict__ W = J.W;
  uint8_t* __restrict__ out = J.out;
  const int rows = J.rows, cols = J.cols, Kp = J.Kp, n_tiles = J.n_tiles;
  const uint32_t sbo = (uint32_t)(Kp / 8) * 128u;
  const int per_tile = 128 * Kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_tiles * per_tile; i += gridDim.x * blockDim.x) {
    const int t = i / per_tile, j = i % per_tile;
    const int r = j / Kp, k = j % Kp;
    const int row = t * 128 + r;
    const float val = (row < rows && k < cols) ? W[(size_t)row * cols + k] : 0.0f;
    *reinterpret_cast<__nv_bfloat16*>(out + (size_t)t * per_tile * 2 + kmajor_off(r, k, 128u, sbo)) = __float2bfloat16(val);
  }
}

// mu_0 of non-zero `inputs` (they are fixed for the whole call, pc_trainer.py:733): mu0[c][u] = b0[u] + sum_k in[c][k] W0[u][k]
// in fp32 -- Linear_0 has no weight tile in the resident kernel, its prediction is a per-chain constant of the launch.
__global__ void __launch_bounds__(256) mu0_kernel(const float* __restrict__ in, const float* __restrict__ W0,
                                                  const float* __restrict__ b0, int B, int d_in, int d0,
                                                  float* __restrict__ mu0) {
  __shared__ float As[16][64 + 1], Ws[16][64 + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int c0 = blockIdx.y * 64, u0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < d_in; k0 += 16) {
    for (int i = tid; i < 64 * 16; i += 256) {
      const int r = i >> 4, k = i & 15;
      As[k][r] = (c0 + r < B && k0 + k < d_in) ? in[(size_t)(c0 + r) * d_in + k0 + k] : 0.0f;
      Ws[k][r] = (u0 + r < d0 && k0 + k < d_in) ? W0[(size_t)(u0 + r) * d_in + k0 + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; w[i] = Ws[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + ty * 4 + i, u = u0 + tx * 4 + j;
      if (c < B && u < d0) mu0[(size_t)c * d0 + u] = acc[i][j] + (b0 != nullptr ? b0[u] : 0.0f);
    }
}

// Adam bias corrections of steps step0+1 .. step0+n (torch.optim.Adam: step_size = lr / (1 - beta1^t), the second
// moment is divided by sqrt(1 - beta2^t)), in double like the host code of torch
__global__ void adam_table_kernel(double beta1, double beta2, double lr, int step0, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double t = (double)(step0 + i + 1);
  out[2 * i] = (float)(lr / (1.0 - pow(beta1, t)));
  out[2 * i + 1] = (float)(1.0 / sqrt(1.0 - pow(beta2, t)));
}

// "the rows of saved step `slot` written by this warp group are in global memory": called by ONE thread after the group's
// end-of-step bar.sync (CTA-scope order over the group's stores), so the gpu-scope release is cumulative over them
__device__ __forceinline__ void signal_saved(unsigned* flag) {
  asm volatile("fence.acq_rel.gpu;\n\tred.release.gpu.global.add.u32 [%0], 1;" ::"l"(flag) : "memory");
}

// one epilogue group finished storing a saved step (called by one thread after the group's end-of-step bar.sync)
__device__ __forceinline__ void saved_step(unsigned* cnt) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(umma::smem_u32(cnt)) : "memory");
}

__device__ __forceinline__ float exp2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// activation / derivative of the tensor-core path (bf16 operands: MUFU tanh.approx is far below their rounding)
__device__ __forceinline__ float act_tc(int kind, float x) {
  return kind == MCPC_ACT_RELU ? fmaxf(x, 0.0f) : (kind == MCPC_ACT_TANH ? tanh_fast(x) : x);
}
__device__ __forceinline__ float dact_tc(int kind, float x, float a) {
  return kind == MCPC_ACT_RELU ? (x > 0.0f ? 1.0f : 0.0f) : (kind == MCPC_ACT_TANH ? fmaf(-a, a, 1.0f) : 1.0f);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float sel4(const float (&n)[4], uint32_t k) {
  return k == 0 ? n[0] : (k == 1 ? n[1] : (k == 2 ? n[2] : n[3]));
}

// NR chains per CTA.  Warp roles (20 warps):
//   warps 8-15  group T: per-tile epilogue (errors of the units a weight tile predicts, G operand for phase B)
//   warps 0-7   group U: latent update of one layer as soon as its back-projection is complete
//   warp 16/17  MMA issuers (one elected thread each runs the whole loop over per-tile tables): predictions / back-projections
//   warp 18     weight-tile loader (bulk async copies for tiles that are not resident)
//   warp 19     signal warp of the opt-in overlapped weight update (idle otherwise)
// Every epilogue thread owns one unit (TMEM lane) and RPT = RV/2 of the chains (two warps per 32-lane quarter).
// RV <= NR is the number of chains the CTA really holds: the epilogues are issue-bound (not the tensor pipe), so a
// batch that would leave SMs idle at RV = NR = 16 runs with RV = 8 on twice as many SMs; the MMA stays N = 16 (the
// minimum at M = 128) and simply carries 8 zero columns.
// Tiles are visited bottom-up (Linear 1 ... L-1, then the output tiles, plan_tc): the update of layer l only needs the
// tiles of Linear l+1 (back-projection) and Linear l (own error), so group U works on the lower layers while the tensor
// pipe and group T are busy with the output tiles, and the lower Linears of the NEXT step overlap with the update of the
// top hidden layer -- the step is pipelined across layers.
// SPEC folds the modes of the two calls that matter most into compile-time constants (the epilogue warps are bound by
// instruction fetch / issue, and every dead branch costs code footprint): 1 = MCPC learning / sampling (SGD + in-kernel
// Philox noise, Bernoulli top, update_x), 2 = deterministic PC / MAP (Adam, no noise, Bernoulli top, update_x),
// 3 = sampling without a sensory gradient (SGD + Philox, zero_fn / no loss: no output tile is ever visited),
// 4 = 3 + trajectory records (thinned read-outs: the output tiles are visited on the recorded steps only),
// 0 = everything read from the parameters.  SPEC 1-3 also mean: no trajectories; SPEC != 0: no x.grad read-out.
template <int NR, int RV, bool TRACE, int SPEC>
__global__ void __launch_bounds__(640, 1) infer_tc_kernel(const __grid_constant__ TcParams p) {
  constexpr int RPT = RV / 2;
  constexpr int SUB = 2;                     // sub-groups of group T on 8-chain CTAs (tiles in flight; 4 measured slower)
  constexpr int kTThreads = SUB * 128;       // threads of group T
  constexpr int kGB = (RV <= 8) ? SUB : 2;   // G operand buffers
  constexpr bool ALT = (RV <= 8);            // group T works on alternate tiles (see there)
  constexpr int kGrp = 256;                  // threads per epilogue group
  constexpr int kTileArr = ALT ? 128 : 256;  // group-T threads that hand one tile over
  constexpr int kDA = (NR == 16) ? 2 * SUB : 2;   // prediction accumulators in flight (TMEM columns permitting)
  constexpr int CH = RPT < 8 ? RPT : 8;      // chains a thread of group U processes at a time
  constexpr bool kNoiseEarly = (RV <= 8);    // draw the Langevin noise before waiting for the back-projection
                                             // (wider chain tiles have no registers to hold it across the wait)
  constexpr int kMmaWarp = 16, kMmaWarpB = 17, kLoadWarp = 18, kSigWarp = 19;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ Barriers bars;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_red[2][2][16][2];       // [group][step parity][warp][energy, loss]
  __shared__ unsigned s_saved[2];            // saved steps whose rows group T / group U have finished storing (signal warp)
  __shared__ int2 s_tile[kMaxTiles];         // x = lin | out_tile << 8 | (h_out + 1) << 16, y = number of units of that Linear

  const NetDev& nd = p.net;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = nd.L, HT = p.HT;
  const int opt_kind = (SPEC == 1 || SPEC == 3 || SPEC == 4) ? (int)MCPC_OPT_SGD : (SPEC == 2 ? (int)MCPC_OPT_ADAM : p.optimizer);
  const int noise_kind = (SPEC == 1 || SPEC == 3 || SPEC == 4) ? (int)MCPC_NOISE_PHILOX : (SPEC == 2 ? (int)MCPC_NOISE_NONE : p.noise_mode);
  const int top_kind = (SPEC == 1 || SPEC == 2) ? (int)MCPC_TOP_BERNOULLI : nd.top;
  const bool top_has_grad = (SPEC == 1 || SPEC == 2) ? true : ((SPEC == 3 || SPEC == 4) ? false : (bool)nd.top_has_grad);
  const bool do_update_x = (SPEC != 0) ? true : (p.update_x != 0);
  // the specialisations also require (the host checks): targets and Adam state resident in TMEM, one unit tile per layer
  const bool y_in_tmem = (SPEC == 1 || SPEC == 2) ? true : ((SPEC == 3 || SPEC == 4) ? false : (p.y_tmem != 0));   // 3, 4: no targets
  const bool adam_in_tmem = (SPEC == 2) ? true : (SPEC != 0 ? false : (p.adam_tmem != 0));
  const int row0 = blockIdx.x * RV;
  // the issuers' tables live in the last kTabBytes of the dynamic allocation (inside the over-read slack of plan_tc:
  // nothing is ever written there, and what an over-reading MMA makes of them lands in accumulator lanes nobody reads)
  MmaA* const s_mA = reinterpret_cast<MmaA*>(smem + p.tab_off);
  MmaB* const s_mB = reinterpret_cast<MmaB*>(smem + p.tab_off + kMaxTiles * sizeof(MmaA));

  if (tid < p.n_hid_tiles + p.n_out_tiles) {
    const Tile& T = p.tiles[tid];
    s_tile[tid] = make_int2(T.lin | (T.out_tile << 8) | ((T.h_out + 1) << 16), T.lin == nd.L ? nd.d_out : nd.dims[T.lin]);
    const uint32_t sb = smem_u32(smem);
    const int in_layer = T.lin - 1;
    const bool has_b = (T.lin < L) || top_has_grad;
    const bool last_of_lin = (tid + 1 == p.n_hid_tiles + p.n_out_tiles) || (p.tiles[tid + 1].lin != T.lin);   // (output tiles come last)
    MmaA A;
    A.ad0 = smem_desc(sb + T.smem_off, 128u, (uint32_t)T.sbo);
    A.bd0 = smem_desc(sb + p.act_off[in_layer], (uint32_t)(NR * 16), 128u);          // MN-major (st_chains_bf16)
    bool first_use = true;                                    // no earlier tile of the step reads this layer's activations
    for (int k = 0; k < tid; ++k) first_use = first_use && (p.tiles[k].lin != T.lin);
    A.meta = in_layer | ((T.slot + 1) << 4) | ((has_b ? 1 : 0) << 6) | ((first_use ? 1 : 0) << 7) | ((T.Kp / 16) << 8);
    A.pad[0] = A.pad[1] = A.pad[2] = 0;
    s_mA[tid] = A;
    // K extent of the back-projection = the tile's valid output units in groups of 16 (rows of G beyond them are never written)
    const int n_units = ((T.lin == L) ? nd.d_out : nd.dims[T.lin]) - T.out_tile * 128;
    const int nkb = n_units >= 128 ? 8 : (n_units + 15) / 16;
    MmaB Bm;
    Bm.ad0 = smem_desc(sb + T.smem_off, (uint32_t)T.sbo, 128u);
    Bm.a_step = (uint32_t)(2 * T.sbo) >> 4;
    const bool first_of_lin = (tid == 0) || (p.tiles[tid - 1].lin != T.lin);   // overwrites the back-projection accumulators
    Bm.meta = (has_b ? 1 : 0) | ((last_of_lin ? 1 : 0) << 1) | ((T.slot + 1) << 2) | (in_layer << 4) | (nkb << 8) |
              (p.ut[in_layer] << 12) | (p.h_off[in_layer] << 16) | ((first_of_lin ? 1 : 0) << 20);
    s_mB[tid] = Bm;
  }
  if (tid == 0) {
    s_saved[0] = 0;
    s_saved[1] = 0;
    mbar_init(&bars.w_res, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.w_full[i], 1);
      mbar_init(&bars.w_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bars.g_full[i], kTileArr / 32);
      mbar_init(&bars.g_empty[i], 1);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&bars.dA_full[i], 1);
      mbar_init(&bars.dA_empty[i], kTileArr / 32);
    }
    mbar_init(&bars.out_read, 1);
    for (int l = 0; l < kMaxL; ++l) {
      mbar_init(&bars.acts_ready[l], kGrp / 32);
      mbar_init(&bars.bp_ready[l], 1);
      mbar_init(&bars.g_ready[l], kTThreads / 32);
      mbar_init(&bars.nz_full[0][l], kGrp / 32);
      mbar_init(&bars.nz_full[1][l], kGrp / 32);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&tmem_base_s, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  // TMEM column map (fp32 columns): [dA0 .. dA(kDA-1) | bp_h ... | x_h ... | gown_h ...], NR columns each
  // then one bias column per weight tile (32 reserved) and, if p.y_tmem, NR target columns per output tile
  const uint32_t col_dA = 0, col_bp = kDA * NR, col_x = (kDA + HT) * NR, col_g = (kDA + 2 * HT) * NR;
  const uint32_t col_bias = (kDA + 3 * HT) * NR, col_y = col_bias + 32;
  // Adam on the latents (deterministic PC / MAP): m and v behind the targets, HT * NR columns each
  const uint32_t col_m = col_y + (y_in_tmem ? (uint32_t)p.n_out_tiles * NR : 0u), col_v = col_m + HT * NR;
  const int n_tiles_all = p.n_hid_tiles + p.n_out_tiles;      // table order: Linear 1 ... L-1, then the output tiles

  // =====================================================================================================
  if (warp == kLoadWarp) {
    if (lane == 0) {
      uint32_t res_bytes = 0;
      for (int t = 0; t < n_tiles_all; ++t)
        if (p.tiles[t].slot < 0) res_bytes += (uint32_t)p.tiles[t].bytes;
      if (res_bytes > 0) {
        mbar_expect_tx(&bars.w_res, res_bytes);
        for (int t = 0; t < n_tiles_all; ++t)
          if (p.tiles[t].slot < 0)
            bulk_g2s(smem + p.tiles[t].smem_off, p.packed + p.tiles[t].gsrc, (uint32_t)p.tiles[t].bytes, &bars.w_res);
      } else {
        mbar_arrive(&bars.w_res);
      }
      uint32_t empty_phase = 3;
      for (int ts = 0; ts < p.n_steps; ++ts) {
        const bool do_traj = (SPEC == 0 || SPEC == 4) && (p.traj_every > 0) && (ts % p.traj_every == 0);
        const bool need_out = top_has_grad || (do_traj && p.traj_out != nullptr);
        for (int t = 0; t < (need_out ? n_tiles_all : p.n_hid_tiles); ++t) {
          const Tile& T = p.tiles[t];
          if (T.slot < 0) continue;
          mbar_wait_parked(&bars.w_empty[T.slot], (empty_phase >> T.slot) & 1u);
          empty_phase ^= 1u << T.slot;
          mbar_expect_tx(&bars.w_full[T.slot], (uint32_t)T.bytes);
          bulk_g2s(smem + T.smem_off, p.packed + T.gsrc, (uint32_t)T.bytes, &bars.w_full[T.slot]);
        }
      }
    }
  } else if (warp == kSigWarp) {
    // ---------------- signal warp (overlapped weight update only) ----------------
    // Tells the concurrent weight-gradient kernel that this CTA's rows of saved step s are in global memory.  The
    // gpu-scope release costs ~1 us (it waits for the write acknowledgements); issued by an epilogue thread it sat on the
    // step's critical path (+2 us per step), so the groups only bump a shared-memory counter after their end-of-step
    // barrier (release.cta) and this otherwise idle warp does the waiting.  Causality: group stores -> bar.sync ->
    // red.release.cta -> ld.acquire.cta here -> fence.acq_rel.gpu + red.release.gpu -> the consumer's ld.acquire.gpu.
    if (lane == 0 && p.ready != nullptr && p.save_g != nullptr) {
      const int s_end = min(p.save_end, p.n_steps);
      const uint32_t a0 = smem_u32(&s_saved[0]), a1 = smem_u32(&s_saved[1]);
      for (int s = 0; s < s_end - p.save_begin; ++s) {
        for (;;) {
          unsigned v0, v1;
          asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v0) : "r"(a0) : "memory");
          asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v1) : "r"(a1) : "memory");
          if (v0 > (unsigned)s && v1 > (unsigned)s) break;
          __nanosleep(200);
        }
        signal_saved(p.ready + s);
      }
    }
  } else if (warp == kMmaWarp) {
    // ---------------- MMA issuer A: predictions (ONE elected thread runs the whole loop) ----------------
    // Two issuer warps: this one runs ahead with the prediction GEMMs (phase A), warp kMmaWarpB issues the
    // back-projections (phase B) as soon as group T hands a G operand over.  Neither waits for the other's barriers,
    // so a late hand-over does not hold back the next tile's prediction.
    // The issuers pace the whole step (ncu r02b: T waits for accumulators half of its time, U a third; each issuer spent
    // ~1,100 cycles and ~170 instructions per tile on descriptor arithmetic over indexed constant-bank loads, elect /
    // reconvergence and re-derived flags).  Everything per tile that does not change from step to step is therefore
    // precomputed into s_mA / s_mB, and one thread runs the loop: per tile two shared-memory loads, the waits, the MMAs.
    if (elect_one()) {
      const uint32_t id_a = idesc_bf16(128, NR, false, true);
      constexpr int kBStep = NR * 2;                           // B descriptor step per 16 units (two rows of NR*16 bytes)
      uint32_t ph_wfull = 0, ph_dAe = (1u << kDA) - 1u;
      mbar_wait_parked(&bars.w_res, 0);
      for (int ts = 0; ts < p.n_steps; ++ts) {
        const bool do_traj = (SPEC == 0 || SPEC == 4) && (p.traj_every > 0) && (ts % p.traj_every == 0);
        const bool need_out = top_has_grad || (do_traj && p.traj_out != nullptr);
        const int t_end = need_out ? n_tiles_all : p.n_hid_tiles;
        TC_STAMP(true, ts, 0);
        for (int t = 0; t < t_end; ++t) {
          const MmaA M = s_mA[t];
          const int in_layer = M.meta & 15, slot = ((M.meta >> 4) & 3) - 1, nk = (M.meta >> 8) & 0xff;
          if (M.meta & 128) mbar_wait_parked(&bars.acts_ready[in_layer], ts & 1);   // act(x_{lin-1}) of THIS step (group U)
          if (slot >= 0) {
            mbar_wait_parked(&bars.w_full[slot], (ph_wfull >> slot) & 1u);
            ph_wfull ^= 1u << slot;
          }
          const int db = t & (kDA - 1);
          mbar_wait_parked(&bars.dA_empty[db], (ph_dAe >> db) & 1u);
          ph_dAe ^= 1u << db;
          fence_after_sync();
          const uint32_t dcol = tmem + col_dA + db * NR;
          // groups of 8 K-steps fully unrolled: the instructions issue back to back (46 instead of 91 cycles each,
          // scripts/umma_timing.py variants 6 / 1)
          int ks = 0;
#pragma unroll 1
          for (; ks + 8 <= nk; ks += 8) {                      // (not unrolled further: the issuer's code must stay small)
            const uint64_t a8 = M.ad0 + (uint64_t)(ks * 16), b8 = M.bd0 + (uint64_t)(ks * kBStep);
            mma_bf16_ss(dcol, a8, b8, id_a, ks > 0);
#pragma unroll
            for (int kk = 1; kk < 8; ++kk) mma_bf16_ss(dcol, a8 + (uint64_t)(kk * 16), b8 + (uint64_t)(kk * kBStep), id_a, true);
          }
#pragma unroll 1
          for (; ks < nk; ++ks) mma_bf16_ss(dcol, M.ad0 + (uint64_t)(ks * 16), M.bd0 + (uint64_t)(ks * kBStep), id_a, ks > 0);
          mma_commit(&bars.dA_full[db]);
          // a streamed tile without back-projection is free again once this prediction has read it
          if (slot >= 0 && !((M.meta >> 6) & 1)) mma_commit(&bars.w_empty[slot]);
          TC_STAMP(true, ts, 1 + t);
        }
        // no back-projection follows read-out only predictions, so nothing else tells group U that the top layer's
        // activations have been read: without this its update of step ts could overwrite act(x_{L-1}) under the last tiles
        if (need_out && !top_has_grad) mma_commit(&bars.out_read);
        TC_STAMP(true, ts, 21);
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarpB) {
    // ---------------- MMA issuer B: back-projections through the MN-major view of the same weight tiles ----------------
    if (elect_one()) {
      const uint32_t id_b = idesc_bf16(128, NR, true, true);
      constexpr int kBStep = NR * 2;
      const uint64_t gd0 = smem_desc(smem_u32(smem) + p.gbuf_off[0], (uint32_t)(NR * 16), 128u);
      const uint32_t gd_step = (uint32_t)(p.gbuf_off[1] - p.gbuf_off[0]) >> 4;
      uint32_t ph_gfull = 0;
      mbar_wait_parked(&bars.w_res, 0);
      for (int ts = 0; ts < p.n_steps; ++ts) {
        const bool do_traj = (SPEC == 0 || SPEC == 4) && (p.traj_every > 0) && (ts % p.traj_every == 0);
        const bool need_out = top_has_grad || (do_traj && p.traj_out != nullptr);
        const int t_end = need_out ? n_tiles_all : p.n_hid_tiles;
        for (int t = 0; t < t_end; ++t) {
          const MmaB M = s_mB[t];
          if (!(M.meta & 1)) continue;                                  // read-out only: nothing flows back
          const int gb = t & (kGB - 1);
          mbar_wait_parked(&bars.g_full[gb], (ph_gfull >> gb) & 1u);    // group T consumed the prediction of this tile,
          ph_gfull ^= 1u << gb;                                         // so phase A has finished reading it as well
          fence_after_sync();
          const int in_layer = (M.meta >> 4) & 15, slot = ((M.meta >> 2) & 3) - 1;
          const int nkb = (M.meta >> 8) & 15, n_ut = (M.meta >> 12) & 15, h0 = (M.meta >> 16) & 15;
          const uint64_t bd0 = gd0 + (uint64_t)((uint32_t)gb * gd_step);
          for (int u = 0; u < n_ut; ++u) {
            const uint64_t ad0 = M.ad0 + (uint64_t)(u * (2048 >> 4));
            const uint32_t dcol = tmem + col_bp + (h0 + u) * NR;
            mma_bf16_ss(dcol, ad0, bd0, id_b, !((M.meta >> 20) & 1));
            if (nkb == 8) {
#pragma unroll
              for (int ks = 1; ks < 8; ++ks) mma_bf16_ss(dcol, ad0 + (uint64_t)(ks * M.a_step), bd0 + (uint64_t)(ks * kBStep), id_b, true);
            } else {
#pragma unroll 1
              for (int ks = 1; ks < nkb; ++ks) mma_bf16_ss(dcol, ad0 + (uint64_t)(ks * M.a_step), bd0 + (uint64_t)(ks * kBStep), id_b, true);
            }
          }
          mma_commit(&bars.g_empty[gb]);
          if ((M.meta >> 1) & 1) mma_commit(&bars.bp_ready[in_layer]);   // back-projection into layer lin-1 is complete
          if (slot >= 0) mma_commit(&bars.w_empty[slot]);
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue groups ----------------
    // The SM's warp arbiter favours higher warp ids: the latency-critical tile epilogues (group T) get warps
    // 8-15, the background layer updates (group U) warps 0-7, the MMA issuer the highest id of all.
    const int grp = (warp < 8) ? 1 : 0;                                // 0 = tiles (T), 1 = update (U)
    const int gw = warp & 7;                                           // warp inside the group
    const int gtid = gw * 32 + lane;
    const int q = warp & 3, cg = (gw >> 2) & 1;
    const int ln = q * 32 + lane;                                      // unit index inside a 128-unit tile
    const int cbase = cg * RPT;
    const int rb = row0 + cbase;
    const int nrow = max(0, min(RPT, p.B - rb));
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t lane_addr = lane_base + (uint32_t)cbase;

    if (grp == 1) {
      // ======================= group U: initial state, then one layer update after the other =======================
      // activation AND G operand buffers: chains RV..NR-1 (if any) stay zero for the whole run
      for (int i = gtid * 16; i < p.gbuf_off[kGB - 1] + NR * 128 * 2 - p.act_off[0]; i += kGrp * 16)
        *reinterpret_cast<uint4*>(smem + p.act_off[0] + i) = make_uint4(0, 0, 0, 0);
      asm volatile("bar.sync 2, 256;" ::: "memory");
      for (int h = 0; h < HT; ++h) {
        const int l = p.h_layer[h], u = p.h_index[h] * 128 + ln, dl = nd.dims[l];
        float xv[RPT], av[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          xv[i] = (u < dl && i < nrow) ? p.x[l][(size_t)(rb + i) * dl + u] : 0.0f;
          av[i] = act_tc(nd.act[l], xv[i]);
        }
        if (u < dl) st_chains_bf16<RPT>(smem + p.act_off[l] + chain_op_off(NR, u, cbase), av);
        __syncwarp();
        tmem_st<RPT>(lane_addr + col_x + h * NR, xv);
        if (adam_in_tmem) {
          float mv[RPT], vv[RPT];
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const bool ok = (u < dl && i < nrow);
            mv[i] = ok ? p.m[l][(size_t)(rb + i) * dl + u] : 0.0f;
            vv[i] = ok ? p.v[l][(size_t)(rb + i) * dl + u] : 0.0f;
          }
          __syncwarp();
          tmem_st<RPT>(lane_addr + col_m + h * NR, mv);
          tmem_st<RPT>(lane_addr + col_v + h * NR, vv);
        }
      }
      tmem_st_wait();
      fence_async_smem();
      fence_before_sync();
      for (int l = 0; l < L; ++l) warp_arrive(&bars.acts_ready[l], lane);

      const bool adam = (opt_kind == MCPC_OPT_ADAM);
      uint32_t ph_out_read = 0;
      const float* mu0 = (SPEC == 0) ? p.mu0 : nullptr;      // the specialised instantiations are zero-input only
      for (int ts = 0; ts < p.n_steps; ++ts) {
        const int t_abs = p.t_begin + ts;
        const bool do_save = (p.save_g != nullptr) && ts >= p.save_begin && ts < p.save_end;
        const int slot = ts - p.save_begin;
        const bool do_traj = (SPEC == 0 || SPEC == 4) && (p.traj_every > 0) && (ts % p.traj_every == 0);
        const int rec = do_traj ? ts / p.traj_every : 0;
        const bool last = (ts == p.n_steps - 1);
        const bool wait_out_read = !top_has_grad && do_traj && p.traj_out != nullptr;   // see Barriers::out_read
        float e_part = 0.0f;
        __nv_bfloat16* sg_row = do_save ? p.save_g + ((size_t)slot * p.B + rb) * p.sg_pitch : nullptr;
        __nv_bfloat16* sf_row = do_save ? p.save_f + ((size_t)slot * p.B + rb) * p.sf_pitch : nullptr;
        float step_size = 0.0f, inv_bc2_sqrt = 1.0f;
        if (adam && do_update_x) {
          // bias corrections come from a table built in fp64 by adam_table_kernel: two divisions and a square root in
          // double per thread and step cost ~6k cycles of the (1/64-rate) FP64 pipe on the step's critical path
          step_size = __ldg(p.adam_tab + 2 * ts);
          inv_bc2_sqrt = __ldg(p.adam_tab + 2 * ts + 1);
        }
        for (int l = 0; l < L; ++l) {                        // bottom-up, the order the tiles complete in
          const bool has_above = (l + 1 < L) || top_has_grad;
          const int kind = nd.act[l];
          const int dl = nd.dims[l];
          const int n_ut = p.ut[l];
          const bool defer = (RV <= 8) && (SPEC != 0 || n_ut == 1);   // one unit tile: global stores go out after the hand-over
                                                             // (wider chain tiles have no registers to spare for it)
          // the deferred stores only need the pre-update latents (single-tile layers keep them in registers across the
          // hand-over); f(x) and the layer-0 error are recomputed from them
          float xold[CH], b0 = 0.0f;
          bool uvalid = false;
          int u = 0;
          auto global_stores = [&](int c0) {
            if (!uvalid) return;
            if (do_traj && p.traj_x[l] != nullptr) {
              float* tx = p.traj_x[l] + ((size_t)rec * p.B + rb) * dl + u;
#pragma unroll
              for (int i = 0; i < CH; ++i)
                if (c0 + i < nrow) tx[(size_t)(c0 + i) * dl] = xold[i];
            }
            if (do_save) {
              __nv_bfloat16* sf = sf_row + p.sg_off[l] + u;
#pragma unroll
              for (int i = 0; i < CH; ++i)
                if (c0 + i < nrow) sf[(size_t)(c0 + i) * p.sf_pitch] = __float2bfloat16(act_tc(kind, xold[i]));
              if (l == 0) {                                  // layer 0 has no weight tile: its G operand is saved here
                __nv_bfloat16* sg = sg_row + u;
#pragma unroll
                for (int i = 0; i < CH; ++i)
                  if (c0 + i < nrow) {
                    const float m0 = (mu0 != nullptr) ? __ldg(mu0 + (size_t)(rb + c0 + i) * dl + u) : b0;
                    sg[(size_t)(c0 + i) * p.sg_pitch] = __float2bfloat16(-nd.gc[0] * (xold[i] - m0));
                  }
              }
            }
          };
          for (int hi = 0; hi < n_ut; ++hi) {
            const int h = p.h_off[l] + hi;
            u = hi * 128 + ln;
            uvalid = u < dl;
            const int gu = nd.off[l] + u;
            // a warp whose 32 units are all padding (e.g. 3 of 4 warps on a 20-unit layer) has nothing to update: its
            // TMEM lanes keep the zeros written at start-up; it only takes part in the waits and the hand-over
            const bool warp_idle = (hi * 128 + q * 32) >= dl;
            auto wait_layer = [&]() {
              if (has_above) mbar_wait_parked(&bars.bp_ready[l], ts & 1);        // all tiles of Linear l+1 back-projected
              if (l > 0) mbar_wait_parked(&bars.g_ready[l], ts & 1);             // group T stored the own-layer error of layer l
              if (l == L - 1 && wait_out_read) mbar_wait_parked(&bars.out_read, ph_out_read);
            };
            if (warp_idle) {
              if (hi == 0) wait_layer();
              continue;
            }
            // the chains of the thread go through in chunks of CH <= 8: bounded register pressure and code size
#pragma unroll 1
            for (int c0 = 0; c0 < RPT; c0 += CH) {
              const int rbc = rb + c0, cbc = cbase + c0;
              const int nrc = nrow - c0;
              const uint32_t lac = lane_addr + (uint32_t)c0;
              const size_t xoffc = (size_t)rbc * dl + u;
              // ---- 1. everything that does not depend on this step's MMAs: noise, Adam state, the layer-0 bias ----
              float mv[CH], vv[CH], nz[CH];
#pragma unroll
              for (int i = 0; i < CH; ++i) { mv[i] = 0.0f; vv[i] = 0.0f; nz[i] = 0.0f; }
              b0 = 0.0f;
              if (uvalid) {
                if (l == 0 && p.b[0] != nullptr) b0 = __ldg(p.b[0] + u);
                if (adam && do_update_x && !adam_in_tmem) {
#pragma unroll
                  for (int i = 0; i < CH; ++i) {
                    mv[i] = (i < nrc) ? p.m[l][xoffc + (size_t)i * dl] : 0.0f;
                    vv[i] = (i < nrc) ? p.v[l][xoffc + (size_t)i * dl] : 0.0f;
                  }
                }
              }
              auto draw_noise = [&]() {
                if (!uvalid) return;
                if (noise_kind == MCPC_NOISE_SUPPLIED) {
                  const float* np_ = p.noise + ((size_t)ts * p.B + rbc) * nd.SD + gu;
#pragma unroll
                  for (int i = 0; i < CH; ++i) nz[i] = (i < nrc) ? __ldg(np_ + (size_t)i * nd.SD) : 0.0f;
                } else if (noise_kind == MCPC_NOISE_PHILOX) {
                  float nrm[4];
                  const uint64_t chain0 = p.chain_offset + (uint64_t)rbc;
                  if ((chain0 & 3) == 0) {                       // the usual case: the chunk is whole groups of four chains
#pragma unroll
                    for (int g = 0; g < CH / 4; ++g) {
                      langevin_normals4(p.seed, (uint32_t)gu, (uint32_t)t_abs, (chain0 >> 2) + g, nrm);
#pragma unroll
                      for (int j = 0; j < 4; ++j) nz[4 * g + j] = p.noise_scale * nrm[j];
                    }
                    return;
                  }
                  uint64_t cur_q = ~0ull;
#pragma unroll
                  for (int i = 0; i < CH; ++i) {
                    const uint64_t chain = p.chain_offset + (uint64_t)(rbc + i);
                    if ((chain >> 2) != cur_q) {
                      cur_q = chain >> 2;
                      langevin_normals4(p.seed, (uint32_t)gu, (uint32_t)t_abs, cur_q, nrm);
                    }
                    nz[i] = p.noise_scale * sel4(nrm, (uint32_t)(chain & 3));
                  }
                }
              };
              // pure sampling: group T (idle but for the hidden tiles) drew this step's noise into shared memory one step ahead
              const bool nz_smem = (SPEC == 3) && (p.nz_off >= 0);
              if (nz_smem) {
                if (hi == 0 && c0 == 0) mbar_wait_parked(&bars.nz_full[ts & 1][l], (ts >> 1) & 1);
                const float* nb = reinterpret_cast<const float*>(smem + p.nz_off) +
                                  ((size_t)((ts & 1) * HT + h) * RV + (size_t)cbc) * 128 + ln;
#pragma unroll
                for (int i = 0; i < CH; ++i) nz[i] = nb[i * 128];
              } else if (kNoiseEarly) {
                draw_noise();                                   // before the wait: off the critical hand-over chain
              }
              // ---- 2. wait for this layer's back-projection and own error, then ONE batch of TMEM loads ----
              if (hi == 0 && c0 == 0) {
                wait_layer();
                fence_after_sync();
                TC_STAMP(gtid == 0, ts, 40 + l);
              }
              __syncwarp();
              float xv[CH], bp[CH], gown[CH], gradv[CH];
              tmem_ld_nw<CH>(lac + col_x + h * NR, xv);
              if (has_above) tmem_ld_nw<CH>(lac + col_bp + h * NR, bp);
              if (l > 0) tmem_ld_nw<CH>(lac + col_g + h * NR, gown);
              const bool adam_t = adam && do_update_x && adam_in_tmem;
              if (adam_t) {
                tmem_ld_nw<CH>(lac + col_m + h * NR, mv);
                tmem_ld_nw<CH>(lac + col_v + h * NR, vv);
              }
              tmem_ld_wait();
              TC_STAMP(gtid == 0 && l == 1, ts, 43);
              tmem_ld_tie(xv);
              if (adam_t) {
                tmem_ld_tie(mv);
                tmem_ld_tie(vv);
              }
              if (has_above) {
                tmem_ld_tie(bp);
              } else {
#pragma unroll
                for (int i = 0; i < CH; ++i) bp[i] = 0.0f;
              }
              if (l > 0) {
                tmem_ld_tie(gown);
              } else {
                // layer 0 is predicted by its bias alone (zero inputs: eps_0 = x_0 - b_0) or by the per-chain constant
                // mu0 = W_0 inputs + b_0 of mu0_kernel
                const float ce = 0.5f * nd.c[0], gc = nd.gc[0];
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                  const float m0 = (mu0 != nullptr && uvalid && i < nrc) ? __ldg(mu0 + xoffc + (size_t)i * dl) : b0;
                  const float eps = xv[i] - m0;
                  gown[i] = -gc * eps;
                  if (uvalid && i < nrc) e_part = fmaf(ce * eps, eps, e_part);
                }
              }
              if (!kNoiseEarly && !nz_smem) draw_noise();
              // ---- 3. update (straight-line: CH independent chains interleave) ----
#pragma unroll
              for (int i = 0; i < CH; ++i) {
                const float x = xv[i];
                const float a = act_tc(kind, x);
                xold[i] = x;
                gradv[i] = fmaf(dact_tc(kind, x, a), bp[i], -gown[i]);
              }
              if (SPEC == 0 && uvalid && last && p.xgrad[l] != nullptr) {
#pragma unroll
                for (int i = 0; i < CH; ++i)
                  if (i < nrc) p.xgrad[l][xoffc + (size_t)i * dl] = gradv[i];
              }
              if (do_update_x) {
                if (!adam) {
#pragma unroll
                  for (int i = 0; i < CH; ++i) xv[i] = fmaf(-p.lr, gradv[i], xv[i]);
                } else {
#pragma unroll
                  for (int i = 0; i < CH; ++i) {
                    mv[i] = fmaf(p.one_minus_b1, gradv[i] - mv[i], mv[i]);
                    vv[i] = fmaf(p.one_minus_b2 * gradv[i], gradv[i], vv[i] * p.beta2f);
                    xv[i] = fmaf(-step_size, __fdividef(mv[i], fmaf(sqrtf(vv[i]), inv_bc2_sqrt, p.adam_eps)), xv[i]);
                  }
                  if (uvalid && !adam_in_tmem) {
#pragma unroll
                    for (int i = 0; i < CH; ++i)
                      if (i < nrc) {
                        p.m[l][xoffc + (size_t)i * dl] = mv[i];
                        p.v[l][xoffc + (size_t)i * dl] = vv[i];
                      }
                  }
                }
              }
              if (noise_kind != MCPC_NOISE_NONE) {
#pragma unroll
                for (int i = 0; i < CH; ++i) xv[i] = fmaf(-p.lr, nz[i], xv[i]);
              }
              TC_STAMP(gtid == 0 && l == 1, ts, 44);
              // ---- 4. next step's operands: bf16 act(x) to shared memory, fp32 x back to TMEM ----
              float av[CH];
#pragma unroll
              for (int i = 0; i < CH; ++i) {
                if (i >= nrc || !uvalid) xv[i] = uvalid ? 0.0f : xold[i];   // chains past the batch keep their zeros
                av[i] = act_tc(kind, xv[i]);
              }
              if (uvalid) st_chains_bf16<CH>(smem + p.act_off[l] + chain_op_off(NR, u, cbc), av);
              __syncwarp();                                    // .sync.aligned store: every lane executes it
              tmem_st<CH>(lac + col_x + h * NR, xv);
              if (adam_t) {
                tmem_st<CH>(lac + col_m + h * NR, mv);
                tmem_st<CH>(lac + col_v + h * NR, vv);
              }
              if (!defer) global_stores(c0);
            }
          }
          TC_STAMP(gtid == 0 && l == 1, ts, 45);
          tmem_st_wait();
          TC_STAMP(gtid == 0 && l == 1, ts, 46);
          fence_async_smem();
          TC_STAMP(gtid == 0 && l == 1, ts, 47);
          fence_before_sync();
          warp_arrive(&bars.acts_ready[l], lane);            // act(x_l) / x_l of step ts+1 are in place
          TC_STAMP(gtid == 0 && l == 1, ts, 48);
          // the proxy fence above waits for every earlier memory operation of the thread: the global stores of a
          // single-tile layer are issued after the hand-over so that they do not delay it
          if (defer) global_stores(0);
          TC_STAMP(gtid == 0 && l == 1, ts, 49);
        }
        TC_STAMP(gtid == 0, ts, 54);
        if (wait_out_read) ph_out_read ^= 1u;
        // layer-0 energy of this step
        e_part = warp_sum_tc(e_part);
        float (*red)[2] = s_red[1][ts & 1];
        if (lane == 0) { red[gw][0] = e_part; red[gw][1] = 0.0f; }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (gtid < 2) {
          float s = 0.0f;
#pragma unroll
          for (int w = 0; w < 8; ++w) s += red[w][gtid];
          p.partials[(((size_t)ts * p.n_ctas + blockIdx.x) * 2 + 1) * 2 + gtid] = s;
        }
        if (gtid == 0 && do_save && p.ready != nullptr) saved_step(&s_saved[1]);
      }
      // ---------- write the latents back (group U owns x) ----------
      for (int h = 0; h < HT; ++h) {
        const int l = p.h_layer[h], u = p.h_index[h] * 128 + ln, dl = nd.dims[l];
        float xv[RPT];
        tmem_ld<RPT>(lane_addr + col_x + h * NR, xv);
#pragma unroll
        for (int i = 0; i < RPT; ++i)
          if (u < dl && i < nrow) p.x[l][(size_t)(rb + i) * dl + u] = xv[i];
        if (adam_in_tmem) {
          float mv[RPT], vv[RPT];
          tmem_ld<RPT>(lane_addr + col_m + h * NR, mv);
          tmem_ld<RPT>(lane_addr + col_v + h * NR, vv);
#pragma unroll
          for (int i = 0; i < RPT; ++i)
            if (u < dl && i < nrow) {
              p.m[l][(size_t)(rb + i) * dl + u] = mv[i];
              p.v[l][(size_t)(rb + i) * dl + u] = vv[i];
            }
        }
      }
    } else {
      // ======================= group T: per-tile epilogues =======================
      // ALT (8-chain CTAs): the two halves of the group (one warp per lane quarter each) take ALTERNATE tiles with all
      // RV chains per thread -- the per-tile chain (wait, TMEM load, MUFU, store, fence, arrive) is latency-bound, so two
      // tiles in flight nearly double the tile rate.  Half h owns accumulator / G buffer h.  Otherwise the two warps of a
      // lane quarter split the chains of every tile.
      constexpr int RT = ALT ? RV : RPT;
      const int half = gw >> 2;                                  // sub-group of this warp (0 .. SUB-1)
      const int cbT = ALT ? 0 : cbase;
      const int rbT = row0 + cbT;
      const int nrT = max(0, min(RT, p.B - rbT));
      const uint32_t laT = lane_base + (uint32_t)cbT;
      const uint32_t gtoT = chain_op_off(NR, ln, cbT);
      const bool stamp_thr = ALT ? ((gtid & 127) == 0) : (gtid == 0);
      uint32_t ph_dAf = 0, ph_ge = (1u << kGB) - 1u;
      const float qnan = __int_as_float(0x7fc00000);
      // Per-tile constants live in TMEM, not in global memory: one bias column per tile and, when they fit, the
      // targets of every output tile (NaN = "this element carries no loss": masked, past the batch, padding).
      // The tile loop then needs no global load and no per-element mask logic.
      auto target_of = [&](int t, float (&yv)[RT]) {
        const int2 ti = s_tile[t];
        const int un = ((ti.x >> 8) & 0xff) * 128 + ln;
        const bool uy = ((ti.x & 0xff) == L) && un < ti.y && un >= nd.mask_start && top_kind >= MCPC_TOP_GAUSS;
        const float* yp = p.target + (size_t)rbT * nd.d_out + un;
#pragma unroll
        for (int i = 0; i < RT; ++i) yv[i] = (uy && i < nrT) ? __ldg(yp + (size_t)i * nd.d_out) : qnan;
      };
      for (int t = 0; t < n_tiles_all; ++t) {
        const int2 ti = s_tile[t];
        const int lin = ti.x & 0xff, un = ((ti.x >> 8) & 0xff) * 128 + ln;
        const float bv = (un < ti.y && p.b[lin] != nullptr) ? __ldg(p.b[lin] + un) : 0.0f;
        __syncwarp();
        tmem_st1(lane_base + col_bias + (uint32_t)t, bv);           // both warps of a lane quarter write the same value
        if (y_in_tmem && lin == L && (!ALT || (t & (SUB - 1)) == half)) {
          float yv[RT];
          target_of(t, yv);
          __syncwarp();
          tmem_st<RT>(laT + col_y + (uint32_t)(t - p.n_hid_tiles) * NR, yv);
        }
      }
      tmem_st_wait();
      // Pure sampling (instantiation 3): this group only has the hidden tiles to do, group U is the bottleneck and half
      // of its time is the Philox + Box-Muller noise.  Group T draws the noise of step `step` for every latent into shared
      // memory ([step & 1][unit tile][chain][unit], the element its twin thread of group U reads) one step ahead.  No
      // "empty" barrier: the buffer of step s+1 was last read in step s-1, and this group has finished the tiles of step
      // s, which needed every layer's update of step s-1 (hence its wait on and its reads of that buffer), before it writes.
      auto produce_noise = [&](int step) {
        float* nbuf = reinterpret_cast<float*>(smem + p.nz_off) + (size_t)((step & 1) * HT) * RV * 128;
        const int cbN = ALT ? half * RPT : cbase;                // chains of this thread, like its twin in group U
        for (int l = 0; l < L; ++l) {
          const int dl = nd.dims[l];
          for (int hi = 0; hi < p.ut[l]; ++hi) {
            if (hi * 128 + q * 32 >= dl) continue;               // warp-uniform: all 32 units are padding
            const int u = hi * 128 + ln;
            if (u < dl) {
              float* nb = nbuf + ((size_t)(p.h_off[l] + hi) * RV + cbN) * 128 + ln;
              const uint32_t gu = (uint32_t)(nd.off[l] + u);
              float nrm[4];
              const uint64_t chain0 = p.chain_offset + (uint64_t)(row0 + cbN);
              if ((chain0 & 3) == 0) {                           // whole groups of four chains (the usual case)
#pragma unroll 2
                for (int g = 0; g < RPT / 4; ++g) {
                  langevin_normals4(p.seed, gu, (uint32_t)(p.t_begin + step), (chain0 >> 2) + g, nrm);
#pragma unroll
                  for (int j = 0; j < 4; ++j) nb[(4 * g + j) * 128] = p.noise_scale * nrm[j];
                }
                continue;
              }
              uint64_t cur_q = ~0ull;
#pragma unroll 4
              for (int i = 0; i < RPT; ++i) {
                const uint64_t chain = p.chain_offset + (uint64_t)(row0 + cbN + i);
                if ((chain >> 2) != cur_q) {
                  cur_q = chain >> 2;
                  langevin_normals4(p.seed, gu, (uint32_t)(p.t_begin + step), cur_q, nrm);
                }
                nb[i * 128] = p.noise_scale * sel4(nrm, (uint32_t)(chain & 3));
              }
            }
          }
          warp_arrive(&bars.nz_full[step & 1][l], lane);         // (release: the stores above are visible to the waiter)
        }
      };
      // bit t: tile t is the last tile of a hidden Linear (the own-layer errors of that layer are complete after it)
      uint32_t last_hid = 0;
      for (int t = 0; t < p.n_hid_tiles; ++t)
        if (t + 1 == n_tiles_all || (s_tile[t + 1].x & 0xff) != (s_tile[t].x & 0xff)) last_hid |= 1u << t;
      const bool nz_producer = (SPEC == 3) && (p.nz_off >= 0);
      if (nz_producer) produce_noise(0);
      for (int ts = 0; ts < p.n_steps; ++ts) {
        const bool do_save = (p.save_g != nullptr) && ts >= p.save_begin && ts < p.save_end;
        const int slot = ts - p.save_begin;
        const bool do_traj = (SPEC == 0 || SPEC == 4) && (p.traj_every > 0) && (ts % p.traj_every == 0);
        const int rec = do_traj ? ts / p.traj_every : 0;
        const bool need_out = top_has_grad || (do_traj && p.traj_out != nullptr);
        const int t_end = need_out ? n_tiles_all : p.n_hid_tiles;
        float e_part = 0.0f, l_part = 0.0f;
        __nv_bfloat16* sg_row = do_save ? p.save_g + ((size_t)slot * p.B + rbT) * p.sg_pitch : nullptr;
        uint32_t x_waited = 0;
        TC_STAMP(gtid == 0, ts, 32);
        for (int t = 0; t < t_end; ++t) {
          const int k = t;
          if (ALT && (k & (SUB - 1)) != half) {               // another sub-group's tile: a handful of instructions
            if ((last_hid >> t) & 1u) warp_arrive(&bars.g_ready[s_tile[t].x & 0xff], lane);
            continue;
          }
          const int2 ti = s_tile[t];
          const int lin = ti.x & 0xff, h = ((ti.x >> 16) & 0xff) - 1, dl = ti.y;
          const int db = k & (kDA - 1), gb = k & (kGB - 1);
          const bool is_out = (lin == L);
          const bool has_b = !is_out || top_has_grad;
          const int u = ((ti.x >> 8) & 0xff) * 128 + ln;
          const bool uvalid = u < dl;
          float yv[RT];
          if (!y_in_tmem && is_out) target_of(t, yv);          // targets do not fit TMEM: plain loads
          TC_STAMP(stamp_thr && (k == 2 || k == 3), ts, 60);
          if (!is_out && !((x_waited >> lin) & 1u)) {         // x_lin of THIS step was written by group U last step
            mbar_wait_parked(&bars.acts_ready[lin], ts & 1);
            x_waited |= 1u << lin;
          }
          mbar_wait_parked(&bars.dA_full[db], (ph_dAf >> db) & 1u);
          ph_dAf ^= 1u << db;
          TC_STAMP(stamp_thr && (k == 2 || k == 3), ts, 61);
          fence_after_sync();
          TC_STAMP(stamp_thr && (k == 2 || k == 3), ts, 55);
          // a warp whose 32 units are all padding (partial last tile of a Linear) only hands the buffers over: the
          // back-projection of a partial tile reads the valid 16-unit groups of G only
          if (((ti.x >> 8) & 0xff) * 128 + q * 32 >= dl) {
            if (has_b) {
              mbar_wait_parked(&bars.g_empty[gb], (ph_ge >> gb) & 1u);
              ph_ge ^= 1u << gb;
            }
            fence_before_sync();
            warp_arrive(&bars.dA_empty[db], lane);
            if (has_b) warp_arrive(&bars.g_full[gb], lane);
            if ((last_hid >> t) & 1u) warp_arrive(&bars.g_ready[lin], lane);
            continue;
          }
          uint8_t* gptr = smem + p.gbuf_off[gb] + gtoT;
          // the G buffer is only needed for the stores at the end: by then the back-projection that read its previous
          // content (two tiles ago) has long completed
          auto wait_g_buffer = [&]() {
            if (has_b) {
              mbar_wait_parked(&bars.g_empty[gb], (ph_ge >> gb) & 1u);
              ph_ge ^= 1u << gb;
            }
          };
          float d[RT], bias1[1];
          __nv_bfloat16 g16[RT];
          __nv_bfloat16* sg_ptr = nullptr;
          tmem_ld_nw<RT>(laT + col_dA + db * NR, d);
          tmem_ld_nw<1>(lane_base + col_bias + (uint32_t)t, bias1);
          if (!is_out) {
            const float ce = 0.5f * nd.c[lin], gc = nd.gc[lin];
            float xv[RT], gv[RT];
            tmem_ld_nw<RT>(laT + col_x + h * NR, xv);
            tmem_ld_wait();
            tmem_ld_tie(d);
            tmem_ld_tie(bias1);
            tmem_ld_tie(xv);
            TC_STAMP(stamp_thr && (k == 2 || k == 3), ts, 56);
            const float bias = bias1[0];
            __nv_bfloat16* sg = (do_save && uvalid) ? sg_row + p.sg_off[lin] + u : nullptr;
            // straight-line math for all RT chains (independent chains interleave); only the stores are predicated
#pragma unroll
            for (int i = 0; i < RT; ++i) {
              const float eps = xv[i] - (d[i] + bias);
              gv[i] = uvalid ? -gc * eps : 0.0f;
              e_part = fmaf((uvalid && i < nrT) ? ce * eps : 0.0f, eps, e_part);
            }
            wait_g_buffer();
#pragma unroll
            for (int i = 0; i < RT; ++i) g16[i] = __float2bfloat16(gv[i]);
            st_chains_bf16<RT>(gptr, gv);
            sg_ptr = sg;
            tmem_st<RT>(laT + col_g + h * NR, gv);
            tmem_st_wait();
          } else {
            if (y_in_tmem) tmem_ld_nw<RT>(laT + col_y + (uint32_t)(t - p.n_hid_tiles) * NR, yv);
            tmem_ld_wait();
            tmem_ld_tie(d);
            tmem_ld_tie(bias1);
            tmem_ld_tie(yv);
            TC_STAMP(stamp_thr && (k == 2 || k == 3), ts, 56);
            const float bias = bias1[0];
            __nv_bfloat16* sg = (do_save && uvalid) ? sg_row + p.sg_off[L] + u : nullptr;
            float* to = (do_traj && p.traj_out != nullptr && uvalid) ? p.traj_out + ((size_t)rec * p.B + rbT) * nd.d_out + u : nullptr;
            const bool bern = top_kind == MCPC_TOP_BERNOULLI;
            float ov[RT], ev[RT];
            if (bern) {
#pragma unroll
              for (int i = 0; i < RT; ++i) {
                const float o = d[i] + bias;
                const float y = yv[i];
                const bool on = (y == y);                      // NaN marks "no loss on this element"
                const float z = exp2_ftz(-1.4426950408889634f * fabsf(o));   // exp(-|o|): MUFU ex2 without the denormal wrapper
                const float lv = fmaxf(o, 0.0f) - o * y + __logf(1.0f + z);
                const float e = __fdividef(o >= 0.0f ? 1.0f : z, 1.0f + z) - y;
                l_part += on ? lv : 0.0f;
                ev[i] = on ? e : 0.0f;
                ov[i] = o;
              }
            } else {
#pragma unroll
              for (int i = 0; i < RT; ++i) {
                const float o = d[i] + bias;
                const float y = yv[i];
                const bool on = (y == y);
                const float dd = on ? o - y : 0.0f;            // (0 * NaN would poison the sum)
                l_part = fmaf(0.5f * nd.inv_var * dd, dd, l_part);
                ev[i] = dd * nd.inv_var;
                ov[i] = o;
              }
            }
            wait_g_buffer();
#pragma unroll
            for (int i = 0; i < RT; ++i) g16[i] = __float2bfloat16(ev[i]);
            if (has_b) st_chains_bf16<RT>(gptr, ev);
            sg_ptr = sg;
            if (to != nullptr) {
#pragma unroll
              for (int i = 0; i < RT; ++i)
                if (i < nrT) to[(size_t)i * nd.d_out] = ov[i];
            }
          }
          TC_STAMP(stamp_thr && (k == 2 || k == 3), ts, 57);
          fence_before_sync();
          warp_arrive(&bars.dA_empty[db], lane);
          TC_STAMP(stamp_thr && (k == 2 || k == 3), ts, 58);
          if (has_b) {
            fence_async_smem();
            warp_arrive(&bars.g_full[gb], lane);
          }
          // the saved dW operand goes out AFTER the hand-over: the proxy fence above waits for every earlier memory
          // operation of the thread, global stores included, so they must not sit in front of it
          if (sg_ptr != nullptr) {
#pragma unroll
            for (int i = 0; i < RT; ++i)
              if (i < nrT) sg_ptr[(size_t)i * p.sg_pitch] = g16[i];
          }
          TC_STAMP(stamp_thr && (k == 2 || k == 3), ts, 59);
          // own-layer errors of layer `lin` are complete after the last tile of Linear lin
          if ((last_hid >> t) & 1u) warp_arrive(&bars.g_ready[lin], lane);
          TC_STAMP(stamp_thr, ts, 22 + k);
        }
        e_part = warp_sum_tc(e_part);
        l_part = warp_sum_tc(l_part);
        float (*red)[2] = s_red[0][ts & 1];
        if (lane == 0) { red[gw][0] = e_part; red[gw][1] = l_part; }
        asm volatile("bar.sync 1, %0;" ::"n"(kTThreads) : "memory");
        if (gtid < 2) {
          float s = 0.0f;
#pragma unroll
          for (int w = 0; w < 4 * SUB; ++w) s += red[w][gtid];
          p.partials[(((size_t)ts * p.n_ctas + blockIdx.x) * 2 + 0) * 2 + gtid] = s;
        }
        if (gtid == 0 && do_save && p.ready != nullptr) saved_step(&s_saved[0]);
        if (nz_producer && ts + 1 < p.n_steps) produce_noise(ts + 1);
      }
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}


inline int pad16(int v) { return (v + 15) & ~15; }

// Fills the tile table + shared-memory plan.  Returns MCPC_OK or MCPC_ERR_UNSUPPORTED (message set).
int plan_tc(const NetDev& nd, int NR, TcParams* p, size_t* smem_bytes, size_t* packed_bytes, bool with_out = true,
            bool shrink_partial = true) {
  const int n_sub = 2;                                         // G operand buffers / accumulator pairs (SUB of the kernel)
  if (nd.L < 1) return MCPC_ERR_INVALID;
  int HT = 0;
  for (int l = 0; l < nd.L; ++l) {
    p->ut[l] = (nd.dims[l] + 127) / 128;
    p->h_off[l] = HT;
    for (int i = 0; i < p->ut[l]; ++i) {
      if (HT >= kMaxHT) {
        set_error("bf16 path: more than %d hidden unit tiles (128 units each)", kMaxHT);
        return MCPC_ERR_UNSUPPORTED;
      }
      p->h_layer[HT] = l;
      p->h_index[HT] = i;
      ++HT;
    }
  }
  p->h_off[nd.L] = HT;
  p->HT = HT;
  const int nda = (NR == 16) ? 2 * n_sub : 2;                  // prediction accumulators (kDA of the kernel)
  if ((nda + 3 * HT) * NR + 32 > 512) {
    set_error("bf16 path: %d hidden unit tiles x %d chains per CTA do not fit the 512 TMEM columns", HT, NR);
    return MCPC_ERR_UNSUPPORTED;
  }
  // operand buffers first, weights after them
  uint32_t off = 0;
  for (int l = 0; l < nd.L; ++l) {
    p->act_kp[l] = pad16(nd.dims[l]);
    p->act_off[l] = (int)off;
    off += (uint32_t)NR * p->act_kp[l] * 2;
  }
  for (int i = 0; i < n_sub; ++i) {
    p->gbuf_off[i] = (int)off;
    off += (uint32_t)NR * 128 * 2;
  }
  off = (off + 1023u) & ~1023u;
  // tile table, in the order a step visits them: Linear 1 ... L-1, then the output tiles (bottom-up).  The update of
  // the top hidden layer closes the step (it needs every output tile); the tiles of the lower Linears of the NEXT step
  // only need the lower layers' updates, which finished long before, so they overlap with it.
  int nt = 0;
  size_t gsrc = 0;
  uint32_t max_tile = 0;
  p->n_hid_tiles = 0;
  for (int lin = 1; lin <= nd.L; ++lin) {
    if (lin == nd.L && (nd.d_out == 0 || !with_out)) continue;   // !with_out: the output Linear is never visited (pure sampling)
    const int rows = (lin == nd.L) ? nd.d_out : nd.dims[lin];
    const int Kp = pad16(nd.dims[lin - 1]);
    if (Kp > 1024) {
      set_error("bf16 path: layer width %d too large for one weight tile", nd.dims[lin - 1]);
      return MCPC_ERR_UNSUPPORTED;
    }
    for (int i = 0; i < (rows + 127) / 128; ++i) {
      if (nt >= kMaxTiles) {
        set_error("bf16 path: more than %d weight tiles", kMaxTiles);
        return MCPC_ERR_UNSUPPORTED;
      }
      Tile& T = p->tiles[nt++];
      T.lin = lin;
      T.out_tile = i;
      T.Kp = Kp;
      T.sbo = (Kp / 8) * 128;
      T.bytes = 128 * Kp * 2;
      // Partial last tile of the OUTPUT Linear (784 = 6 x 128 + 16): in the canonical layout the 8-row groups are
      // contiguous, so only the valid rows (in groups of 16, the back-projection's K step) are copied and kept; the
      // prediction MMA (M = 128) reads whatever follows in shared memory into accumulator lanes nobody looks at (targets of
      // padding units are NaN = "no loss", the stores are predicated).  mcpc_ml: 28 KB that keep one more tile resident.
      if (shrink_partial && lin == nd.L && rows - i * 128 < 128) T.bytes = pad16(rows - i * 128) * Kp * 2;
      T.h_out = (lin < nd.L) ? p->h_off[lin] + i : -1;
      T.gsrc = gsrc;
      T.slot = -1;
      gsrc += (size_t)128 * Kp * 2;                                // the packed copy in global memory keeps whole tiles
      if ((uint32_t)T.bytes > max_tile) max_tile = (uint32_t)T.bytes;
    }
    if (lin < nd.L) p->n_hid_tiles = nt;
  }
  p->n_out_tiles = nt - p->n_hid_tiles;
  p->y_tmem = ((nda + 3 * HT) * NR + 32 + p->n_out_tiles * NR <= 512) ? 1 : 0;
  *packed_bytes = gsrc;
  // residency: everything if it fits, else as many leading tiles as fit beside a 2-slot ring
  const uint32_t slack = 4096;             // MN-major reads of narrow tiles overrun their 128 x Kp footprint
  uint32_t total = 0;
  for (int t = 0; t < nt; ++t) total += (uint32_t)p->tiles[t].bytes;
  // resident tiles: the shrunken partial tiles first, so that what the prediction MMA reads beyond them is weight data
  auto place_resident = [&](const bool* streamed) {
    for (int pass = 0; pass < 2; ++pass)
      for (int t = 0; t < nt; ++t) {
        Tile& T = p->tiles[t];
        if ((streamed != nullptr && streamed[t]) || (T.bytes < 128 * T.Kp * 2) != (pass == 0)) continue;
        T.smem_off = (int)off;
        off += (uint32_t)T.bytes;
      }
  };
  if (off + total + slack <= kSmemBudget) {
    place_resident(nullptr);
  } else {
    const uint32_t ring = 2 * max_tile;
    if (off + ring + slack > kSmemBudget) {
      set_error("bf16 path: weight tiles of %u B do not fit shared memory", max_tile);
      return MCPC_ERR_UNSUPPORTED;
    }
    const uint32_t ring_off = off;
    off += ring;
    bool streamed[kMaxTiles] = {};
    uint32_t res_bytes = 0;
    for (int t = 0; t < nt; ++t) {
      const Tile& T = p->tiles[t];
      if (off + res_bytes + (uint32_t)T.bytes + slack <= kSmemBudget) res_bytes += (uint32_t)T.bytes;
      else streamed[t] = true;
    }
    place_resident(streamed);
    // Visiting order of the OUTPUT tiles (any order is valid: their back-projections sum into the same accumulator):
    // streamed and resident tiles alternate, starting with a streamed one.  A streamed tile's bulk copy can only start
    // when the previous user of its ring slot has finished its back-projection; with the streamed tiles at the end of the
    // step (r01) the third one waited ~4,800 cycles for the first one's slot, on the step's critical path (cycle trace,
    // profiles/r02_tc_trace.txt).  Interleaved, a slot is free 3 tile times before it is needed again.
    if (p->n_out_tiles > 1) {
      Tile res[kMaxTiles], str[kMaxTiles];
      int n_res = 0, n_str = 0;
      for (int t = p->n_hid_tiles; t < nt; ++t) {
        if (streamed[t]) str[n_str++] = p->tiles[t];
        else res[n_res++] = p->tiles[t];
      }
      int t = p->n_hid_tiles, i_r = 0, i_s = 0;
      while (i_r < n_res || i_s < n_str) {
        if (i_s < n_str) { p->tiles[t] = str[i_s++]; streamed[t++] = true; }
        if (i_r < n_res) { p->tiles[t] = res[i_r++]; streamed[t++] = false; }
      }
    }
    // 2 ring slots, used round-robin in visiting order: a streamed tile finds its slot free again before the same slot
    // is needed twice in one step by construction
    int next_slot = 0;
    for (int t = 0; t < nt; ++t) {
      if (!streamed[t]) continue;
      Tile& T = p->tiles[t];
      T.slot = next_slot;
      T.smem_off = (int)(ring_off + next_slot * max_tile);
      next_slot ^= 1;
    }
  }
  *smem_bytes = off + slack;
  // the M = 128 prediction MMA of a shrunken tile must stay inside the allocation; otherwise plan with whole tiles
  for (int t = 0; t < nt; ++t)
    if ((size_t)p->tiles[t].smem_off + (size_t)128 * p->tiles[t].Kp * 2 > *smem_bytes) {
      if (!shrink_partial) return MCPC_ERR_INVALID;
      return plan_tc(nd, NR, p, smem_bytes, packed_bytes, with_out, false);
    }
  return MCPC_OK;
}

// Chains per CTA (RV) and MMA N extent (NR).  One wave of 8-chain CTAs while the batch allows it (the epilogues,
// not the MMAs, bound the step: more SMs beat fuller MMAs), 32-chain CTAs once 16-chain ones exceed two waves.
struct RowsChoice { int nr, rv; };
RowsChoice choose_rows(int B) {
  if (const char* env = getenv("MCPC_TC_ROWS")) {
    const int v = atoi(env);
    if (v == 8) return {16, 8};
    if (v == 16) return {16, 16};
    if (v == 32) return {32, 32};
  }
  if ((B + 7) / 8 <= 148) return {16, 8};
  if ((B + 15) / 16 >= 2 * 148) return {32, 32};
  return {16, 16};
}

}  // namespace

bool infer_tc_fits(const NetDev& nd, int B) {
  if (getenv("MCPC_FORCE_STREAMING") != nullptr) return false;     // testing hook: exercise the streaming path on small nets
  TcParams p{};
  size_t smem = 0, packed = 0;
  return plan_tc(nd, 16, &p, &smem, &packed) == MCPC_OK;
}

int infer_tc_workspace(const NetDev& nd, int B, int n_steps, size_t* bytes) {
  TcParams p{};
  size_t smem = 0, packed = 0;
  RowsChoice rc_ = choose_rows(B);
  int rc = plan_tc(nd, rc_.nr, &p, &smem, &packed);
  if (rc != MCPC_OK && rc_.nr == 32) {
    rc_ = {16, 16};
    rc = plan_tc(nd, rc_.nr, &p, &smem, &packed);
  }
  if (rc != MCPC_OK) return rc;
  const int n_ctas = (B + rc_.rv - 1) / rc_.rv;
  *bytes = ((packed + 255) & ~(size_t)255) + (((size_t)n_steps * n_ctas * 4 * sizeof(float) + 255) & ~(size_t)255) +
           (((size_t)n_steps * 2 * sizeof(float) + 255) & ~(size_t)255) +
           (((size_t)B * nd.dims[0] * sizeof(float) + 255) & ~(size_t)255) + (size_t)n_steps * sizeof(unsigned) + 512;
  return MCPC_OK;
}

int launch_infer_tc(const NetDev& nd, const McpcIO* io, const McpcOpts* o, int B, void* ws, size_t ws_bytes,
                    cudaStream_t stream) {
  TcParams p{};
  size_t smem = 0, packed = 0;
  RowsChoice rows = choose_rows(B);
  int rc = plan_tc(nd, rows.nr, &p, &smem, &packed);
  if (rc != MCPC_OK && rows.nr == 32) {
    rows = {16, 16};
    p = TcParams{};
    rc = plan_tc(nd, rows.nr, &p, &smem, &packed);
  }
  if (rc != MCPC_OK) return rc;
  size_t plan_smem = smem;                                     // end of the shared-memory plan in use (incl. its slack)
  p.net = nd;
  p.B = B;
  p.n_ctas = (B + rows.rv - 1) / rows.rv;
  size_t need = 0;
  infer_tc_workspace(nd, B, o->n_steps, &need);
  if (ws == nullptr || ws_bytes < need) {
    set_error("workspace too small: %zu B given, %zu B needed", ws_bytes, need);
    return MCPC_ERR_WORKSPACE;
  }
  uint8_t* wsb = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  p.packed = wsb;
  p.partials = reinterpret_cast<float*>(wsb + ((packed + 255) & ~(size_t)255));
  float* adam_tab = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p.partials) +
                                             (((size_t)o->n_steps * p.n_ctas * 4 * sizeof(float) + 255) & ~(size_t)255));
  p.adam_tab = adam_tab;
  float* mu0_buf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(adam_tab) +
                                            (((size_t)o->n_steps * 2 * sizeof(float) + 255) & ~(size_t)255));
  const size_t mu0_bytes = ((size_t)B * nd.dims[0] * sizeof(float) + 255) & ~(size_t)255;
  unsigned* ready_buf = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(mu0_buf) + mu0_bytes);
  if (io->inputs != nullptr) {
    // non-zero inputs: Linear_0's prediction is a per-chain constant of this launch (fp32, exact operands)
    float* mu0 = mu0_buf;
    mu0_kernel<<<dim3((nd.dims[0] + 63) / 64, (B + 63) / 64), 256, 0, stream>>>(io->inputs, io->W[0], io->b[0], B, nd.d_in,
                                                                               nd.dims[0], mu0);
    count_launch();
    p.mu0 = mu0;
  }
  if (o->optimizer == MCPC_OPT_ADAM && o->update_x) {
    adam_table_kernel<<<(o->n_steps + 127) / 128, 128, 0, stream>>>(o->adam_beta1, o->adam_beta2, o->lr, o->adam_step0, o->n_steps,
                                                                     adam_tab);
    count_launch();
  }
  // pack the weights (fp32 nn.Linear layout -> bf16 canonical tiles); they change once per learning step
  {
    PackJobs jobs{};
    int n_jobs = 0, max_blocks = 1;
    for (int t = 0; t < p.n_hid_tiles + p.n_out_tiles;) {
      const Tile& T = p.tiles[t];
      const int lin = T.lin;
      const int rows = (lin == nd.L) ? nd.d_out : nd.dims[lin];
      const int n_lin_tiles = (rows + 127) / 128;
      const int n_elems = n_lin_tiles * 128 * T.Kp;                   // 2 elements per thread
      size_t gbase = T.gsrc;                                          // packed tiles of a Linear are contiguous in out_tile
      for (int k = t; k < t + n_lin_tiles; ++k) gbase = std::min(gbase, p.tiles[k].gsrc);   // order; the table is in visiting order
      jobs.job[n_jobs++] = PackJob{io->W[lin], wsb + gbase, rows, nd.dims[lin - 1], T.Kp, n_lin_tiles};
      max_blocks = std::max(max_blocks, (n_elems + 511) / 512);
      t += n_lin_tiles;
    }
    if (n_jobs > 0) {
      pack_weights_kernel<<<dim3(max_blocks, n_jobs), 256, 0, stream>>>(jobs);
      count_launch();
    }
  }
  for (int l = 0; l <= nd.L; ++l) p.b[l] = io->b[l];
  for (int l = 0; l < nd.L; ++l) {
    p.x[l] = io->x[l];
    p.m[l] = io->adam_m[l];
    p.v[l] = io->adam_v[l];
    p.xgrad[l] = io->x_grad[l];
    p.traj_x[l] = io->traj_x[l];
  }
  p.traj_out = io->traj_out;
  p.save_g = reinterpret_cast<__nv_bfloat16*>(io->save_g);
  p.save_f = reinterpret_cast<__nv_bfloat16*>(io->save_f);
  {
    int f_off[kMaxL + 1];
    save_layout_bf16(nd, p.sg_off, &p.sg_pitch, f_off, &p.sf_pitch);
  }
  p.target = io->target;
  p.noise = io->noise;
  p.n_steps = o->n_steps;
  p.t_begin = o->t_begin;
  p.optimizer = o->optimizer;
  p.update_x = o->update_x;
  {
    const int nda = (rows.nr == 16) ? 4 : 2;
    const int used = (nda + 3 * p.HT) * rows.nr + 32 + (p.y_tmem ? p.n_out_tiles * rows.nr : 0);
    p.adam_tmem = (o->optimizer == MCPC_OPT_ADAM && o->update_x && io->adam_m[0] != nullptr &&
                   used + 2 * p.HT * rows.nr <= 512) ? 1 : 0;
  }
  p.lr = (float)o->lr;
  p.lr_d = o->lr;
  p.beta1 = o->adam_beta1;
  p.beta2 = o->adam_beta2;
  p.one_minus_b1 = (float)(1.0 - o->adam_beta1);
  p.one_minus_b2 = (float)(1.0 - o->adam_beta2);
  p.beta2f = (float)o->adam_beta2;
  p.adam_eps = (float)o->adam_eps;
  p.b1_pow0 = pow(o->adam_beta1, (double)o->adam_step0);
  p.b2_pow0 = pow(o->adam_beta2, (double)o->adam_step0);
  p.noise_mode = o->noise_mode;
  p.noise_scale = (float)o->noise_scale;
  p.seed = o->seed;
  p.chain_offset = o->chain_offset;
  bool any_traj = io->traj_out != nullptr;
  for (int l = 0; l < nd.L; ++l) any_traj = any_traj || io->traj_x[l] != nullptr;
  p.traj_every = any_traj ? (o->traj_every > 0 ? o->traj_every : 1) : 0;
  p.save_begin = o->save_begin;
  p.save_end = o->save_end;
  const bool timing = getenv("MCPC_TC_TIMING") != nullptr;     // debug only: allocates + synchronises
  if (timing) {
    p.dbg_t0 = atoi(getenv("MCPC_TC_TIMING"));
    cudaMalloc(&p.dbg, 8 * 64 * sizeof(long long));
    cudaMemsetAsync(p.dbg, 0, 8 * 64 * sizeof(long long), stream);
  }
  // ---- the weight update of the saved steps, when the caller passed accumulators (McpcIO.gW/gb) --------------------
  // Overlapped: B <= 1184 chains occupy 8 chains x <= 148 CTAs of one SM each; when at least one CTA per output tile of the
  // weight-gradient kernel fits the SMs left over, it is launched on a side stream NEXT to the inference kernel and
  // consumes each saved step as soon as every inference CTA has signalled it (TcParams.ready) -- its operands come from L2
  // and only the last step's tail is left when the inference kernel ends.  Otherwise it runs after the kernel.
  bool want_dw = false;
  for (int l = 0; l <= nd.L; ++l) want_dw = want_dw || io->gW[l] != nullptr || io->gb[l] != nullptr;
  want_dw = want_dw && io->save_g != nullptr && o->save_end > o->save_begin;
  McpcGradIO gio{};
  bool overlap = false;
  int free_sms = 0;
  const int n_save = o->save_end - o->save_begin;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  if (want_dw) {
    gio.save_g = io->save_g;
    gio.save_f = io->save_f;
    gio.inputs = io->inputs;
    for (int l = 0; l <= nd.L; ++l) {
      gio.gW[l] = io->gW[l];
      gio.gb[l] = io->gb[l];
    }
    gio.scratch = mu0_buf;                                       // free again once the inference kernel has finished
    gio.scratch_bytes = mu0_bytes;
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    free_sms = n_sm - p.n_ctas;
    const int nt = weight_grad_tc_tiles(nd, &gio);
    const char* e = getenv("MCPC_TC_DW_OVERLAP");                // opt-in, see infer_tc_overlaps_weight_grad
    overlap = io->inputs == nullptr && nt > 0 && free_sms >= nt && (e != nullptr && e[0] == '1') && !timing && dev < 16;
    if (overlap) {
      static cudaStream_t s_side[16] = {};
      static cudaEvent_t s_fork[16] = {}, s_join[16] = {};
      if (s_side[dev] == nullptr) {
        MCPC_CUDA_CHECK(cudaStreamCreateWithFlags(&s_side[dev], cudaStreamNonBlocking));
        MCPC_CUDA_CHECK(cudaEventCreateWithFlags(&s_fork[dev], cudaEventDisableTiming));
        MCPC_CUDA_CHECK(cudaEventCreateWithFlags(&s_join[dev], cudaEventDisableTiming));
      }
      side = s_side[dev];
      ev_fork = s_fork[dev];
      ev_join = s_join[dev];
      p.ready = ready_buf;
      MCPC_CUDA_CHECK(cudaMemsetAsync(ready_buf, 0, (size_t)n_save * sizeof(unsigned), stream));
      // everything the consumer depends on besides the flags (zeroed accumulators, the flags' reset) is in `stream` by now
      MCPC_CUDA_CHECK(cudaEventRecord(ev_fork, stream));
      MCPC_CUDA_CHECK(cudaStreamWaitEvent(side, ev_fork, 0));
    }
  }
  auto launch = [&](auto kernel) -> int {
    p.tab_off = (int)((plan_smem - kTabBytes) & ~(size_t)15);    // inside the slack at the end of the PLAN (see s_mA)
    MCPC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<p.n_ctas, 640, smem, stream>>>(p);
    return MCPC_OK;
  };
  bool plain = (p.traj_every == 0) && o->update_x;           // what every specialisation assumes
  for (int l = 0; l < nd.L; ++l) plain = plain && io->x_grad[l] == nullptr && p.ut[l] == 1;
  plain = plain && (p.y_tmem || !nd.top_has_grad) && (p.adam_tmem || o->optimizer != MCPC_OPT_ADAM);
  const bool bern_grad = nd.top == MCPC_TOP_BERNOULLI && nd.top_has_grad;
  const bool sgd_philox = o->optimizer == MCPC_OPT_SGD && o->noise_mode == MCPC_NOISE_PHILOX;
  int spec = 0;
  if (getenv("MCPC_TC_NOSPEC") != nullptr) plain = false;      // testing hook: the generic instantiation
  if (p.mu0 != nullptr) plain = false;                         // non-zero inputs: generic instantiation only
  if (plain && bern_grad && sgd_philox) spec = 1;
  else if (plain && bern_grad && o->optimizer == MCPC_OPT_ADAM && o->noise_mode == MCPC_NOISE_NONE) spec = 2;
  else if (plain && !nd.top_has_grad && sgd_philox) spec = 3;
  else if (!nd.top_has_grad && sgd_philox && o->update_x && p.mu0 == nullptr && getenv("MCPC_TC_NOSPEC") == nullptr) {
    // sampling with thinned read-outs (SURVEY 8f N2 / C3 "read out every k-th step"): instantiation 4 = 3 + trajectory
    // records (it visits the output tiles on the recorded steps only); kept apart from 3, whose pure-sampling loop is 4 %
    // faster without the recording code (C3: 181 vs 190 us per step)
    bool ok4 = true;
    for (int l = 0; l < nd.L; ++l) ok4 = ok4 && io->x_grad[l] == nullptr && p.ut[l] == 1;
    if (ok4) spec = 4;
  }
  p.nz_off = -1;
  if (spec == 3 && getenv("MCPC_TC_NOISE_BY_U") == nullptr) {
    // pure sampling: no output tile is ever visited -- plan without them (everything resident) and give the shared memory
    // to the two noise buffers group T fills one step ahead (see infer_tc_kernel); if they do not fit, group U keeps
    // drawing its own noise
    TcParams q = p;
    size_t smem3 = 0, packed3 = 0;
    if (plan_tc(nd, rows.nr, &q, &smem3, &packed3, false) == MCPC_OK) {
      const size_t nz_bytes = (size_t)2 * q.HT * rows.rv * 128 * sizeof(float);
      const size_t off = (smem3 + 127) & ~(size_t)127;
      if (off + nz_bytes <= kSmemBudget) {
        for (int t = 0; t < kMaxTiles; ++t) p.tiles[t] = q.tiles[t];
        p.n_hid_tiles = q.n_hid_tiles;
        p.n_out_tiles = q.n_out_tiles;
        p.y_tmem = q.y_tmem;
        p.nz_off = (int)off;
        plan_smem = smem3;
        smem = off + nz_bytes;
      }
    }
  }
  if (rows.rv == 32) {
    rc = spec == 3 ? launch(infer_tc_kernel<32, 32, false, 3>)
                   : (spec == 4 ? launch(infer_tc_kernel<32, 32, false, 4>) : launch(infer_tc_kernel<32, 32, false, 0>));
  } else if (rows.rv == 16) {
    rc = spec == 1 ? launch(infer_tc_kernel<16, 16, false, 1>) : launch(infer_tc_kernel<16, 16, false, 0>);
  } else if (timing) {
    rc = launch(infer_tc_kernel<16, 8, true, 0>);    // the cycle trace exists for the generic 8-chain variant only
  } else if (spec == 1) {
    rc = launch(infer_tc_kernel<16, 8, false, 1>);
  } else if (spec == 2) {
    rc = launch(infer_tc_kernel<16, 8, false, 2>);
  } else if (spec == 3) {
    rc = launch(infer_tc_kernel<16, 8, false, 3>);
  } else if (spec == 4) {
    rc = launch(infer_tc_kernel<16, 8, false, 4>);
  } else {
    rc = launch(infer_tc_kernel<16, 8, false, 0>);
  }
  if (rc != MCPC_OK) return rc;
  MCPC_CUDA_CHECK(cudaGetLastError());
  count_launch();
  if (want_dw && overlap) {
    rc = launch_weight_grad_tc_overlapped(nd, &gio, B, n_save, ready_buf, (unsigned)p.n_ctas, free_sms, side);
    if (rc != MCPC_OK) return rc;
    MCPC_CUDA_CHECK(cudaEventRecord(ev_join, side));
  }
  if (timing) {
    long long h[8 * 64];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    for (int ts = 0; ts < 8 && ts + p.dbg_t0 < o->n_steps; ++ts) {
      const long long base = h[ts * 64 + 32];
      fprintf(stderr, "[tc timing] step %d (abs %lld; cycles rel. to epilogue step start):", ts + p.dbg_t0, base);
      for (int i = 0; i < 64; ++i)
        if (h[ts * 64 + i] != 0) fprintf(stderr, " %d:%lld", i, h[ts * 64 + i] - base);
      fprintf(stderr, "\n");
    }
  }
  if (io->energy != nullptr || io->loss != nullptr) {
    rc = launch_reduce_partials(p.partials, o->n_steps, 2 * p.n_ctas, io->energy, io->loss, stream);
    if (rc != MCPC_OK) return rc;
  }
  if (want_dw) {
    if (overlap) {
      MCPC_CUDA_CHECK(cudaStreamWaitEvent(stream, ev_join, 0));      // the caller's stream continues after the weight update
    } else {
      rc = launch_weight_grad_tc(nd, &gio, B, n_save, stream);
      if (rc != MCPC_OK) return rc;
    }
  }
  return MCPC_OK;
}

// Will mcpc_infer overlap the weight update with the inference kernel for this (net, B)?  (all accumulators given)
bool infer_tc_overlaps_weight_grad(const NetDev& nd, int B, bool has_inputs) {
  if (has_inputs) return false;
  // Opt-in (MCPC_TC_DW_OVERLAP=1).  Measured on C2 (B = 1024, T = 150, r02): the weight update then ends 10 us after the
  // inference kernel instead of 85 us, but a caller that reads the results of every call is bound by its own latency
  // from "scalars ready" to "next kernel launched" (~115 us of Python), which the sequential order hides behind the
  // weight-gradient and optimizer kernels: 1.23 ms per call overlapped vs 1.17 ms sequential.  It pays only for callers
  // that enqueue calls without reading results in between.
  const char* e = getenv("MCPC_TC_DW_OVERLAP");
  if (e == nullptr || e[0] != '1') return false;
  McpcGradIO gio{};
  float* dummy = reinterpret_cast<float*>(0x100);
  for (int l = 0; l <= nd.L; ++l) {
    gio.gW[l] = dummy;
    gio.gb[l] = dummy;
  }
  const int nt = weight_grad_tc_tiles(nd, &gio);
  int dev = 0, n_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return false;
  const RowsChoice rows = choose_rows(B);
  const int n_ctas = (B + rows.rv - 1) / rows.rv;
  return nt > 0 && n_sm - n_ctas >= nt && dev < 16;
}

}  // namespace mcpc
