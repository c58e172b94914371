// mcpc_infer, MCPC_PREC_FP32: reference-exact persistent Langevin/PC inference kernel.
//
// One CTA owns R chains (batch rows) for ALL n_steps steps.  Their latents x, activations
// act(x) and error signals G live in shared memory for the whole launch; weights are read
// through L1/L2 (coalesced: phase A reads a transposed copy, phase B the nn.Linear layout).
// Per step (reference predictive_coding/pc_trainer.py:733-918):
//   phase A  every (layer, unit) column in parallel (Jacobi):
//            mu = act(x_below) W^T + b          utils/model.py:54-65 (nn.Linear forward)
//            eps = x - mu, E += c/2 eps^2       pc_layer.py:17-18,272,295
//            out -> loss, e_out = dloss/dout    utils/model.py:17-33
//   phase B  every latent unit in parallel:
//            g = coef*c*eps - act'(x) * (G_above W_above)     autograd of pc_trainer.py:862
//            x <- SGD / Adam step                             pc_trainer.py:877
//            x <- x - lr * noise                              utils/model.py:42-44 (random_step)
// fp32 FMA on CUDA cores throughout, accurate expf/log1pf/tanhf: this is the mode the 1e-5
// per-step parity tests run in.  The tensor-core path lives in infer_tc.cu.
#include <cstdlib>

#include "mcpc_common.cuh"
#include "philox.cuh"

namespace mcpc {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct RowsParams {
  NetDev net;
  const float* W[kMaxL + 1];
  const float* WT[kMaxL + 1];
  const float* b[kMaxL + 1];
  float* x[kMaxL];
  float* m[kMaxL];
  float* v[kMaxL];
  float* xgrad[kMaxL];
  float* traj_x[kMaxL];
  float* traj_out;
  float* save_g;
  float* save_f;
  const float* inputs;
  const float* target;
  const float* noise;
  float* partials;   // [n_steps][n_tiles][2]
  int B, n_tiles, n_steps, t_begin;
  int optimizer, update_x;
  float lr, adam_eps, one_minus_b1, one_minus_b2, beta2f;
  double lr_d, beta1, beta2, b1_pow0, b2_pow0;
  int noise_mode;
  float noise_scale;
  uint64_t seed, chain_offset;
  int traj_every, save_begin, save_end;
  // shared-memory layout (floats): per row [x: SDp][a: d_in_p + SDp][g: SDp + d_out_p]
  int poff[kMaxL + 1];   // layer offsets padded to 4 floats; poff[L] = SDp
  int d_in_p, d_out_p;
  int x_stride, a_stride, g_stride;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// acc[r] += sum_{k in [k0,K)} a[r*astride + k] * w[k*wstride]   (k0 a multiple of 4).
// The weight column is strided in memory (one element per k, coalesced ACROSS the warp), so its loads are
// software-pipelined 8 deep: the next 8 weights are in flight while the current 8 feed R*8 FMAs.
template <int R>
__device__ __forceinline__ void dot_strided(const float* __restrict__ w, size_t wstride, int k0, int K,
                                            const float* a, int astride, float (&acc)[R]) {
  int k = k0;
  if (k + 8 <= K) {
    float wc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) wc[i] = __ldg(w + (size_t)(k + i) * wstride);
    for (; k + 8 <= K; k += 8) {
      float wn[8];
      const bool more = (k + 16 <= K);
#pragma unroll
      for (int i = 0; i < 8; ++i) wn[i] = more ? __ldg(w + (size_t)(k + 8 + i) * wstride) : 0.0f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 a0 = *reinterpret_cast<const float4*>(a + r * astride + k);
        const float4 a1 = *reinterpret_cast<const float4*>(a + r * astride + k + 4);
        float s = acc[r];
        s = fmaf(a0.x, wc[0], s); s = fmaf(a0.y, wc[1], s); s = fmaf(a0.z, wc[2], s); s = fmaf(a0.w, wc[3], s);
        s = fmaf(a1.x, wc[4], s); s = fmaf(a1.y, wc[5], s); s = fmaf(a1.z, wc[6], s); s = fmaf(a1.w, wc[7], s);
        acc[r] = s;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) wc[i] = wn[i];
    }
  }
  for (; k < K; ++k) {
    const float w0 = __ldg(w + (size_t)k * wstride);
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = fmaf(a[r * astride + k], w0, acc[r]);
  }
}

template <int R>
__global__ void __launch_bounds__(kThreads) infer_rows_kernel(const __grid_constant__ RowsParams p) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                                  // [R][x_stride]
  float* as = xs + R * p.x_stride;                   // [R][a_stride]   (inputs | act(x_0) | act(x_1) ...)
  float* gs = as + R * p.a_stride;                   // [R][g_stride]   (G_0 | G_1 | ... | e_out)
  __shared__ float s_red[kWarps][2];

  const NetDev& nd = p.net;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int row0 = tile * R;
  const int nvalid = min(R, p.B - row0);
  const int L = nd.L;

  // ---- load the tile: latents, their activations, the constant bottom activations (inputs)
  for (int u = tid; u < nd.SD; u += kThreads) {
    int l = 0;
    while (u >= nd.off[l + 1]) ++l;
    const int k = u - nd.off[l];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float xv = (r < nvalid) ? p.x[l][(size_t)(row0 + r) * nd.dims[l] + k] : 0.0f;
      xs[r * p.x_stride + p.poff[l] + k] = xv;
      as[r * p.a_stride + p.d_in_p + p.poff[l] + k] = act_apply(nd.act[l], xv);
    }
  }
  for (int k = tid; k < p.d_in_p; k += kThreads) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      as[r * p.a_stride + k] =
          (p.inputs != nullptr && k < nd.d_in && r < nvalid) ? p.inputs[(size_t)(row0 + r) * nd.d_in + k] : 0.0f;
  }
  // zero the padding lanes of a/g rows once so vectorised reads never see garbage
  for (int i = tid; i < R * p.g_stride; i += kThreads) gs[i] = 0.0f;
  __syncthreads();

  double b1p = p.b1_pow0, b2p = p.b2_pow0;

  for (int ts = 0; ts < p.n_steps; ++ts) {
    const int t_abs = p.t_begin + ts;
    const bool do_save = (p.save_g != nullptr) && ts >= p.save_begin && ts < p.save_end;
    const int slot = ts - p.save_begin;
    const bool do_traj = (p.traj_every > 0) && (ts % p.traj_every == 0);
    const int rec = do_traj ? ts / p.traj_every : 0;

    // ================= phase A: predictions, errors, energy, loss =================
    float e_part = 0.0f, l_part = 0.0f;
    // the output Linear only matters when its loss has a gradient or its value is recorded
    const int n_cols = (nd.top_has_grad || (do_traj && p.traj_out != nullptr)) ? nd.NG : nd.SD;
    for (int c = tid; c < n_cols; c += kThreads) {
      int l = 0;
      while (l < L && c >= nd.off[l + 1]) ++l;
      const bool is_out = (l == L);
      const int j = c - (is_out ? nd.SD : nd.off[l]);
      const int K = (l == 0) ? nd.d_in : nd.dims[l - 1];
      const int N = is_out ? nd.d_out : nd.dims[l];
      const float bias = (p.b[l] != nullptr) ? __ldg(p.b[l] + j) : 0.0f;
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = bias;
      if (l > 0 || p.inputs != nullptr) {
        const float* __restrict__ wt = p.WT[l] + j;                       // [K][N]
        const float* a = as + ((l == 0) ? 0 : p.d_in_p + p.poff[l - 1]);
        dot_strided<R>(wt, (size_t)N, 0, K, a, p.a_stride, acc);
      }
      if (!is_out) {
        const float ce = 0.5f * nd.c[l], gc = nd.gc[l];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float eps = xs[r * p.x_stride + p.poff[l] + j] - acc[r];
          const float G = -gc * eps;                       // d overall / d mu_l
          gs[r * p.g_stride + p.poff[l] + j] = G;
          if (r < nvalid) {
            e_part = fmaf(ce * eps, eps, e_part);
            if (do_save) p.save_g[((size_t)slot * p.B + row0 + r) * nd.NG + c] = G;
          }
        }
      } else {
        const bool in_mask = j >= nd.mask_start;
        float yv[R];
        if (in_mask && nd.top >= MCPC_TOP_GAUSS) {       // independent loads, not serialised behind the stores below
#pragma unroll
          for (int r = 0; r < R; ++r) yv[r] = (r < nvalid) ? __ldg(p.target + (size_t)(row0 + r) * nd.d_out + j) : 0.0f;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float o = acc[r];
          float e_out = 0.0f;
          if (r < nvalid) {
            if (in_mask && nd.top >= MCPC_TOP_GAUSS) {
              const float y = yv[r];
              if (nd.top == MCPC_TOP_GAUSS) {
                const float d = o - y;
                l_part = fmaf(0.5f * nd.inv_var * d, d, l_part);
                e_out = d * nd.inv_var;
              } else {
                const float z = expf(-fabsf(o));
                l_part += fmaxf(o, 0.0f) - o * y + log1pf(z);
                const float s = (o >= 0.0f ? 1.0f : z) / (1.0f + z);
                e_out = s - y;
              }
            }
            if (do_traj && p.traj_out != nullptr) p.traj_out[((size_t)rec * p.B + row0 + r) * nd.d_out + j] = o;
            if (do_save) p.save_g[((size_t)slot * p.B + row0 + r) * nd.NG + c] = e_out;
          }
          gs[r * p.g_stride + p.poff[L] + j] = e_out;
        }
      }
    }
    e_part = warp_sum(e_part);
    l_part = warp_sum(l_part);
    if (lane == 0) { s_red[warp][0] = e_part; s_red[warp][1] = l_part; }
    __syncthreads();
    if (tid < 2) {
      float s = 0.0f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += s_red[w][tid];
      p.partials[((size_t)ts * p.n_tiles + tile) * 2 + tid] = s;
    }

    // ================= phase B: back-projection, latent gradient, update =================
    const bool last = (ts == p.n_steps - 1);
    float step_size = 0.0f, bc2_sqrt = 1.0f;
    if (p.optimizer == MCPC_OPT_ADAM && p.update_x) {
      b1p *= p.beta1;
      b2p *= p.beta2;
      step_size = (float)(p.lr_d / (1.0 - b1p));
      bc2_sqrt = (float)sqrt(1.0 - b2p);
    }
    for (int u = tid; u < nd.SD; u += kThreads) {
      int l = 0;
      while (u >= nd.off[l + 1]) ++l;
      const int k = u - nd.off[l];
      const int dl = nd.dims[l];
      float bp[R];
#pragma unroll
      for (int r = 0; r < R; ++r) bp[r] = 0.0f;
      const bool has_above = (l + 1 < L) || nd.top_has_grad;
      if (has_above) {
        const int Nup = (l + 1 < L) ? nd.dims[l + 1] : nd.d_out;
        const float* __restrict__ w = p.W[l + 1] + k;                     // [Nup][dl]
        const float* g = gs + p.poff[l + 1];
        const int j0 = (l + 1 == L) ? (nd.mask_start & ~3) : 0;           // masked-out e_out are exactly 0
        dot_strided<R>(w, (size_t)dl, j0, Nup, g, p.g_stride, bp);
      }
      float nrm[4];
      uint64_t cur_q = ~0ull;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r >= nvalid) break;
        const int row = row0 + r;
        float* xp = xs + r * p.x_stride + p.poff[l] + k;
        float* ap = as + r * p.a_stride + p.d_in_p + p.poff[l] + k;
        float xv = *xp;
        const float av = *ap;
        const float grad = fmaf(act_deriv(nd.act[l], xv, av), bp[r], -gs[r * p.g_stride + p.poff[l] + k]);
        if (do_traj && p.traj_x[l] != nullptr) p.traj_x[l][((size_t)rec * p.B + row) * dl + k] = xv;
        if (do_save) p.save_f[((size_t)slot * p.B + row) * nd.SD + u] = av;
        if (last && p.xgrad[l] != nullptr) p.xgrad[l][(size_t)row * dl + k] = grad;
        if (p.update_x) {
          if (p.optimizer == MCPC_OPT_SGD) {
            xv = fmaf(-p.lr, grad, xv);
          } else {
            const size_t si = (size_t)row * dl + k;
            float mv = p.m[l][si], vv = p.v[l][si];
            mv = fmaf(p.one_minus_b1, grad - mv, mv);        // exp_avg.lerp_(grad, 1-beta1)
            vv = vv * p.beta2f;                               // exp_avg_sq.mul_(beta2)
            vv = fmaf(p.one_minus_b2 * grad, grad, vv);       //   .addcmul_(grad, grad, value=1-beta2)
            p.m[l][si] = mv;
            p.v[l][si] = vv;
            const float denom = sqrtf(vv) / bc2_sqrt + p.adam_eps;
            xv = fmaf(-step_size, mv / denom, xv);
          }
        }
        if (p.noise_mode == MCPC_NOISE_SUPPLIED) {
          xv = fmaf(-p.lr, p.noise[((size_t)ts * p.B + row) * nd.SD + u], xv);
        } else if (p.noise_mode == MCPC_NOISE_PHILOX) {
          const uint64_t chain = p.chain_offset + (uint64_t)row;
          if ((chain >> 2) != cur_q) {
            cur_q = chain >> 2;
            langevin_normals4(p.seed, (uint32_t)u, (uint32_t)t_abs, cur_q, nrm);
          }
          xv = fmaf(-p.lr, p.noise_scale * nrm[chain & 3], xv);
        }
        *xp = xv;
        *ap = act_apply(nd.act[l], xv);
      }
    }
    __syncthreads();
  }

  // ---- write the latents back (PCLayer._x storage is updated in place)
  for (int u = tid; u < nd.SD; u += kThreads) {
    int l = 0;
    while (u >= nd.off[l + 1]) ++l;
    const int k = u - nd.off[l];
    for (int r = 0; r < nvalid; ++r) p.x[l][(size_t)(row0 + r) * nd.dims[l] + k] = xs[r * p.x_stride + p.poff[l] + k];
  }
}

// out[c][r] = in[r][c]
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = by + i, c = bx + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * cols + c] : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = bx + i, r = by + threadIdx.x;
    if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partials, int n_steps, int n_tiles,
                                       double* __restrict__ energy, double* __restrict__ loss) {
  const int t = blockIdx.x * blockDim.y + threadIdx.y;
  if (t >= n_steps) return;
  double e = 0.0, l = 0.0;
  for (int i = threadIdx.x; i < n_tiles; i += 32) {
    e += (double)partials[((size_t)t * n_tiles + i) * 2 + 0];
    l += (double)partials[((size_t)t * n_tiles + i) * 2 + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e += __shfl_xor_sync(0xffffffffu, e, o);
    l += __shfl_xor_sync(0xffffffffu, l, o);
  }
  if (threadIdx.x == 0) {
    if (energy != nullptr) energy[t] = e;
    if (loss != nullptr) loss[t] = l;
  }
}

__global__ void fill_noise_kernel(uint64_t seed, int t_begin, int n_steps, uint64_t chain_offset, int B, int n_units,
                                  float noise_scale, float* __restrict__ out) {
  const size_t total = (size_t)n_steps * B * n_units;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int u = (int)(i % n_units);
    const size_t rb = i / n_units;
    const int row = (int)(rb % B);
    const int ts = (int)(rb / B);
    const uint64_t chain = chain_offset + (uint64_t)row;
    float nrm[4];
    langevin_normals4(seed, (uint32_t)u, (uint32_t)(t_begin + ts), chain >> 2, nrm);
    out[i] = noise_scale * nrm[chain & 3];
  }
}

inline int pad4(int v) { return (v + 3) & ~3; }

struct Layout {
  int R, n_tiles;
  size_t smem_bytes;
  RowsParams p;
};

int choose_rows(int B, size_t floats_per_row, size_t smem_limit) {
  const int cand[5] = {16, 8, 4, 2, 1};
  if (const char* env = getenv("MCPC_ROWS")) {       // tuning override: rows per CTA
    const int R = atoi(env);
    for (int i = 0; i < 5; ++i)
      if (cand[i] == R && (size_t)R * floats_per_row * sizeof(float) <= smem_limit) return R;
  }
  for (int i = 0; i < 5; ++i) {
    const int R = cand[i];
    if ((size_t)R * floats_per_row * sizeof(float) > smem_limit) continue;
    if (R == 1 || (B + R - 1) / R >= 120) return R;
  }
  return 0;
}

size_t wt_floats(const NetDev& nd) {
  size_t n = 0;
  for (int l = 0; l <= nd.L; ++l) {
    if (l == nd.L && nd.d_out == 0) break;
    const int K = (l == 0) ? nd.d_in : nd.dims[l - 1];
    const int N = (l == nd.L) ? nd.d_out : nd.dims[l];
    n += (size_t)pad4(K * N);
  }
  return n;
}

void fill_layout(const NetDev& nd, RowsParams* p) {
  int o = 0;
  for (int l = 0; l < nd.L; ++l) {
    p->poff[l] = o;
    o += pad4(nd.dims[l]);
  }
  p->poff[nd.L] = o;
  p->d_in_p = pad4(nd.d_in);
  p->d_out_p = pad4(nd.d_out);
  p->x_stride = o;
  p->a_stride = p->d_in_p + o;
  p->g_stride = o + p->d_out_p;
}

constexpr size_t kSmemLimit = 200 * 1024;

}  // namespace

int infer_rows_workspace(const NetDev& nd, int B, int n_steps, size_t* bytes) {
  RowsParams p;
  fill_layout(nd, &p);
  const size_t fpr = (size_t)p.x_stride + p.a_stride + p.g_stride;
  const int R = choose_rows(B, fpr, kSmemLimit);
  if (R == 0) {
    set_error("fp32 resident kernel: one chain needs %zu B of shared memory (> %zu B); "
              "use the wide (non-resident) path", fpr * sizeof(float), kSmemLimit);
    return MCPC_ERR_UNSUPPORTED;
  }
  const int n_tiles = (B + R - 1) / R;
  *bytes = (wt_floats(nd) + (size_t)n_steps * n_tiles * 2) * sizeof(float) + 256;
  return MCPC_OK;
}

template <int R>
static int launch_R(const RowsParams& p, size_t smem, cudaStream_t stream) {
  MCPC_CUDA_CHECK(cudaFuncSetAttribute(infer_rows_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  infer_rows_kernel<R><<<p.n_tiles, kThreads, smem, stream>>>(p);
  MCPC_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return MCPC_OK;
}

int launch_infer_rows(const NetDev& nd, const McpcIO* io, const McpcOpts* o, int B, void* ws, size_t ws_bytes,
                      cudaStream_t stream) {
  RowsParams p{};
  p.net = nd;
  fill_layout(nd, &p);
  const size_t fpr = (size_t)p.x_stride + p.a_stride + p.g_stride;
  const int R = choose_rows(B, fpr, kSmemLimit);
  if (R == 0) {
    set_error("fp32 resident kernel: network too wide for shared memory (%zu B per chain)", fpr * sizeof(float));
    return MCPC_ERR_UNSUPPORTED;
  }
  p.B = B;
  p.n_tiles = (B + R - 1) / R;
  size_t need = 0;
  infer_rows_workspace(nd, B, o->n_steps, &need);
  if (ws == nullptr || ws_bytes < need) {
    set_error("workspace too small: %zu B given, %zu B needed", ws_bytes, need);
    return MCPC_ERR_WORKSPACE;
  }
  // workspace: transposed weights, then per-tile scalar partials
  float* wsf = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  size_t cursor = 0;
  const int n_lin = nd.L + (nd.d_out > 0 ? 1 : 0);
  for (int l = 0; l < n_lin; ++l) {
    const int K = (l == 0) ? nd.d_in : nd.dims[l - 1];
    const int N = (l == nd.L) ? nd.d_out : nd.dims[l];
    p.W[l] = io->W[l];
    p.b[l] = io->b[l];
    float* wt = wsf + cursor;
    cursor += (size_t)pad4(K * N);
    p.WT[l] = wt;
    if (l == 0 && io->inputs == nullptr) continue;     // mu_0 = b_0: Linear_0's weight is never read
    dim3 grid((K + 31) / 32, (N + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, stream>>>(io->W[l], wt, N, K);   // W [N][K] -> WT [K][N]
    count_launch();
  }
  p.partials = wsf + cursor;
  for (int l = 0; l < nd.L; ++l) {
    p.x[l] = io->x[l];
    p.m[l] = io->adam_m[l];
    p.v[l] = io->adam_v[l];
    p.xgrad[l] = io->x_grad[l];
    p.traj_x[l] = io->traj_x[l];
  }
  p.traj_out = io->traj_out;
  p.save_g = reinterpret_cast<float*>(io->save_g);
  p.save_f = reinterpret_cast<float*>(io->save_f);
  p.inputs = io->inputs;
  p.target = io->target;
  p.noise = io->noise;
  p.n_steps = o->n_steps;
  p.t_begin = o->t_begin;
  p.optimizer = o->optimizer;
  p.update_x = o->update_x;
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
  const size_t smem = (size_t)R * fpr * sizeof(float);
  int rc = MCPC_OK;
  switch (R) {
    case 16: rc = launch_R<16>(p, smem, stream); break;
    case 8: rc = launch_R<8>(p, smem, stream); break;
    case 4: rc = launch_R<4>(p, smem, stream); break;
    case 2: rc = launch_R<2>(p, smem, stream); break;
    default: rc = launch_R<1>(p, smem, stream); break;
  }
  if (rc != MCPC_OK) return rc;
  if (io->energy != nullptr || io->loss != nullptr) {
    dim3 block(32, 4);
    reduce_partials_kernel<<<(o->n_steps + 3) / 4, block, 0, stream>>>(p.partials, o->n_steps, p.n_tiles, io->energy,
                                                                      io->loss);
    MCPC_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  return MCPC_OK;
}

int launch_reduce_partials(const float* partials, int n_steps, int n_tiles, double* energy, double* loss,
                           cudaStream_t stream) {
  dim3 block(32, 4);
  reduce_partials_kernel<<<(n_steps + 3) / 4, block, 0, stream>>>(partials, n_steps, n_tiles, energy, loss);
  MCPC_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return MCPC_OK;
}

int launch_fill_noise(uint64_t seed, int t_begin, int n_steps, uint64_t chain_offset, int B, int n_units,
                      float noise_scale, float* out, cudaStream_t stream) {
  const size_t total = (size_t)n_steps * B * n_units;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  fill_noise_kernel<<<blocks > 0 ? blocks : 1, 256, 0, stream>>>(seed, t_begin, n_steps, chain_offset, B, n_units,
                                                                noise_scale, out);
  MCPC_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return MCPC_OK;
}

}  // namespace mcpc
