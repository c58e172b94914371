// extern "C" entry points of libmcpc_b200.so (declared in include/mcpc_b200.h).
// Argument validation + dispatch only; the kernels live in infer_rows.cu / weight_grad.cu /
// infer_tc.cu.  No entry point allocates, frees or synchronises.
#include <atomic>
#include <cstring>

#include "mcpc_common.cuh"

namespace mcpc {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_net(const McpcNet* net, NetDev* nd) {
  if (net == nullptr) {
    set_error("net is NULL");
    return MCPC_ERR_INVALID;
  }
  if (net->n_layers < 1 || net->n_layers > kMaxL) {
    set_error("n_layers=%d outside [1,%d]", net->n_layers, kMaxL);
    return MCPC_ERR_INVALID;
  }
  if (net->d_in < 1 || net->d_out < 0) {
    set_error("bad d_in=%d / d_out=%d", net->d_in, net->d_out);
    return MCPC_ERR_INVALID;
  }
  if (net->top < MCPC_TOP_NONE || net->top > MCPC_TOP_BERNOULLI) {
    set_error("unknown top=%d", net->top);
    return MCPC_ERR_INVALID;
  }
  std::memset(nd, 0, sizeof(*nd));
  nd->L = net->n_layers;
  nd->d_in = net->d_in;
  nd->d_out = net->d_out;
  int o = 0;
  for (int l = 0; l < nd->L; ++l) {
    if (net->dims[l] < 1) {
      set_error("dims[%d]=%d", l, net->dims[l]);
      return MCPC_ERR_INVALID;
    }
    if (net->act[l] < MCPC_ACT_IDENTITY || net->act[l] > MCPC_ACT_TANH) {
      set_error("unknown act[%d]=%d", l, net->act[l]);
      return MCPC_ERR_INVALID;
    }
    nd->dims[l] = net->dims[l];
    nd->act[l] = net->act[l];
    nd->off[l] = o;
    o += net->dims[l];
    nd->c[l] = net->energy_scale[l];
    nd->gc[l] = net->energy_scale[l] * net->energy_coefficient;
  }
  nd->off[nd->L] = o;
  nd->SD = o;
  nd->NG = o + nd->d_out;
  nd->top = net->top;
  nd->inv_var = net->top_inv_var;
  nd->mask_start = net->mask_start_col;
  if (nd->mask_start < 0 || nd->mask_start > nd->d_out) {
    set_error("mask_start_col=%d outside [0,%d]", nd->mask_start, nd->d_out);
    return MCPC_ERR_INVALID;
  }
  nd->top_has_grad = nd->d_out > 0 && (nd->top == MCPC_TOP_GAUSS || nd->top == MCPC_TOP_BERNOULLI);
  if ((nd->top == MCPC_TOP_GAUSS || nd->top == MCPC_TOP_BERNOULLI) && nd->d_out == 0) {
    set_error("a loss on top needs an output Linear (d_out > 0)");
    return MCPC_ERR_INVALID;
  }
  return MCPC_OK;
}

}  // namespace mcpc

using namespace mcpc;

extern "C" {

int mcpc_version(void) { return MCPC_ABI_VERSION; }

const char* mcpc_last_error(void) { return g_err; }

uint64_t mcpc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int mcpc_workspace_bytes(const McpcNet* net, int32_t B, int32_t n_steps, int32_t precision, size_t* out_bytes) {
  NetDev nd;
  int rc = check_net(net, &nd);
  if (rc != MCPC_OK) return rc;
  if (B < 1 || n_steps < 1 || out_bytes == nullptr) {
    set_error("bad B=%d / n_steps=%d / out_bytes", B, n_steps);
    return MCPC_ERR_INVALID;
  }
  if (precision == MCPC_PREC_FP32) return infer_rows_workspace(nd, B, n_steps, out_bytes);
  if (precision == MCPC_PREC_BF16)
    return infer_tc_fits(nd, B) ? infer_tc_workspace(nd, B, n_steps, out_bytes) : infer_wide_workspace(nd, B, n_steps, out_bytes);
  set_error("precision %d not implemented", precision);
  return MCPC_ERR_UNSUPPORTED;
}

int mcpc_save_layout(const McpcNet* net, int32_t precision, int32_t* g_width, int32_t* f_width, int32_t* elem_bytes) {
  NetDev nd;
  int rc = check_net(net, &nd);
  if (rc != MCPC_OK) return rc;
  if (g_width == nullptr || f_width == nullptr || elem_bytes == nullptr) {
    set_error("mcpc_save_layout: NULL output");
    return MCPC_ERR_INVALID;
  }
  if (precision == MCPC_PREC_FP32) {
    *g_width = nd.NG;
    *f_width = nd.SD;
    *elem_bytes = 4;
    return MCPC_OK;
  }
  if (precision == MCPC_PREC_BF16) {
    int g_off[kMaxL + 1], f_off[kMaxL + 1], gw = 0, fw = 0;
    save_layout_bf16(nd, g_off, &gw, f_off, &fw);
    *g_width = gw;
    *f_width = fw;
    *elem_bytes = 2;
    return MCPC_OK;
  }
  set_error("precision %d not implemented", precision);
  return MCPC_ERR_UNSUPPORTED;
}

int mcpc_infer_mode(const McpcNet* net, int32_t B, int32_t precision, int32_t* mode) {
  NetDev nd;
  int rc = check_net(net, &nd);
  if (rc != MCPC_OK) return rc;
  if (mode == nullptr || B < 1) {
    set_error("mcpc_infer_mode: bad arguments");
    return MCPC_ERR_INVALID;
  }
  if (precision == MCPC_PREC_FP32) {
    size_t bytes = 0;
    rc = infer_rows_workspace(nd, B, 1, &bytes);
    if (rc != MCPC_OK) return rc;
    *mode = MCPC_MODE_RESIDENT_FP32;
    return MCPC_OK;
  }
  if (precision == MCPC_PREC_BF16) {
    if (infer_tc_fits(nd, B)) {
      *mode = MCPC_MODE_RESIDENT_BF16;
      return MCPC_OK;
    }
    size_t bytes = 0;
    rc = infer_wide_workspace(nd, B, 1, &bytes);
    if (rc != MCPC_OK) return rc;
    *mode = MCPC_MODE_STREAMING_BF16;
    return MCPC_OK;
  }
  set_error("precision %d not implemented", precision);
  return MCPC_ERR_UNSUPPORTED;
}

int mcpc_infer_fuses_weight_grad(const McpcNet* net, int32_t B, int32_t precision, int32_t has_inputs, int32_t* out) {
  NetDev nd;
  int rc = check_net(net, &nd);
  if (rc != MCPC_OK) return rc;
  if (out == nullptr || B < 1) {
    set_error("bad arguments");
    return MCPC_ERR_INVALID;
  }
  *out = 0;
  if (precision == MCPC_PREC_BF16)
    *out = infer_tc_fits(nd, B) ? (infer_tc_overlaps_weight_grad(nd, B, has_inputs != 0) ? 1 : 0) : 1;
  return MCPC_OK;
}

int mcpc_infer(const McpcNet* net, const McpcIO* io, const McpcOpts* o, int32_t B, void* workspace,
               size_t workspace_bytes, void* stream) {
  NetDev nd;
  int rc = check_net(net, &nd);
  if (rc != MCPC_OK) return rc;
  if (io == nullptr || o == nullptr || B < 1 || o->n_steps < 1) {
    set_error("bad io/opts/B/n_steps");
    return MCPC_ERR_INVALID;
  }
  for (int l = 0; l < nd.L; ++l) {
    if (io->x[l] == nullptr || io->W[l] == nullptr) {
      set_error("x[%d] / W[%d] is NULL", l, l);
      return MCPC_ERR_INVALID;
    }
    if (o->optimizer == MCPC_OPT_ADAM && o->update_x && (io->adam_m[l] == nullptr || io->adam_v[l] == nullptr)) {
      set_error("Adam needs adam_m/adam_v for layer %d", l);
      return MCPC_ERR_INVALID;
    }
  }
  if (nd.d_out > 0 && io->W[nd.L] == nullptr) {
    set_error("output Linear weight is NULL");
    return MCPC_ERR_INVALID;
  }
  if (nd.top_has_grad && io->target == nullptr) {
    set_error("loss on top needs a target");
    return MCPC_ERR_INVALID;
  }
  if (o->optimizer != MCPC_OPT_SGD && o->optimizer != MCPC_OPT_ADAM) {
    set_error("unknown optimizer %d", o->optimizer);
    return MCPC_ERR_INVALID;
  }
  if (o->noise_mode == MCPC_NOISE_SUPPLIED && io->noise == nullptr) {
    set_error("NOISE_SUPPLIED without a noise tensor");
    return MCPC_ERR_INVALID;
  }
  if (io->save_g != nullptr && o->precision == MCPC_PREC_BF16 && !infer_tc_fits(nd, B)) {
    set_error("streaming bf16 path: pass McpcIO.gW/gb instead of save_g/save_f (see mcpc_infer_mode)");
    return MCPC_ERR_INVALID;
  }
  if ((io->save_g != nullptr) != (io->save_f != nullptr)) {
    set_error("save_g and save_f must be given together");
    return MCPC_ERR_INVALID;
  }
  if (io->save_g != nullptr && (o->save_begin < 0 || o->save_end > o->n_steps || o->save_begin > o->save_end)) {
    set_error("bad save window [%d,%d) for n_steps=%d", o->save_begin, o->save_end, o->n_steps);
    return MCPC_ERR_INVALID;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (o->precision == MCPC_PREC_FP32) {
    rc = launch_infer_rows(nd, io, o, B, workspace, workspace_bytes, s);
    if (rc != MCPC_OK) return rc;
    // McpcIO.gW/gb given: the weight update of the saved steps is part of this call (same contract as the bf16 modes)
    bool want_dw = false;
    for (int l = 0; l <= nd.L; ++l) want_dw = want_dw || io->gW[l] != nullptr || io->gb[l] != nullptr;
    if (want_dw && io->save_g != nullptr && o->save_end > o->save_begin) {
      McpcGradIO gio{};
      gio.save_g = io->save_g;
      gio.save_f = io->save_f;
      gio.inputs = io->inputs;
      for (int l = 0; l <= nd.L; ++l) {
        gio.gW[l] = io->gW[l];
        gio.gb[l] = io->gb[l];
      }
      return launch_weight_grad_fp32(nd, &gio, B, o->save_end - o->save_begin, s);
    }
    return MCPC_OK;
  }
  if (o->precision == MCPC_PREC_BF16)
    return infer_tc_fits(nd, B) ? launch_infer_tc(nd, io, o, B, workspace, workspace_bytes, s)
                                : launch_infer_wide(nd, io, o, B, workspace, workspace_bytes, s);
  set_error("precision %d not implemented", o->precision);
  return MCPC_ERR_UNSUPPORTED;
}

int mcpc_weight_grad(const McpcNet* net, const McpcGradIO* io, int32_t B, int32_t n_save, int32_t precision,
                     void* stream) {
  NetDev nd;
  int rc = check_net(net, &nd);
  if (rc != MCPC_OK) return rc;
  if (io == nullptr || io->save_g == nullptr || io->save_f == nullptr || B < 1 || n_save < 1) {
    set_error("bad weight_grad arguments");
    return MCPC_ERR_INVALID;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (precision == MCPC_PREC_FP32) return launch_weight_grad_fp32(nd, io, B, n_save, s);
  if (precision == MCPC_PREC_BF16) return launch_weight_grad_tc(nd, io, B, n_save, s);
  set_error("precision %d not implemented", precision);
  return MCPC_ERR_UNSUPPORTED;
}

int mcpc_fill_noise(uint64_t seed, int32_t t_begin, int32_t n_steps, uint64_t chain_offset, int32_t B,
                    int32_t n_units, float noise_scale, float* out, void* stream) {
  if (out == nullptr || n_steps < 1 || B < 1 || n_units < 1) {
    set_error("bad fill_noise arguments");
    return MCPC_ERR_INVALID;
  }
  return launch_fill_noise(seed, t_begin, n_steps, chain_offset, B, n_units, noise_scale, out,
                           reinterpret_cast<cudaStream_t>(stream));
}

int mcpc_marginal_ll_workspace_bytes(int32_t N, int32_t S, int32_t D, size_t* bytes) {
  if (bytes == nullptr) {
    set_error("mcpc_marginal_ll_workspace_bytes: NULL argument");
    return MCPC_ERR_INVALID;
  }
  return marginal_ll_workspace(N, S, D, bytes);
}

int mcpc_marginal_ll_bernoulli(const float* logits, int32_t S, const float* data, int32_t N, int32_t D, float clamp_abs,
                               void* workspace, size_t workspace_bytes, double* ml_out, float* row_ll, void* stream) {
  return launch_marginal_ll(logits, S, data, N, D, clamp_abs, workspace, workspace_bytes, ml_out, row_ll,
                            reinterpret_cast<cudaStream_t>(stream));
}

int mcpc_traj_stats_update(const float* traj, int32_t n_rec, uint64_t n_elems, uint64_t count_before, float* mean, float* m2,
                           void* stream) {
  if (traj == nullptr || mean == nullptr || m2 == nullptr || n_rec < 1 || n_elems < 1) {
    set_error("mcpc_traj_stats_update: bad arguments");
    return MCPC_ERR_INVALID;
  }
  return launch_traj_stats(traj, n_rec, (size_t)n_elems, (double)count_before, mean, m2, reinterpret_cast<cudaStream_t>(stream));
}

int mcpc_p_step(const McpcPStep* s, void* stream) {
  if (s == nullptr || s->n_tensors < 1 || s->n_tensors > MCPC_MAX_PTENSORS) {
    set_error("mcpc_p_step: bad tensor count");
    return MCPC_ERR_INVALID;
  }
  if (s->kind != MCPC_OPT_SGD && s->kind != MCPC_OPT_ADAM) {
    set_error("mcpc_p_step: unknown optimizer %d", s->kind);
    return MCPC_ERR_INVALID;
  }
  for (int i = 0; i < s->n_tensors; ++i) {
    if (s->param[i] == nullptr || s->grad[i] == nullptr || (s->kind == MCPC_OPT_ADAM && (s->state1[i] == nullptr || s->state2[i] == nullptr))) {
      set_error("mcpc_p_step: NULL param / grad / state for tensor %d", i);
      return MCPC_ERR_INVALID;
    }
  }
  if (s->kind == MCPC_OPT_ADAM && s->step < 1) {
    set_error("mcpc_p_step: Adam step index must be >= 1");
    return MCPC_ERR_INVALID;
  }
  return launch_p_step(s, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
