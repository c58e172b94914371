// extern "C" entry points of libmcpc_b200_probes.so (include/mcpc_b200_probes.h): validation-only known-answer tests of
// the tcgen05 / TMEM / TMA primitives.  Built as its own library; nothing here is linked into libmcpc_b200.so.
#include <atomic>

#include "mcpc_b200_probes.h"
#include "mcpc_common.cuh"

namespace mcpc {

static thread_local char g_perr[512] = "";

void count_launch(int) {}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_perr, sizeof(g_perr), fmt, ap);
  va_end(ap);
}

}  // namespace mcpc

using namespace mcpc;

extern "C" {

const char* mcpc_probes_last_error(void) { return g_perr; }

int mcpc_debug_umma(const float* Wt, const float* Bx, const float* G, int32_t Kin, int32_t N, float* D1, float* D2,
                    void* ws, void* stream) {
  if (Wt == nullptr || Bx == nullptr || G == nullptr || D1 == nullptr || D2 == nullptr || ws == nullptr) {
    set_error("mcpc_debug_umma: NULL argument");
    return MCPC_ERR_INVALID;
  }
  return launch_umma_probe(Wt, Bx, G, Kin, N, D1, D2, ws, reinterpret_cast<cudaStream_t>(stream));
}

int mcpc_debug_tma(const float* A, const float* B, int32_t N, int32_t a_mn, int32_t b_mn, float* D, void* ws, void* stream) {
  if (A == nullptr || B == nullptr || D == nullptr || ws == nullptr) {
    set_error("mcpc_debug_tma: NULL argument");
    return MCPC_ERR_INVALID;
  }
  return launch_tma_probe(A, B, N, a_mn, b_mn, D, ws, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
