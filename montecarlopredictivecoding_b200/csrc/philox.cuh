// Counter-based Langevin noise: Philox4x32-10 + Box-Muller.
//
// Replaces `x.grad.normal_(0, sqrt(var/lr))` of utils/model.py:42-43 (reference).  The counter
// layout is part of the ABI (include/mcpc_b200.h, McpcOpts.seed / chain_offset):
//     counter = (unit, t, chain>>2 lo, chain>>2 hi), key = (seed lo, seed hi)
// and the four outputs give four normals, one per chain of an aligned group of four, so a
// draw depends only on (seed, global chain, unit, step) -- never on tiling or GPU count.
// The CPU restatement is oracle/mcpc_oracle.py:langevin_normals.
#pragma once
#include <cstdint>

namespace mcpc {

struct Philox4 {
  uint32_t v[4];
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{{c0, c1, c2, c3}};
}

__device__ __forceinline__ float u01_24(uint32_t bits) {
  // (top 24 bits + 0.5) / 2^24: strictly inside (0,1), exactly representable
  return (static_cast<float>(bits >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

// MUFU lg2 without the denormal-input wrapper of __log2f (the arguments are >= 2^-25): same value, three instructions less
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Four standard normals for chains 4q..4q+3 of `unit` at step `t`.
__device__ __forceinline__ void langevin_normals4(uint64_t seed, uint32_t unit, uint32_t t, uint64_t q, float out[4]) {
  const Philox4 r = philox4x32_10(unit, t, static_cast<uint32_t>(q), static_cast<uint32_t>(q >> 32),
                                  static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  // -2 ln u = (-2 ln 2) log2 u; MUFU lg2 / sqrt (approx, ~1 ulp): the radius of a N(0,1) draw needs no IEEE rounding,
  // and sqrtf's correctly rounded slow path was a function call per draw
  const float ra = sqrt_approx(-1.38629436111989062f * lg2_approx(u01_24(r.v[0])));
  const float rb = sqrt_approx(-1.38629436111989062f * lg2_approx(u01_24(r.v[2])));
  // MUFU sin/cos on an argument in (-pi, pi): absolute error ~2^-21, far below the noise it shapes
  float sa, ca, sb, cb;
  __sincosf(6.28318530717958648f * (u01_24(r.v[1]) - 0.5f), &sa, &ca);
  __sincosf(6.28318530717958648f * (u01_24(r.v[3]) - 0.5f), &sb, &cb);
  out[0] = ra * ca; out[1] = ra * sa; out[2] = rb * cb; out[3] = rb * sb;
}

}  // namespace mcpc
