// dmip_rng.cuh — Philox4x32-10 keyed Gaussian stream (mirrored by oracle/philox.py).
// The reference draws from torch's global RNG (models/diffusion.py:32,42); a counter-based generator keyed by
// the GLOBAL particle index makes samples independent of tiling and of the number of GPUs.
//   counter = (gidx_lo, gidx_hi, step, (stream << 16) | quad)    key = (seed_lo, seed_hi)
#pragma once
#include <stdint.h>

namespace dmip {

constexpr uint32_t kPhiloxStepInit = 0xFFFFFFFFu;  // "step" used for the initial draw x0
constexpr uint32_t kStreamState = 0;               // noise on the state x
constexpr uint32_t kStreamObs = 1;                 // CDiffE: noise re-diffusing the observation y

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float u01(uint32_t r) { return (static_cast<float>(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// four standard normals for elements 4*quad .. 4*quad+3 of particle gidx at `step` in `stream`
__device__ __forceinline__ void philox_normal4(uint64_t gidx, uint32_t step, uint32_t stream, uint32_t quad,
                                               uint64_t seed, float (&z)[4]) {
  uint32_t r[4];
  philox4x32_10(static_cast<uint32_t>(gidx), static_cast<uint32_t>(gidx >> 32), step, (stream << 16) | quad,
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float u1 = u01(r[2 * j]), u2 = u01(r[2 * j + 1]);
    float R;   // sqrt.approx (1 ulp, branch-free): keeps the draw one straight-line block the scheduler can interleave
    // u01 can round to exactly 1 and __logf is only specified to an ABSOLUTE error there: clamp the radicand so a
    // slightly positive log can never turn into NaN (one FMNMX)
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(R) : "f"(fmaxf(-2.0f * __logf(u1), 0.0f)));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    z[2 * j] = R * c;
    z[2 * j + 1] = R * s;
  }
}

}  // namespace dmip
