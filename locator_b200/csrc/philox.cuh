// Philox4x32-10 (Salmon et al. 2011). Layout contract mirrored by oracle/philox_ref.py:
//   counter = (idx_lo, idx_hi, stream, 0), key = (seed_lo, seed_hi);
//   element e -> block e/4, lane e%4; uniform = (x >> 8) * 2^-24.
#pragma once
#include <stdint.h>

namespace loc {

struct Philox4 {
  uint32_t v[4];
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint64_t idx, uint32_t stream, uint64_t seed) {
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = stream, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox4 out;
  out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
  return out;
}

__host__ __device__ __forceinline__ float philox_uniform(uint64_t e, uint32_t stream, uint64_t seed) {
  Philox4 r = philox4x32_10(e >> 2, stream, seed);
  return (float)(r.v[e & 3] >> 8) * 5.9604644775390625e-08f;  // 2^-24
}

constexpr uint32_t kDropoutStreamBase = 0x40000000u;

}  // namespace loc
