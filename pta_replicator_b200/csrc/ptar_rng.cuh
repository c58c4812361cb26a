// Counter-based normal stream of the throughput mode (replaces the reference's global
// np.random.randn stream): Philox4x32-10 (Salmon et al. 2011) + Box-Muller
// evaluated in fp32.  Restated in numpy by oracle/philox.py; tests compare the two.
#pragma once
#include <stdint.h>

namespace ptar {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// Two standard normals from two 32-bit words.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u1 = (static_cast<float>(a) + 0.5f) * 2.3283064365386963e-10f;  // (0, 1]
  const float u2 = (static_cast<float>(b) + 0.5f) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.2831853071795865f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// Counter layout: c.x = block index, c.y = kind | psr << 8, c.z = low word of the
// realization field, c.w = high word.  White noise: block = idx >> 2, realization field =
// global realization id, the 4 outputs are idx&~3 .. +3.  Everything else: block = idx,
// realization field = id >> 2, the 4 outputs are realizations id&~3 .. +3.
__device__ __forceinline__ void normals4(float n[4], uint32_t block, uint32_t kind, uint32_t psr,
                                         uint64_t rfield, uint64_t seed) {
  const uint4 w = philox4x32_10(make_uint4(block, kind | (psr << 8), static_cast<uint32_t>(rfield),
                                           static_cast<uint32_t>(rfield >> 32)),
                                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  box_muller(w.x, w.y, n[0], n[1]);
  box_muller(w.z, w.w, n[2], n[3]);
}

}  // namespace ptar
