// Philox4x32-10 counter-based generator (Salmon et al., SC'11) and the repo's random-stream contract
// (oracle/philox.py):
//   element e of stream t -> counter (lo32(e>>1), hi32(e>>1), t, 0), key (lo32(seed), hi32(seed))      [normals]
//   element e of stream t -> counter (lo32(e),    hi32(e),    t, 0)                                     [uniforms]
//   u = (((x>>5)<<26 | (y>>6)) + 0.5) 2^-53;  Box-Muller; even element takes r cos(theta), odd r sin(theta).
// The raw words are bit-identical to the numpy oracle for any sharding of the logical index space.
#pragma once
#include <cstdint>

namespace gpp {

enum PhiloxStream { STREAM_OMEGA = 0, STREAM_PHASE = 1, STREAM_PRIOR_W = 2, STREAM_U_EPS = 3, STREAM_UPDATE_XI = 4, STREAM_X0 = 5 };

struct Philox4 { uint32_t v[4]; };

__host__ __device__ inline Philox4 philox4x32_10(uint64_t index, uint32_t stream, uint64_t seed) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = stream, c3 = 0;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  Philox4 out;
  out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
  return out;
}

__host__ __device__ inline double philox_u53(uint32_t x, uint32_t y) {
  uint64_t bits = ((uint64_t)(x >> 5) << 26) | (uint64_t)(y >> 6);
  return ((double)bits + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ inline double philox_normal(uint64_t element, uint32_t stream, uint64_t seed) {
  Philox4 w = philox4x32_10(element >> 1, stream, seed);
  double u1 = philox_u53(w.v[0], w.v[1]), u2 = philox_u53(w.v[2], w.v[3]);
  double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincos(6.283185307179586476925 * u2, &s, &c);
  return (element & 1) ? r * s : r * c;
}

__device__ inline double philox_uniform(uint64_t element, uint32_t stream, uint64_t seed) {
  Philox4 w = philox4x32_10(element, stream, seed);
  return philox_u53(w.v[0], w.v[1]);
}

}  // namespace gpp
