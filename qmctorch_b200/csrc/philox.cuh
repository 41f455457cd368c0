// Counter-based RNG (Philox4x32-10) for the in-kernel Metropolis draws.  Every draw is a
// pure function of (seed, offset=step counter, element index, substream), so results do
// not depend on the tiling or the number of GPUs.  Replaces the torch CPU/GPU generator
// calls at sampler/metropolis.py:247,266-275,295 when no draws are injected.
#pragma once
#include <cstdint>

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__device__ __forceinline__ void philox4(uint64_t seed, uint64_t offset, uint64_t idx, uint32_t sub,
                                        uint32_t (&c)[4]) {
  c[0] = (uint32_t)idx; c[1] = (uint32_t)(idx >> 32);
  c[2] = (uint32_t)offset; c[3] = ((uint32_t)(offset >> 32) & 0x0fffffffu) | (sub << 28);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__device__ __forceinline__ uint32_t philox_u32(uint64_t seed, uint64_t offset, uint64_t idx, uint32_t sub) {
  uint32_t c[4];
  philox4(seed, offset, idx, sub, c);
  return c[0];
}

__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// uniform in [0,1)
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t offset, uint64_t idx, uint32_t sub) {
  uint32_t c[4];
  philox4(seed, offset, idx, sub, c);
  return u53(c[0], c[1]);
}

// Four standard normals per Philox call: element quad q -> (z[0..3]) = two Box-Muller pairs.
// The draws only steer a symmetric Metropolis proposal, so they are formed in FP32 on the SFU
// (MUFU.LG2 / RSQ / SIN / COS) instead of FP64 libdevice log/sincospi/sqrt (~9 x fewer
// instructions per normal, 2 x fewer Philox calls).  Detailed balance needs q(z) = q(-z) exactly:
// the angle covers [0, pi) and one random bit flips the sign of the pair, so -z is produced by
// the same arithmetic as z.  Resolution: radius from 23 uniform bits (|z| < 5.8), angle 31 bits.
__device__ __forceinline__ void philox_pair(uint32_t a, uint32_t b, double &z0, double &z1) {
  const float u1 = ((float)(a >> 9) + 0.5f) * 1.1920928955078125e-7f;        // (0,1), exact in FP32
  const float th = (float)(b >> 1) * 1.4629180792671596e-9f;                 // pi * 2^-31 -> [0, pi]
  const float t = -2.0f * __logf(u1);
  float r = t * rsqrtf(t);
  r = (b & 1u) ? -r : r;
  float s, c;
  __sincosf(th, &s, &c);
  z0 = (double)(r * c);
  z1 = (double)(r * s);
}
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t offset, uint64_t quad, double (&z)[4]) {
  uint32_t c[4];
  philox4(seed, offset, quad, 3u, c);
  philox_pair(c[0], c[1], z[0], z[1]);
  philox_pair(c[2], c[3], z[2], z[3]);
}
