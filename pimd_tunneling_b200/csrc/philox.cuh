// Counter-based RNG for the thermostats (replaces MKL VSL MT19937 streams,
// verletmodule.f90:350-368 and the vdrnggaussian / virngpoisson call sites :106,202,215,230,598,641).
// RNG contract:
//   Philox4x32-10, key = (seed lo, seed hi),
//   counter = (pair, step lo, traj_gid, (stream << 24) | (step hi & 0xffffff)),
//   u = ((hi32<<32 | lo32) >> 11 + 0.5) * 2^-53, Box-Muller with log/sincos from pimdk_detmath.h,
//   normal #idx lives in pair idx>>1, slot idx&1.
#pragma once
#include <cstdint>

#include "../../include/pimdk_detmath.h"

namespace pimdk {

enum { STREAM_INIT = 0, STREAM_LANGEVIN = 1, STREAM_ANDERSEN = 2, STREAM_POISSON = 3, STREAM_READHESS = 4 };

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t* out) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double normal_at(uint64_t seed, int stream, uint64_t step, uint32_t gid, uint64_t idx) {
  uint32_t r[4];
  philox4x32_10((uint32_t)(idx >> 1), (uint32_t)(step & 0xffffffffu), gid,
                ((uint32_t)stream << 24) | (uint32_t)((step >> 32) & 0xffffffu), (uint32_t)(seed & 0xffffffffu),
                (uint32_t)(seed >> 32), r);
  const double two53 = 1.0 / 9007199254740992.0;
  const double u1 = ((double)((((uint64_t)r[0] << 32) | r[1]) >> 11) + 0.5) * two53;
  const double u2 = ((double)((((uint64_t)r[2] << 32) | r[3]) >> 11) + 0.5) * two53;
  const double rad = sqrt(-2.0 * pimdk_log(u1));
  const double ang = 6.283185307179586 * u2;
  double sn, cs;
  pimdk_sincos(ang, &sn, &cs);
  return (idx & 1) ? rad * sn : rad * cs;
}

// Both normals of Box-Muller pair `pair` (normal #2*pair and #2*pair+1): one Philox block, one log, one sincos for two
// numbers; bit-identical to normal_at(.., 2*pair) and normal_at(.., 2*pair + 1).
__device__ __forceinline__ void normal_pair_at(uint64_t seed, int stream, uint64_t step, uint32_t gid, uint64_t pair,
                                               double& z0, double& z1) {
  uint32_t r[4];
  philox4x32_10((uint32_t)pair, (uint32_t)(step & 0xffffffffu), gid,
                ((uint32_t)stream << 24) | (uint32_t)((step >> 32) & 0xffffffu), (uint32_t)(seed & 0xffffffffu),
                (uint32_t)(seed >> 32), r);
  const double two53 = 1.0 / 9007199254740992.0;
  const double u1 = ((double)((((uint64_t)r[0] << 32) | r[1]) >> 11) + 0.5) * two53;
  const double u2 = ((double)((((uint64_t)r[2] << 32) | r[3]) >> 11) + 0.5) * two53;
  const double rad = sqrt(-2.0 * pimdk_log(u1));
  const double ang = 6.283185307179586 * u2;
  double sn, cs;
  pimdk_sincos(ang, &sn, &cs);
  z0 = rad * cs;
  z1 = rad * sn;
}

// Poisson(lambda), normal approximation (the reference requests VSL_RNG_METHOD_POISSON_POISNORM,
// verletmodule.f90:361): floor(lambda + sqrt(lambda) z + 0.5), clamped at 0.
__device__ __forceinline__ int poisson_norm(uint64_t seed, uint64_t step, uint32_t gid, double lambda) {
  const double z = normal_at(seed, STREAM_POISSON, step, gid, 0);
  const double k = floor(lambda + sqrt(lambda) * z + 0.5);
  return k < 0.0 ? 0 : (int)k;
}

}  // namespace pimdk
