// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// The reference draws its thermostat noise from MKL VSL MT19937 streams seeded from the
// wall clock (verletmodule.f90:350-368, :106,202,215,230,598,641) — unpinnable.  The new
// build replaces that by a counter-based generator whose stream layout is part of the
// specification (DESIGN.md "RNG contract"); this file restates that contract on the CPU so
// that thermostatted trajectories can be compared draw for draw.
//
//   generator : Philox4x32-10 (Salmon et al. 2011; constants below)
//   key       : (seed & 0xffffffff, seed >> 32)
//   counter   : (pair, step & 0xffffffff, traj_gid, (stream << 24) | ((step >> 32) & 0xffffff))
//   uniforms  : u1 = ((r0<<32 | r1) >> 11 + 0.5) * 2^-53,  u2 likewise from r2,r3
//   log/sin/cos: include/pimdk_detmath.h (bit-identical on host and device)
//   normals   : Box-Muller  z0 = sqrt(-2 ln u1) cos(2 pi u2),  z1 = sqrt(-2 ln u1) sin(2 pi u2)
//   normal #idx of a (stream, step, traj) lives in pair idx>>1, slot idx&1
//   streams   : 0 init_path momenta, 1 Langevin O-step, 2 Andersen resample, 3 Poisson interval
#pragma once
#include <cmath>
#include <cstdint>

#include "../include/pimdk_detmath.h"

namespace oracle {

enum RngStream { STREAM_INIT = 0, STREAM_LANGEVIN = 1, STREAM_ANDERSEN = 2, STREAM_POISSON = 3 };

inline void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

inline void normal_pair(uint64_t seed, int stream, uint64_t step, uint32_t traj_gid, uint32_t pair,
                        double& z0, double& z1) {
  uint32_t key[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
  uint32_t ctr[4] = {pair, (uint32_t)(step & 0xffffffffu), traj_gid,
                     ((uint32_t)stream << 24) | (uint32_t)((step >> 32) & 0xffffffu)};
  uint32_t r[4];
  philox4x32_10(ctr, key, r);
  const double two53 = 1.0 / 9007199254740992.0;
  double u1 = ((double)((((uint64_t)r[0] << 32) | r[1]) >> 11) + 0.5) * two53;
  double u2 = ((double)((((uint64_t)r[2] << 32) | r[3]) >> 11) + 0.5) * two53;
  double rad = std::sqrt(-2.0 * pimdk_log(u1));
  double ang = 6.283185307179586 * u2;
  double sn, cs;
  pimdk_sincos(ang, &sn, &cs);
  z0 = rad * cs;
  z1 = rad * sn;
}

inline double normal_at(uint64_t seed, int stream, uint64_t step, uint32_t traj_gid, uint64_t idx) {
  double z0, z1;
  normal_pair(seed, stream, step, traj_gid, (uint32_t)(idx >> 1), z0, z1);
  return (idx & 1) ? z1 : z0;
}

// Poisson(lambda) in the normal approximation (the reference asks VSL for POISNORM,
// verletmodule.f90:361): k = floor(lambda + sqrt(lambda) z + 0.5), clamped at 0.
inline int poisson_norm(uint64_t seed, uint64_t step, uint32_t traj_gid, double lambda) {
  double z = normal_at(seed, STREAM_POISSON, step, traj_gid, 0);
  double k = std::floor(lambda + std::sqrt(lambda) * z + 0.5);
  return k < 0.0 ? 0 : (int)k;
}

}  // namespace oracle
