// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Forward-mode dual numbers with 18 derivative components: fed to the same templates as `double` and `Counted`, they
// give the ANALYTIC gradient of one `ccpol` energy with respect to the 18 Cartesian coordinates.  The reference has no
// analytic gradient (mcmod_waterdimer_ccpol.f90:40-58 is a central difference, eps = 1e-4 bohr); this measures that
// difference's truncation error and is the yardstick for an analytic-gradient mode.
#pragma once
#include <cmath>

#include "../include/pimdk_detmath.h"

namespace oracle {

struct Dual {
  static constexpr int N = 18;
  double v;
  double d[N];
  Dual() : v(0.0) { for (double& t : d) t = 0.0; }
  explicit Dual(double x) : v(x) { for (double& t : d) t = 0.0; }
};

template <class F>
inline Dual dual_map(const Dual& a, double value, F slope) {   // f(a): value f(a.v), derivative slope * a'
  Dual r;
  r.v = value;
  const double s = slope;
  for (int i = 0; i < Dual::N; ++i) r.d[i] = a.d[i] == 0.0 ? 0.0 : s * a.d[i];   // constants stay constants: acos(-1), sqrt(0)
  return r;
}
inline Dual operator+(const Dual& a, const Dual& b) { Dual r; r.v = a.v + b.v; for (int i = 0; i < Dual::N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
inline Dual operator-(const Dual& a, const Dual& b) { Dual r; r.v = a.v - b.v; for (int i = 0; i < Dual::N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
inline Dual operator*(const Dual& a, const Dual& b) { Dual r; r.v = a.v * b.v; for (int i = 0; i < Dual::N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
inline Dual operator/(const Dual& a, const Dual& b) {
  Dual r;
  r.v = a.v / b.v;
  for (int i = 0; i < Dual::N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
  return r;
}
inline Dual operator-(const Dual& a) { Dual r; r.v = -a.v; for (int i = 0; i < Dual::N; ++i) r.d[i] = -a.d[i]; return r; }
inline bool operator<(const Dual& a, const Dual& b) { return a.v < b.v; }
inline bool operator>(const Dual& a, const Dual& b) { return a.v > b.v; }
inline bool operator==(const Dual& a, const Dual& b) { return a.v == b.v; }
inline bool operator!=(const Dual& a, const Dual& b) { return a.v != b.v; }
inline Dual sqrt(const Dual& a) { const double s = std::sqrt(a.v); return dual_map(a, s, 0.5 / s); }
inline Dual exp(const Dual& a) { const double e = pimdk_exp(a.v); return dual_map(a, e, e); }
inline Dual pow(const Dual& a, const Dual& b) {   // a > 0
  const double p = pimdk_pow(a.v, b.v);
  Dual r;
  r.v = p;
  const double da = b.v * p / a.v, db = p * pimdk_log(a.v);
  for (int i = 0; i < Dual::N; ++i) r.d[i] = (a.d[i] == 0.0 ? 0.0 : da * a.d[i]) + (b.d[i] == 0.0 ? 0.0 : db * b.d[i]);
  return r;
}
inline Dual fabs(const Dual& a) { return a.v < 0.0 ? -a : a; }
inline Dual sin(const Dual& a) { return dual_map(a, pimdk_sin(a.v), pimdk_cos(a.v)); }
inline Dual cos(const Dual& a) { return dual_map(a, pimdk_cos(a.v), -pimdk_sin(a.v)); }
inline Dual acos(const Dual& a) { return dual_map(a, pimdk_acos(a.v), -1.0 / std::sqrt(1.0 - a.v * a.v)); }
inline Dual atan(const Dual& a) { return dual_map(a, pimdk_atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
inline Dual tanh(const Dual& a) { const double t = pimdk_tanh(a.v); return dual_map(a, t, 1.0 - t * t); }

}  // namespace oracle
