// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Scalar type that counts the source-level FP64 operations of whatever template it is fed to.
// Used to replace SURVEY §8(d)'s static estimate of F_V (flops per `ccpol` energy) by an exact
// census: add/sub, mul, div, sqrt, exp, pow, other transcendentals (sin cos acos atan tanh).
#pragma once
#include <cmath>

#include "../include/pimdk_detmath.h"

namespace oracle {

struct Counted {
  enum Kind { ADD = 0, MUL, DIV, SQRT, EXP, POW, TRIG, NKIND };
  static inline long long cnt[NKIND] = {};
  static void reset() { for (auto& c : cnt) c = 0; }
  static const char* name(int i) {
    static const char* n[NKIND] = {"add", "mul", "div", "sqrt", "exp", "pow", "trig"};
    return (i >= 0 && i < NKIND) ? n[i] : "";
  }
  double v;
  Counted() : v(0.0) {}
  explicit Counted(double x) : v(x) {}
};

inline Counted operator+(Counted a, Counted b) { Counted::cnt[Counted::ADD]++; return Counted(a.v + b.v); }
inline Counted operator-(Counted a, Counted b) { Counted::cnt[Counted::ADD]++; return Counted(a.v - b.v); }
inline Counted operator*(Counted a, Counted b) { Counted::cnt[Counted::MUL]++; return Counted(a.v * b.v); }
inline Counted operator/(Counted a, Counted b) { Counted::cnt[Counted::DIV]++; return Counted(a.v / b.v); }
inline Counted operator-(Counted a) { return Counted(-a.v); }
inline bool operator<(Counted a, Counted b) { return a.v < b.v; }
inline bool operator>(Counted a, Counted b) { return a.v > b.v; }
inline bool operator==(Counted a, Counted b) { return a.v == b.v; }
inline bool operator!=(Counted a, Counted b) { return a.v != b.v; }
inline Counted sqrt(Counted a) { Counted::cnt[Counted::SQRT]++; return Counted(std::sqrt(a.v)); }
inline Counted exp(Counted a) { Counted::cnt[Counted::EXP]++; return Counted(pimdk_exp(a.v)); }
inline Counted pow(Counted a, Counted b) { Counted::cnt[Counted::POW]++; return Counted(pimdk_pow(a.v, b.v)); }
inline Counted fabs(Counted a) { return Counted(std::fabs(a.v)); }
inline Counted sin(Counted a) { Counted::cnt[Counted::TRIG]++; return Counted(pimdk_sin(a.v)); }
inline Counted cos(Counted a) { Counted::cnt[Counted::TRIG]++; return Counted(pimdk_cos(a.v)); }
inline Counted acos(Counted a) { Counted::cnt[Counted::TRIG]++; return Counted(pimdk_acos(a.v)); }
inline Counted atan(Counted a) { Counted::cnt[Counted::TRIG]++; return Counted(pimdk_atan(a.v)); }
inline Counted tanh(Counted a) { Counted::cnt[Counted::TRIG]++; return Counted(pimdk_tanh(a.v)); }

}  // namespace oracle
