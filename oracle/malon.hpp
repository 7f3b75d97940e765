// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// CPU restatement of the malonaldehyde surface behind mcmod_malon.f90:10-72 (V, Vprime, Vdoubleprime):
// `subroutine pes(x, iopt, e, g, h)` of pes_malonaldehyde.f90:4-9592 with v_morse / f_morse / h_morse (:9598-9645),
// v_gauss / f_gauss / h_gauss (:9650-9734) and iorder (:9737-9745).  The fit's numbers are the DATA statements of
// pes_malonaldehyde.f90:45-9347, read from pimd_tunneling_b200/data/malonaldehyde.tbl (tools/pack_malon_tables.py).
// Parity pin: the reference file's own header lists the minimum-energy structure (pes_malonaldehyde.f90:12-21) and says the
// energy is "above equilibrium": tests/test_oracle.py checks V = 0 and a vanishing gradient there.
// dexp is the repository's deterministic exp (include/pimdk_detmath.h), shared with the kernels.
#pragma once
#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/pimdk_detmath.h"

namespace oracle {

struct Malonaldehyde {
  static constexpr int natoms = 9, ndist = 36, nx = 27;
  static constexpr int nmorse = 9, ng1d = 80, ng2d = 1728, ng3d = 1741;
  double shift = 0.0;
  std::vector<int> imorse, ig1d, ig2d, ig3d;          // 1-based distance indices, Fortran storage order
  std::vector<double> morse, g1d, g2d, g3d;           // morse(3,9), g1d(4,80), g2d(6,1728), g3d(8,1741)
  bool loaded = false;

  void load(const std::string& path) {   // C stdio on purpose: the library is loaded into processes that carry another libstdc++
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f) throw std::runtime_error("cannot open " + path);
    char key[64];
    long cnt;
    int c;
    while ((c = std::fgetc(f)) != EOF && c != '\n') {}   // header line
    bool ok = true;
    while (ok && std::fscanf(f, "%63s %ld", key, &cnt) == 2) {
      const std::string k(key);
      if (k[0] == 'i') {
        std::vector<int>& v = k == "imorse" ? imorse : k == "ig1d" ? ig1d : k == "ig2d" ? ig2d : ig3d;
        v.resize(cnt);
        for (long i = 0; i < cnt && ok; ++i) ok = std::fscanf(f, "%d", &v[i]) == 1;
      } else if (k == "shift") {
        ok = std::fscanf(f, "%lf", &shift) == 1;
      } else {
        std::vector<double>& v = k == "morse" ? morse : k == "g1d" ? g1d : k == "g2d" ? g2d : g3d;
        v.resize(cnt);
        for (long i = 0; i < cnt && ok; ++i) ok = std::fscanf(f, "%lf", &v[i]) == 1;
      }
    }
    std::fclose(f);
    if (!ok || (int)imorse.size() != nmorse || (int)morse.size() != 3 * nmorse || (int)ig1d.size() != ng1d ||
        (int)g1d.size() != 4 * ng1d || (int)ig2d.size() != 2 * ng2d || (int)g2d.size() != 6 * ng2d ||
        (int)ig3d.size() != 3 * ng3d || (int)g3d.size() != 8 * ng3d)
      throw std::runtime_error("malonaldehyde table file is incomplete: " + path);
    loaded = true;
  }

  // v_gauss / f_gauss / h_gauss share their first lines (:9660-9665, :9686-9691, :9715-9720)
  static double gauss_arg(int nd, const double* r, const double* x, const double* alpha) {
    double v = 0.0;
    for (int i = 0; i < nd; ++i) v = v + ((r[i] - x[i]) * (r[i] - x[i])) * alpha[i];
    return v * 0.5;
  }

  // dist(ij), ij = pairs (i, j < i) in the reference's loop order (:9356-9364)
  void distances(const double* x, double* dist) const {
    int ij = 0;
    for (int i = 0; i < natoms; ++i)
      for (int j = 0; j < i; ++j) {
        double r0 = x[3 * i] - x[3 * j], r1 = x[3 * i + 1] - x[3 * j + 1], r2 = x[3 * i + 2] - x[3 * j + 2];
        double rij = r0 * r0 + r1 * r1 + r2 * r2;
        dist[ij++] = std::sqrt(rij);
      }
  }
  // the B matrix dr/dx (:9368-9383): row ij holds +-(xi - xj) * (1 / |xi - xj|)
  void bmatrix(const double* x, double* B /* [ndist][nx] */, double* rr /* [ndist]: 1/r */, double* u /* [ndist][3] */) const {
    for (int q = 0; q < ndist * nx; ++q) B[q] = 0.0;
    int ij = 0;
    for (int i = 0; i < natoms; ++i)
      for (int j = 0; j < i; ++j) {
        double r[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
        double rrij = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        rrij = 1.0 / std::sqrt(rrij);
        for (int k = 0; k < 3; ++k) {
          r[k] = rrij * r[k];
          B[ij * nx + 3 * i + k] = r[k];
          B[ij * nx + 3 * j + k] = -r[k];
          if (u) u[ij * 3 + k] = r[k];
        }
        if (rr) rr[ij] = rrij;
        ++ij;
      }
  }

  // iopt = 0 (:9408-9446)
  double energy(const double* x) const {
    double dist[ndist], r[3];
    distances(x, dist);
    double e = shift;
    for (int i = 0; i < nmorse; ++i) {  // v_morse (:9598-9610)
      const double re = morse[3 * i], al = morse[3 * i + 1], de = morse[3 * i + 2];
      double v = al * (re - dist[imorse[i] - 1]);
      v = pimdk_exp(v) - 1.0;
      v = (de * v) * (de * v);
      e = e + v;
    }
    for (int i = 0; i < ng1d; ++i) {  // v_gauss(1, r, g1d(1,i), g1d(2,i), g1d(3,i), g1d(4,i))
      r[0] = dist[ig1d[i] - 1];
      double v = gauss_arg(1, r, &g1d[4 * i], &g1d[4 * i + 1]);
      v = pimdk_exp(-v) - g1d[4 * i + 3];
      e = e + v * g1d[4 * i + 2];
    }
    for (int i = 0; i < ng2d; ++i) {  // v_gauss(2, r, g2d(1,i), g2d(3,i), g2d(5,i), g2d(6,i))
      for (int j = 0; j < 2; ++j) r[j] = dist[ig2d[2 * i + j] - 1];
      double v = gauss_arg(2, r, &g2d[6 * i], &g2d[6 * i + 2]);
      v = pimdk_exp(-v) - g2d[6 * i + 5];
      e = e + v * g2d[6 * i + 4];
    }
    for (int i = 0; i < ng3d; ++i) {  // v_gauss(3, r, g3d(1,i), g3d(4,i), g3d(7,i), g3d(8,i))
      for (int j = 0; j < 3; ++j) r[j] = dist[ig3d[3 * i + j] - 1];
      double v = gauss_arg(3, r, &g3d[8 * i], &g3d[8 * i + 3]);
      v = pimdk_exp(-v) - g3d[8 * i + 7];
      e = e + v * g3d[8 * i + 6];
    }
    return e;
  }

  // the internal gradient gint(ndist) of iopt >= 1 (:9450-9489)
  void internal_gradient(const double* dist, double* gint) const {
    double r[3];
    for (int k = 0; k < ndist; ++k) gint[k] = 0.0;
    for (int i = 0; i < nmorse; ++i) {  // f_morse (:9614-9628)
      const double re = morse[3 * i], al = morse[3 * i + 1], de = morse[3 * i + 2];
      const int ii = imorse[i] - 1;
      double f = al * (re - dist[ii]);
      f = pimdk_exp(f);
      f = f * (f - 1.0);
      f = f * 2.0 * al * (de * de);
      gint[ii] = gint[ii] - f;
    }
    auto fg = [&](int nd, const int* idx, const double* x0, const double* alpha, double d) {  // f_gauss (:9673-9698)
      for (int j = 0; j < nd; ++j) r[j] = dist[idx[j] - 1];
      double vv = gauss_arg(nd, r, x0, alpha);
      vv = pimdk_exp(-vv);
      vv = vv * d;
      double gg[3];
      for (int j = 0; j < nd; ++j) gg[j] = -vv * (r[j] - x0[j]) * alpha[j];
      for (int j = 0; j < nd; ++j) gint[idx[j] - 1] = gint[idx[j] - 1] + gg[j];
    };
    for (int i = 0; i < ng1d; ++i) fg(1, &ig1d[i], &g1d[4 * i], &g1d[4 * i + 1], g1d[4 * i + 2]);
    for (int i = 0; i < ng2d; ++i) fg(2, &ig2d[2 * i], &g2d[6 * i], &g2d[6 * i + 2], g2d[6 * i + 4]);
    for (int i = 0; i < ng3d; ++i) fg(3, &ig3d[3 * i], &g3d[8 * i], &g3d[8 * i + 3], g3d[8 * i + 6]);
  }

  // iopt = 1: g(27) = B^T gint, the distances added in ascending order (:9491-9495)
  void gradient(const double* x, double* g) const {
    double dist[ndist], gint[ndist];
    std::vector<double> B(ndist * nx);
    distances(x, dist);
    bmatrix(x, B.data(), nullptr, nullptr);
    internal_gradient(dist, gint);
    for (int i = 0; i < nx; ++i) {
      double s = 0.0;
      for (int j = 0; j < ndist; ++j) s = s + B[j * nx + i] * gint[j];
      g[i] = s;
    }
  }

  // iopt = 2 (:9500-9590): h packed lower triangle (i = 1..27, j = 1..i)
  void hessian_packed(const double* x, double* h) const {
    double dist[ndist], gint[ndist], rr[ndist], u[ndist * 3];
    std::vector<double> B(ndist * nx), dB((size_t)ndist * nx * nx, 0.0), hint((ndist + 1) * ndist / 2, 0.0), W(ndist * nx, 0.0);
    distances(x, dist);
    bmatrix(x, B.data(), rr, u);
    auto DB = [&](int k, int a, int b) -> double& { return dB[((size_t)b * nx + a) * ndist + k]; };  // dB(k, a, b)
    {  // dB/dx (:9384-9400)
      int ij = 0;
      for (int i = 0; i < natoms; ++i)
        for (int j = 0; j < i; ++j) {
          const int ix = 3 * i, jx = 3 * j;
          const double rrij = rr[ij];
          for (int k = 0; k < 3; ++k) {
            DB(ij, ix + k, ix + k) = rrij;
            DB(ij, jx + k, jx + k) = rrij;
            DB(ij, ix + k, jx + k) = -rrij;
            DB(ij, jx + k, ix + k) = -rrij;
          }
          for (int k = 0; k < 3; ++k)
            for (int l = 0; l < 3; ++l) {
              const double d0 = u[ij * 3 + k] * u[ij * 3 + l] * rrij;
              DB(ij, ix + k, ix + l) = DB(ij, ix + k, ix + l) - d0;
              DB(ij, jx + k, jx + l) = DB(ij, jx + k, jx + l) - d0;
              DB(ij, ix + k, jx + l) = DB(ij, ix + k, jx + l) + d0;
              DB(ij, jx + k, ix + l) = DB(ij, jx + k, ix + l) + d0;
            }
          ++ij;
        }
    }
    internal_gradient(dist, gint);
    double r[3];
    for (int i = 0; i < nmorse; ++i) {  // h_morse (:9632-9645)
      const double re = morse[3 * i], al = morse[3 * i + 1], de = morse[3 * i + 2];
      const int ii = imorse[i];
      double hh = al * (re - dist[ii - 1]);
      hh = pimdk_exp(hh);
      hh = hh * (2.0 * hh - 1.0);
      hh = hh * 2.0 * (al * al) * (de * de);
      const int ij = (ii - 1) * ii / 2 + ii;
      hint[ij - 1] = hint[ij - 1] + hh;
    }
    auto hg = [&](int nd, const int* idx, const double* x0, const double* alpha, double d) {  // h_gauss (:9702-9734)
      for (int j = 0; j < nd; ++j) r[j] = dist[idx[j] - 1];
      double vv = gauss_arg(nd, r, x0, alpha);
      vv = pimdk_exp(-vv);
      vv = vv * d;
      double hh[6];
      int ij = 0;
      for (int i = 0; i < nd; ++i) {
        const double fi = (r[i] - x0[i]) * alpha[i];
        for (int j = 0; j <= i; ++j) hh[ij++] = vv * fi * (r[j] - x0[j]) * alpha[j];
        hh[ij - 1] = hh[ij - 1] - alpha[i] * vv;
      }
      ij = 0;
      for (int j = 0; j < nd; ++j)
        for (int k = 0; k <= j; ++k) {
          int ii = idx[j], jj = idx[k];
          if (jj > ii) { const int t = ii; ii = jj; jj = t; }   // iorder(jj, ii)
          const int kl = (ii - 1) * ii / 2 + jj;
          hint[kl - 1] = hint[kl - 1] + hh[ij++];
        }
    };
    for (int i = 0; i < ng1d; ++i) hg(1, &ig1d[i], &g1d[4 * i], &g1d[4 * i + 1], g1d[4 * i + 2]);
    for (int i = 0; i < ng2d; ++i) hg(2, &ig2d[2 * i], &g2d[6 * i], &g2d[6 * i + 2], g2d[6 * i + 4]);
    for (int i = 0; i < ng3d; ++i) hg(3, &ig3d[3 * i], &g3d[8 * i], &g3d[8 * i + 3], g3d[8 * i + 6]);
    int ij = 0;
    for (int i = 0; i < nx; ++i)   // h(ij) = sum_k dB(k,j,i) gint(k)  (:9555-9562)
      for (int j = 0; j <= i; ++j) {
        double s = 0.0;
        for (int k = 0; k < ndist; ++k) s = s + DB(k, j, i) * gint[k];
        h[ij++] = s;
      }
    for (int j = 0; j < nx; ++j) {   // W(k,j) = sum hint B, in the reference's interleaved order (:9565-9576)
      int kl = 0;
      for (int k = 0; k < ndist; ++k) {
        for (int l = 0; l < k; ++l) {
          W[k * nx + j] = W[k * nx + j] + hint[kl] * B[l * nx + j];
          W[l * nx + j] = W[l * nx + j] + hint[kl] * B[k * nx + j];
          ++kl;
        }
        W[k * nx + j] = W[k * nx + j] + hint[kl] * B[k * nx + j];
        ++kl;
      }
    }
    ij = 0;
    for (int i = 0; i < nx; ++i)   // h(ij) += sum_k B(k,i) W(k,j)  (:9578-9586)
      for (int j = 0; j <= i; ++j) {
        double s = h[ij];
        for (int k = 0; k < ndist; ++k) s = s + B[k * nx + i] * W[k * nx + j];
        h[ij++] = s;
      }
  }
};

}  // namespace oracle
