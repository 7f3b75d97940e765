// Malonaldehyde surface (pes_malonaldehyde.f90 `pes`, iopt = 0 / 1 / 2, behind mcmod_malon.f90's V, Vprime, Vdoubleprime).
//   malon_kernel          thread = geometry.  The 36 distances and the 36 internal-gradient sums live in shared memory,
//                         slot-major [slot][thread] (conflict-free; the slot index of a term is warp-uniform), the 3558 term
//                         records are read warp-uniformly through the read-only path (one 16-byte load per two numbers, the
//                         207 KB table stays in L1/L2), every sum is formed in the reference's order.
//   malon_hessian_kernel  CTA = geometry.  Thread 0 walks the terms (internal gradient and the 666 packed internal second
//                         derivatives are sequential sums per entry), then the CTA forms B, the hint·B product in the
//                         reference's interleaved order and the 378 packed Cartesian entries, one thread per entry.
// Built with -fmad=false: bit-identical to the oracle (oracle/malon.hpp).
#include "kernels.h"
#include "malon.cuh"

namespace pimdk {
namespace {

constexpr int kMalBlock = 128;

__device__ __forceinline__ void ld2(const double* p, double& a, double& b) {
  const double2 t = __ldg(reinterpret_cast<const double2*>(p));
  a = t.x;
  b = t.y;
}

// distances in the reference's loop order (:9356-9364); S: slot k of this thread at dist[k * stride]
__device__ __forceinline__ void mal_distances(const double* xs, double* dist, int stride) {
  int ij = 0;
#pragma unroll
  for (int i = 0; i < kMalAtoms; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) {
      const double r0 = xs[3 * i] - xs[3 * j], r1 = xs[3 * i + 1] - xs[3 * j + 1], r2 = xs[3 * i + 2] - xs[3 * j + 2];
      const double rij = r0 * r0 + r1 * r1 + r2 * r2;
      dist[(ij++) * stride] = sqrt(rij);
    }
}

// iopt = 0 (:9408-9446)
__device__ __noinline__ double mal_energy(const MalonTab* __restrict__ T, const double* dist, int stride) {
  double e = __ldg(&T->shift);
#pragma unroll 1
  for (int i = 0; i < kMalMorse; ++i) {   // v_morse (:9598-9610)
    double re, al, de, z;
    ld2(&T->morse[4 * i], re, al);
    ld2(&T->morse[4 * i + 2], de, z);
    double v = al * (re - dist[__ldg(&T->imorse[i]) * stride]);
    v = pimdk_exp(v) - 1.0;
    v = (de * v) * (de * v);
    e = e + v;
  }
#pragma unroll 1
  for (int i = 0; i < kMalG1; ++i) {      // v_gauss (:9650-9670), ndim = 1
    double p[4];
    ld2(&T->g1d[4 * i], p[0], p[1]);
    ld2(&T->g1d[4 * i + 2], p[2], p[3]);
    const double r[1] = {dist[__ldg(&T->ig1d[i]) * stride]};
    double v = mal_gauss_arg<1>(r, p, p + 1);
    v = pimdk_exp(-v) - p[3];
    e = e + v * p[2];
  }
#pragma unroll 1
  for (int i = 0; i < kMalG2; ++i) {      // ndim = 2
    double p[6];
    ld2(&T->g2d[6 * i], p[0], p[1]);
    ld2(&T->g2d[6 * i + 2], p[2], p[3]);
    ld2(&T->g2d[6 * i + 4], p[4], p[5]);
    const uint32_t w = __ldg(&T->ig2d[i]);
    const double r[2] = {dist[(w & 255) * stride], dist[((w >> 8) & 255) * stride]};
    double v = mal_gauss_arg<2>(r, p, p + 2);
    v = pimdk_exp(-v) - p[5];
    e = e + v * p[4];
  }
#pragma unroll 1
  for (int i = 0; i < kMalG3; ++i) {      // ndim = 3
    double p[8];
    ld2(&T->g3d[8 * i], p[0], p[1]);
    ld2(&T->g3d[8 * i + 2], p[2], p[3]);
    ld2(&T->g3d[8 * i + 4], p[4], p[5]);
    ld2(&T->g3d[8 * i + 6], p[6], p[7]);
    const uint32_t w = __ldg(&T->ig3d[i]);
    const double r[3] = {dist[(w & 255) * stride], dist[((w >> 8) & 255) * stride], dist[((w >> 16) & 255) * stride]};
    double v = mal_gauss_arg<3>(r, p, p + 3);
    v = pimdk_exp(-v) - p[7];
    e = e + v * p[6];
  }
  return e;
}

// one Gaussian's share of the internal gradient: f_gauss (:9673-9698) and the accumulation of :9459-9489
template <int ND>
__device__ __forceinline__ void mal_fgauss(const double* p /* x0(ND), alpha(ND), d */, uint32_t w, const double* dist, double* gint,
                                           int stride) {
  int idx[ND];
  double r[ND], gg[ND];
#pragma unroll
  for (int j = 0; j < ND; ++j) {
    idx[j] = (int)((w >> (8 * j)) & 255) * stride;
    r[j] = dist[idx[j]];
  }
  double vv = mal_gauss_arg<ND>(r, p, p + ND);
  vv = pimdk_exp(-vv);
  vv = vv * p[2 * ND];
#pragma unroll
  for (int j = 0; j < ND; ++j) gg[j] = -vv * (r[j] - p[j]) * p[ND + j];
#pragma unroll
  for (int j = 0; j < ND; ++j) gint[idx[j]] = gint[idx[j]] + gg[j];   // sequential: a term may name one distance twice
}

__device__ __noinline__ void mal_internal_gradient(const MalonTab* __restrict__ T, const double* dist, double* gint, int stride) {
#pragma unroll
  for (int k = 0; k < kMalDist; ++k) gint[k * stride] = 0.0;
#pragma unroll 1
  for (int i = 0; i < kMalMorse; ++i) {   // f_morse (:9614-9628)
    double re, al, de, z;
    ld2(&T->morse[4 * i], re, al);
    ld2(&T->morse[4 * i + 2], de, z);
    const int ii = (int)__ldg(&T->imorse[i]) * stride;
    double f = al * (re - dist[ii]);
    f = pimdk_exp(f);
    f = f * (f - 1.0);
    f = f * 2.0 * al * (de * de);
    gint[ii] = gint[ii] - f;
  }
#pragma unroll 1
  for (int i = 0; i < kMalG1; ++i) {
    double p[4];
    ld2(&T->g1d[4 * i], p[0], p[1]);
    ld2(&T->g1d[4 * i + 2], p[2], p[3]);
    mal_fgauss<1>(p, __ldg(&T->ig1d[i]), dist, gint, stride);
  }
#pragma unroll 1
  for (int i = 0; i < kMalG2; ++i) {
    double p[6];
    ld2(&T->g2d[6 * i], p[0], p[1]);
    ld2(&T->g2d[6 * i + 2], p[2], p[3]);
    ld2(&T->g2d[6 * i + 4], p[4], p[5]);
    mal_fgauss<2>(p, __ldg(&T->ig2d[i]), dist, gint, stride);
  }
#pragma unroll 1
  for (int i = 0; i < kMalG3; ++i) {
    double p[8];
    ld2(&T->g3d[8 * i], p[0], p[1]);
    ld2(&T->g3d[8 * i + 2], p[2], p[3]);
    ld2(&T->g3d[8 * i + 4], p[4], p[5]);
    ld2(&T->g3d[8 * i + 6], p[6], p[7]);
    mal_fgauss<3>(p, __ldg(&T->ig3d[i]), dist, gint, stride);
  }
}

// row `ij` of the B matrix restricted to its two atoms (:9368-9383): u = (xi - xj) * (1 / |xi - xj|)
__device__ __forceinline__ void mal_unit(const double* xs, int i, int j, double* u, double* rr) {
  double r[3] = {xs[3 * i] - xs[3 * j], xs[3 * i + 1] - xs[3 * j + 1], xs[3 * i + 2] - xs[3 * j + 2]};
  double rrij = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  rrij = 1.0 / sqrt(rrij);
#pragma unroll
  for (int k = 0; k < 3; ++k) u[k] = rrij * r[k];
  if (rr) *rr = rrij;
}

__global__ void __launch_bounds__(kMalBlock)
malon_kernel(const MalonTab* __restrict__ tab, GeomLayout L, const double* __restrict__ x, double V0, double* __restrict__ v,
             double* __restrict__ grad, long ngeom, int* __restrict__ flags) {
  extern __shared__ double mal_smem[];
  double* dist = mal_smem + threadIdx.x;                            // dist[k * kMalBlock]
  double* gint = mal_smem + kMalDist * kMalBlock + threadIdx.x;
  const long g = (long)blockIdx.x * kMalBlock + threadIdx.x;
  if (g >= ngeom) return;
  const long base = L.base(g);
  double xs[kMalDof];
#pragma unroll
  for (int d = 0; d < kMalDof; ++d) xs[d] = x[base + (long)d * L.stride_dof];
  mal_distances(xs, dist, kMalBlock);
  if (v) v[g] = mal_energy(tab, dist, kMalBlock) - V0;                // mcmod_malon.f90:15-23
  if (grad) {                                                        // mcmod_malon.f90:26-40: the same storage order
    mal_internal_gradient(tab, dist, gint, kMalBlock);
    // g(c) = sum_j B(j, c) gint(j), j ascending (:9491-9495).  B(j, c) is zero unless distance j involves the atom of c; a
    // zero entry adds +-0 to a sum that starts at +0 and can never become -0, so those terms are left out: same bits.
    bool bad = false;
#pragma unroll
    for (int i = 0; i < kMalAtoms; ++i) {
      double s[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int m = 0; m < kMalAtoms; ++m) {
        if (m == i) continue;
        double u[3];
        if (m < i) mal_unit(xs, i, m, u, nullptr); else mal_unit(xs, m, i, u, nullptr);
        const double gi = gint[(m < i ? mal_pair(i, m) : mal_pair(m, i)) * kMalBlock];
#pragma unroll
        for (int k = 0; k < 3; ++k) s[k] = s[k] + (m < i ? u[k] : -u[k]) * gi;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        grad[base + (long)(3 * i + k) * L.stride_dof] = s[k];
        bad = bad || s[k] != s[k];
      }
    }
    if (bad) atomicOr(flags, PIMDK_FLAG_NAN);
  }
}

// ---- Hessian (iopt = 2, :9500-9590; mcmod_malon.f90:43-70) --------------------------------------------------------
template <int ND>
__device__ __forceinline__ void mal_hgauss(const double* p, uint32_t w, const double* dist, double* hint) {   // h_gauss (:9702-9734)
  int idx[ND];
  double r[ND], hh[ND * (ND + 1) / 2];
#pragma unroll
  for (int j = 0; j < ND; ++j) {
    idx[j] = (int)((w >> (8 * j)) & 255);
    r[j] = dist[idx[j]];
  }
  double vv = mal_gauss_arg<ND>(r, p, p + ND);
  vv = pimdk_exp(-vv);
  vv = vv * p[2 * ND];
  int ij = 0;
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    const double fi = (r[i] - p[i]) * p[ND + i];
#pragma unroll
    for (int j = 0; j <= i; ++j) hh[ij++] = vv * fi * (r[j] - p[j]) * p[ND + j];
    hh[ij - 1] = hh[ij - 1] - p[ND + i] * vv;
  }
  ij = 0;
#pragma unroll
  for (int j = 0; j < ND; ++j)
#pragma unroll
    for (int k = 0; k <= j; ++k) {
      int ii = idx[j] + 1, jj = idx[k] + 1;               // 1-based like the reference's kl arithmetic
      if (jj > ii) { const int t = ii; ii = jj; jj = t; }   // iorder(jj, ii)
      const int kl = (ii - 1) * ii / 2 + jj - 1;
      hint[kl] = hint[kl] + hh[ij++];
    }
}

// dB(k, a, b) of :9384-9400 for distance k = (i, j): assigned +-rr on the coordinate diagonal, then -+ u u rr
__device__ __forceinline__ double mal_db(int pi, int pj, const double* u, double rr, int a, int b) {
  const int aa = a / 3, ka = a - 3 * aa, ab = b / 3, kb = b - 3 * ab;
  const bool ina = aa == pi || aa == pj, inb = ab == pi || ab == pj;
  if (!ina || !inb) return 0.0;
  const double d0 = u[ka] * u[kb] * rr;
  if (aa == ab) return (ka == kb ? rr : 0.0) - d0;
  return (ka == kb ? -rr : 0.0) + d0;
}

constexpr int kMalHessBlock = 128;
__global__ void __launch_bounds__(kMalHessBlock)
malon_hessian_kernel(const MalonTab* __restrict__ tab, GeomLayout L, const double* __restrict__ x, double* __restrict__ hess, long ngeom) {
  __shared__ double xs[kMalDof], dist[kMalDist], gint[kMalDist], rr[kMalDist], u[kMalDist * 3], hint[kMalHint];
  __shared__ double B[kMalDist * kMalDof], W[kMalDist * kMalDof];
  __shared__ unsigned char pa[kMalDist], pb[kMalDist];
  const long g = blockIdx.x;
  const long base = L.base(g);
  const int t = threadIdx.x;
  if (t < kMalDof) xs[t] = x[base + (long)t * L.stride_dof];
  for (int q = t; q < kMalHint; q += kMalHessBlock) hint[q] = 0.0;
  for (int q = t; q < kMalDist * kMalDof; q += kMalHessBlock) { B[q] = 0.0; W[q] = 0.0; }
  __syncthreads();
  if (t < kMalDist) {   // distance t = (i, j), j < i
    int i = 1;
    while (i * (i + 1) / 2 <= t) ++i;
    const int j = t - i * (i - 1) / 2;
    pa[t] = (unsigned char)i;
    pb[t] = (unsigned char)j;
    const double r0 = xs[3 * i] - xs[3 * j], r1 = xs[3 * i + 1] - xs[3 * j + 1], r2 = xs[3 * i + 2] - xs[3 * j + 2];
    dist[t] = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
    double uu[3], rrij;
    mal_unit(xs, i, j, uu, &rrij);
    rr[t] = rrij;
    for (int k = 0; k < 3; ++k) {
      u[3 * t + k] = uu[k];
      B[t * kMalDof + 3 * i + k] = uu[k];
      B[t * kMalDof + 3 * j + k] = -uu[k];
    }
  }
  __syncthreads();
  if (t == 0) {
    mal_internal_gradient(tab, dist, gint, 1);
    for (int i = 0; i < kMalMorse; ++i) {   // h_morse (:9632-9645)
      const double re = tab->morse[4 * i], al = tab->morse[4 * i + 1], de = tab->morse[4 * i + 2];
      const int ii = (int)tab->imorse[i] + 1;
      double hh = al * (re - dist[ii - 1]);
      hh = pimdk_exp(hh);
      hh = hh * (2.0 * hh - 1.0);
      hh = hh * 2.0 * (al * al) * (de * de);
      const int ij = (ii - 1) * ii / 2 + ii - 1;
      hint[ij] = hint[ij] + hh;
    }
#pragma unroll 1
    for (int i = 0; i < kMalG1; ++i) mal_hgauss<1>(&tab->g1d[4 * i], tab->ig1d[i], dist, hint);
#pragma unroll 1
    for (int i = 0; i < kMalG2; ++i) mal_hgauss<2>(&tab->g2d[6 * i], tab->ig2d[i], dist, hint);
#pragma unroll 1
    for (int i = 0; i < kMalG3; ++i) mal_hgauss<3>(&tab->g3d[8 * i], tab->ig3d[i], dist, hint);
  }
  __syncthreads();
  if (t < kMalDof) {   // W(:, j) = hint B(:, j) in the reference's interleaved order (:9565-9576)
    const int j = t;
    int kl = 0;
    for (int k = 0; k < kMalDist; ++k) {
      for (int l = 0; l < k; ++l) {
        W[k * kMalDof + j] = W[k * kMalDof + j] + hint[kl] * B[l * kMalDof + j];
        W[l * kMalDof + j] = W[l * kMalDof + j] + hint[kl] * B[k * kMalDof + j];
        ++kl;
      }
      W[k * kMalDof + j] = W[k * kMalDof + j] + hint[kl] * B[k * kMalDof + j];
      ++kl;
    }
  }
  __syncthreads();
  double* H = hess + g * kMalDof * kMalDof;
  for (int e = t; e < kMalDof * (kMalDof + 1) / 2; e += kMalHessBlock) {   // packed entry (i, j <= i)
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    const int j = e - i * (i + 1) / 2;
    double s = 0.0;
    for (int k = 0; k < kMalDist; ++k) s = s + mal_db(pa[k], pb[k], &u[3 * k], rr[k], j, i) * gint[k];   // (:9555-9562)
    for (int k = 0; k < kMalDist; ++k) s = s + B[k * kMalDof + i] * W[k * kMalDof + j];                  // (:9578-9586)
    H[i * kMalDof + j] = s;   // mcmod_malon.f90:56-67: hess(i1,j1,i2,j2) and its transpose
    H[j * kMalDof + i] = s;
  }
}

}  // namespace

cudaError_t launch_malon(const MalonTab* tab, GeomLayout L, const double* x, double V0, double* v, double* grad, long ngeom, int* flags,
                         cudaStream_t st) {
  static unsigned long long attr_mask = 0;   // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t smem = (size_t)2 * kMalDist * kMalBlock * sizeof(double);
  if (!(attr_mask & (1ull << (dev & 63)))) {
    cudaError_t e = cudaFuncSetAttribute(malon_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_mask |= 1ull << (dev & 63);
  }
  malon_kernel<<<(unsigned)((ngeom + kMalBlock - 1) / kMalBlock), kMalBlock, smem, st>>>(tab, L, x, V0, v, grad, ngeom, flags);
  return cudaGetLastError();
}
cudaError_t launch_malon_hessian(const MalonTab* tab, GeomLayout L, const double* x, double* hess, long ngeom, cudaStream_t st) {
  malon_hessian_kernel<<<(unsigned)ngeom, kMalHessBlock, 0, st>>>(tab, L, x, hess, ngeom);
  return cudaGetLastError();
}

}  // namespace pimdk
