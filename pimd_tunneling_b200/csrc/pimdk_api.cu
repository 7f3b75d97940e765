// C ABI of the hot path (include/pimdk.h): context, table upload, workspace, step loop.
// Host-side arithmetic that feeds device tables (lam, beadmass, transmatrix, cos/sin of the free
// ring-polymer rotation, PILE coefficients) uses the reference's own expressions
// (verletmodule.f90:306-338, 381-386, 515-531) evaluated once per call in FP64 on the host.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/pimdk.h"
#include "ccpol_grad.cuh"
#include "kernels.h"
#include "nm_device.cuh"
#include "malon.cuh"
#include "watmeth.cuh"

using namespace pimdk;

namespace {

const double PI_TRUNC = 3.14159265358979;  // instantonmod.f90:4 (truncated literal, used by init_nm and gauleg)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {  // page-locked host staging: asynchronous copies in both directions with one synchronisation
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct Prof {
  double ms = 0.0;
  long launches = 0;
};

struct Ctx {
  bool inited = false;
  int device = 0, num_sms = 148;
  cudaStream_t stream = 0;
  std::string data_dir = ".";
  double malon_V0 = 0.0;   // module variable V0 of mcmod_malon
  std::string err;
  int mode = PIMDK_MODE_STRICT;
  bool fused = true;  // small systems: one persistent warp-per-ring-polymer kernel
  long restart = 0, restartnmc = 0;  // module verletint's restart / restartnmc (verletmodule.f90:10)
  long sums_n = 0;                   // trajectories in wDhSum (running sums of the last propagate call)
  bool andersen_carry = false;       // pimdk_set_andersen_carry: the next Andersen call continues the collision clocks
  long clock_n = 0;                  // trajectories whose (count, rkick) the last Andersen call left in wCount / wKick
  double dhdrlimit = -1.0;           // pimdk_set_dhdrlimit (namelist dHdrlimit, pimd_par.f90:88): < 0 = no outlier guard
  int rp_npath = 0;                  // spline path and per-trajectory xi for the re-initialisation (wPath, wXi)
  long rp_ntraj = 0;
  long chunk_traj = 0;               // pimdk_set_propagate_chunk: trajectories per chunk of the host-buffer propagate (0 = automatic)
  long sum_off = 0, sum_total = 0;   // chunked host-buffer propagate: this chunk's offset into wDhSum / whole batch
  cudaStream_t copy_stream = nullptr; // host<->device copies of the chunked propagate, overlapped with compute
  cudaEvent_t ev_in[2] = {nullptr, nullptr};
  // PES
  PesKind pes = PES_NONE;
  int ndim = 0, natom = 0;
  SimplePesParams sp{};
  bool tab_loaded = false;
  CcpolHost htab;
  CcpolDev hdev;
  agrad::CcpolGradTab hgrad;   // rigid-body coefficients and sweep tables of the analytic-gradient mode
  DevBuf dtab, dgtab, dwm, dmal;   // dwm: WatMethTab, dmal: MalonTab
  // normal modes
  bool nm_ready = false;
  int n = 0, nm_ndim = 0, nm_natom = 0;
  double betan = 0.0, tau = 1.0;
  std::vector<double> mass, lam, beadmass, T;
  DevBuf dT, dsA, dsB, dlamb2, dmass, dtabs;  // dtabs: 8 per-call (natom,n) tables
  // workspaces
  DevBuf wCc, wP, wQ, wG, wGn, wV, wX, wAux, wCount, wKick, wFlags, wGid, wA, wB, wDbdl, wDhdr, wPp, wMisc, wDhSum, wX2, wPp2, wUmIn, wUmOut, wHgp, wHgm, wHess, wBand, wDense, wEig, wWork, wSums, wBV, wPath, wXi, wReinit;
  PinBuf hUm;
  // profiling
  bool profiling = false;
  std::map<std::string, Prof> prof;
  struct Span { std::string fam; cudaEvent_t a, b; };
  std::vector<Span> spans;
  long launch_count = 0;
  long last_nan_traj = -1;
};

Ctx g;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g.err = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) return fail(PIMDK_ECUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e__), \
                                        __FILE__, __LINE__, cudaGetErrorString(e__));              \
  } while (0)
#define NEED_INIT()                                                          \
  do {                                                                       \
    if (!g.inited) return fail(PIMDK_EINVAL, "pimdk_init has not been called"); \
  } while (0)

// profiling spans: events on the library stream around a kernel family
struct Scope {
  const char* fam;
  bool on;
  cudaEvent_t a{}, b{};
  int nlaunch;
  Scope(const char* f, int nl = 1) : fam(f), on(g.profiling), nlaunch(nl) {
    g.launch_count += nl;
    if (on) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, g.stream);
    }
  }
  ~Scope() {
    if (on) {
      cudaEventRecord(b, g.stream);
      g.spans.push_back({fam, a, b});
      g.prof[fam].launches += nlaunch;
    }
  }
};

void resolve_spans() {
  for (auto& s : g.spans) {
    float ms = 0.f;
    if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess)
      g.prof[s.fam].ms += ms;
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  g.spans.clear();
}

int check_flags(bool sync_first) {
  int fl = 0;
  if (sync_first) CU(cudaStreamSynchronize(g.stream));
  CU(cudaMemcpyAsync(&fl, g.wFlags.p, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  if (fl & PIMDK_FLAG_BADGID)
    return fail(PIMDK_EINVAL, "trajectory id outside 0 .. 2^32-1 (the Philox counter carries 32 bits of it: RNG contract, pimdk.h)");
  if (fl & PIMDK_FLAG_NOCONV) return fail(PIMDK_ENOCONV, "No convergence in indN_iter");
  if (fl & PIMDK_FLAG_NAN) return fail(PIMDK_ENAN, "NaN in pot propagation");
  return PIMDK_OK;
}

// wFlags: [int flags][int pad][long long slot] — the slot serves the NaN-trajectory lookup (no allocation on the error path)
int clear_flags() {
  CU(g.wFlags.ensure(16));
  CU(cudaMemsetAsync(g.wFlags.p, 0, sizeof(int), g.stream));
  return PIMDK_OK;
}

int ensure_ccpol_tables(int isurf) {
  if (g.tab_loaded && g.htab.isurf == isurf) return PIMDK_OK;
  g.tab_loaded = false;
  const char* m = load_ccpol_tables(g.data_dir.c_str(), isurf, &g.htab);
  if (m[0]) return fail(PIMDK_EDATA, "%s", m);
  g.tab_loaded = true;
  return PIMDK_OK;
}

int upload_ccpol_dev() {
  ccpol_host_tables_strict(&g.hdev);   // the pair-sum, rigid and sweep kernels take their tables as kernel parameters
  ccpol_host_tables_fast(&g.hdev);
  ccpol_host_tables_analytic(&g.hdev);
  CU(g.dtab.ensure(sizeof(CcpolDev)));
  CU(cudaMemcpyAsync(g.dtab.p, &g.hdev, sizeof(CcpolDev), cudaMemcpyHostToDevice, g.stream));
  CU(g.dgtab.ensure(sizeof(agrad::CcpolGradTab)));
  CU(cudaMemcpyAsync(g.dgtab.p, &g.hgrad, sizeof(agrad::CcpolGradTab), cudaMemcpyHostToDevice, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  return PIMDK_OK;
}

// energy and/or gradient of `ngeom` geometries laid out per L (device pointers)
int pes_eval_dev(GeomLayout L, double* x, double* v, double* grad, long ngeom, int write_drift) {
  if (g.pes == PES_NONE) return fail(PIMDK_EINVAL, "no PES selected (pimdk_pes_select)");
  if (ngeom <= 0) return PIMDK_OK;
  int* flags = g.wFlags.as<int>();
  if (g.pes == PES_CCPOL) {
    const CcpolDev* tab = g.dtab.as<CcpolDev>();
    if (g.mode == PIMDK_MODE_ANALYTIC && grad) {
      // opt-in: analytic gradient (and the energy with it) in one pipeline run; x is not perturbed, so there is no drift
      if (g.hdev.iembed != 2 || g.hdev.potparts_old)
        return fail(PIMDK_EINVAL, "the analytic-gradient mode covers the Radau-embedded surfaces with potparts (isurf 3 and 10)");
      const long cap = 262144;
      const size_t wb = (size_t)(ngeom < cap ? ngeom : cap) * ccpol_analytic_bytes_per_geom();
      CU(g.wCc.ensure(wb));
      Scope s("pes", (int)ccpol_analytic_launches(ngeom, g.hdev.icc, wb));
      CU(launch_ccpol_analytic(tab, g.dgtab.as<agrad::CcpolGradTab>(), g.hdev.iemonomer, g.hdev.icc, g.hdev.V0, L, x, v, grad, ngeom,
                               flags, g.wCc.as<double>(), wb, g.num_sms, g.stream));
      return PIMDK_OK;
    }
    const bool fast = g.mode == PIMDK_MODE_FAST;
    for (int pass = 0; pass < 2; ++pass) {  // energies, then gradients (two pipeline runs, like V then Vprime)
      double* vv = pass == 0 ? v : nullptr;
      double* gg = pass == 1 ? grad : nullptr;
      if (!vv && !gg) continue;
      const size_t wb = fast ? ccpol_work_bytes_fast(ngeom, gg != nullptr) : ccpol_work_bytes_strict(ngeom, gg != nullptr);
      CU(g.wCc.ensure(wb));
      Scope s("pes", (int)(fast ? ccpol_launches_fast(ngeom, gg != nullptr, g.hdev.icc, g.wCc.cap) : ccpol_launches_strict(ngeom, gg != nullptr, g.hdev.icc, g.wCc.cap)));
      CU(fast ? launch_ccpol_fast(tab, g.hdev.iemonomer, g.hdev.iembed, g.hdev.icc, g.hdev.potparts_old, g.hdev.V0, L, x, vv, gg, ngeom, write_drift, flags,
                                  g.wCc.as<double>(), g.wCc.cap, g.stream)
              : launch_ccpol_strict(tab, g.hdev.iemonomer, g.hdev.iembed, g.hdev.icc, g.hdev.potparts_old, g.hdev.V0, L, x, vv, gg, ngeom, write_drift, flags,
                                    g.wCc.as<double>(), g.wCc.cap, g.stream));
    }
  } else if (g.pes == PES_WATMETH) {
    Scope s("pes");
    CU(launch_watmeth(g.dwm.as<WatMethTab>(), L, x, v, grad, ngeom, flags, g.stream));
  } else if (g.pes == PES_MALON) {
    Scope s("pes");
    CU(launch_malon(g.dmal.as<MalonTab>(), L, x, g.malon_V0, v, grad, ngeom, flags, g.stream));
  } else {
    Scope s("pes");
    CU(launch_simple_pes(g.pes, g.sp, L, x, v, grad, ngeom, flags, g.stream));
  }
  return PIMDK_OK;
}

NmTables nm_tables_base() {
  NmTables nm{};
  nm.T = g.dT.as<double>();
  nm.sA = g.dsA.as<double>();
  nm.sB = g.dsB.as<double>();
  nm.lamb2 = g.dlamb2.as<double>();
  nm.mass = g.dmass.as<double>();
  nm.n = g.n;
  nm.ndim = g.nm_ndim;
  nm.natom = g.nm_natom;
  nm.ndof = g.nm_ndim * g.nm_natom;
  nm.norm = std::sqrt(2.0 / (double)(g.n + 1));
  nm.stdev = std::sqrt(1.0 / g.betan);
  return nm;
}

// per-call (natom,n) tables that depend on dt, gamma, cayley
int build_step_tables(NmTables* nm, double dt, double gamma, int cayley) {
  const int n = g.n, natom = g.nm_natom;
  const size_t cnt = (size_t)natom * n;
  std::vector<double> h(8 * cnt);
  double *cosw = &h[0], *sinw = &h[cnt], *omega = &h[2 * cnt], *bmass = &h[3 * cnt], *wbm = &h[4 * cnt],
         *c1sq = &h[5 * cnt], *cnoise = &h[6 * cnt], *sigp = &h[7 * cnt];
  const double time = 0.5 * dt;
  for (int a = 0; a < natom; ++a)
    for (int k = 0; k < n; ++k) {
      const size_t i = (size_t)a * n + k;
      const double bm = g.beadmass[(size_t)k * natom + a];  // beadmass(atom,k)
      const double om = std::sqrt(g.mass[a] / bm) * g.lam[k];  // step_nm :519
      omega[i] = om;
      bmass[i] = bm;
      cosw[i] = std::cos(time * om);
      sinw[i] = std::sin(om * time);
      wbm[i] = om * bm;
      const double c1 = std::exp(-gamma * dt * g.lam[k] * std::sqrt(g.mass[a] / bm));  // :383
      const double c2 = std::sqrt(1.0 - c1 * c1);                                       // :384
      c1sq[i] = c1 * c1;
      cnoise[i] = std::sqrt(bm / g.betan) * c2 * std::sqrt(1.0 + c1 * c1);  // :651-652
      sigp[i] = std::sqrt(bm);
    }
  CU(g.dtabs.ensure(h.size() * sizeof(double)));
  CU(cudaMemcpyAsync(g.dtabs.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  double* d = g.dtabs.as<double>();
  nm->cosw = d;
  nm->sinw = d + cnt;
  nm->omega = d + 2 * cnt;
  nm->bmass = d + 3 * cnt;
  nm->wbm = d + 4 * cnt;
  nm->c1sq = d + 5 * cnt;
  nm->cnoise = d + 6 * cnt;
  nm->sigp = d + 7 * cnt;
  nm->cayley = cayley;
  nm->time = time;
  return PIMDK_OK;
}

__global__ void first_nan_kernel(const double* __restrict__ p, long per_traj, long ntraj, long long* out) {
  const long total = per_traj * ntraj;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
    if (p[e] != p[e]) atomicMin(out, (long long)(e / per_traj));
}

// RNG contract: the Philox counter carries the low 32 bits of the global trajectory id; larger ids would alias streams
__global__ void check_gid_kernel(const int64_t* __restrict__ gid, long ntraj, int* __restrict__ flags) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < ntraj && (gid[t] < 0 || gid[t] > 0xffffffffLL)) atomicOr(flags, PIMDK_FLAG_BADGID);
}

// init_path positions: x(k,dof,traj) = splint(lampath, path(:,dof), splinepath(:,dof), (k-1)*xi/(n-1))
// (verletmodule.f90:39-48; splint/locate instantonmod.f90:500-524, 560-596)
__global__ void init_path_kernel(int n, int ndof, int npath, const double* __restrict__ lampath,
                                 const double* __restrict__ path, const double* __restrict__ spl,
                                 const double* __restrict__ xi, long ntraj, double* __restrict__ x) {
  const long total = ntraj * (long)ndof * n;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int k = (int)(e % n);
    const long r = e / n;
    const int dof = (int)(r % ndof);
    const long traj = r / ndof;
    const double xv = (double)k * xi[traj] / (double)(n - 1);
    x[e] = splint_at(lampath, path + (long)dof * npath, spl + (long)dof * npath, npath, xv);
  }
}


// Per-lambda statistics of the local shard on the device (pimd_par.f90:379, 397-409): one CTA per lambda point sums
// I = dHdr/betan**2, I**2 and 1 over the shard's trajectories with that lambda index (global id / nrep) in a fixed
// order (thread-strided partial sums, then a shared-memory tree), so a given shard always produces the same bits.
constexpr int kTiThreads = 256;
__global__ void __launch_bounds__(kTiThreads)
ti_partial_sums_kernel(const double* __restrict__ dHdr, const int64_t* __restrict__ gid, long ntraj, long nrep, long nintegral,
                       double betan, double* __restrict__ sums, int* __restrict__ flags) {
  __shared__ double sh[3][kTiThreads];
  const long il = blockIdx.x;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (long t = threadIdx.x; t < ntraj; t += kTiThreads) {
    const long id = gid ? (long)gid[t] : t;
    const long l = id / nrep;
    if (il == 0 && (id < 0 || l >= nintegral)) atomicOr(flags, PIMDK_FLAG_BADGID);
    if (l != il) continue;
    const double I = dHdr[t] / (betan * betan);
    s0 = s0 + I;
    s1 = s1 + I * I;
    s2 = s2 + 1.0;
  }
  sh[0][threadIdx.x] = s0;
  sh[1][threadIdx.x] = s1;
  sh[2][threadIdx.x] = s2;
  __syncthreads();
  for (int w = kTiThreads / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w)
      for (int q = 0; q < 3; ++q) sh[q][threadIdx.x] = sh[q][threadIdx.x] + sh[q][threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x < 3) sums[3 * il + threadIdx.x] = sh[threadIdx.x][0];
}

// NCCL is bound at run time (like cuSOLVER): libpimdk.so itself needs only the CUDA runtime, and a single-GPU user
// needs no NCCL at all.  The communicator belongs to the library (SURVEY 8(b) ownership row).
struct NcclId { char internal[128]; };
struct Nccl {
  void* lib = nullptr;
  void* comm = nullptr;
  int rank = 0, nranks = 1;
  int (*get_unique_id)(NcclId*) = nullptr;
  int (*comm_init_rank)(void**, int, NcclId, int) = nullptr;
  int (*all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*comm_destroy)(void*) = nullptr;
  const char* (*error_string)(int) = nullptr;
  int (*get_version)(int*) = nullptr;
} nc;
int nccl_load() {
  if (nc.lib) return PIMDK_OK;
  const char* env = getenv("PIMDK_NCCL_LIB");
  const char* names[] = {env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    nc.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (nc.lib) break;
  }
  if (!nc.lib) return fail(PIMDK_ECUDA, "multi-GPU needs NCCL (libnccl.so.2 not found: %s; set PIMDK_NCCL_LIB)", dlerror());
  nc.get_unique_id = (int (*)(NcclId*))dlsym(nc.lib, "ncclGetUniqueId");
  nc.comm_init_rank = (int (*)(void**, int, NcclId, int))dlsym(nc.lib, "ncclCommInitRank");
  nc.all_reduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(nc.lib, "ncclAllReduce");
  nc.comm_destroy = (int (*)(void*))dlsym(nc.lib, "ncclCommDestroy");
  nc.error_string = (const char* (*)(int))dlsym(nc.lib, "ncclGetErrorString");
  nc.get_version = (int (*)(int*))dlsym(nc.lib, "ncclGetVersion");
  if (!nc.get_unique_id || !nc.comm_init_rank || !nc.all_reduce || !nc.comm_destroy || !nc.error_string) {
    nc.lib = nullptr;
    return fail(PIMDK_ECUDA, "NCCL symbols missing");
  }
  return PIMDK_OK;
}
#define NCCLCHK(call)                                                                                     \
  do {                                                                                                    \
    int r__ = (call);                                                                                     \
    if (r__ != 0) return fail(PIMDK_ECUDA, "NCCL error %d at %s:%d (%s)", r__, __FILE__, __LINE__, nc.error_string(r__)); \
  } while (0)

}  // namespace

extern "C" {

const char* pimdk_last_error(void) { return g.err.c_str(); }

int pimdk_init(pimdk_int device, const char* data_dir) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(PIMDK_ENODEV, "no CUDA device available (%s); this library has no CPU path",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device >= 0) {
    if (device >= ndev) return fail(PIMDK_EINVAL, "device %lld out of range (%d devices)", (long long)device, ndev);
    CU(cudaSetDevice((int)device));
  }
  CU(cudaGetDevice(&g.device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, g.device));
  g.num_sms = prop.multiProcessorCount;
  if (prop.major < 10)
    return fail(PIMDK_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", g.device, prop.major,
                prop.minor);
  g.data_dir = data_dir ? data_dir : ".";
  g.inited = true;
  g.err.clear();
  return clear_flags();
}

int pimdk_finalize(void) {
  if (!g.inited) return PIMDK_OK;
  cudaStreamSynchronize(g.stream);
  resolve_spans();
  DevBuf* bufs[] = {&g.dtab, &g.dgtab, &g.dwm, &g.dmal, &g.dT, &g.dsA, &g.dsB, &g.dlamb2, &g.dmass, &g.dtabs, &g.wCc, &g.wP, &g.wQ, &g.wG, &g.wGn,
                    &g.wV, &g.wX, &g.wAux, &g.wCount, &g.wKick, &g.wFlags, &g.wGid, &g.wA, &g.wB, &g.wDbdl,
                    &g.wDhdr, &g.wPp, &g.wMisc, &g.wDhSum, &g.wX2, &g.wPp2, &g.wUmIn, &g.wUmOut, &g.wHgp, &g.wHgm, &g.wHess, &g.wBand,
                    &g.wDense, &g.wEig, &g.wWork, &g.wSums, &g.wBV, &g.wPath, &g.wXi, &g.wReinit};
  for (DevBuf* b : bufs) b->release();
  g.hUm.release();
  if (nc.comm) {
    nc.comm_destroy(nc.comm);
    nc.comm = nullptr;
    nc.rank = 0;
    nc.nranks = 1;
  }
  if (g.copy_stream) cudaStreamDestroy(g.copy_stream);
  g.copy_stream = nullptr;
  for (cudaEvent_t& e : g.ev_in) {
    if (e) cudaEventDestroy(e);
    e = nullptr;
  }
  g.inited = false;
  g.andersen_carry = false;
  g.clock_n = 0;
  g.dhdrlimit = -1.0;
  g.rp_npath = 0;
  g.rp_ntraj = 0;
  g.nm_ready = false;
  g.pes = PES_NONE;
  g.tab_loaded = false;
  g.stream = 0;
  return PIMDK_OK;
}

int pimdk_set_stream(void* s) {
  g.stream = reinterpret_cast<cudaStream_t>(s);
  return PIMDK_OK;
}

int pimdk_set_gemm(pimdk_int kind) {
  if (kind < 0 || kind > 3) return fail(PIMDK_EINVAL, "gemm kind must be 0 (DFMA), 1 (DMMA, tile chosen by size), 2 (DMMA 128x64) or 3 (DMMA 128x128)");
  set_nm_gemm_dmma((int)kind);
  return PIMDK_OK;
}

int pimdk_set_fused(pimdk_int enable) {
  g.fused = enable != 0;
  return PIMDK_OK;
}

int pimdk_set_mode(pimdk_int mode) {
  if (mode != PIMDK_MODE_STRICT && mode != PIMDK_MODE_FAST && mode != PIMDK_MODE_ANALYTIC) return fail(PIMDK_EINVAL, "unknown mode");
  g.mode = (int)mode;
  return PIMDK_OK;
}

int pimdk_pes_select(const char* name, const double* pp, pimdk_int np) {
  NEED_INIT();
  std::string s(name ? name : "");
  if (s == "1d") {  // mcmod_1d.f90:8-12
    g.pes = PES_1D;
    g.ndim = 1;
    g.natom = 1;
    g.sp = SimplePesParams{};
    g.sp.Vheight = np > 0 ? pp[0] : 1.0;
    g.sp.x0 = np > 1 ? pp[1] : 1.0;
    g.sp.ndof = 1;
    return PIMDK_OK;
  }
  if (s == "2dtest") {  // mcmod_2dtest.f90:11-27
    g.pes = PES_2DTEST;
    g.ndim = 2;
    g.natom = 1;
    g.sp = SimplePesParams{};
    g.sp.a0 = np > 0 ? pp[0] : 2.0;
    g.sp.b0 = np > 1 ? pp[1] : 0.2;
    const double rho0 = np > 2 ? pp[2] : 3.0;
    const int m = 6;
    for (int k = 1; k <= m; ++k) {
      g.sp.wx[k - 1] = rho0 * std::cos((double)k * 2.0 * PI_TRUNC / (double)m);
      g.sp.wy[k - 1] = rho0 * std::sin((double)k * 2.0 * PI_TRUNC / (double)m);
    }
    g.sp.V0 = 0.0;
    g.sp.ndof = 2;
    return PIMDK_OK;
  }
  if (s == "so2") {  // mcmod_so2.f90:10-15: harmonic ring, omegaforce = 10000, r0 = 20
    g.pes = PES_SO2;
    g.ndim = 2;
    g.natom = 1;
    g.sp = SimplePesParams{};
    g.sp.omegaforce = np > 0 ? pp[0] : 10000.0;
    g.sp.r0 = np > 1 ? pp[1] : 20.0;
    g.sp.V0 = 0.0;
    g.sp.ndof = 2;
    return PIMDK_OK;
  }
  if (s == "watmeth") {  // mcmod_watmeth.f90:10-13 + the unit conversions of wmrb (watermethane.f90:279-287)
    WatMethTab t;
    build_watmeth_tab(&t);
    CU(g.dwm.ensure(sizeof(WatMethTab)));
    CU(cudaMemcpyAsync(g.dwm.p, &t, sizeof(WatMethTab), cudaMemcpyHostToDevice, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    g.pes = PES_WATMETH;
    g.ndim = 3;
    g.natom = kWmSites;
    return PIMDK_OK;
  }
  if (s == "malon") {  // mcmod_malon.f90:10-13: V_init does nothing, the fit is in pes' DATA statements (here: data/malonaldehyde.tbl)
    std::vector<unsigned char> hb(sizeof(MalonTab));
    MalonTab* t = reinterpret_cast<MalonTab*>(hb.data());
    const char* m = load_malon_tab(g.data_dir.c_str(), t);
    if (m[0]) return fail(PIMDK_EDATA, "%s", m);
    CU(g.dmal.ensure(sizeof(MalonTab)));
    CU(cudaMemcpyAsync(g.dmal.p, t, sizeof(MalonTab), cudaMemcpyHostToDevice, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    g.pes = PES_MALON;
    g.ndim = 3;
    g.natom = kMalAtoms;
    g.malon_V0 = 0.0;
    return PIMDK_OK;
  }
  if (s == "ccpol8sf") {  // mcmod_waterdimer_ccpol.f90:9-16 -> init_ccpol(3,1,1,0)
    const int iemon = np > 0 ? (int)pp[0] : 1;
    const int isurf = np > 1 ? (int)pp[1] : 3;
    if (iemon != 0 && iemon != 1) return fail(PIMDK_EINVAL, "wrong value of iemonomer");
    if (isurf < 1 || isurf > 10) return fail(PIMDK_EINVAL, "wrong value of isurf");
    int rc = ensure_ccpol_tables(isurf);
    if (rc) return rc;
    const char* m = build_ccpol_dev(g.htab, iemon, &g.hdev);
    if (m[0]) return fail(PIMDK_EDATA, "%s", m);
    m = agrad::build_grad_tab(g.hdev, &g.hgrad);
    if (m[0]) return fail(PIMDK_EDATA, "%s", m);
    rc = upload_ccpol_dev();
    if (rc) return rc;
    g.pes = PES_CCPOL;
    g.ndim = 3;
    g.natom = 6;
    return PIMDK_OK;
  }
  return fail(PIMDK_EINVAL, "unknown PES '%s' (1d, 2dtest, so2, watmeth, malon, ccpol8sf)", s.c_str());
}

int pimdk_pes_info(pimdk_int* ndim, pimdk_int* natom) {
  if (g.pes == PES_NONE) return fail(PIMDK_EINVAL, "no PES selected");
  if (ndim) *ndim = g.ndim;
  if (natom) *natom = g.natom;
  return PIMDK_OK;
}

int pimdk_pes_set_v0(double v0) {
  NEED_INIT();
  if (g.pes == PES_CCPOL) {
    g.hdev.V0 = v0;
    return upload_ccpol_dev();
  }
  if (g.pes == PES_2DTEST || g.pes == PES_SO2) g.sp.V0 = v0;  // mcmod_1d's V ignores V0 (mcmod_1d.f90:20)
  if (g.pes == PES_MALON) g.malon_V0 = v0;                     // mcmod_malon.f90:21
  return PIMDK_OK;
}

static int check_dims(pimdk_int ndim, pimdk_int natom) {
  if (g.pes == PES_NONE) return fail(PIMDK_EINVAL, "no PES selected (pimdk_pes_select)");
  if (g.pes == PES_1D) {  // any shape: the 1D surface sums over all components (mcmod_1d.f90:20)
    g.sp.ndof = (int)(ndim * natom);
    g.ndim = (int)ndim;
    g.natom = (int)natom;
    return PIMDK_OK;
  }
  if (ndim != g.ndim || natom != g.natom)
    return fail(PIMDK_EINVAL, "PES expects ndim=%d natom=%d, got %lld %lld", g.ndim, g.natom, (long long)ndim,
                (long long)natom);
  return PIMDK_OK;
}

int pimdk_pes_eval_dev(pimdk_int nbatch, pimdk_int ndim, pimdk_int natom, const double* x, double* v, double* grad) {
  NEED_INIT();
  int rc = check_dims(ndim, natom);
  if (rc) return rc;
  const long ndof = ndim * natom;
  GeomLayout L{1, ndof, 0, 1};
  rc = clear_flags();
  if (rc) return rc;
  double* xw = const_cast<double*>(x);
  if (grad && g.pes == PES_CCPOL) {  // keep the caller's x untouched: FD perturbation works on a copy
    CU(g.wX.ensure(sizeof(double) * nbatch * ndof));
    CU(cudaMemcpyAsync(g.wX.p, x, sizeof(double) * nbatch * ndof, cudaMemcpyDeviceToDevice, g.stream));
    xw = g.wX.as<double>();
  }
  rc = pes_eval_dev(L, xw, v, grad, nbatch, 0);
  if (rc) return rc;
  return check_flags(false);
}

static int pes_eval_host(pimdk_int nbatch, pimdk_int ndim, pimdk_int natom, double* x, double* v, double* grad,
                         int inplace) {
  NEED_INIT();
  int rc = check_dims(ndim, natom);
  if (rc) return rc;
  if (nbatch <= 0) return PIMDK_OK;
  const size_t nx = (size_t)nbatch * ndim * natom;
  CU(g.wX.ensure(nx * sizeof(double)));
  CU(g.wG.ensure(nx * sizeof(double)));
  CU(g.wV.ensure((size_t)nbatch * sizeof(double)));
  CU(cudaMemcpyAsync(g.wX.p, x, nx * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  rc = clear_flags();
  if (rc) return rc;
  GeomLayout L{1, (long)(ndim * natom), 0, 1};
  rc = pes_eval_dev(L, g.wX.as<double>(), v ? g.wV.as<double>() : nullptr, grad ? g.wG.as<double>() : nullptr, nbatch,
                    inplace);
  if (rc) return rc;
  if (v) CU(cudaMemcpyAsync(v, g.wV.p, (size_t)nbatch * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  if (grad) CU(cudaMemcpyAsync(grad, g.wG.p, nx * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  if (inplace) CU(cudaMemcpyAsync(x, g.wX.p, nx * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  return check_flags(false);
}

int pimdk_pes_eval(pimdk_int nbatch, pimdk_int ndim, pimdk_int natom, const double* x, double* v, double* grad) {
  return pes_eval_host(nbatch, ndim, natom, const_cast<double*>(x), v, grad, 0);
}

int pimdk_pes_vprime_inplace(pimdk_int nbatch, pimdk_int ndim, pimdk_int natom, double* x, double* grad) {
  if (!grad) return fail(PIMDK_EINVAL, "grad must not be NULL");
  return pes_eval_host(nbatch, ndim, natom, x, nullptr, grad, 1);
}

// UMforceenergy for npoly independent ring polymers x(n,ndim,natom,npoly) with end points a (shared) and
// b(ndim,natom,npoly): one PES pass over all npoly*n beads, spring terms and the ordered UM sum per polymer.
static int um_forceenergy_impl(pimdk_int npoly, pimdk_int n, pimdk_int ndim, pimdk_int natom, const double* x,
                               const double* a, const double* b, const double* mass, double betan, pimdk_int fixedends,
                               double* f, double* gout) {
  int rc = check_dims(ndim, natom);
  if (rc) return rc;
  if (n < 2) return fail(PIMDK_EINVAL, "n must be >= 2");
  if (npoly < 1 || npoly > 65535) return fail(PIMDK_EINVAL, "npoly must be in 1..65535");
  if (fixedends && (!a || !b)) return fail(PIMDK_EINVAL, "fixedends needs a and b");
  // This call sits inside L-BFGS-B's reverse-communication loop (instantonmod.f90:741-767): small problems, so its
  // cost is latency.  Inputs are packed into one page-locked block (one H2D copy), outputs and the flag word come
  // back through it as well, and there is a single stream synchronisation.
  const long ndof = ndim * natom;
  const size_t nx = (size_t)n * ndof, nxt = nx * (size_t)npoly;
  const size_t nin = nxt + ndof + (size_t)npoly * ndof + natom, nout = nxt + (size_t)npoly + 1;
  //   [x | a | b | mass]  /  [g | UM(npoly) | flags]
  CU(g.wUmIn.ensure(nin * sizeof(double)));
  CU(g.wUmOut.ensure(nout * sizeof(double)));
  CU(g.wG.ensure(nxt * sizeof(double)));
  CU(g.wV.ensure((size_t)n * npoly * sizeof(double)));
  CU(g.hUm.ensure((nin + nout) * sizeof(double)));
  double* hin = g.hUm.as<double>();
  double* hout = hin + nin;
  std::memcpy(hin, x, nxt * sizeof(double));
  if (fixedends) {
    std::memcpy(hin + nxt, a, ndof * sizeof(double));
    std::memcpy(hin + nxt + ndof, b, (size_t)npoly * ndof * sizeof(double));
  }
  std::memcpy(hin + nxt + ndof + (size_t)npoly * ndof, mass, natom * sizeof(double));
  double* dx = g.wUmIn.as<double>();
  double *da = dx + nxt, *db = da + ndof, *dm = db + (size_t)npoly * ndof;
  double* dg = g.wUmOut.as<double>();
  double* dum = dg + nxt;
  CU(cudaMemcpyAsync(dx, hin, nin * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  rc = clear_flags();
  if (rc) return rc;
  GeomLayout L{n, ndof * n, 1, n};
  // x is intent(in) in UM*: the FD perturbation of ccpol works on a copy and its drift is dropped
  double* xw = dx;
  if (gout && g.pes == PES_CCPOL) {
    CU(g.wAux.ensure(nxt * sizeof(double)));
    CU(cudaMemcpyAsync(g.wAux.p, dx, nxt * sizeof(double), cudaMemcpyDeviceToDevice, g.stream));
    xw = g.wAux.as<double>();
  }
  rc = pes_eval_dev(L, xw, f ? g.wV.as<double>() : nullptr, gout ? g.wG.as<double>() : nullptr, (long)n * npoly, 0);
  if (rc) return rc;
  {
    Scope s("um", (f ? 1 : 0) + (gout ? 1 : 0));
    CU(launch_um((int)npoly, (int)n, (int)ndim, (int)natom, dx, da, fixedends ? db : nullptr, dm, betan, fixedends != 0,
                 g.wV.as<double>(), g.wG.as<double>(), f ? dum : nullptr, gout ? dg : nullptr, g.stream));
  }
  CU(cudaMemcpyAsync(hout, dg, (nxt + npoly) * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaMemcpyAsync(hout + nxt + npoly, g.wFlags.p, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  if (f) std::memcpy(f, hout + nxt, (size_t)npoly * sizeof(double));
  if (gout) std::memcpy(gout, hout, nxt * sizeof(double));
  int fl = 0;
  std::memcpy(&fl, hout + nxt + npoly, sizeof(int));
  if (fl & PIMDK_FLAG_NOCONV) return fail(PIMDK_ENOCONV, "No convergence in indN_iter");
  if (fl & PIMDK_FLAG_NAN) return fail(PIMDK_ENAN, "NaN in pot propagation");
  return PIMDK_OK;
}

int pimdk_um_forceenergy(pimdk_int n, pimdk_int ndim, pimdk_int natom, const double* x, const double* a,
                         const double* b, const double* mass, double betan, pimdk_int fixedends, double* f,
                         double* gout) {
  NEED_INIT();
  return um_forceenergy_impl(1, n, ndim, natom, x, a, b, mass, betan, fixedends, f, gout);
}

int pimdk_um_forceenergy_batch(pimdk_int npoly, pimdk_int n, pimdk_int ndim, pimdk_int natom, const double* x,
                               const double* a, const double* b, const double* mass, double betan, pimdk_int fixedends,
                               double* f, double* gout) {
  NEED_INIT();
  return um_forceenergy_impl(npoly, n, ndim, natom, x, a, b, mass, betan, fixedends, f, gout);
}

// ---- second derivatives: Vdoubleprime, UMhessian, detJ ------------------------------------------
// Hessians of `ngeom` geometries laid out per L (device pointers); x is perturbed in place like the reference's.
static int pes_hessian_dev(GeomLayout L, double* x, double* hess, long ngeom, int ndim, int natom) {
  if (g.pes == PES_NONE) return fail(PIMDK_EINVAL, "no PES selected (pimdk_pes_select)");
  if (ngeom <= 0) return PIMDK_OK;
  const int nd = ndim * natom;
  if (g.pes == PES_WATMETH) {
    Scope s("hess");
    CU(launch_watmeth_hessian(g.dwm.as<WatMethTab>(), L, x, hess, ngeom, g.stream));
    return PIMDK_OK;
  }
  if (g.pes == PES_MALON) {   // mcmod_malon.f90:43-70: analytic (pes, iopt = 2)
    Scope s("hess");
    CU(launch_malon_hessian(g.dmal.as<MalonTab>(), L, x, hess, ngeom, g.stream));
    return PIMDK_OK;
  }
  if (g.pes != PES_CCPOL) {
    if (nd > 4) return fail(PIMDK_EINVAL, "Vdoubleprime of the model surfaces supports ndim*natom <= 4");
    Scope s("hess");
    CU(launch_simple_hessian(g.pes, g.sp, ndim, natom, L, x, hess, ngeom, g.stream));
    return PIMDK_OK;
  }
  // mcmod_waterdimer_ccpol.f90:59-76: central difference (eps = 1e-5) of the finite-difference Vprime.  Every
  // Vprime call perturbs all 18 coordinates in place and leaves its drift behind, so the 36 gradient passes are
  // sequential in x by construction; each pass runs the whole batch through the gradient pipeline.
  const double eps = 1e-5;
  // extent of the coordinate array addressed by L (gradient buffers share its layout)
  const long ext = L.base(ngeom - 1) + (long)(nd - 1) * L.stride_dof + 1;
  CU(g.wHgp.ensure(sizeof(double) * ext));
  CU(g.wHgm.ensure(sizeof(double) * ext));
  double *gp = g.wHgp.as<double>(), *gm = g.wHgm.as<double>();
  for (int i = 0; i < ndim; ++i)
    for (int j = 0; j < natom; ++j) {
      const int d1 = j * ndim + i;
      CU(launch_perturb(L, x, ngeom, d1, eps, g.stream));
      int rc = pes_eval_dev(L, x, nullptr, gp, ngeom, 1);
      if (rc) return rc;
      CU(launch_perturb(L, x, ngeom, d1, -2.0 * eps, g.stream));
      rc = pes_eval_dev(L, x, nullptr, gm, ngeom, 1);
      if (rc) return rc;
      CU(launch_perturb(L, x, ngeom, d1, eps, g.stream));
      CU(launch_hess_column(L, gp, gm, ngeom, nd, d1, eps, hess, g.stream));
    }
  return PIMDK_OK;
}

int pimdk_pes_hessian(pimdk_int nbatch, pimdk_int ndim, pimdk_int natom, double* x, double* hess) {
  NEED_INIT();
  int rc = check_dims(ndim, natom);
  if (rc) return rc;
  if (nbatch <= 0) return PIMDK_OK;
  if (!x || !hess) return fail(PIMDK_EINVAL, "x and hess must not be NULL");
  const long nd = ndim * natom;
  CU(g.wX.ensure(sizeof(double) * nbatch * nd));
  CU(g.wHess.ensure(sizeof(double) * nbatch * nd * nd));
  CU(cudaMemcpyAsync(g.wX.p, x, sizeof(double) * nbatch * nd, cudaMemcpyHostToDevice, g.stream));
  rc = clear_flags();
  if (rc) return rc;
  GeomLayout L{1, nd, 0, 1};
  rc = pes_hessian_dev(L, g.wX.as<double>(), g.wHess.as<double>(), nbatch, (int)ndim, (int)natom);
  if (rc) return rc;
  CU(cudaMemcpyAsync(x, g.wX.p, sizeof(double) * nbatch * nd, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaMemcpyAsync(hess, g.wHess.p, sizeof(double) * nbatch * nd * nd, cudaMemcpyDeviceToHost, g.stream));
  return check_flags(false);
}

// UMhessian into g.wBand (device); x (n,ndim,natom) on the host is updated with the perturbation drift
static int um_hessian_dev(pimdk_int n, pimdk_int ndim, pimdk_int natom, double* x, const double* mass, double betan,
                          pimdk_int singlewell) {
  int rc = check_dims(ndim, natom);
  if (rc) return rc;
  if (n < 1 || !x || !mass || !(betan > 0.0)) return fail(PIMDK_EINVAL, "bad UMhessian arguments");
  const long nd = ndim * natom, nx = n * nd;
  const long ngeom = singlewell ? 1 : n;   // singlewell: Vdoubleprime is called for bead 1 only (instantonmod.f90:183)
  CU(g.wX.ensure(sizeof(double) * nx));
  CU(g.wHess.ensure(sizeof(double) * ngeom * nd * nd));
  CU(g.wBand.ensure(sizeof(double) * (nd + 1) * nx));
  CU(g.wMisc.ensure(sizeof(double) * natom));
  CU(cudaMemcpyAsync(g.wX.p, x, sizeof(double) * nx, cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(g.wMisc.p, mass, sizeof(double) * natom, cudaMemcpyHostToDevice, g.stream));
  rc = clear_flags();
  if (rc) return rc;
  GeomLayout L{n, nd * n, 1, n};   // x(n,ndim,natom): bead index fastest
  rc = pes_hessian_dev(L, g.wX.as<double>(), g.wHess.as<double>(), ngeom, (int)ndim, (int)natom);
  if (rc) return rc;
  {
    Scope s("hess");
    CU(launch_um_band((int)n, (int)ndim, (int)natom, g.wHess.as<double>(), g.wMisc.as<double>(), betan, singlewell != 0,
                      g.wBand.as<double>(), g.stream));
  }
  CU(cudaMemcpyAsync(x, g.wX.p, sizeof(double) * nx, cudaMemcpyDeviceToHost, g.stream));
  return check_flags(false);
}

int pimdk_um_hessian(pimdk_int n, pimdk_int ndim, pimdk_int natom, double* x, const double* mass, double betan,
                     pimdk_int singlewell, double* band) {
  NEED_INIT();
  if (!band) return fail(PIMDK_EINVAL, "band must not be NULL");
  int rc = um_hessian_dev(n, ndim, natom, x, mass, betan, singlewell);
  if (rc) return rc;
  const long nd = ndim * natom;
  CU(cudaMemcpyAsync(band, g.wBand.p, sizeof(double) * (nd + 1) * n * nd, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  return PIMDK_OK;
}

// cuSOLVER (dense symmetric eigensolver) is bound at run time so that libpimdk.so itself only needs the CUDA runtime
namespace {
struct Cusolver {
  void* lib = nullptr;
  void* handle = nullptr;
  int (*create)(void**) = nullptr;
  int (*destroy)(void*) = nullptr;
  int (*set_stream)(void*, cudaStream_t) = nullptr;
  int (*bufsize)(void*, int, int, int, const double*, int, const double*, int*) = nullptr;
  int (*syevd)(void*, int, int, int, double*, int, double*, double*, int, int*) = nullptr;
} cs;
int cusolver_load() {
  if (cs.handle) return PIMDK_OK;
  const char* names[] = {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11",
                         "/usr/local/cuda/lib64/libcusolver.so"};
  for (const char* nm : names) {
    cs.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (cs.lib) break;
  }
  if (!cs.lib) return fail(PIMDK_ECUDA, "detJ needs cuSOLVER (libcusolver.so.11 not found: %s)", dlerror());
  cs.create = (int (*)(void**))dlsym(cs.lib, "cusolverDnCreate");
  cs.destroy = (int (*)(void*))dlsym(cs.lib, "cusolverDnDestroy");
  cs.set_stream = (int (*)(void*, cudaStream_t))dlsym(cs.lib, "cusolverDnSetStream");
  cs.bufsize = (int (*)(void*, int, int, int, const double*, int, const double*, int*))dlsym(cs.lib, "cusolverDnDsyevd_bufferSize");
  cs.syevd = (int (*)(void*, int, int, int, double*, int, double*, double*, int, int*))dlsym(cs.lib, "cusolverDnDsyevd");
  if (!cs.create || !cs.destroy || !cs.set_stream || !cs.bufsize || !cs.syevd) return fail(PIMDK_ECUDA, "cuSOLVER symbols missing");
  if (cs.create(&cs.handle) != 0) {
    cs.handle = nullptr;
    return fail(PIMDK_ECUDA, "cusolverDnCreate failed");
  }
  return PIMDK_OK;
}
}  // namespace

// detJ's work on the device: UMhessian (x comes back with the Hessian's finite-difference drift, like the reference's),
// dense symmetric eigensolver.  Leaves the eigenvalues (ascending) in wEig and, if asked, the eigenvectors (column-major,
// one per column) in wDense; the stream is drained on return.
static int detj_core(pimdk_int n, pimdk_int ndim, pimdk_int natom, double* x, const double* mass, double betan,
                     pimdk_int singlewell, bool vectors) {
  int rc = um_hessian_dev(n, ndim, natom, x, mass, betan, singlewell);   // "Hessian is cooked."
  if (rc) return rc;
  rc = cusolver_load();
  if (rc) return rc;
  const long nd = ndim * natom, N = n * nd;
  if (N > 46000) return fail(PIMDK_EINVAL, "totdof too large for the dense eigensolver");
  CU(g.wDense.ensure(sizeof(double) * N * N));
  CU(g.wEig.ensure(sizeof(double) * N + sizeof(int)));
  {
    Scope s("hess");
    CU(launch_band_to_dense(N, (int)nd, g.wBand.as<double>(), g.wDense.as<double>(), g.stream));
  }
  // DSBEVD(jobz, 'L', totdof, ndof, H, ndof+1, etasquared, ...) (instantonmod.f90:819-823): all eigenvalues, ascending
  const int jobz = vectors ? 1 : 0;   // CUSOLVER_EIG_MODE_VECTOR / NOVECTOR
  const int uplo = 0;                 // CUBLAS_FILL_MODE_LOWER
  int lwork = 0;
  if (cs.set_stream(cs.handle, g.stream) != 0) return fail(PIMDK_ECUDA, "cusolverDnSetStream failed");
  if (cs.bufsize(cs.handle, jobz, uplo, (int)N, g.wDense.as<double>(), (int)N, g.wEig.as<double>(), &lwork) != 0)
    return fail(PIMDK_ECUDA, "cusolverDnDsyevd_bufferSize failed");
  CU(g.wWork.ensure(sizeof(double) * (size_t)lwork));
  int* dinfo = reinterpret_cast<int*>(g.wEig.as<double>() + N);
  const int st = cs.syevd(cs.handle, jobz, uplo, (int)N, g.wDense.as<double>(), (int)N, g.wEig.as<double>(),
                          g.wWork.as<double>(), lwork, dinfo);
  if (st != 0) return fail(PIMDK_ECUDA, "cusolverDnDsyevd failed (status %d)", st);
  int info = 0;
  CU(cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  if (info != 0) return fail(PIMDK_ECUDA, "eigensolver did not converge (info = %d)", info);
  return PIMDK_OK;
}

int pimdk_detj(pimdk_int n, pimdk_int ndim, pimdk_int natom, double* x, const double* mass, double betan,
               pimdk_int singlewell, double* etasquared, double* eigvecs) {
  NEED_INIT();
  if (!etasquared) return fail(PIMDK_EINVAL, "etasquared must not be NULL");
  int rc = detj_core(n, ndim, natom, x, mass, betan, singlewell, eigvecs != nullptr);
  if (rc) return rc;
  const long N = (long)n * ndim * natom;
  CU(cudaMemcpyAsync(etasquared, g.wEig.p, sizeof(double) * N, cudaMemcpyDeviceToHost, g.stream));
  if (eigvecs) CU(cudaMemcpyAsync(eigvecs, g.wDense.p, sizeof(double) * N * N, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  return PIMDK_OK;
}

int pimdk_readhess_displace(pimdk_int n, pimdk_int ndim, pimdk_int natom, double* x, const double* mass, double betan,
                            double beta, uint64_t seed, pimdk_int traj_gid, double* etasquared) {
  NEED_INIT();
  if (!(beta > 0.0)) return fail(PIMDK_EINVAL, "beta must be positive");
  int rc = detj_core(n, ndim, natom, x, mass, betan, 0, true);   // detJ(x, etasquared, .false., interphess, eigvecs)
  if (rc) return rc;
  const long N = (long)n * ndim * natom;
  CU(g.wAux.ensure(sizeof(double) * N));
  {
    Scope s("hess", 2);
    // wX holds x as UMhessian left it; wMisc the masses (um_hessian_dev)
    CU(launch_readhess_displace((int)n, (int)ndim, (int)natom, g.wEig.as<double>(), g.wDense.as<double>(),
                                g.wMisc.as<double>(), std::sqrt(1.0 / beta), seed, (uint32_t)traj_gid, g.wAux.as<double>(),
                                g.wX.as<double>(), g.stream));
  }
  CU(cudaMemcpyAsync(x, g.wX.p, sizeof(double) * N, cudaMemcpyDeviceToHost, g.stream));
  if (etasquared) CU(cudaMemcpyAsync(etasquared, g.wEig.p, sizeof(double) * N, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  return PIMDK_OK;
}

int pimdk_nm_setup(pimdk_int n, pimdk_int ndim, pimdk_int natom, const double* mass, double betan, double tau) {
  NEED_INIT();
  if (n < 2 || ndim < 1 || natom < 1 || !mass || !(betan > 0.0)) return fail(PIMDK_EINVAL, "bad nm_setup arguments");
  g.n = (int)n;
  g.nm_ndim = (int)ndim;
  g.nm_natom = (int)natom;
  g.betan = betan;
  g.tau = tau;
  g.mass.assign(mass, mass + natom);
  g.lam.assign(n, 0.0);
  g.beadmass.assign((size_t)natom * n, 0.0);
  g.T.assign((size_t)n * n, 0.0);
  std::vector<double> sA(n), sB(n), lamb2(n);
  // init_nm, verletmodule.f90:306-338
  for (long i = 1; i <= n; ++i) {
    g.lam[i - 1] = 2.0 * std::sin((double)i * PI_TRUNC / (double)(2 * n + 2)) / betan;
    for (long j = 1; j <= natom; ++j)
      g.beadmass[(size_t)(i - 1) * natom + (j - 1)] = mass[j - 1] * ((g.lam[i - 1] * tau) * (g.lam[i - 1] * tau));
    for (long l = i; l <= n; ++l) {
      const double t = std::sin((double)(i * l) * PI_TRUNC / (double)(n + 1)) * std::sqrt(2.0 / (double)(n + 1));
      if (t != t) return fail(PIMDK_ENAN, "Nan!");
      g.T[(size_t)(l - 1) * n + (i - 1)] = t;
      g.T[(size_t)(i - 1) * n + (l - 1)] = t;
    }
    sA[i - 1] = std::sin((double)i * PI_TRUNC / (double)(n + 1));
    sB[i - 1] = std::sin((double)(n * i) * PI_TRUNC / (double)(n + 1));
    lamb2[i - 1] = (g.lam[i - 1] * betan) * (g.lam[i - 1] * betan);
  }
  CU(g.dT.ensure(g.T.size() * sizeof(double)));
  CU(g.dsA.ensure(n * sizeof(double)));
  CU(g.dsB.ensure(n * sizeof(double)));
  CU(g.dlamb2.ensure(n * sizeof(double)));
  CU(g.dmass.ensure(natom * sizeof(double)));
  CU(cudaMemcpyAsync(g.dT.p, g.T.data(), g.T.size() * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(g.dsA.p, sA.data(), n * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(g.dsB.p, sB.data(), n * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(g.dlamb2.p, lamb2.data(), n * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(g.dmass.p, mass, natom * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  g.nm_ready = true;
  return PIMDK_OK;
}

int pimdk_nm_get(double* T, double* lam, double* beadmass) {
  if (!g.nm_ready) return fail(PIMDK_EINVAL, "pimdk_nm_setup has not been called");
  if (T) std::memcpy(T, g.T.data(), g.T.size() * sizeof(double));
  if (lam) std::memcpy(lam, g.lam.data(), g.lam.size() * sizeof(double));
  if (beadmass) std::memcpy(beadmass, g.beadmass.data(), g.beadmass.size() * sizeof(double));
  return PIMDK_OK;
}

int pimdk_nm_transform(pimdk_int forward, pimdk_int nvec, const double* vin, const double* beadvec, double* vout) {
  NEED_INIT();
  if (!g.nm_ready) return fail(PIMDK_EINVAL, "pimdk_nm_setup has not been called");
  const size_t cnt = (size_t)nvec * g.n;
  std::vector<double> tmp;
  const double* src = vin;
  if (!forward && beadvec) {  // nmtransform_backward: qprop + beadvec before the product (:274-278)
    tmp.resize(cnt);
    for (size_t i = 0; i < cnt; ++i) tmp[i] = vin[i] + beadvec[i];
    src = tmp.data();
  }
  CU(g.wX.ensure(cnt * sizeof(double)));
  CU(g.wG.ensure(cnt * sizeof(double)));
  CU(cudaMemcpyAsync(g.wX.p, src, cnt * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  NmTables nm = nm_tables_base();
  {
    Scope s("gemm");
    CU(launch_nm_gemm(nm, GEMM_PLAIN, g.wX.as<double>(), g.wG.as<double>(), nvec, nullptr, nullptr, g.stream));
  }
  CU(cudaMemcpyAsync(vout, g.wG.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  if (forward && beadvec)  // nmtransform_forward: qprop - beadvec after the product (:260-264)
    for (size_t i = 0; i < cnt; ++i) vout[i] = vout[i] - beadvec[i];
  return PIMDK_OK;
}

static int upload_gid(const pimdk_int* gid, pimdk_int ntraj, const int64_t** dgid) {
  *dgid = nullptr;
  if (!gid) return PIMDK_OK;
  CU(g.wGid.ensure(sizeof(int64_t) * ntraj));
  CU(cudaMemcpyAsync(g.wGid.p, gid, sizeof(int64_t) * ntraj, cudaMemcpyHostToDevice, g.stream));
  *dgid = g.wGid.as<int64_t>();
  return PIMDK_OK;
}

int pimdk_init_path(pimdk_int ntraj, pimdk_int npath, const double* lampath, const double* path,
                    const double* splinepath, const double* xi, uint64_t seed, const pimdk_int* traj_gid, double* x,
                    double* p) {
  NEED_INIT();
  if (!g.nm_ready) return fail(PIMDK_EINVAL, "pimdk_nm_setup has not been called");
  if (ntraj <= 0) return PIMDK_OK;
  if (npath < 2) return fail(PIMDK_EINVAL, "npath must be >= 2");
  const int ndof = g.nm_ndim * g.nm_natom;
  const size_t tot = (size_t)ntraj * ndof * g.n;
  CU(g.wX.ensure(tot * sizeof(double)));
  CU(g.wP.ensure(tot * sizeof(double)));
  CU(g.wPp.ensure(tot * sizeof(double)));
  const size_t np = (size_t)npath, npd = np * ndof;
  CU(g.wMisc.ensure((np + 2 * npd + ntraj) * sizeof(double)));
  double* d = g.wMisc.as<double>();
  double *dl = d, *dpth = d + np, *dspl = d + np + npd, *dxi = d + np + 2 * npd;
  CU(cudaMemcpyAsync(dl, lampath, np * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(dpth, path, npd * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(dspl, splinepath, npd * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(dxi, xi, ntraj * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  const int64_t* dgid;
  int rc = upload_gid(traj_gid, ntraj, &dgid);
  if (rc) return rc;
  NmTables nm = nm_tables_base();
  rc = build_step_tables(&nm, 0.0, 0.0, 0);
  if (rc) return rc;
  {
    Scope s("init", 3);
    long blocks = (long)((tot + 255) / 256);
    if (blocks > 148L * 32) blocks = 148L * 32;
    init_path_kernel<<<(unsigned)blocks, 256, 0, g.stream>>>(g.n, ndof, (int)npath, dl, dpth, dspl, dxi, ntraj,
                                                            g.wX.as<double>());
    CU(cudaGetLastError());
    CU(launch_sample_momenta(nm, g.wP.as<double>(), ntraj, seed, 0, 0, dgid, g.stream));
    CU(launch_nm_gemm(nm, GEMM_PLAIN, g.wP.as<double>(), g.wPp.as<double>(), (long)ntraj * ndof, nullptr, nullptr,
                      g.stream));
  }
  CU(cudaMemcpyAsync(x, g.wX.p, tot * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaMemcpyAsync(p, g.wPp.p, tot * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  return PIMDK_OK;
}

int pimdk_propagate_dev(pimdk_int thermostat, pimdk_int ntraj, double* x, double* p, const double* a, const double* b,
                        const double* dbdl, double dt, double gamma, pimdk_int NMC, pimdk_int imin, pimdk_int Noutput,
                        pimdk_int cayley, uint64_t seed, const pimdk_int* traj_gid, double* dHdr) {
  NEED_INIT();
  if (!g.nm_ready) return fail(PIMDK_EINVAL, "pimdk_nm_setup has not been called");
  if (g.pes == PES_NONE) return fail(PIMDK_EINVAL, "no PES selected (pimdk_pes_select)");
  if (thermostat != PIMDK_THERMOSTAT_ANDERSEN && thermostat != PIMDK_THERMOSTAT_PILE)
    return fail(PIMDK_EINVAL, "Incorrect thermostat option.");
  int rc = check_dims(g.nm_ndim, g.nm_natom);
  if (rc) return rc;
  if (ntraj <= 0) return PIMDK_OK;
  if (NMC < 0 || imin < 0 || NMC - imin <= 0) return fail(PIMDK_EINVAL, "need NMC > imin >= 0");
  const int keep_sum = g.restart == 2;                 // restart = 2: dHdr arrives holding the running sums
  const long step0 = keep_sum ? g.restartnmc : 0;      // steps already done; also offsets the RNG step counter
  CU(g.wDhSum.ensure(sizeof(double) * (g.sum_total > 0 ? g.sum_total : ntraj)));   // (the chunked caller sized it before its first chunk)
  g.sums_n = g.sum_total > 0 ? g.sum_total : ntraj;
  double* dsum = g.wDhSum.as<double>() + (g.sum_total > 0 ? g.sum_off : 0);
  const int n = g.n, ndof = g.nm_ndim * g.nm_natom;
  const long rows = (long)ntraj * ndof;
  const size_t tot = (size_t)rows * n;
  CU(g.wP.ensure(tot * sizeof(double)));
  CU(g.wQ.ensure(tot * sizeof(double)));
  CU(g.wG.ensure(tot * sizeof(double)));
  CU(g.wGn.ensure(tot * sizeof(double)));
  // Andersen collision clocks (count, rkick per trajectory: verletmodule.f90:196-234).  They outlive the call so that a
  // run cut into several calls (restart = 1 writes its files every Noutput steps) can continue them
  // (pimdk_set_andersen_carry) instead of drawing a new interval at every cut, which the reference does not do.
  const long clock_total = g.sum_total > 0 ? g.sum_total : (long)ntraj;
  const long clock_off = g.sum_total > 0 ? g.sum_off : 0;
  bool carry = false;
  if (thermostat == PIMDK_THERMOSTAT_ANDERSEN) {
    carry = g.andersen_carry && g.clock_n == clock_total && g.wCount.cap >= sizeof(int) * clock_total;
    if (g.andersen_carry && !carry)
      return fail(PIMDK_EINVAL, "andersen carry: no collision clocks of a previous call for %ld trajectories", clock_total);
    if (!carry) {
      CU(g.wCount.ensure(sizeof(int) * clock_total));
      CU(g.wKick.ensure(sizeof(int) * clock_total));
    }
  }
  if (!traj_gid && ntraj > 0x100000000LL) return fail(PIMDK_EINVAL, "more than 2^32 trajectories need explicit ids below 2^32");
  // dHdrlimit (verletmodule.f90:404-409): propagate_pimd_pile only
  const bool guard = thermostat == PIMDK_THERMOSTAT_PILE && g.dhdrlimit >= 0.0;
  const double* rp_lam = nullptr; const double* rp_path = nullptr; const double* rp_spl = nullptr; const double* rp_xi = nullptr;
  if (guard) {
    if (g.rp_ntraj != clock_total)
      return fail(PIMDK_EINVAL, "dHdrlimit >= 0: pimdk_set_dhdrlimit was given xi for %ld trajectories, the call has %ld", g.rp_ntraj, clock_total);
    const size_t np_ = (size_t)g.rp_npath, npd = np_ * (size_t)(g.nm_ndim * g.nm_natom);
    rp_lam = g.wPath.as<double>();
    rp_path = rp_lam + np_;
    rp_spl = rp_path + npd;
    rp_xi = g.wXi.as<double>() + clock_off;
    CU(g.wReinit.ensure(sizeof(int) * ntraj));
  }
  NmTables nm = nm_tables_base();
  rc = build_step_tables(&nm, dt, gamma, (int)cayley);
  if (rc) return rc;
  rc = clear_flags();
  if (rc) return rc;
  const int64_t* dgid = reinterpret_cast<const int64_t*>(traj_gid);
  if (dgid) {
    check_gid_kernel<<<(unsigned)((ntraj + 255) / 256), 256, 0, g.stream>>>(dgid, (long)ntraj, g.wFlags.as<int>());
    CU(cudaGetLastError());
  }
  int* const clk_count = thermostat == PIMDK_THERMOSTAT_ANDERSEN ? g.wCount.as<int>() + clock_off : nullptr;
  int* const clk_kick = thermostat == PIMDK_THERMOSTAT_ANDERSEN ? g.wKick.as<int>() + clock_off : nullptr;
  if (thermostat == PIMDK_THERMOSTAT_ANDERSEN && (g.sum_total == 0 || clock_off + ntraj >= clock_total)) g.clock_n = clock_total;
  if (g.fused && fused_small_supported(g.pes, n, g.nm_ndim, g.nm_natom)) {
    {
      Scope s("fused");
      CU(launch_fused_small(nm, g.pes, g.sp, (int)thermostat, ntraj, x, p, a, b, dbdl, dt, NMC, imin, (double)Noutput, seed,
                            dgid, dHdr, g.wFlags.as<int>(), step0, keep_sum, dsum, clk_count, clk_kick, carry ? 1 : 0,
                            guard ? g.dhdrlimit : -1.0, g.rp_npath, rp_lam, rp_path, rp_spl, rp_xi, g.stream));
    }
    rc = check_flags(false);
    g.last_nan_traj = -1;
    if (g.profiling) resolve_spans();
    return rc;
  }
  double *P = g.wP.as<double>(), *Q = g.wQ.as<double>(), *G = g.wG.as<double>(), *Gn = g.wGn.as<double>();
  int* flags = g.wFlags.as<int>();
  GeomLayout L{n, (long)ndof * n, 1, n};
  if (!keep_sum) CU(cudaMemsetAsync(dHdr, 0, sizeof(double) * ntraj, g.stream));  // restart < 2: dHdr = 0 (:200,388)
  // beadvec (init_nm :328-333) depends on a and b only: formed once per call for all rows; the update kernel then writes
  // Q + beadvec next to Q (into G's buffer, which is dead between the gradient transform and the next gradient), and
  // the back-transform is a plain product of that array
  const bool use_bv = nm_uses_beadvec_array(nm);
  double* BV = nullptr;
  double* QB = use_bv ? G : nullptr;
  if (use_bv) {
    CU(g.wBV.ensure(tot * sizeof(double)));
    BV = g.wBV.as<double>();
    Scope s("update");
    CU(launch_beadvec(nm, a, b, rows, BV, g.stream));
  }
  auto back_transform = [&]() -> cudaError_t {   // x = T (Q + beadvec)
    return use_bv ? launch_nm_gemm(nm, GEMM_PLAIN, QB, x, rows, a, b, g.stream) : launch_nm_gemm(nm, GEMM_ADD_BEADVEC, Q, x, rows, a, b, g.stream);
  };
  {
    Scope s("gemm", 2);
    CU(launch_nm_gemm(nm, GEMM_PLAIN, p, P, rows, a, b, g.stream));
    CU(launch_nm_gemm(nm, GEMM_SUB_BEADVEC, x, Q, rows, a, b, g.stream, BV));
  }
  if (thermostat == PIMDK_THERMOSTAT_PILE) {
    for (pimdk_int ii = 1; ii <= NMC; ++ii) {
      rc = pes_eval_dev(L, x, nullptr, G, (long)ntraj * n, 1);  // step_v's Vprime per bead
      if (rc) return rc;
      {
        Scope s("gemm");
        CU(launch_nm_gemm(nm, GEMM_PLAIN, G, Gn, rows, a, b, g.stream));
      }
      {
        Scope s("update");
        CU(launch_nm_update(nm, P, Q, Gn, dt, ntraj, 1, 2, 1, seed, (uint64_t)(ii + step0), dgid, flags, g.stream, BV, QB));
      }
      {
        Scope s("gemm");
        CU(back_transform());
      }
      if (ii > imin && guard) {
        Scope s("estimator", 3);
        int* reinit = g.wReinit.as<int>();
        CU(launch_estimator_limit(nm, x, dbdl, dHdr, ntraj, g.dhdrlimit, reinit, g.stream));
        CU(launch_reinit(nm, g.rp_npath, rp_lam, rp_path, rp_spl, rp_xi, x, P, Q, a, b, ntraj, seed, (uint64_t)(ii + step0), dgid,
                         reinit, g.stream));
      } else if (ii > imin) {
        Scope s("estimator");
        CU(launch_estimator(nm, x, dbdl, dHdr, ntraj, g.stream));
      }
    }
  } else {
    int* count = clk_count;
    int* rkick = clk_kick;
    const bool fuse_andersen = nm_update_fuses_andersen(nm, ntraj);
    if (!carry) {
      Scope s("update");
      CU(launch_andersen_init(ntraj, seed, (uint64_t)step0, (double)Noutput, dgid, count, rkick, g.stream));
    }
    // with a beadvec array the step's last kernel is the estimator of this step AND the first update of the next one
    const bool fuse_eu = fuse_andersen && use_bv && ndof <= 8;   // (a CTA of the fused kernel owns whole trajectories: 8 rows)
    for (pimdk_int ii = 1; ii <= NMC; ++ii) {
      if (fuse_eu && ii > 1) {
        // (done by the previous turn's estimator + update kernel)
      } else if (fuse_andersen) {   // resampling of the fired trajectories inside the first update kernel, clocks advanced by the second
        Scope s("update");
        CU(launch_nm_update(nm, P, Q, Gn, dt, ntraj, 0, 1, 0, seed, (uint64_t)(ii + step0), dgid, flags, g.stream, BV, QB, 1, count,
                            rkick, (double)Noutput));
      } else {
        Scope s("update", 3);
        CU(launch_andersen(nm, P, ntraj, seed, (uint64_t)(ii + step0), (double)Noutput, dgid, count, rkick, g.stream));
        CU(launch_nm_update(nm, P, Q, Gn, dt, ntraj, 0, 1, 0, seed, (uint64_t)(ii + step0), dgid, flags, g.stream, BV, QB));
      }
      // model surfaces: the bead gradient is formed in the back-transform's epilogue (positions never stored).  Q + beadvec
      // lives in G's buffer, so the gradient goes to x's, which nothing else reads inside this loop.
      const double* grad_src = G;
      if (use_bv && nm_gemm_fuses_model_pes(nm, rows, (int)g.pes)) {
        Scope s("gemm");
        CU(launch_nm_gemm_model_pes(nm, QB, rows, (int)g.pes, g.sp, x, flags, g.stream));
        grad_src = x;
      } else {
        {
          Scope s("gemm");
          CU(back_transform());
        }
        rc = pes_eval_dev(L, x, nullptr, G, (long)ntraj * n, 1);
        if (rc) return rc;
      }
      if (fuse_andersen && nm_gemm_fuses_kick_rotate(nm, rows)) {   // kick + rotation (+ clocks) in the transform's epilogue
        Scope s("gemm");
        CU(launch_nm_gemm_kick_rotate(nm, grad_src, rows, P, Q, dt, 1, seed, (uint64_t)(ii + step0), dgid, flags, count, rkick,
                                      (double)Noutput, g.stream));
      } else {
        {
          Scope s("gemm");
          CU(launch_nm_gemm(nm, GEMM_PLAIN, grad_src, Gn, rows, a, b, g.stream));
        }
        Scope s("update");
        CU(launch_nm_update(nm, P, Q, Gn, dt, ntraj, 1, 1, 0, seed, (uint64_t)(ii + step0), dgid, flags, g.stream, nullptr, nullptr,
                            fuse_andersen ? 2 : 0, count, rkick, (double)Noutput));
      }
      if (fuse_eu) {
        Scope s("estimator");
        CU(launch_estimator_update(nm, P, Q, BV, QB, dbdl, dHdr, ntraj, ii > imin, ii < NMC, seed, (uint64_t)(ii + 1 + step0), dgid, flags,
                                   count, rkick, g.stream));
      } else if (ii > imin && ndof <= 32) {   // the estimator reads the last bead only: one contraction per (trajectory, dof), not a transform
        Scope s("estimator");
        CU(launch_estimator_modes(nm, Q, a, b, dbdl, dHdr, ntraj, g.stream, BV));
      } else if (ii > imin) {          // many-site surfaces (water-methane, 51 dof): the full back-transform, then the plain estimator
        Scope s("estimator", use_bv ? 3 : 2);
        if (use_bv) CU(launch_add(Q, BV, QB, (long)tot, g.stream));
        CU(back_transform());
        CU(launch_estimator(nm, x, dbdl, dHdr, ntraj, g.stream));
      }
    }
    {
      Scope s("gemm", use_bv ? 2 : 1);   // positions at the end of the call
      if (use_bv) CU(launch_add(Q, BV, QB, (long)tot, g.stream));
      CU(back_transform());
    }
  }
  {
    Scope s("gemm");
    CU(launch_nm_gemm(nm, GEMM_PLAIN, P, p, rows, a, b, g.stream));
  }
  {
    Scope s("estimator");
    // the running sum is what write_restart stores (:171, 246, 412); then dHdr/dble(NMC+restartnmc-imin) (:247,413)
    CU(cudaMemcpyAsync(dsum, dHdr, sizeof(double) * ntraj, cudaMemcpyDeviceToDevice, g.stream));
    CU(launch_scale(dHdr, (double)(NMC + step0 - imin), ntraj, g.stream));
  }
  rc = check_flags(false);
  if (rc == PIMDK_ENAN) {
    // everything on the library stream (a caller's non-blocking stream is not ordered against the legacy default stream)
    long long first = (long long)ntraj;
    long long* dfirst = reinterpret_cast<long long*>(g.wFlags.as<char>() + 8);
    bool ok = cudaMemcpyAsync(dfirst, &first, sizeof(long long), cudaMemcpyHostToDevice, g.stream) == cudaSuccess;
    if (ok) {
      first_nan_kernel<<<148, 256, 0, g.stream>>>(P, (long)ndof * n, ntraj, dfirst);
      ok = cudaGetLastError() == cudaSuccess &&
           cudaMemcpyAsync(&first, dfirst, sizeof(long long), cudaMemcpyDeviceToHost, g.stream) == cudaSuccess &&
           cudaStreamSynchronize(g.stream) == cudaSuccess;
    }
    g.last_nan_traj = (ok && first < ntraj) ? (long)first : -1;
  } else {
    g.last_nan_traj = -1;
  }
  if (g.profiling) resolve_spans();
  return rc;
}

// Host-buffer form.  Large batches are cut into chunks of whole trajectories (independent units; results depend on the
// global trajectory id only) and pipelined: while chunk c is propagated on the compute stream, chunk c+1 is copied in and
// chunk c-1 copied out on a second stream, through two device buffers.  Only the first copy-in and the last copy-out
// are exposed.
static int propagate_chunked(pimdk_int thermostat, pimdk_int ntraj, long chunk, double* x, double* p, const double* a,
                             const double* b, const double* dbdl, double dt, double gamma, pimdk_int NMC, pimdk_int imin,
                             pimdk_int Noutput, pimdk_int cayley, uint64_t seed, const pimdk_int* traj_gid, double* dHdr) {
  const int ndof = g.nm_ndim * g.nm_natom;
  const size_t per = (size_t)ndof * g.n;                       // doubles of x (or p) per trajectory
  const size_t nb = (size_t)ntraj * ndof;
  if (!g.copy_stream) CU(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
  for (cudaEvent_t& e : g.ev_in)
    if (!e) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  DevBuf* bx[2] = {&g.wX, &g.wX2};
  DevBuf* bp[2] = {&g.wPp, &g.wPp2};
  for (int i = 0; i < 2; ++i) {
    CU(bx[i]->ensure((size_t)chunk * per * sizeof(double)));
    CU(bp[i]->ensure((size_t)chunk * per * sizeof(double)));
  }
  CU(g.wA.ensure(ndof * sizeof(double)));
  CU(g.wB.ensure(nb * sizeof(double)));
  CU(g.wDbdl.ensure(nb * sizeof(double)));
  CU(g.wDhdr.ensure(ntraj * sizeof(double)));
  CU(g.wDhSum.ensure(sizeof(double) * ntraj));
  CU(cudaMemcpyAsync(g.wA.p, a, ndof * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(g.wB.p, b, nb * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(g.wDbdl.p, dbdl, nb * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  if (g.restart == 2) CU(cudaMemcpyAsync(g.wDhdr.p, dHdr, ntraj * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  const int64_t* dgid;
  int rc = upload_gid(traj_gid, ntraj, &dgid);
  if (rc) return rc;
  const long nchunk = (ntraj + chunk - 1) / chunk;
  auto copy_in = [&](long c) -> cudaError_t {
    const long t0 = c * chunk, nt = (ntraj - t0 < chunk) ? ntraj - t0 : chunk;
    cudaError_t e = cudaMemcpyAsync(bx[c & 1]->p, x + (size_t)t0 * per, (size_t)nt * per * sizeof(double), cudaMemcpyHostToDevice, g.copy_stream);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(bp[c & 1]->p, p + (size_t)t0 * per, (size_t)nt * per * sizeof(double), cudaMemcpyHostToDevice, g.copy_stream);
    if (e != cudaSuccess) return e;
    return cudaEventRecord(g.ev_in[c & 1], g.copy_stream);
  };
  int result = PIMDK_OK;
  long first_nan = -1;
  std::string keep;
  CU(copy_in(0));
  g.sum_total = ntraj;
  for (long c = 0; c < nchunk; ++c) {
    const long t0 = c * chunk, nt = (ntraj - t0 < chunk) ? ntraj - t0 : chunk;
    if (c + 1 < nchunk) {
      cudaError_t e = copy_in(c + 1);   // behind the copy-out of chunk c-1 in stream order: its buffer is free by then
      if (e != cudaSuccess) { g.sum_total = 0; return fail(PIMDK_ECUDA, "%s", cudaGetErrorString(e)); }
    }
    cudaStreamWaitEvent(g.stream, g.ev_in[c & 1], 0);
    g.sum_off = t0;
    rc = pimdk_propagate_dev(thermostat, nt, bx[c & 1]->as<double>(), bp[c & 1]->as<double>(), g.wA.as<double>(),
                             g.wB.as<double>() + (size_t)t0 * ndof, g.wDbdl.as<double>() + (size_t)t0 * ndof, dt, gamma, NMC,
                             imin, Noutput, cayley, seed, dgid ? reinterpret_cast<const pimdk_int*>(dgid + t0) : nullptr,
                             g.wDhdr.as<double>() + t0);   // returns with the compute stream drained
    if (rc != PIMDK_OK && rc != PIMDK_ENAN) { g.sum_total = 0; return rc; }
    if (rc == PIMDK_ENAN && result == PIMDK_OK) {
      result = rc;
      keep = g.err;
      first_nan = g.last_nan_traj >= 0 ? t0 + g.last_nan_traj : -1;
    }
    cudaError_t e = cudaMemcpyAsync(x + (size_t)t0 * per, bx[c & 1]->p, (size_t)nt * per * sizeof(double), cudaMemcpyDeviceToHost, g.copy_stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(p + (size_t)t0 * per, bp[c & 1]->p, (size_t)nt * per * sizeof(double), cudaMemcpyDeviceToHost, g.copy_stream);
    if (e != cudaSuccess) { g.sum_total = 0; return fail(PIMDK_ECUDA, "%s", cudaGetErrorString(e)); }
  }
  g.sum_total = 0;
  g.sum_off = 0;
  CU(cudaMemcpyAsync(dHdr, g.wDhdr.p, ntraj * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.copy_stream));
  CU(cudaStreamSynchronize(g.stream));
  if (result == PIMDK_ENAN) {
    g.err = keep;
    g.last_nan_traj = first_nan;
  }
  return result;
}

int pimdk_propagate(pimdk_int thermostat, pimdk_int ntraj, double* x, double* p, const double* a, const double* b,
                    const double* dbdl, double dt, double gamma, pimdk_int NMC, pimdk_int imin, pimdk_int Noutput,
                    pimdk_int cayley, uint64_t seed, const pimdk_int* traj_gid, double* dHdr) {
  NEED_INIT();
  if (!g.nm_ready) return fail(PIMDK_EINVAL, "pimdk_nm_setup has not been called");
  if (ntraj <= 0) return PIMDK_OK;
  const int ndof = g.nm_ndim * g.nm_natom;
  const size_t tot = (size_t)ntraj * ndof * g.n, nb = (size_t)ntraj * ndof;
  {
    // automatic: about eight chunks per call, each at least 256 trajectories and at most 128 MB of state (x and p); worth it
    // from three chunks on.  Only the first copy-in and the last copy-out are exposed, so a call of 1024 trajectories (C4 on 8
    // GPUs) runs as four chunks rather than unpipelined, and 8192 trajectories keep chunks of 1024 (smaller ones cost more in
    // per-chunk launches than they hide: 6.13 against 6.16 M bead-steps/s end to end at 256).
    long chunk = g.chunk_traj;
    if (chunk <= 0) {
      const size_t per_traj = 2 * (size_t)ndof * g.n * sizeof(double);
      long cmax = (long)(((size_t)128 << 20) / per_traj) + 1;
      if (cmax < 256) cmax = 256;
      cmax = (cmax + 63) / 64 * 64;
      // (the many-site surfaces only: 256 trajectories of a model surface would not fill the machine, those keep 128 MB chunks)
      const bool heavy = g.pes == PES_CCPOL || g.pes == PES_WATMETH || g.pes == PES_MALON;
      chunk = heavy ? ((long)ntraj / 8 + 63) / 64 * 64 : cmax;
      if (chunk < 256) chunk = 256;
      if (chunk > cmax) chunk = cmax;
    }
    if ((long)ntraj >= 3 * chunk)
      return propagate_chunked(thermostat, ntraj, chunk, x, p, a, b, dbdl, dt, gamma, NMC, imin, Noutput, cayley, seed,
                               traj_gid, dHdr);
  }
  CU(g.wX.ensure(tot * sizeof(double)));
  CU(g.wPp.ensure(tot * sizeof(double)));
  CU(g.wA.ensure(ndof * sizeof(double)));
  CU(g.wB.ensure(nb * sizeof(double)));
  CU(g.wDbdl.ensure(nb * sizeof(double)));
  CU(g.wDhdr.ensure(ntraj * sizeof(double)));
  CU(cudaMemcpyAsync(g.wX.p, x, tot * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(g.wPp.p, p, tot * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(g.wA.p, a, ndof * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(g.wB.p, b, nb * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(g.wDbdl.p, dbdl, nb * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  if (g.restart == 2) CU(cudaMemcpyAsync(g.wDhdr.p, dHdr, ntraj * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  const int64_t* dgid;
  int rc = upload_gid(traj_gid, ntraj, &dgid);
  if (rc) return rc;
  rc = pimdk_propagate_dev(thermostat, ntraj, g.wX.as<double>(), g.wPp.as<double>(), g.wA.as<double>(),
                           g.wB.as<double>(), g.wDbdl.as<double>(), dt, gamma, NMC, imin, Noutput, cayley, seed,
                           reinterpret_cast<const pimdk_int*>(dgid), g.wDhdr.as<double>());
  if (rc != PIMDK_OK && rc != PIMDK_ENAN) return rc;
  std::string keep = g.err;
  CU(cudaMemcpyAsync(x, g.wX.p, tot * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaMemcpyAsync(p, g.wPp.p, tot * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaMemcpyAsync(dHdr, g.wDhdr.p, ntraj * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  g.err = keep;
  return rc;
}

pimdk_int pimdk_last_nan_trajectory(void) { return g.last_nan_traj; }

int pimdk_set_propagate_chunk(pimdk_int ntraj_per_chunk) {
  if (ntraj_per_chunk < 0) return fail(PIMDK_EINVAL, "chunk size must be >= 0 (0 = automatic)");
  g.chunk_traj = (long)ntraj_per_chunk;
  return PIMDK_OK;
}

int pimdk_set_dhdrlimit(double limit, pimdk_int npath, const double* lampath, const double* path, const double* splinepath,
                        pimdk_int ntraj, const double* xi) {
  NEED_INIT();
  g.dhdrlimit = limit;
  if (!(limit >= 0.0)) {
    g.dhdrlimit = -1.0;
    return PIMDK_OK;
  }
  if (!g.nm_ready) return fail(PIMDK_EINVAL, "pimdk_nm_setup has not been called");
  if (npath < 2 || ntraj < 1 || !lampath || !path || !splinepath || !xi) return fail(PIMDK_EINVAL, "dHdrlimit >= 0 needs the spline path and xi of every trajectory");
  const size_t np_ = (size_t)npath, npd = np_ * (size_t)(g.nm_ndim * g.nm_natom);
  CU(g.wPath.ensure((np_ + 2 * npd) * sizeof(double)));
  CU(g.wXi.ensure((size_t)ntraj * sizeof(double)));
  double* d = g.wPath.as<double>();
  CU(cudaMemcpyAsync(d, lampath, np_ * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(d + np_, path, npd * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(d + np_ + npd, splinepath, npd * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaMemcpyAsync(g.wXi.p, xi, (size_t)ntraj * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  g.rp_npath = (int)npath;
  g.rp_ntraj = (long)ntraj;
  return PIMDK_OK;
}

int pimdk_set_andersen_carry(pimdk_int enable) {
  g.andersen_carry = enable != 0;
  return PIMDK_OK;
}

int pimdk_set_restart(pimdk_int restart, pimdk_int restartnmc) {
  if (restart < 0 || restart > 2 || restartnmc < 0) return fail(PIMDK_EINVAL, "restart must be 0, 1 or 2 and restartnmc >= 0");
  g.restart = (long)restart;
  g.restartnmc = (long)restartnmc;
  return PIMDK_OK;
}

int pimdk_get_dhdr_sums(pimdk_int ntraj, double* sums) {
  NEED_INIT();
  if (ntraj <= 0 || ntraj > g.sums_n || !sums) return fail(PIMDK_EINVAL, "no running sums for that many trajectories");
  CU(cudaMemcpyAsync(sums, g.wDhSum.p, sizeof(double) * ntraj, cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  return PIMDK_OK;
}

int pimdk_ti_partial_sums(pimdk_int ntraj, const double* dHdr, const pimdk_int* gid, pimdk_int nrep,
                          pimdk_int nintegral, double betan, double* sums) {
  if (nrep <= 0 || nintegral <= 0) return fail(PIMDK_EINVAL, "nrep and nintegral must be positive");
  for (pimdk_int i = 0; i < 3 * nintegral; ++i) sums[i] = 0.0;
  for (pimdk_int t = 0; t < ntraj; ++t) {
    const pimdk_int id = gid ? gid[t] : t;
    const pimdk_int il = id / nrep;  // global id = nrep*(ilambda-1) + irep (pimd_par.f90:249-253)
    if (il < 0 || il >= nintegral) return fail(PIMDK_EINVAL, "trajectory id %lld outside nintegral*nrep", (long long)id);
    const double I = dHdr[t] / (betan * betan);  // integrand(ii)=dHdr/(betan**2) (pimd_par.f90:379)
    sums[3 * il + 0] += I;
    sums[3 * il + 1] += I * I;
    sums[3 * il + 2] += 1.0;
  }
  return PIMDK_OK;
}

int pimdk_ti_finish(pimdk_int nintegral, const double* sums, const double* weights, double betan, double* mean,
                    double* var, double* deltaA, double* sigmaA, double* qq0) {
  double answer = 0.0, sA = 0.0;
  for (pimdk_int i = 0; i < nintegral; ++i) {  // pimd_par.f90:401-418
    const double cnt = sums[3 * i + 2];
    if (!(cnt > 0.0)) return fail(PIMDK_EINVAL, "lambda point %lld has no trajectories", (long long)i);
    const double m = sums[3 * i] / cnt;
    double s = sums[3 * i + 1] / cnt;
    s = s - m * m;
    if (mean) mean[i] = m;
    if (var) var[i] = s;
    if (m == m) answer = answer + weights[i] * m;
    sA = sA + s * weights[i] * weights[i];
  }
  if (deltaA) *deltaA = answer;
  if (sigmaA) *sigmaA = std::sqrt(sA);
  if (qq0) *qq0 = std::exp(-answer * betan);
  return PIMDK_OK;
}

// ---- multi-GPU: the single collective of the path ------------------------------------------------------------------
int pimdk_comm_unique_id(void* id) {
  if (!id) return fail(PIMDK_EINVAL, "id must point to PIMDK_UNIQUE_ID_BYTES bytes");
  int rc = nccl_load();
  if (rc) return rc;
  static_assert(sizeof(NcclId) == PIMDK_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
  NcclId u;
  NCCLCHK(nc.get_unique_id(&u));
  std::memcpy(id, &u, sizeof u);
  return PIMDK_OK;
}

int pimdk_comm_init(pimdk_int rank, pimdk_int nranks, const void* id) {
  NEED_INIT();
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(PIMDK_EINVAL, "need 0 <= rank < nranks");
  if (nc.comm) return fail(PIMDK_EINVAL, "communicator already initialised (pimdk_comm_finalize first)");
  if (nranks == 1) {   // a one-rank job needs no NCCL
    nc.rank = 0;
    nc.nranks = 1;
    return PIMDK_OK;
  }
  if (!id) return fail(PIMDK_EINVAL, "id must not be NULL for nranks > 1");
  int rc = nccl_load();
  if (rc) return rc;
  NcclId u;
  std::memcpy(&u, id, sizeof u);
  CU(cudaSetDevice(g.device));
  NCCLCHK(nc.comm_init_rank(&nc.comm, (int)nranks, u, (int)rank));
  nc.rank = (int)rank;
  nc.nranks = (int)nranks;
  return PIMDK_OK;
}

int pimdk_comm_finalize(void) {
  if (nc.comm) {
    cudaStreamSynchronize(g.stream);
    nc.comm_destroy(nc.comm);
  }
  nc.comm = nullptr;
  nc.rank = 0;
  nc.nranks = 1;
  return PIMDK_OK;
}

int pimdk_comm_info(pimdk_int* rank, pimdk_int* nranks, pimdk_int* nccl_version) {
  if (rank) *rank = nc.rank;
  if (nranks) *nranks = nc.nranks;
  if (nccl_version) {
    int v = 0;
    if (nc.lib && nc.get_version) nc.get_version(&v);
    *nccl_version = v;
  }
  return PIMDK_OK;
}

// sum over ranks of `count` doubles that live on the device, in place, on the library stream
static int allreduce_dev(double* d, size_t count) {
  if (nc.nranks == 1) return PIMDK_OK;
  if (!nc.comm) return fail(PIMDK_EINVAL, "pimdk_comm_init has not been called");
  NCCLCHK(nc.all_reduce(d, d, count, /*ncclDouble*/ 8, /*ncclSum*/ 0, nc.comm, g.stream));
  g.launch_count += 1;
  return PIMDK_OK;
}

int pimdk_ti_allreduce(pimdk_int nintegral, double* sums) {
  NEED_INIT();
  if (nintegral < 1 || !sums) return fail(PIMDK_EINVAL, "bad pimdk_ti_allreduce arguments");
  if (nc.nranks == 1) return PIMDK_OK;
  const size_t cnt = 3 * (size_t)nintegral;
  CU(g.wSums.ensure(cnt * sizeof(double)));
  CU(cudaMemcpyAsync(g.wSums.p, sums, cnt * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  int rc = allreduce_dev(g.wSums.as<double>(), cnt);
  if (rc) return rc;
  CU(cudaMemcpyAsync(sums, g.wSums.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  return PIMDK_OK;
}

int pimdk_ti_reduce_dev(pimdk_int ntraj, const double* dHdr, const pimdk_int* traj_gid, pimdk_int nrep, pimdk_int nintegral,
                        double betan, double* sums) {
  NEED_INIT();
  if (nrep <= 0 || nintegral <= 0 || ntraj < 0 || !sums || (ntraj > 0 && !dHdr) || !(betan > 0.0))
    return fail(PIMDK_EINVAL, "bad pimdk_ti_reduce_dev arguments");
  const size_t cnt = 3 * (size_t)nintegral;
  CU(g.wSums.ensure(cnt * sizeof(double)));
  int rc = clear_flags();
  if (rc) return rc;
  {
    Scope s("estimator");
    ti_partial_sums_kernel<<<(unsigned)nintegral, kTiThreads, 0, g.stream>>>(dHdr, reinterpret_cast<const int64_t*>(traj_gid), (long)ntraj,
                                                                            (long)nrep, (long)nintegral, betan, g.wSums.as<double>(),
                                                                            g.wFlags.as<int>());
    CU(cudaGetLastError());
  }
  rc = allreduce_dev(g.wSums.as<double>(), cnt);
  if (rc) return rc;
  CU(cudaMemcpyAsync(sums, g.wSums.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  int fl = 0;
  CU(cudaMemcpyAsync(&fl, g.wFlags.p, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  CU(cudaStreamSynchronize(g.stream));
  if (fl & PIMDK_FLAG_BADGID) return fail(PIMDK_EINVAL, "trajectory id outside nintegral*nrep");
  return PIMDK_OK;
}

int pimdk_gauleg(double x1, double x2, pimdk_int nintegral, double* x, double* w) {
  if (nintegral < 1) return fail(PIMDK_EINVAL, "nintegral must be >= 1");
  const double EPS = 3.e-14;
  const pimdk_int m = (nintegral + 1) / 2;
  const double xm = 0.5 * (x2 + x1), xl = 0.5 * (x2 - x1);
  for (pimdk_int i = 1; i <= m; ++i) {
    double z = std::cos(PI_TRUNC * (i - 0.25) / (nintegral + 0.5));
    double z1, pp;
    do {
      double p1 = 1.0, p2 = 0.0;
      for (pimdk_int j = 1; j <= nintegral; ++j) {
        const double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
      }
      pp = nintegral * (z * p1 - p2) / (z * z - 1.0);
      z1 = z;
      z = z1 - p1 / pp;
    } while (std::fabs(z - z1) > EPS);
    x[i - 1] = xm - xl * z;
    x[nintegral - i] = xm + xl * z;
    w[i - 1] = 2.0 * xl / ((1.0 - z * z) * pp * pp);
    w[nintegral - i] = w[i - 1];
  }
  return PIMDK_OK;
}

int pimdk_profile(pimdk_int enable) {
  g.profiling = enable != 0;
  return PIMDK_OK;
}
int pimdk_profile_reset(void) {
  resolve_spans();
  g.prof.clear();
  g.launch_count = 0;
  return PIMDK_OK;
}
int pimdk_profile_get(const char* family, double* ms, pimdk_int* launches) {
  resolve_spans();
  auto it = g.prof.find(family ? family : "");
  if (ms) *ms = it == g.prof.end() ? 0.0 : it->second.ms;
  if (launches) *launches = it == g.prof.end() ? 0 : it->second.launches;
  return PIMDK_OK;
}
int pimdk_fp64_peak(double* tflops) {
  NEED_INIT();
  CU(fp64_peak_probe(g.num_sms, tflops, g.stream));
  return PIMDK_OK;
}
pimdk_int pimdk_launch_count(void) { return g.launch_count; }
int pimdk_selftest_fastmath(pimdk_int* mismatches) {
  NEED_INIT();
  unsigned long long bad = 0;
  CU(fastmath_selftest(&bad, g.stream));
  *mismatches = (pimdk_int)bad;
  return PIMDK_OK;
}
int pimdk_selftest_math(pimdk_int kind, pimdk_int n, const double* x, double* y) {
  NEED_INIT();
  if (kind < 0 || kind > 9 || n <= 0 || !x || !y) return fail(PIMDK_EINVAL, "bad selftest_math arguments");
  CU(math_eval((int)kind, (long)n, x, y, g.stream));
  return PIMDK_OK;
}
int pimdk_selftest_division(pimdk_int* mismatches) {
  NEED_INIT();
  unsigned long long bad = 0;
  CU(div_selftest(&bad, g.stream));
  *mismatches = (pimdk_int)bad;
  return PIMDK_OK;
}

}  // extern "C"
