// Issue-slot microbenchmark for sm_100a: does a non-FP64 instruction issued next to a stream of FP64 instructions cost
// time?  Each warp runs F independent DFMA chains and, per DFMA, M independent instructions of another kind (integer add,
// IMAD, LDS, FSEL/selp, FP32 FFMA).  Model A (pipe-bound): cycles/iteration/SMSP = max(2 F W, (F + F M) W); model B (an FP64
// instruction holds the issue port for 2 cycles): (2 F + F M) W, W = warps per sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND, int M>
__global__ void __launch_bounds__(512) k(double* out, int iters, double a, double b, int ia, long long* cyc) {
  __shared__ double sh[1024];
  sh[threadIdx.x] = a; sh[threadIdx.x + 512] = b;
  __syncthreads();
  double x0 = a + threadIdx.x, x1 = a + 1, x2 = a + 2, x3 = a + 3;
  const int tv = ia + threadIdx.x;
  int i0 = tv, i1 = tv + 1, i2 = tv + 2, i3 = tv + 3;
  float f0 = tv, f1 = tv + 1, f2 = tv + 2, f3 = tv + 3;
  const float fb = (float)b + threadIdx.x;
  double l0 = 0, l1 = 0, l2 = 0, l3 = 0;
  unsigned saddr = (unsigned)__cvta_generic_to_shared(&sh[threadIdx.x]);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#define OTHER(iv, fv, lv)                                                                              \
  _Pragma("unroll") for (int m = 0; m < M; ++m) {                                                      \
    if (KIND == 0) asm volatile("add.s32 %0, %0, %1;" : "+r"(iv) : "r"(tv));                           \
    if (KIND == 1) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(iv) : "r"(tv));                    \
    if (KIND == 2) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(lv) : "r"(saddr + 8 * (u & 3)));      \
    if (KIND == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(fv) : "f"(fb));                      \
    if (KIND == 4) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(fv) : "f"(fb));              \
    if (KIND == 5) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(iv) : "r"(tv));                \
  }
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x0) : "d"(b), "d"(a));
      OTHER(i0, f0, l0)
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x1) : "d"(b), "d"(a));
      OTHER(i1, f1, l1)
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x2) : "d"(b), "d"(a));
      OTHER(i2, f2, l2)
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x3) : "d"(b), "d"(a));
      OTHER(i3, f3, l3)
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + i0 + i1 + i2 + i3 + f0 + f1 + f2 + f3 + l0 + l1 + l2 + l3;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int KIND, int M>
void run(double* out, long long* cyc, int threads) {
  const int iters = 4000;
  k<KIND, M><<<148, threads>>>(out, 10, 1.0, 1.0000001, 3, cyc);
  cudaDeviceSynchronize();
  k<KIND, M><<<148, threads>>>(out, iters, 1.0, 1.0000001, 3, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const int W = threads / 128;  // warps per sub-partition
  const double per = (double)c / (iters * 8.0);   // cycles per group of 4 DFMA + 4 M others, per warp's turn
  static const char* names[] = {"IADD", "IMAD", "LDS.64", "FMNMX", "FFMA", "LOP3"};
  printf("%-10s M=%d warps/SMSP %d: %.2f cycles per (4 DFMA + %d other) x %d warps;  model A %.0f  model B %.0f\n", names[KIND], M, W, per,
         4 * M, W, (double)W * (8 > 4 + 4 * M ? 8 : 4 + 4 * M), (double)W * (8 + 4 * M));
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 8);
  for (int threads : {512, 256}) {
    run<0, 0>(out, cyc, threads); run<0, 1>(out, cyc, threads); run<0, 2>(out, cyc, threads); run<0, 3>(out, cyc, threads);
    run<1, 1>(out, cyc, threads); run<1, 2>(out, cyc, threads);
    run<2, 1>(out, cyc, threads); run<2, 2>(out, cyc, threads);
    run<3, 1>(out, cyc, threads); run<3, 2>(out, cyc, threads);
    run<4, 1>(out, cyc, threads); run<4, 2>(out, cyc, threads);
    run<5, 1>(out, cyc, threads); run<5, 2>(out, cyc, threads);
  }
  return 0;
}
