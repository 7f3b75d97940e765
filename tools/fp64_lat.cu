// FP64 pipe microbenchmark for sm_100a: dependent-issue latency of DADD/DMUL/DFMA and throughput as a
// function of warps per SM and independent chains per warp.  nvcc -arch=sm_100a -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void k(double* out, int iters, double a, double b, long long* cyc) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (OP == 0) x[i] = __dadd_rn(x[i], b);
        if (OP == 1) x[i] = __dmul_rn(x[i], b);
        if (OP == 2) x[i] = __fma_rn(x[i], b, a);
      }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int OP>
void run(int warps_per_sm, double* out, long long* cyc) {
  const int iters = 2000;
  int threads = 32 * (warps_per_sm < 32 ? warps_per_sm : 32);
  int blocks = 148 * ((warps_per_sm + 31) / 32);
  k<ILP, OP><<<blocks, threads>>>(out, 10, 1.0, 1.0000001, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<ILP, OP><<<blocks, threads>>>(out, iters, 1.0, 1.0000001, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double ninst = (double)iters * 8 * ILP;
  double rate = ninst * blocks * threads / (ms * 1e-3) / 1e12;
  printf("op %d ilp %d warps/SM %2d: %.2f cycles per dependent op, %.2f Tinst/s (x2 flop for FMA)\n", OP, ILP, warps_per_sm,
         (double)c / (iters * 8), rate);
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 2 * 1024 * 8); cudaMalloc(&cyc, 8);
  for (int w : {1, 4, 8, 16, 32, 64}) { run<1, 0>(w, out, cyc); }
  for (int w : {1, 4, 8, 16, 32, 64}) { run<1, 2>(w, out, cyc); }
  for (int w : {4, 8, 16, 32}) { run<2, 2>(w, out, cyc); }
  for (int w : {4, 8, 16, 32}) { run<4, 2>(w, out, cyc); }
  for (int w : {4, 8, 16}) { run<8, 1>(w, out, cyc); }
  return 0;
}
