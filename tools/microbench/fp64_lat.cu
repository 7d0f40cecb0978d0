// Microbenchmark: dependent-chain latency and per-SM throughput of FP64 ops on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __device__ __forceinline__ double step(double x, double c) {
  if (OP == 0) return x + c;            // DADD
  if (OP == 1) return x * c;            // DMUL
  if (OP == 2) return fma(x, c, c);     // DFMA
  if (OP == 3) return x / c;            // division
  if (OP == 4) return x > c ? x : c;    // DSETP + select
  return x;
}
template <int OP> __global__ void lat(double* out, double c, long long* clk, int n) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) x = step<OP>(x, c);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
// throughput: 4 independent chains per thread, many warps
template <int OP> __global__ void thr(double* out, double c, int n) {
  double x0 = out[threadIdx.x], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
#pragma unroll 4
  for (int i = 0; i < n; ++i) { x0 = step<OP>(x0, c); x1 = step<OP>(x1, c); x2 = step<OP>(x2, c); x3 = step<OP>(x3, c); }
  out[blockIdx.x*blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
int main() {
  double* out; long long* clk; cudaMalloc(&out, 1 << 24); cudaMalloc(&clk, 8);
  cudaMemset(out, 0, 1 << 24);
  const char* names[] = {"DADD", "DMUL", "DFMA", "DDIV", "DSETP+SEL"};
  const int n = 4096;
  long long h;
#define RUN(OP) lat<OP><<<1, 32>>>(out, 1.0000001, clk, n); lat<OP><<<1, 32>>>(out, 1.0000001, clk, n); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("%-10s dependent latency %.1f clk\n", names[OP], (double) h/n);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
#define THR(OP) { thr<OP><<<148*8, 256>>>(out, 1.0000001, 2048); cudaEventRecord(e0); thr<OP><<<148*8, 256>>>(out, 1.0000001, 2048); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); double ops = 148.0*8*256*2048*4; printf("%-10s throughput %.1f Gop/s = %.1f lanes/clk/SM @1.965GHz\n", names[OP], ops/ms/1e6, ops/(ms*1e-3)/148/1.965e9); }
  THR(0) THR(1) THR(2) THR(3) THR(4)
  return 0;
}
