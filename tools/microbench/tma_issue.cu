// How long does one thread take to ISSUE cp.async.bulk copies (global -> shared, ~4.7 KB
// each, the ring kernel's row size), and does spreading the issue over warps help?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_issue tma_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned s32 (const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
__global__ void k (const double* src, long long stride, int nops, int nwarps_issue, int bytes,
                   unsigned long long* out) {
  extern __shared__ __align__(16) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
  double* buf = reinterpret_cast<double*>(sm + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nslot = 16;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(bar)), "r"(nops*bytes) : "memory");
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps_issue && lane == 0) {
    for (int i = warp; i < nops; i += nwarps_issue) {
      const double* g = src + ((long long) blockIdx.x*nops + i)*stride;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(s32(buf + (i % nslot)*(bytes/8))), "l"(g), "r"(bytes), "r"(s32(bar)) : "memory");
    }
  }
  long long t1 = clock64();
  // wait for completion
  unsigned ok = 0;
  while ( ! ok)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(s32(bar)), "r"(0) : "memory");
  long long t2 = clock64();
  if (lane == 0 && warp < nwarps_issue && blockIdx.x == 0) {
    out[2*warp] = t1 - t0;
    out[2*warp + 1] = t2 - t0;
  }
}
int main () {
  const int bytes = 4736, nops = 64, grid = 148;
  const long long stride = 86400;
  double* src; unsigned long long* out;
  cudaMalloc(&src, sizeof(double)*stride*nops*grid + (1 << 20));
  cudaMemset(src, 0, sizeof(double)*stride*nops*grid);
  cudaMalloc(&out, 64*8);
  const size_t smem = 64 + 16*bytes;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  for (int nw : {1, 2, 4, 8}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(out, 0, 64*8);
      k<<<grid, 256, smem>>>(src, stride, nops, nw, bytes, out);
      cudaDeviceSynchronize();
    }
    unsigned long long h[16];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("issue warps %d: %d ops of %d B per CTA: issue %llu cycles (%.0f per op per warp), all landed %llu cycles (%.1f B/clk/SM)  %s\n",
           nw, nops, bytes, h[0], (double) h[0]/(nops/nw), h[1], (double) nops*bytes/h[1],
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
