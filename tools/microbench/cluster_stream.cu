// Go / no-go for a cluster-resident CAAS: how many 16-CTA clusters with ~216 KB of shared
// memory per CTA are co-resident on a B200, and what aggregate HBM bandwidth do they reach
// when every CTA streams 4 rows in (cp.async.bulk into a 5-slot ring) and 1 row out per
// "tracer", with one cluster barrier per tracer (the exchange of the partial sums)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_stream cluster_stream.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned s32 (const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait (uint64_t* bar, unsigned parity) {
  unsigned ok = 0;
  while ( ! ok)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
}
__global__ void k (const double* src, double* dst, long long row_ld, int cells_per_cta, int ntr,
                   int nclusters, int cs, int use_cluster_barrier) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);          // [2]
  double* ring = reinterpret_cast<double*>(sm + 128);      // 5 slots
  const int tid = threadIdx.x;
  const int cluster = blockIdx.x / cs, rank = blockIdx.x % cs;
  const unsigned bytes = 8u*cells_per_cta;
  const long long cell0 = (long long) rank*cells_per_cta;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(&bar[0])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(&bar[1])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // slots: 0,1 = q (by parity), 2,3,4 = lo, hi, prev (transient)
  auto issue = [&] (int i, int t) {
    uint64_t* b = &bar[i & 1];
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(b)), "r"(4*bytes) : "memory");
    for (int f = 0; f < 4; ++f) {
      const int slot = f == 1 ? (i & 1) : (f == 0 ? 2 : f == 2 ? 3 : 4);
      const double* g = src + ((long long) t*4 + f)*row_ld + cell0;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(s32(ring + (long long) slot*cells_per_cta)), "l"(g), "r"(bytes), "r"(s32(b)) : "memory");
    }
  };
  int i = 0;
  if (tid == 0 && cluster < ntr) issue(0, cluster);
  for (int t = cluster; t < ntr; t += nclusters, ++i) {
    mwait(&bar[i & 1], (i >> 1) & 1);
    __syncthreads();      // "sums done": transient slots free
    if (tid == 0) {
      if (t + nclusters < ntr) issue(i + 1, t + nclusters);
    }
    if (use_cluster_barrier) {
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (tid == 0) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // store of i-2.. done reading
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   :: "l"(dst + (long long) t*row_ld + cell0), "r"(s32(ring + (long long) (i & 1)*cells_per_cta)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main () {
  const int ncells = 86400, ntr = 1280;
  const long long ld = ncells;
  double *src, *dst;
  cudaMalloc(&src, sizeof(double)*ld*4*ntr);
  cudaMalloc(&dst, sizeof(double)*ld*ntr);
  cudaMemset(src, 0, sizeof(double)*ld*4*ntr);
  for (int cs : {16, 8, 4}) {
    const int cells = ncells/16;     // the CTA's slice stays 5400 cells: smem is the limit
    const size_t smem = 128 + 5*8*(size_t) cells;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int threads : {128, 512}) {
      cudaLaunchConfig_t cfg = {};
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.attrs = at; cfg.numAttrs = 1;
      cfg.gridDim = dim3(cs);
      int maxc = 0;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, k, &cfg);
      printf("cluster %d x %d threads, %zu B smem: max active clusters %d (%s)\n", cs, threads, smem, maxc, cudaGetErrorString(e));
      if (maxc < 1) continue;
      for (int cb : {0, 1}) {
        const int nclusters = maxc;
        cfg.gridDim = dim3(nclusters*cs);
        // a "tracer" here = cs*cells cells; scale the count so the volume is the same
        const int ntr_eff = ntr;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          cudaLaunchKernelEx(&cfg, k, (const double*) src, dst, ld, cells, ntr_eff, nclusters, cs, cb);
          cudaEventRecord(e1);
          cudaError_t e2 = cudaDeviceSynchronize();
          if (e2 != cudaSuccess) { printf("  error %s\n", cudaGetErrorString(e2)); return 1; }
          float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double bytes = 5.0*8*cells*cs*(double) ntr_eff;
        printf("  cluster barrier %d: %.3f ms, %.0f GB/s (%d CTAs)\n", cb, best, bytes/best/1e6, nclusters*cs);
      }
    }
  }
  return 0;
}
