// CAAS::run as ONE pass over HBM: a thread-block CLUSTER holds a whole tracer on chip.
//
// The two-pass CAAS (up_kernel<CLS_CAAS> -> tier sweep -> caas_adjust_kernel) reads every
// cell twice because a tracer's redistribution factor depends on sums over all of its
// cells (cedr_caas.cpp:129-253), and 86,400 cells x 4 rows = 2.76 MB do not fit one SM. They
// do fit the shared memory of 16 SMs: a cluster of up to 16 CTAs (non-portable size) takes
// one tracer at a time, each CTA a contiguous slice of whole tier-0 blocks (<= 8 of them):
//   1. four bulk-TMA copies bring the slice's rows (min, Qm, max, prev) into a ring of five
//      row slots (two Qm slots by parity, three transient);
//   2. each 128-thread group sums its blocks in the tree order of up_kernel<CLS_CAAS> --
//      a thread is a depth-7 node -- KEEPING its leaves' bounds in registers, and stores the
//      block-root records into every CTA of the cluster through distributed shared memory;
//   3. the transient slots are free again: the NEXT tracer's rows start streaming in, under
//      everything that follows;
//   4. one cluster barrier; every CTA then sums the block records up the tier-1 tree in the
//      fixed tree order (BfbTreeAllReducer's order, cedr_bfb_tree_allreduce.cpp:86-124) and
//      forms the tracer's redistribution scalars (CAAS::finish_locally, cedr_caas.cpp:211-253);
//   5. the threads adjust their leaves in the Qm slot, which leaves by one bulk-TMA store.
// HBM traffic is the algorithmic 40 B per cell x tracer (4 rows in, 1 out). Same operations
// in the same order as the two-pass kernels: bit-identical results.
//
// tools/microbench/cluster_stream.cu measures the ceiling of this schedule on a B200: seven
// 16-CTA clusters with 216 KB of shared memory per CTA are co-resident and stream 4 rows in /
// 1 row out with a cluster barrier per tracer at 7.0 TB/s.
#ifndef CEDR_B200_CLUSTER_CAAS_CUH
#define CEDR_B200_CLUSTER_CAAS_CUH

#include <cstdio>

#include "fast_kernels.cuh"

namespace cedr_b200 {
namespace ccaas {

constexpr int kMaxBlocksPerGroup = 2;     // blocks one 128-thread group sums per tracer

struct Args {
  const BlockDev* blocks;       // tier-0 blocks, consecutive leaf ranges
  int nblocks;
  const unsigned short* dtab;
  const unsigned short* perm;   // FastArgs::perm (the up-sweep's bank-spreading deal)
  const double* const* rowaddr; // [4 t + role]
  const int* trcr_prob;
  int ntr;
  int bpc;                      // blocks per CTA
  int cs;                       // CTAs per cluster
  int nclusters;
  int cap;                      // doubles per ring slot (even)
  int need_prev;                // some tracer conserves: there is a Qm_prev row
  BlockDev top;                 // the tier-1 block: its leaves are the tier-0 block roots
  const int* lvlptr;
  const int* kid0;
  const int* kid1;
};

// Head of the shared memory, in bytes: 2 mbarriers, the scalars, the warp roots, the block
// records (leaves of the tier-1 tree, by tracer parity), the tier-1 tree's internal sums and
// its tables (kids, level pointers).
__host__ __device__ inline size_t head_bytes (const Args& a, const int ngroups) {
  const size_t n = 64 + sizeof(double)*(static_cast<size_t>(ngroups)*kMaxBlocksPerGroup*16 +
                                        4*(2*static_cast<size_t>(a.top.nl) + a.top.ni)) +
    sizeof(int)*(2*static_cast<size_t>(a.top.ni) + a.top.nlev + 1);
  return (n + 127) & ~static_cast<size_t>(127);
}
inline size_t smem_bytes (const Args& a, const int ngroups) {
  return head_bytes(a, ngroups) + sizeof(double)*5*static_cast<size_t>(a.cap);
}

__device__ __forceinline__ unsigned cluster_ctarank () {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all () {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Store a double into the same shared-memory location of CTA `rank` of this cluster.
__device__ __forceinline__ void st_cluster (double* local, const unsigned rank, const double v) {
  unsigned remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
               : "=r"(remote) : "r"(fast::smem_u32(local)), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" :: "r"(remote), "d"(v) : "memory");
}

#ifdef CEDR_CCAAS_CLOCKS
# define CCAAS_T(k) do { const long long c1 = clock64(); pc[k] += c1 - c0; c0 = c1; } while (0)
#else
# define CCAAS_T(k) do {} while (0)
#endif

template <int NG>
__global__ void __launch_bounds__(128*NG, 1)
run_kernel (const Args a) {
#ifdef CEDR_CCAAS_CLOCKS
  long long pc[8] = {0}, c0 = clock64();
#endif
  using namespace fast;
  extern __shared__ __align__(128) unsigned char smraw[];
  uint64_t* const mbar = reinterpret_cast<uint64_t*>(smraw);              // [2]
  double* const scal = reinterpret_cast<double*>(smraw + 16);             // mode, fac
  double* const wroot = reinterpret_cast<double*>(smraw + 64);            // [NG][2][4 warps][4]
  double* const recs = wroot + NG*kMaxBlocksPerGroup*16;                  // [2][4][nl1]
  const int nl1 = a.top.nl, ni1 = a.top.ni;
  double* const reci = recs + 2*4*nl1;                                    // [4][ni1]
  int* const tkid0 = reinterpret_cast<int*>(reci + 4*ni1);                // [ni1]
  int* const tkid1 = tkid0 + ni1;                                         // [ni1]
  int* const tlvl = tkid1 + ni1;                                          // [nlev + 1]
  double* const ring = reinterpret_cast<double*>(smraw + head_bytes(a, NG));   // [5][cap]
  const int cap = a.cap;

  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 7, node = tid & 127,
    w = (tid >> 5) & 3;
  const unsigned rank = cluster_ctarank();
  const int cluster = blockIdx.x/a.cs;
  const int b0 = static_cast<int>(rank)*a.bpc;
  const int nb = max(0, min(a.bpc, a.nblocks - b0));
  // This CTA's slice of the rows: leaves [L0, L1), staged from the even index below L0.
  int L0 = 0, L1 = 0;
  if (nb > 0) {
    L0 = a.blocks[b0].leaf0;
    L1 = a.blocks[b0 + nb - 1].leaf0 + a.blocks[b0 + nb - 1].nl;
  }
  const int src0 = L0 & ~1, shift = L0 - src0;
  const unsigned bytes = 8u*static_cast<unsigned>(((L1 + 1) & ~1) - src0);
  const int nrows = a.need_prev ? 4 : 3;

  // This thread's depth-7 node in each of its group's blocks: slot offsets of its four
  // depth-9 nodes (bit 15: a pair of leaves) and the block's leaf index in the tier above.
  int off[kMaxBlocksPerGroup][4], gidx[kMaxBlocksPerGroup], src_lane[kMaxBlocksPerGroup];
  bool pr[kMaxBlocksPerGroup][4], have[kMaxBlocksPerGroup];
#pragma unroll
  for (int jj = 0; jj < kMaxBlocksPerGroup; ++jj) {
    const int j = g + jj*NG;
    have[jj] = j < nb;
    gidx[jj] = 0;
    src_lane[jj] = lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) { off[jj][k] = 0; pr[jj][k] = false; }
    if (have[jj]) {
      const BlockDev B = a.blocks[b0 + j];
      // The thread's node is one of its own warp's 32, dealt so that the half-warps' leaf
      // offsets spread over the shared-memory banks; lane l gets node 32 w + l's record
      // back by one shuffle per field before the shuffle tree (as in up_kernel).
      const unsigned short* const pup = a.perm + B.fperm_up_off;
      const int mynode = pup[node];
      src_lane[jj] = pup[128 + node] & 31;
      const ushort4 e = reinterpret_cast<const ushort4*>(a.dtab + B.ftab_off)[mynode];
      const int base = B.leaf0 - L0 + shift;
      off[jj][0] = (e.x & 0x7fff) + base; pr[jj][0] = (e.x >> 15) != 0;
      off[jj][1] = (e.y & 0x7fff) + base; pr[jj][1] = (e.y >> 15) != 0;
      off[jj][2] = (e.z & 0x7fff) + base; pr[jj][2] = (e.z >> 15) != 0;
      off[jj][3] = (e.w & 0x7fff) + base; pr[jj][3] = (e.w >> 15) != 0;
      gidx[jj] = B.gidx;
    }
  }

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
  }
  for (int j = tid; j < ni1; j += 128*NG) {
    tkid0[j] = a.kid0[a.top.kid_off + j];
    tkid1[j] = a.kid1[a.top.kid_off + j];
  }
  for (int l = tid; l <= a.top.nlev; l += 128*NG) tlvl[l] = a.lvlptr[a.top.lvlptr_off + l];
  __syncthreads();
  // Everyone's barriers and shared memory exist before anyone stores into a peer.
  cluster_sync_all();

  // Slots: 0, 1 = Qm by parity; 2 = min, 3 = max, 4 = prev.
  auto issue = [&] (const int i, const int t) {
    if (nb == 0) return;
    const double* const* const ra = a.rowaddr + 4*t;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&mbar[i & 1], nrows*bytes);
    tma_load(ring + 2*cap, ra[0] + src0, bytes, &mbar[i & 1]);
    tma_load(ring + (i & 1)*cap, ra[1] + src0, bytes, &mbar[i & 1]);
    tma_load(ring + 3*cap, ra[2] + src0, bytes, &mbar[i & 1]);
    if (nrows == 4) tma_load(ring + 4*cap, ra[3] + src0, bytes, &mbar[i & 1]);
  };
  if (tid == 0 && cluster < a.ntr) issue(0, cluster);

  const double* const s_lo = ring + 2*cap;
  const double* const s_hi = ring + 3*cap;
  const double* const s_pv = ring + 4*cap;
  int i = 0;
  for (int t = cluster; t < a.ntr; t += a.nclusters, ++i) {
    const int p = i & 1;
    double* const s_q = ring + p*cap;
    double* const rec = recs + p*4*nl1;
    const bool conserve = a.need_prev && (a.trcr_prob[t] & 1);
    CCAAS_T(0);
    if (nb > 0) mbar_wait(&mbar[p], (i >> 1) & 1);
    CCAAS_T(1);

    // ---- 2. block sums, tree order of up_kernel<CLS_CAAS>; bounds stay in registers.
    double LO[kMaxBlocksPerGroup][8], HI[kMaxBlocksPerGroup][8];
    double r[kMaxBlocksPerGroup][4];
#pragma unroll
    for (int jj = 0; jj < kMaxBlocksPerGroup; ++jj) {
      double n[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        double v[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          LO[jj][2*k + j] = 0; HI[jj][2*k + j] = 0;
#pragma unroll
          for (int f = 0; f < 4; ++f) v[j][f] = 0;
          if ( ! have[jj] || (j == 1 && ! pr[jj][k])) continue;
          const int o = off[jj][k] + j;
          const double lo = s_lo[o], q = s_q[o], hi = s_hi[o];
          const double term = conserve ? s_pv[o] : q;
          const double clip = dev::rmin(hi, dev::rmax(lo, q));
          LO[jj][2*k + j] = lo; HI[jj][2*k + j] = hi;
          v[j][0] = 0.0 + lo; v[j][1] = 0.0 + clip; v[j][2] = 0.0 + hi; v[j][3] = 0.0 + term;
        }
#pragma unroll
        for (int f = 0; f < 4; ++f) n[k][f] = pr[jj][k] ? v[0][f] + v[1][f] : v[0][f];
      }
#pragma unroll
      for (int f = 0; f < 4; ++f) r[jj][f] = (n[0][f] + n[1][f]) + (n[2][f] + n[3][f]);
    }
#pragma unroll
    for (int jj = 0; jj < kMaxBlocksPerGroup; ++jj)
#pragma unroll
      for (int f = 0; f < 4; ++f) r[jj][f] = __shfl_sync(0xffffffffu, r[jj][f], src_lane[jj]);
    // Depths 6..2 inside the warp: lane l (l % 2^(L+1) == 0) takes left + right.
#pragma unroll
    for (int L = 0; L < 5; ++L)
#pragma unroll
      for (int jj = 0; jj < kMaxBlocksPerGroup; ++jj)
#pragma unroll
        for (int f = 0; f < 4; ++f)
          r[jj][f] = r[jj][f] + __shfl_down_sync(0xffffffffu, r[jj][f], 1 << L);
    double* const wr = wroot + g*kMaxBlocksPerGroup*16;     // [jj][warp][f]
    if (lane == 0) {
#pragma unroll
      for (int jj = 0; jj < kMaxBlocksPerGroup; ++jj)
#pragma unroll
        for (int f = 0; f < 4; ++f) wr[jj*16 + w*4 + f] = r[jj][f];
    }
    // The group's four warps (named barrier 1 + g).
    asm volatile("bar.sync %0, 128;" :: "r"(1 + g) : "memory");
    {
      // Depths 1 and 0, then the record goes to every CTA of the cluster: thread `node`
      // takes block jj = node / 64, field (node / 16) % 4, destination CTA node % 16.
      const int jj = node >> 6, f = (node >> 4) & 3, dst_cta = node & 15;
      const bool mine = jj == 0 ? have[0] : have[kMaxBlocksPerGroup - 1];
      if (mine && dst_cta < a.cs) {
        const double* const x = wr + jj*16 + f;
        const double v = (x[0] + x[4]) + (x[8] + x[12]);
        st_cluster(rec + f*nl1 + (jj == 0 ? gidx[0] : gidx[kMaxBlocksPerGroup - 1]),
                   static_cast<unsigned>(dst_cta), v);
      }
    }
    CCAAS_T(2);
    __syncthreads();
    // ---- 3. the transient slots are free: the next tracer streams in. Its Qm goes to the
    // other Qm slot, which the bulk store of the tracer before this one has left by now.
    if (tid == 0 && t + a.nclusters < a.ntr) {
      tma_store_wait_read();
      issue(i + 1, t + a.nclusters);
    }
    // ---- 4. all block records are in every CTA: tier-1 sums in tree order, scalars.
    CCAAS_T(3);
    cluster_sync_all();
    CCAAS_T(4);
    {
      const int* const lvlptr = tlvl;
      const int* const kid0 = tkid0;
      const int* const kid1 = tkid1;
      // Tier-1 node id < nl1: a block record; else internal node id - nl1.
      auto at = [&] (const int f, const int id) -> double& {
        return id < nl1 ? rec[f*nl1 + id] : reci[f*ni1 + id - nl1];
      };
      // Warp f sums field f up the levels on its own (no block-wide barrier per level).
      if (tid < 128) {
        const int f = tid >> 5;
        for (int l = 0; l < a.top.nlev; ++l) {
          const int je = lvlptr[l + 1];
          for (int j = lvlptr[l] + lane; j < je; j += 32)
            reci[f*ni1 + j] = (0.0 + at(f, kid0[j])) + at(f, kid1[j]);
          __syncwarp();
        }
      }
      __syncthreads();
      if (tid == 0) {
        const int root = ni1 ? nl1 + ni1 - 1 : 0;
        const double clip_sum = at(1, root), term_sum = at(3, root);
        const double m = term_sum - clip_sum;
        double mode = 0, fac = 0;
        if (m < 0) {
          fac = clip_sum - at(0, root);
          if (fac > 0) { fac = m/fac; mode = -1; }
        } else if (m > 0) {
          fac = at(2, root) - clip_sum;
          if (fac > 0) { fac = m/fac; mode = 1; }
        }
        scal[0] = mode;
        scal[1] = fac;
      }
      __syncthreads();
    }
    CCAAS_T(5);
    // ---- 5. CAAS::finish_locally per cell (caas_adjust_kernel), in the Qm slot.
    if (nb > 0) {
      const double mode = scal[0], fac = scal[1];
#pragma unroll
      for (int jj = 0; jj < kMaxBlocksPerGroup; ++jj) {
        if ( ! have[jj]) continue;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (j == 1 && ! pr[jj][k]) break;
            const int o = off[jj][k] + j;
            const double lo = LO[jj][2*k + j], hi = HI[jj][2*k + j];
            double q = dev::rmin(hi, dev::rmax(lo, s_q[o]));
            if (mode < 0) {
              q += fac*(q - lo);
              q = dev::rmax(lo, q);
            } else if (mode > 0) {
              q += fac*(hi - q);
              q = dev::rmin(hi, q);
            }
            s_q[o] = q;
          }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && nb > 0) {
      double* const o = const_cast<double*>(a.rowaddr[4*t + 1]) + L0;
      const int n = L1 - L0, q0 = L0 & 1;
      const int nint = (n - q0) & ~1;
      if (nint) tma_store(o + q0, s_q + shift + q0, 8u*static_cast<unsigned>(nint));
      tma_store_commit();
      if (q0) o[0] = s_q[shift];
      if (q0 + nint < n) o[n - 1] = s_q[shift + n - 1];
    }
  }
  CCAAS_T(6);
#ifdef CEDR_CCAAS_CLOCKS
  if (tid == 0 && blockIdx.x == 0)
    printf("ccaas clocks: loop %lld wait-load %lld sums %lld sync+issue %lld cluster-barrier %lld "
           "top %lld adjust+store %lld (iterations %d)\n", pc[0], pc[1], pc[2], pc[3], pc[4],
           pc[5], pc[6], i);
#endif
  if (tid == 0) tma_store_wait_read();
  // No CTA may exit while a peer can still store into its shared memory.
  cluster_sync_all();
}

} // namespace ccaas
} // namespace cedr_b200

#endif
