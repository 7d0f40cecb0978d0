// Fused persistent run() kernel for plans whose tier-0 blocks all have the fast shape
// (see fast_kernels.cuh: recursive bisection, 513..1024 leaves, perfect to depth 9) and
// whose block roots form one tier-1 block: every cubed-sphere config of BASELINE.json.
//
// One launch does the whole of QLT::run (cedr_qlt.cpp:618-640) for a problem class, or
// the whole of CAAS::run (cedr_caas.cpp:258-270):
//
//   CTA (block b, lane l) walks tracers l, l + L, l + 2L, ... of the class. For tracer k
//     UP(k)    TMA-stage the block's leaf rows, sum them in tree order to the block-root
//              record, publish it, and count the arrival on the tracer's counter;
//     TOP(k)   the CTA whose arrival completes the tracer sweeps the tier-1 block
//              (up-sweep over the block roots, root_compute, down-sweep) and raises
//              the tracer's flag;
//     DOWN(k)  `depth` tracers later, re-stage (min, Qm, max) -- an L2 hit, the window
//              of in-flight tracers is a few tens of MB -- wait for the flag, and solve
//              every node problem of the block from its root mass down to the leaves.
//
// So the leaf data cross HBM once on the way in (32 B per cell x tracer) and once on
// the way out (8 B), nothing waits on a kernel boundary, and the only grid-wide
// dependency (all blocks of a tracer -> its root) is hidden behind the next tracers'
// up-sweeps. All CTAs must be co-resident: the host launches cooperatively.
//
// Node arithmetic is node_solve.cuh in the reference's tree order: bit-identical to the
// generic kernels and to the reference.
#ifndef CEDR_B200_FUSED_KERNELS_CUH
#define CEDR_B200_FUSED_KERNELS_CUH

#include <cstdint>

#include "fast_kernels.cuh"

namespace cedr_b200 {
namespace fused {

using fast::kD9;
using fast::kHeapNodes;
using fast::mbar_init;
using fast::mbar_fence_init;
using fast::mbar_expect_tx;
using fast::mbar_wait;
using fast::tma_load;

constexpr int kThreads = 128;   // one thread per depth-7 node of the block
constexpr int kMinCtasPerSm = 4;

struct Args {
  const BlockDev* blocks;       // tier 0
  int nblocks;
  const unsigned short* dtab;   // per shape: 512 depth-9 entries, off | (pair << 15)
  const unsigned short* ptab;   // per shape: depth-9 positions of the pairs, increasing
  const dev::NodeWQ* wq;        // per block, fast order (heap nodes, then pairs)
  const dev::NodeRh* rh;
  const double* in;             // tier-0 rows
  long long in_ld;
  const int* trcr_row;
  const int* trcr_prob;
  double* rec;                  // block-root records [(4 t + f) rec_ld + block]
  long long rec_ld;
  const double* sol;            // solved block-root masses [t sol_ld + block]
  long long sol_ld;
  double* out;                  // [t out_ld + leaf]; CAAS: the `in` buffer itself
  long long out_ld;
  const int* tracers;           // tracer ids of the class
  int ntr;
  int sbuf;                     // doubles per staged row (even, >= max_nl + 2)
  int prefer_mass_con;
  int depth;                    // tracers between UP(k) and DOWN(k), >= 1
  unsigned* cnt;                // [ntr] arrivals (zeroed before the launch)
  unsigned* flag;               // [ntr] raised when the tracer's tier 1 is solved
  int* status;                  // set nonzero if a flag wait gives up
  const double* caas_scal;      // CAAS: [2t] mode, [2t+1] fac (written by TOP)
  SweepArgs top;                // the tier-1 block sweep
};

__device__ __forceinline__ void fence_proxy_async () {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ unsigned ld_acquire (const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release (unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// Shared-memory carve-up, in doubles from the (16-byte aligned) base.
struct Smem {
  double* bufU;    // [4][sbuf]   UP staging: min, Qm, max, prev
  double* un;      // [4][256]    heap nodes 0..254: min, Qm, max sums; solved mass
  double* d9x;     // [512]       solved masses of the depth-9 pair nodes
  double* bufD;    // [3][sbuf]   DOWN staging: min, Qm (-> solved leaf masses), max
  double* wroot;   // [16]        warp-root records of the up-sweep
  uint64_t* mbar;  // [2]         bufU full, bufD full
  int* misc;       // [4]
  dev::NodeWQ* topc; // [128]     node constants of heap nodes 0..126 (depths 0..6)
  __device__ Smem (unsigned char* raw, const int sbuf) {
    bufU = reinterpret_cast<double*>(raw);
    un = bufU + 4*sbuf;
    d9x = un + 4*256;
    bufD = d9x + kD9;
    wroot = bufD + 3*sbuf;
    mbar = reinterpret_cast<uint64_t*>(wroot + 16);
    misc = reinterpret_cast<int*>(mbar + 2);
    topc = reinterpret_cast<dev::NodeWQ*>(misc + 4);
  }
};
inline size_t smem_bytes (const int sbuf) {
  return sizeof(double)*(7*static_cast<size_t>(sbuf) + 4*256 + kD9 + 16) + 16 + 16 +
    128*sizeof(dev::NodeWQ);
}
// Doubles of scratch the in-kernel tier-1 sweep may use (bufU + un + d9x, contiguous).
inline size_t top_scratch_doubles (const int sbuf) {
  return 4*static_cast<size_t>(sbuf) + 4*256 + kD9;
}

template <int CLS>
__global__ void __launch_bounds__(kThreads, kMinCtasPerSm)
run_kernel (const Args a) {
  static_assert(CLS == CLS_ST || CLS == CLS_CST || CLS == CLS_CAAS, "fused classes");
  constexpr bool caas = CLS == CLS_CAAS;
  constexpr bool prev = CLS != CLS_ST;
  constexpr int nrowsU = prev ? 4 : 3;
  extern __shared__ __align__(16) unsigned char smraw[];
  const Smem sm(smraw, a.sbuf);
  const int sbuf = a.sbuf;

  const int b = blockIdx.x % a.nblocks, lane_id = blockIdx.x / a.nblocks;
  const int nlanes = gridDim.x / a.nblocks;
  const int K = lane_id < a.ntr ? (a.ntr - lane_id + nlanes - 1)/nlanes : 0;
  const int D = a.depth;
  const BlockDev B = a.blocks[b];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int src0 = B.leaf0 & ~1, shift = B.leaf0 - src0;
  const unsigned bytes = 8u*static_cast<unsigned>(((B.leaf0 + B.nl + 1) & ~1) - src0);
  const unsigned short* const dtab = a.dtab + B.ftab_off;
  const unsigned short* const ptab = a.ptab + B.fpair_off;
  const dev::NodeWQ* const wq = a.wq + B.fbase;
  const dev::NodeRh* const rh = a.rh + B.fbase;

  // This thread's four depth-9 nodes: leaf offset in the staged rows, pair or leaf.
  const ushort4 e = reinterpret_cast<const ushort4*>(dtab)[tid];
  const int off[4] = {(e.x & 0x7fff) + shift, (e.y & 0x7fff) + shift,
                      (e.z & 0x7fff) + shift, (e.w & 0x7fff) + shift};
  const bool pr[4] = {(e.x >> 15) != 0, (e.y >> 15) != 0, (e.z >> 15) != 0,
                      (e.w >> 15) != 0};
  // This warp's pairs are ptab[ps .. pe): depth-9 positions in [128 warp, 128 warp + 128).
  int ps = 0, pe = 0;
  if ( ! caas) {
    for (int j = lane; j < B.npairs; j += 32) {
      const int p = ptab[j];
      ps += p < 128*warp;
      pe += p < 128*warp + 128;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      ps += __shfl_xor_sync(0xffffffffu, ps, o);
      pe += __shfl_xor_sync(0xffffffffu, pe, o);
    }
  }
  // Node constants of this thread's micro-subtree (depth 7: heap 127 + tid; depth 8:
  // heap 255 + 2 tid, + 1): the same for every tracer.
  dev::NodeWQ c7, c8a, c8b;
  // Where this lane's (up to two) pairs of the warp's list sit: leaf offset in the
  // staged rows | depth-9 position << 16; 0xffffffff if none.
  unsigned pair_slot[2] = {0xffffffffu, 0xffffffffu};
  if ( ! caas) {
    c7 = wq[127 + tid]; c8a = wq[255 + 2*tid]; c8b = wq[256 + 2*tid];
    if (tid < kHeapNodes/4) sm.topc[tid] = wq[tid];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int j = ps + lane + 32*q;
      if (j < pe) {
        const unsigned p = ptab[j];
        pair_slot[q] = ((dtab[p] & 0x7fffu) + shift) | (p << 16);
      }
    }
  }

  if (tid == 0) {
    mbar_init(&sm.mbar[0], 1);
    mbar_init(&sm.mbar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto tracer_of = [&] (const int k) { return a.tracers[lane_id + k*nlanes]; };
  auto issue_up = [&] (const int k) {
    const int t = tracer_of(k);
    const double* src = a.in + static_cast<long long>(a.trcr_row[t])*a.in_ld + src0;
    fence_proxy_async();
    mbar_expect_tx(&sm.mbar[0], nrowsU*bytes);
#pragma unroll
    for (int f = 0; f < nrowsU; ++f)
      tma_load(sm.bufU + f*sbuf, src + f*a.in_ld, bytes, &sm.mbar[0]);
  };
  auto issue_down = [&] (const int k) {
    const int t = tracer_of(k);
    const double* src = a.in + static_cast<long long>(a.trcr_row[t])*a.in_ld + src0;
    fence_proxy_async();
    mbar_expect_tx(&sm.mbar[1], 3*bytes);
#pragma unroll
    for (int f = 0; f < 3; ++f)
      tma_load(sm.bufD + f*sbuf, src + f*a.in_ld, bytes, &sm.mbar[1]);
  };
  // Wait until tracer index i's tier 1 is solved (one thread).
  auto wait_flag = [&] (const int i) {
    unsigned spins = 0;
    while (ld_acquire(a.flag + i) == 0) {
      __nanosleep(32);
      // Watchdog: never hang the device. Once any wait has given up, none waits again.
      if ((++spins & 1023u) == 0 &&
          (spins > (1u << 23) || *reinterpret_cast<volatile int*>(a.status))) {
        atomicExch(a.status, 1);
        break;
      }
    }
  };

  if (tid == 0 && K > 0) issue_up(0);

  for (int s = 0; s < K + D; ++s) {
    const bool do_up = s < K, do_down = s >= D;
    // bufD is free: DOWN(s - D - 1) ended with a barrier.
    if (tid == 0 && do_down) issue_down(s - D);

    if (do_up) {
      // ------------------------------------------------------------------ UP(s)
      const int i = lane_id + s*nlanes, t = a.tracers[i];
      mbar_wait(&sm.mbar[0], s & 1);
      const double* const u = sm.bufU;
      double r[4];   // (min, Qm | clip, max, prev | term) of this thread's depth-7 node
      {
        double n[4][4];
        bool conserve = true;
        if (caas) conserve = a.trcr_prob[t] & 1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          double v[2][4];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            // A lone leaf's right neighbour is read but not used (in-bounds: sbuf has
            // room for max_nl + 2).
            const int o = off[k] + j;
            if (caas) {
              const double lo = u[o], q = u[sbuf + o], hi = u[2*sbuf + o];
              const double term = conserve ? u[3*sbuf + o] : q;
              const double clip = dev::rmin(hi, dev::rmax(lo, q));
              v[j][0] = 0.0 + lo; v[j][1] = 0.0 + clip; v[j][2] = 0.0 + hi;
              v[j][3] = 0.0 + term;
            } else {
              v[j][0] = u[o]; v[j][1] = u[sbuf + o]; v[j][2] = u[2*sbuf + o];
              if (prev) v[j][3] = u[3*sbuf + o];
            }
          }
#pragma unroll
          for (int f = 0; f < nrowsU; ++f) n[k][f] = pr[k] ? v[0][f] + v[1][f] : v[0][f];
        }
#pragma unroll
        for (int f = 0; f < nrowsU; ++f) r[f] = (n[0][f] + n[1][f]) + (n[2][f] + n[3][f]);
      }
      // Depths 6..2 inside the warp: lane l (l % 2^(L+1) == 0) takes left + right.
#pragma unroll
      for (int Lv = 0; Lv < 5; ++Lv) {
#pragma unroll
        for (int f = 0; f < nrowsU; ++f)
          r[f] = r[f] + __shfl_down_sync(0xffffffffu, r[f], 1 << Lv);
      }
      if (lane == 0) {
#pragma unroll
        for (int f = 0; f < nrowsU; ++f) sm.wroot[warp*4 + f] = r[f];
      }
      __syncthreads();   // bufU consumed by everyone; warp roots visible
      if (tid == 0) {
        // Depths 1 and 0, then the record for tier 1.
        double* rec = a.rec + static_cast<long long>(t)*4*a.rec_ld + B.gidx;
#pragma unroll
        for (int f = 0; f < nrowsU; ++f)
          rec[f*a.rec_ld] = (sm.wroot[f] + sm.wroot[4 + f]) + (sm.wroot[8 + f] + sm.wroot[12 + f]);
        __threadfence();
        const unsigned old = atomicAdd(a.cnt + i, 1u);
        sm.misc[0] = old + 1 == static_cast<unsigned>(a.nblocks);
      }
      __syncthreads();
      if (sm.misc[0]) {
        // ---------------------------------------------------------------- TOP(s)
        // Every block of tracer t has published its record: sweep tier 1 here.
        __threadfence();
        sweep_block<CLS, MODE_TOP>(a.top, 0, t, sm.bufU);
        __syncthreads();
        if (tid == 0) {
          __threadfence();
          st_release(a.flag + i, 1u);
        }
      }
      if (tid == 0 && s + 1 < K) issue_up(s + 1);
    }

    if (do_down) {
      // ------------------------------------------------------------- DOWN(s - D)
      const int kd = s - D;
      const int i = lane_id + kd*nlanes, t = a.tracers[i];
      mbar_wait(&sm.mbar[1], kd & 1);
      double* const d = sm.bufD;
      if (caas) {
        // CAAS::finish_locally, cedr_caas.cpp:211-253, on the clipped values
        // (reduce_locally stores the clip in place, :177).
        if (tid == 0) {
          wait_flag(i);
          sm.wroot[0] = __ldcg(a.caas_scal + 2*t);
          sm.wroot[1] = __ldcg(a.caas_scal + 2*t + 1);
        }
        __syncthreads();
        const double mode = sm.wroot[0], fac = sm.wroot[1];
        double* const o = a.out + (static_cast<long long>(a.trcr_row[t]) + 1)*a.out_ld + B.leaf0;
        for (int k = tid; k < B.nl; k += kThreads) {
          const double lo = d[shift + k], hi = d[2*sbuf + shift + k];
          double q = dev::rmin(hi, dev::rmax(lo, d[sbuf + shift + k]));
          if (mode < 0) {
            q += fac*(q - lo);
            q = dev::rmax(lo, q);
          } else if (mode > 0) {
            q += fac*(hi - q);
            q = dev::rmin(hi, q);
          }
          o[k] = q;
        }
        __syncthreads();
        continue;
      }
      const bool prefer = a.prefer_mass_con != 0;
      auto solve = [&] (const dev::NodeWQ& c, const int cpos, const double* nd,
                        const double bm, const double* k0, const double* k1, double& x0,
                        double& x1) {
        if (prefer)
          dev::solve_bounded_lean<true>(c, 0.0, rh + cpos, nd[0], nd[1], nd[2], bm, k0[0], k0[1],
                                        k0[2], k1[0], k1[1], k1[2], x0, x1);
        else
          dev::solve_bounded_lean<false>(c, 0.0, rh + cpos, nd[0], nd[1], nd[2], bm, k0[0], k0[1],
                                         k0[2], k1[0], k1[1], k1[2], x0, x1);
      };
      // Sums of the depth-9 node k of this thread (a leaf or a pair), from the rows.
      auto node9 = [&] (const int k, double* n9) {
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          const double v0 = d[f*sbuf + off[k]], v1 = d[f*sbuf + off[k] + 1];
          n9[f] = pr[k] ? v0 + v1 : v0;
        }
      };
      double* const un = sm.un;
      // Sums of this thread's two depth-8 nodes (nothing of the micro-subtree is kept in
      // registers across the block-top phase; it is re-summed from the staged rows).
      auto node8 = [&] (double n8[2][3]) {
        double n9[4][3];
#pragma unroll
        for (int k = 0; k < 4; ++k) node9(k, n9[k]);
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          n8[0][f] = n9[0][f] + n9[1][f];
          n8[1][f] = n9[2][f] + n9[3][f];
        }
      };
      {
        double n8[2][3];
        node8(n8);
#pragma unroll
        for (int f = 0; f < 3; ++f) un[f*256 + 127 + tid] = n8[0][f] + n8[1][f];
      }
      __syncthreads();
      // Sums of depths 6..0 (heap node h has kids 2h+1, 2h+2).
      for (int dd = 6; dd >= 0; --dd) {
        if (tid < (1 << dd)) {
          const int h = (1 << dd) - 1 + tid;
#pragma unroll
          for (int f = 0; f < 3; ++f)
            un[f*256 + h] = un[f*256 + 2*h + 1] + un[f*256 + 2*h + 2];
        }
        if (dd > 5) __syncthreads(); else __syncwarp();
      }
      // The block root's mass comes from tier 1.
      if (tid == 0) {
        wait_flag(i);
        un[3*256] = __ldcg(a.sol + static_cast<long long>(t)*a.sol_ld + B.gidx);
      }
      __syncwarp();
      // Node problems of depths 0..6; depths 0..5 fit in warp 0.
      for (int dd = 0; dd <= 6; ++dd) {
        if (dd == 6) __syncthreads();
        if (tid < (1 << dd)) {
          const int h = (1 << dd) - 1 + tid;
          const double nd[3] = {un[h], un[256 + h], un[512 + h]};
          const double k0[3] = {un[2*h + 1], un[256 + 2*h + 1], un[512 + 2*h + 1]};
          const double k1[3] = {un[2*h + 2], un[256 + 2*h + 2], un[512 + 2*h + 2]};
          double x0, x1;
          solve(sm.topc[h], h, nd, un[768 + h], k0, k1, x0, x1);
          un[768 + 2*h + 1] = x0;
          un[768 + 2*h + 2] = x1;
        }
        if (dd < 5) __syncwarp();
      }
      __syncthreads();
      // The micro-subtree in registers: depth 7, then the two depth-8 nodes.
      double* const xout = d + sbuf;   // solved leaf masses replace the Qm row
      // Constants of this lane's pairs: issued now, used after the depth-8 solves.
      dev::NodeWQ cp[2];
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (pair_slot[q] != 0xffffffffu) cp[q] = wq[kHeapNodes + ps + lane + 32*q];
      {
        double n8[2][3], x8[2];
        node8(n8);
        const double n7[3] = {un[127 + tid], un[256 + 127 + tid], un[512 + 127 + tid]};
        solve(c7, 127 + tid, n7, un[768 + 127 + tid], n8[0], n8[1], x8[0], x8[1]);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          double n9[2][3], x9[2];
          node9(2*hf, n9[0]);
          node9(2*hf + 1, n9[1]);
          solve(hf ? c8b : c8a, 255 + 2*tid + hf, n8[hf], x8[hf], n9[0], n9[1], x9[0], x9[1]);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int k = 2*hf + j;
            if (pr[k]) sm.d9x[4*tid + k] = x9[j];
            else xout[off[k]] = x9[j];
          }
        }
      }
      __syncwarp();
      // This warp's depth-9 pairs, densely over its lanes.
      auto solve_pair = [&] (const dev::NodeWQ& c, const int j, const int o, const int p) {
        double k0[3], k1[3], nd[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          k0[f] = d[f*sbuf + o];
          k1[f] = d[f*sbuf + o + 1];
          nd[f] = k0[f] + k1[f];
        }
        double x0, x1;
        solve(c, kHeapNodes + j, nd, sm.d9x[p], k0, k1, x0, x1);
        xout[o] = x0;
        xout[o + 1] = x1;
      };
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (pair_slot[q] != 0xffffffffu)
          solve_pair(cp[q], ps + lane + 32*q, pair_slot[q] & 0xffff, pair_slot[q] >> 16);
      for (int j = ps + lane + 64; j < pe; j += 32) {   // blocks with > 64 pairs per warp
        const int p = ptab[j];
        solve_pair(wq[kHeapNodes + j], j, (dtab[p] & 0x7fff) + shift, p);
      }
      __syncthreads();
      // Coalesced write-back of the block's solved leaf masses.
      {
        double* const o = a.out + static_cast<long long>(t)*a.out_ld + B.leaf0;
        for (int k = tid; k < B.nl; k += kThreads) o[k] = xout[shift + k];
      }
      __syncthreads();
    }
  }
}

} // namespace fused
} // namespace cedr_b200

#endif
