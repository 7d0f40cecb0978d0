// The down-sweep of the fast-path blocks as TWO kernels (the default for the st / cst
// classes when the block tops come from the tier above, FastArgs::split = 2 or 3).
//
// down2_kernel (fast_kernels.cuh) does a whole block x tracer in one CTA: a top warp solves
// depths 0..6 one tracer ahead of four leaf warps. That top warp is the kernel's critical
// resource -- its levels are narrower than a warp (8, 16, 32, 64 nodes: 75 % of its lanes
// work), every level is a shared-memory round trip, the leaf warps wait for it (13-20 % of
// their time) and it costs the CTA 32 threads x 128 registers, so only three CTAs fit an SM.
//
//   midT_kernel   depths S..6 with LANES = TRACERS: a warp takes one depth-4 node and 32
//                 tracers, so the tree structure and the node constants are warp-uniform,
//                 all lanes work at every level and a thread walks its subtree (8 depth-7
//                 sums from up_kernel's n7buf as leaves) alone: no barrier, shuffle or
//                 hand-off. The one or two levels above depth 4 are solved redundantly by
//                 the 2 / 4 threads below them. The depth-7 sums reach the lanes through a
//                 shared-memory transposition: coalesced 8-byte cp.async copies with lanes
//                 along nodes, then reads with lanes along tracers (row pitch 33 doubles,
//                 tracer pitch 99: conflict-free). Masses go to
//                 x7[(tracer*nblocks + block)*128 + depth-7 node], 16 bytes per store.
//   down3_kernel  depths 7..9: down2_kernel's four leaf warps (a lane is a depth-7 node, one
//                 tracer at a time, TMA-staged rows) without the top warp: the depth-7
//                 masses arrive as a fourth staged row. 128 threads and 38 KB of shared
//                 memory per CTA, so four CTAs fit an SM, and no leaf warp waits for a top
//                 warp.
//
// Same node arithmetic (node_solve.cuh), same tree order: bit-identical to down2_kernel.
// Measured at ne120 x 1280 cst tracers: midT 0.23 + down3 0.95 ms against down2 1.28 ms.
// (A leaf-level kernel with lanes = tracers was also built -- 32 tracers x the <= 32 leaves
// of four depth-7 nodes transposed through shared memory, double-buffered: 1.41 ms, half of
// its instructions and two thirds of its time in the transposition, so the leaf levels keep
// the lane = node form.)
#ifndef CEDR_B200_TRANSPOSED_KERNELS_CUH
#define CEDR_B200_TRANSPOSED_KERNELS_CUH

#include "fast_kernels.cuh"

namespace cedr_b200 {
namespace fast {

constexpr int kTLanes = 32;            // tracers per CTA
constexpr int kTRow = 33;              // doubles per staged row (<= 32 leaves + 1: odd pitch)
constexpr int kTPitch = 3*kTRow;       // doubles per tracer: min, Qm, max

struct TArgs {
  double* x7;                          // depth-7 masses, [(tracer*nblocks + block)*128 + node]
};

__device__ __forceinline__ void cp_async8 (double* sdst, const double* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;"
               :: "r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all () {
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// ------------------------------------------------------------------------ MID
template <int CLS, bool PREFER>
__global__ void __launch_bounds__(128, 4)
midT_kernel (const FastArgs a, const TArgs ta) {
  static_assert(CLS == CLS_ST || CLS == CLS_CST, "transposed down-sweep: st / cst only");
  __shared__ double tile[kTLanes*kTPitch];
  __shared__ dev::NodeWQ cst[32];
  const int b = blockIdx.x >> 2, d2 = blockIdx.x & 3;
  const BlockDev B = a.blocks[b];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g0 = blockIdx.y*kTLanes;
  const int gn = min(kTLanes, a.ntr - g0);
  const dev::NodeWQ* const wq = a.wq + B.fbase;
  const dev::NodeRh* const rh = a.rh + B.fbase;
  const int S = a.split;

  // Node constants of this depth-2 subtree, relative heap order: depth 2 + dd, position
  // (d2 << dd) + p -> cst[2^dd - 1 + p].
  if (tid < 31) {
    const int dd = 31 - __clz(tid + 1), p = tid + 1 - (1 << dd);
    cst[tid] = wq[(4 << dd) - 1 + (d2 << dd) + p];
  }
  // The 32 depth-7 sums below this depth-2 node, three fields, for each tracer.
  const int tmine = lane < gn ? a.tracers[g0 + lane] : 0;
  for (int r = w; r < 3*gn; r += 4) {
    const int i = r/3, f = r - 3*i;
    const int t = __shfl_sync(0xffffffffu, tmine, i);
    cp_async8(tile + i*kTPitch + f*kTRow + lane,
              a.n7buf + (static_cast<long long>(t)*a.nblocks + b)*384 + f*128 + 32*d2 + lane);
  }
  // The mass of this thread's sub-root, from the tier above.
  double xin = 0;
  if (lane < gn) {
    const int E = 1 << S;
    const int j = S == 3 ? 2*d2 + (w >> 1) : d2;
    xin = __ldcg(a.sol_in + static_cast<long long>(tmine)*a.sol_in_ld +
                 static_cast<long long>(B.gidx)*E + j);
  }
  cp_async_wait_all();
  __syncthreads();
  if (lane >= gn) return;

  const double* const s = tile + lane*kTPitch;
  // Sums of the own depth-4 node's subtree (leaves 8w..8w+7), in tree order.
  double n6[4][3], n5[2][3], n4[3];
#pragma unroll
  for (int f = 0; f < 3; ++f) {
#pragma unroll
    for (int q = 0; q < 4; ++q) n6[q][f] = s[f*kTRow + 8*w + 2*q] + s[f*kTRow + 8*w + 2*q + 1];
    n5[0][f] = n6[0][f] + n6[1][f];
    n5[1][f] = n6[2][f] + n6[3][f];
    n4[f] = n5[0][f] + n5[1][f];
  }
  // A depth-4 sum of another node of this depth-2 subtree.
  auto sum4 = [&] (const int k, double* o) {
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const double* const l = s + f*kTRow + 8*k;
      o[f] = ((l[0] + l[1]) + (l[2] + l[3])) + ((l[4] + l[5]) + (l[6] + l[7]));
    }
  };
  double sib[3], n3[3];
  sum4(w ^ 1, sib);
  const bool right4 = w & 1, right3 = w & 2;
#pragma unroll
  for (int f = 0; f < 3; ++f) n3[f] = right4 ? sib[f] + n4[f] : n4[f] + sib[f];

  double x3 = xin;
  if (S == 2) {
    // Depth 2, solved by all four warps alike.
    double ca[3], cb[3], c3[3], n2[3];
    sum4((w ^ 2) & 2, ca);
    sum4(((w ^ 2) & 2) + 1, cb);
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      c3[f] = ca[f] + cb[f];
      n2[f] = right3 ? c3[f] + n3[f] : n3[f] + c3[f];
    }
    const double* const k0 = right3 ? c3 : n3;
    const double* const k1 = right3 ? n3 : c3;
    double x0, x1;
    dev::solve_bounded_lean<PREFER>(cst[0], 0.0, rh + 3 + d2, n2[0], n2[1], n2[2], xin, k0[0],
                                    k0[1], k0[2], k1[0], k1[1], k1[2], x0, x1);
    x3 = right3 ? x1 : x0;
  }
  double x4;
  { // Depth 3, solved by both warps below it.
    const int e = w >> 1;
    const double* const k0 = right4 ? sib : n4;
    const double* const k1 = right4 ? n4 : sib;
    double x0, x1;
    dev::solve_bounded_lean<PREFER>(cst[1 + e], 0.0, rh + 7 + 2*d2 + e, n3[0], n3[1], n3[2], x3,
                                    k0[0], k0[1], k0[2], k1[0], k1[1], k1[2], x0, x1);
    x4 = right4 ? x1 : x0;
  }
  const int m = 4*d2 + w;     // this thread's depth-4 node
  double x5[2];
  dev::solve_bounded_lean<PREFER>(cst[3 + w], 0.0, rh + 15 + m, n4[0], n4[1], n4[2], x4, n5[0][0],
                                  n5[0][1], n5[0][2], n5[1][0], n5[1][1], n5[1][2], x5[0], x5[1]);
  double* const xo = ta.x7 + (static_cast<long long>(tmine)*a.nblocks + b)*128 + 8*m;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    double x6[2];
    dev::solve_bounded_lean<PREFER>(cst[7 + 2*w + h], 0.0, rh + 31 + 2*m + h, n5[h][0], n5[h][1],
                                    n5[h][2], x5[h], n6[2*h][0], n6[2*h][1], n6[2*h][2],
                                    n6[2*h + 1][0], n6[2*h + 1][1], n6[2*h + 1][2], x6[0], x6[1]);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int q = 2*h + j;
      const double* const l = s + 8*w + 2*q;
      double x0, x1;
      dev::solve_bounded_lean<PREFER>(cst[15 + 4*w + q], 0.0, rh + 63 + 4*m + q, n6[q][0],
                                      n6[q][1], n6[q][2], x6[j], l[0], l[kTRow], l[2*kTRow],
                                      l[1], l[kTRow + 1], l[2*kTRow + 1], x0, x1);
      *reinterpret_cast<double2*>(xo + 2*q) = make_double2(x0, x1);
    }
  }
}

// ----------------------------------------------------------------------- DOWN
//
// Shared memory, in doubles: stage[2][3][sbuf] leaf rows, x7s[2][128] depth-7 masses,
// d9x[512] solved masses of the depth-9 pairs; then 2 mbarriers.
inline size_t down3_smem_bytes (const int sbuf) {
  return sizeof(double)*(6*static_cast<size_t>(sbuf) + 2*128 + kD9) + 16;
}

// (Measured at ne120 x 1280 tracers, 0.95 ms as is: 5 CTAs/SM at 96-102 registers 1.07-1.23 ms
// even with the depth-9 sums re-formed instead of kept; node constants fetched at their use
// 0.98 ms.)
template <int CLS>
__global__ void __launch_bounds__(kLeafThreads, 4)
down3_kernel (const FastArgs a, const TArgs ta) {
  static_assert(CLS == CLS_ST || CLS == CLS_CST, "fast down-sweep: st / cst only");
  extern __shared__ __align__(16) unsigned char smraw[];
  const int sbuf = a.sbuf;
  double* const stage = reinterpret_cast<double*>(smraw);            // [2][3][sbuf]
  double* const x7s = stage + 6*sbuf;                                // [2][128]
  double* const d9x = x7s + 2*128;                                   // [4][128]
  uint64_t* const mbar = reinterpret_cast<uint64_t*>(d9x + kD9);     // [2]
  constexpr int BAR_LEAF = 1;

  const int b = blockIdx.x % a.nblocks, grp = blockIdx.x / a.nblocks;
  const BlockDev B = a.blocks[b];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int src0 = B.leaf0 & ~1, shift = B.leaf0 - src0;
  const unsigned bytes = 8u*static_cast<unsigned>(((B.leaf0 + B.nl + 1) & ~1) - src0);
  const int g0 = grp*a.group;
  const int gn = min(a.group, a.ntr - g0);
  const unsigned short* const dtab = a.dtab + B.ftab_off;
  const dev::NodeWQ* const wq = a.wq + B.fbase;
  const dev::NodeRh* const rh = a.rh + B.fbase;
  const double* const rqv = a.rq + B.fbase;
  const bool prefer = a.prefer_mass_con != 0;
  auto solve = [&] (const dev::NodeWQ& c, const double rq, const int cpos, const double* nd,
                    const double bm, const double* k0, const double* k1, double& x0,
                    double& x1) {
    if (prefer)
      dev::solve_bounded_lean<true>(c, rq, rh + cpos, nd[0], nd[1], nd[2], bm, k0[0], k0[1],
                                    k0[2], k1[0], k1[1], k1[2], x0, x1);
    else
      dev::solve_bounded_lean<false>(c, rq, rh + cpos, nd[0], nd[1], nd[2], bm, k0[0], k0[1],
                                     k0[2], k1[0], k1[1], k1[2], x0, x1);
  };

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // This thread's depth-7 node: dealt out so that the leaf offsets of a half-warp's 16
  // nodes are (nearly) distinct mod 16, the 8-byte shared-memory banks.
  const int node = a.perm[B.fperm_off + tid];
  const ushort4 e = reinterpret_cast<const ushort4*>(dtab)[node];
  const int off[4] = {(e.x & 0x7fff) + shift, (e.y & 0x7fff) + shift,
                      (e.z & 0x7fff) + shift, (e.w & 0x7fff) + shift};
  const bool pr[4] = {(e.x >> 15) != 0, (e.y >> 15) != 0, (e.z >> 15) != 0,
                      (e.w >> 15) != 0};
  const unsigned* const pent = a.pent + B.fpent_off;
  const int ps = pent[warp], pe = pent[warp + 1];
  const dev::NodeWQ c7 = wq[127 + node], c8a = wq[255 + 2*node], c8b = wq[256 + 2*node];

  auto issue = [&] (const int i) {
    const int t = a.tracers[g0 + i];
    double* dst = stage + (i & 1)*3*sbuf;
    const double* const* const ra = a.rowaddr + 4*t;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&mbar[i & 1], 3*bytes + 1024);
#pragma unroll
    for (int f = 0; f < 3; ++f) tma_load(dst + f*sbuf, ra[f] + src0, bytes, &mbar[i & 1]);
    tma_load(x7s + (i & 1)*128, ta.x7 + (static_cast<long long>(t)*a.nblocks + b)*128, 1024,
             &mbar[i & 1]);
  };
  if (tid == 0) {
    issue(0);
    if (gn > 1) issue(1);
  }

  for (int k = 0; k < gn; ++k) {
    const int t = a.tracers[g0 + k];
    double* const s = stage + (k & 1)*3*sbuf;
    double* const xout = s + sbuf;                 // solved leaves replace the Qm row
    mbar_wait(&mbar[k & 1], (k >> 1) & 1);
    const double x7 = x7s[(k & 1)*128 + node];
    // Sums of this thread's depth-9 nodes (a leaf or a pair), depth-8 and depth-7 nodes.
    double n9[4][3], n8[2][3], n7[3];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double* const r0 = s + off[q];
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        const double v0 = r0[f*sbuf];
        n9[q][f] = v0;
        if (pr[q]) n9[q][f] = v0 + r0[f*sbuf + 1];
      }
    }
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      n8[0][f] = n9[0][f] + n9[1][f];
      n8[1][f] = n9[2][f] + n9[3][f];
      n7[f] = n8[0][f] + n8[1][f];
    }
    // Refill the other stage with tracer k+1: the bulk store of tracer k-1 (issued at the
    // end of the last iteration) has read it by now, and every thread has read its x7 of
    // tracer k-1 before that iteration's leaf barrier.
    if (tid == 0 && k >= 1) {
      tma_store_wait_read();
      if (k + 1 < gn) issue(k + 1);
    }
    double x8[2];
    solve(c7, 0.0, 127 + node, n7, x7, n8[0], n8[1], x8[0], x8[1]);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      double x9[2];
      solve(hf ? c8b : c8a, 0.0, 255 + 2*node + hf, n8[hf], x8[hf], n9[2*hf], n9[2*hf + 1],
            x9[0], x9[1]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int q = 2*hf + j;
        if (pr[q]) d9x[q*128 + tid] = x9[j];
        else xout[off[q]] = x9[j];
      }
    }
    __syncwarp();
    auto solve_pair = [&] (const dev::NodeWQ& c, const double rq, const int j, const int o,
                           const int slot) {
      double k0[3], k1[3], nd[3];
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        k0[f] = s[f*sbuf + o];
        k1[f] = s[f*sbuf + o + 1];
        nd[f] = k0[f] + k1[f];
      }
      double x0, x1;
      solve(c, rq, kHeapNodes + j, nd, d9x[slot], k0, k1, x0, x1);
      xout[o] = x0;
      xout[o + 1] = x1;
    };
#pragma unroll 1
    for (int j = ps + lane; j < pe; j += 32) {
      const unsigned pe_j = pent[j];
      const int r = pe_j >> 20;
      solve_pair(wq[kHeapNodes + r], rqv[kHeapNodes + r], r, (pe_j & 0x7ff) + shift,
                 (pe_j >> 11) & 0x1ff);
    }
    // Write-back: one TMA bulk store moves the 16-byte aligned interior, thread 0 stores
    // the (at most two) edge elements; the stage is refilled in the next iteration.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    bar_sync_i<BAR_LEAF, kLeafThreads>();
    if (tid == 0) {
      double* const o = a.out + static_cast<long long>(t)*a.out_ld + B.leaf0;
      const int q0 = B.leaf0 & 1;
      const int nint = (B.nl - q0) & ~1;
      if (nint) tma_store(o + q0, xout + shift + q0, 8u*static_cast<unsigned>(nint));
      tma_store_commit();
      if (q0) o[0] = xout[shift];
      if (q0 + nint < B.nl) o[B.nl - 1] = xout[shift + B.nl - 1];
    }
  }
  if (tid == 0) tma_store_wait_read();
}

} // namespace fast
} // namespace cedr_b200

#endif
