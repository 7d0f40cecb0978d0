// Persistent run() kernel ("ring" kernel) for plans whose tier-0 blocks all have the fast
// shape (fast_kernels.cuh: recursive bisection, 513..1024 leaves, perfect to depth 9) under
// one tier-1 block: every cubed-sphere config of BASELINE.json.
//
// One cooperative launch does the whole of QLT::run (cedr_qlt.cpp:618-640) for a problem
// class, or the whole of CAAS::run (cedr_caas.cpp:258-270). The leaf data cross HBM once in
// each direction (32 B in, 8 B out per cell x tracer); the second read of the leaves, for
// the down-sweep, follows the first by a few tracers and is served by L2.
//
//   * The leaves are cut into one PIECE per CTA: a run of consecutive depth-S subtrees of
//     the tier-0 blocks ("sub-blocks"; S is chosen so that there are about 7 per CTA, which
//     balances the CTAs to ~1%). A CTA owns its piece for the whole launch, so everything
//     that depends on the tree only -- leaf offsets, node constants -- is loaded once.
//   * A UNIT is the piece x a batch of TB tracers (TB = 1 at ne120; small pieces batch
//     several tracers so that a 128-thread group still has one depth-7 node per thread).
//     Every unit makes two passes through small rings of shared-memory slots, worked on by
//     specialised warps:
//       P  (1 warp)       all TMA traffic: bulk loads of a unit's four rows (UP pass), bulk
//                         re-loads of its three rows once its tracers' flags are up (DOWN
//                         pass, L2 hits), bulk stores of finished units;
//       L  (4 warps x N)  UP: micro-subtree sums in registers, the levels up to the
//                         sub-roots by shuffles, records (and, QLT, the depth-7 sums) to
//                         global memory;
//       T  (1 warp x N)   UP: one arrival per tracer on a global counter;
//       S  (2-4 warps)    the CTA that owns tracer t (round robin) waits for all CTAs'
//                         arrivals and sweeps everything above the sub-roots
//                         (l2r_combine_kid_data, root_compute, r2l_solve_qp of the top of the
//                         tree; CAAS: the four global sums and the redistribution scalars),
//                         then raises the tracer's flag;
//       T                 DOWN: solves from the sub-roots to depth 7 (from the depth-7 sums
//                         the UP pass left in an L2-resident ring);
//       L                 DOWN: solves depths 7..9 in registers, the pairs densely over the
//                         group, results into the slot (CAAS: the clip-and-redistribute
//                         pass, cedr_caas.cpp:211-253).
//     Nothing waits on a kernel boundary: the grid-wide hand-off (a few microseconds per
//     tracer) is hidden behind the UP pass of the next tracers, whose leaves wait in L2.
//
// Node arithmetic is node_solve.cuh in the reference's tree order: results are bit-identical
// to the generic kernels and to the reference.
#ifndef CEDR_B200_RING_KERNELS_CUH
#define CEDR_B200_RING_KERNELS_CUH

#include <cstdint>

#include "fast_kernels.cuh"

namespace cedr_b200 {
namespace ring {

using fast::mbar_init;
using fast::mbar_fence_init;
using fast::mbar_expect_tx;
using fast::tma_load;
using fast::tma_store;
using fast::tma_store_commit;
using fast::smem_u32;

constexpr int kGroup = 128;      // threads of an L group = depth-7 entries of a unit
constexpr int kMaxPipes = 4;
constexpr int kMaxSlots = 16;    // per ring
constexpr int kMaxTB = 32;
constexpr int kWin = 1024;       // tracers in the shared-memory window of ktab

struct PieceDev {
  int leaf0;      // first leaf of the piece (local cell index)
  int nl;         // leaves
  int nsub;       // sub-blocks
  int sub0;       // global index of the first sub-root
  int nd7;        // depth-7 nodes = nsub << (7 - S)
  int npairs;     // depth-9 pairs
  int d7_off;     // into d7tab / d7c
  int pair_off;   // into pairtab
  int top_off;    // into topc: nd7 entries, sub-block-major, heap order within a sub-block
};

// Everything above the sub-roots, swept by the S warps of the tracer's owner CTA: the 8
// sub-roots under a "micro-root" in registers, the tree over the M micro-roots (the rest of
// the blocks' tops and the tier-1 block) level by level in shared memory.
struct TopArgs {
  int M;                    // micro-roots = nblocks << (S - 3)
  int ni, nlev;             // internal nodes / levels of the tree over the micro-roots
  const int* lvlptr;        // [nlev + 1]
  const int* kid0;          // [ni] node ids: < M micro-root, else M + internal index
  const int* kid1;
  const dev::NodeWQ* mwq;   // [ni] constants of those nodes
  const dev::NodeRh* mrh;
  const int* micro_c;       // [M] index into wq / rh of the micro-root's own node
  const int* micro_h;       // [M] its heap index within its block (kids: 2h+1, 2h+2)
};

struct Args {
  const PieceDev* pieces;   // [gridDim.x]
  const ushort4* d7tab;     // per depth-7 node: leaf offsets of its 4 depth-9 nodes | pair << 15
  const int2* d7c;          // per depth-7 node: index of its constants, of its first kid's
  const uint2* pairtab;     // per pair: off | q << 16 | node << 18, constants index
  const int* topc;          // per (sub-block, heap position): constants index
  const dev::NodeWQ* wq;    // tier-0 node constants, fast order
  const dev::NodeRh* rh;
  const double* in;
  long long in_ld;
  const int* trcr_row;
  const int* trcr_prob;
  double* out;              // QLT: [t out_ld + leaf]; CAAS: unused (in place)
  long long out_ld;
  double* rec;              // sub-root records [(4 t + f) rec_ld + sub-root]
  long long rec_ld;
  double* sol;              // solved sub-root masses [t sol_ld + sub-root]
  long long sol_ld;
  double* scal;             // CAAS: [2t] mode, [2t+1] fac
  double* n7ring;           // QLT: depth-7 sums [(unit % n7len) gridDim + cta][3][128]
  int n7len;                //      n7len >= maxlag + ndslots
  const int* tracers;       // tracer ids of the class
  // The same per class-local tracer index k, packed for the kernel's shared-memory window:
  // x = tracer id, y = its first row | conserve << 30.
  const int2* ktab;
  int ntr;
  int S;                    // sub-root depth within a block, 3..7
  int npn;                  // depth-7 nodes of the largest piece
  int npairs_max;           // pairs of the piece with the most
  int TB;                   // tracers per unit: TB npn <= 128
  int plen;                 // doubles per staged row (even, >= piece leaves + 2)
  int nuslots, ndslots;     // UP ring / DOWN ring depth
  int maxlag;               // units the UP pass may run ahead of the DOWN pass (L2 window)
  int prefer_mass_con;
  int caas_rows;            // CAAS: 3 or 4 rows per tracer
  unsigned* cnt;            // [ntr] arrivals, zero before the launch
  unsigned* flag;           // [ntr] zero before the launch
  int* status;              // nonzero: a wait gave up (results invalid)
  unsigned long long spin_limit;   // nanoseconds a wait may poll without progress
  // Debug (CEDR_B200_RING_TRACE): globaltimer stamps, [(cta U + unit) 8 + stage] and, for
  // the S warps, [gridDim U 8 + 2 k + {0, 1}]; null in production.
  unsigned long long* trace;
  unsigned long long* clk;   // debug: [64] accumulated clock64 per phase, CTA 0 (null in production)
  TopArgs top;
};

// ---- small device helpers

__device__ __forceinline__ void fence_proxy_async () {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ unsigned ld_acquire (const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Polling load: no L1 invalidation per poll (an acquire load costs a CCTL.IVALL each time);
// the poller issues one fence_acquire() once the value is what it waits for.
__device__ __forceinline__ unsigned ld_relaxed (const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acquire () {
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ void st_release (unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add (unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive (uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test (uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n"
               " mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
               " selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Named barrier over `count` threads (a multiple of 32), id in a register.
__device__ __forceinline__ void bar_sync_dyn (const int id, const int count) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all () {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// A polling loop's watchdog: true once *status is set or after spin_limit nanoseconds of
// fruitless polling -- a persistent kernel must never hang the device.
__device__ __forceinline__ unsigned long long global_ns () {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct Watchdog {
  unsigned long long spins = 0, t0 = 0;
  __device__ __forceinline__ bool expired (const Args& a) {
    if ((++spins & 31u) != 0) return false;
    if (*reinterpret_cast<volatile int*>(a.status)) return true;
    const unsigned long long now = global_ns();
    if (spins == 32u) { t0 = now; return false; }
    if (now - t0 > a.spin_limit) { atomicExch(a.status, 3); return true; }
    return false;
  }
};

#define CEDR_RING_TRACE(u, stage) do { if (a.trace && lane == 0) \
    a.trace[(static_cast<size_t>(blockIdx.x)*U + (u))*8 + (stage)] = global_ns(); } while (0)

// Debug phase clocks (CTA 0, one thread per role): CEDR_CLK_BEGIN once, CEDR_CLK(i) adds the
// clocks since the previous mark to phase i and counts it.
#define CEDR_CLK_BEGIN long long clk_t0 = clock64()
#define CEDR_CLK(i, cond) do { if (a.clk && blockIdx.x == 0 && (cond)) { const long long clk_t1 = clock64(); \
    atomicAdd(a.clk + 2*(i), static_cast<unsigned long long>(clk_t1 - clk_t0)); atomicAdd(a.clk + 2*(i) + 1, 1ull); \
    clk_t0 = clock64(); } } while (0)

// Shared-memory carve-up, identical on the host (sizes) and the device (pointers).
struct SmemLayout {
  size_t uslots, dslots, td7, tsum, tx, d9x, lscal, mt, mwq, mkid, mlvl, tconst, tgi, pconst,
    tcidx, kwin, bars, ltask, total;
  int uslot_doubles, dslot_doubles;
};

__host__ __device__ inline size_t align16 (size_t x) {
  return (x + 15) & ~static_cast<size_t>(15);
}

__host__ __device__ inline SmemLayout
smem_layout (const int TB, const int plen, const int nuslots, const int ndslots,
             const int npipes, const int M, const int mni, const int mnlev, const int npn,
             const int npairs_max, const bool caas) {
  SmemLayout l;
  size_t o = 0;
  // UP slot: rows [TB][4][plen] (min, Qm, max, prev). DOWN slot: rows [TB][3][plen], x7 [128].
  l.uslot_doubles = TB*4*plen;
  l.dslot_doubles = TB*3*plen + kGroup;
  l.uslots = o; o += sizeof(double)*static_cast<size_t>(nuslots)*l.uslot_doubles;
  l.dslots = o; o += sizeof(double)*static_cast<size_t>(ndslots)*l.dslot_doubles;
  l.td7 = o; o += caas ? 0 : sizeof(double)*npipes*3*kGroup;    // T: depth-7 sums of a unit
  l.tsum = o; o += caas ? 0 : sizeof(double)*npipes*3*kGroup;   // node z + i, kids 2(z + i), +1
  l.tx = o; o += caas ? 0 : sizeof(double)*npipes*kGroup;
  l.d9x = o; o += caas ? 0 : sizeof(double)*npipes*4*kGroup;
  l.lscal = o; o += sizeof(double)*npipes*2*kMaxTB;
  l.mt = o; o += sizeof(double)*4*2*static_cast<size_t>(M);
  l.kwin = o; o += sizeof(int2)*kWin;
  o = align16(o);
  l.mwq = o; o += caas ? 0 : sizeof(dev::NodeWQ)*static_cast<size_t>(mni);
  l.tconst = o; o += caas ? 0 : sizeof(dev::NodeWQ)*static_cast<size_t>(npn);
  l.pconst = o; o += caas ? 0 : sizeof(dev::NodeWQ)*static_cast<size_t>(npairs_max);
  l.tgi = o; o += caas ? 0 : sizeof(int)*static_cast<size_t>(npn);
  l.mlvl = o; o += sizeof(int)*static_cast<size_t>(mnlev + 1);
  l.mkid = o; o += sizeof(unsigned short)*2*static_cast<size_t>(mni);
  l.tcidx = o; o += sizeof(unsigned short)*kGroup;
  o = align16(o);
  l.bars = o; o += sizeof(uint64_t)*(3*static_cast<size_t>(nuslots) + 4*static_cast<size_t>(ndslots));
  l.ltask = o; o += sizeof(int)*(2*kMaxPipes + 2) + kMaxSlots;   // + P's freed rounds
  l.total = align16(o);
  return l;
}

// ------------------------------------------------------------------ the S warps
//
// Everything above the sub-roots for tracer index k (class-local), once every CTA's records
// have arrived. QLT: l2r_combine_kid_data (cedr_qlt.cpp:339-430), root_compute (:441-476)
// and r2l_solve_qp (:490-604) for those nodes; CAAS: the four tree-ordered global sums
// (cedr_caas.cpp:129-209 with cedr_bfb_tree_allreduce.cpp:86-124) and the scalars of
// finish_locally (:211-227). st / snt: this thread's index among the S threads / their
// number. The tree over the micro-roots (topology, constants) sits in shared memory.
struct TopSmem {
  double* mt;                  // [4][2 M]
  const dev::NodeWQ* mwq;      // [ni]
  const unsigned short* kid0;  // [ni]
  const unsigned short* kid1;
  const int* lvl;              // [nlev + 1]
};

template <int CLS>
__device__ __forceinline__ void
serve_top (const Args& a, const TopSmem& ts, const bool sum4, const int k, const int st,
           const int snt) {
  constexpr bool caas = CLS == CLS_CAAS;
  const TopArgs& T = a.top;
  const int t = a.tracers[k];
  const int M = T.M;
  double* const f0 = ts.mt;
  double* const f1 = f0 + 2*M;
  double* const f2 = f1 + 2*M;
  double* const f3 = f2 + 2*M;
  const bool prefer = a.prefer_mass_con != 0;
  const double* const rbase = a.rec + static_cast<long long>(t)*4*a.rec_ld;
  auto sbar = [&] () { if (snt > 32) bar_sync_dyn(15, snt); else __syncwarp(); };
  const int nf = sum4 ? 4 : 3;
  // One micro-root per thread keeps its 8 sub-root records in registers between the sums
  // and the solves; otherwise they are read again (L2 hits).
  const bool keep = M <= snt;

  // The 8 sub-roots of micro-root m, all fields in flight at once (written by other SMs:
  // bypass L1).
  CEDR_CLK_BEGIN;
  const bool clk_me = st == 0;
  double sub[8][4];
  auto load_sub = [&] (const int m, const int nfield) {
    double2 v[4][4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      if (f >= nfield) break;
      const double2* const r = reinterpret_cast<const double2*>(rbase + f*a.rec_ld + 8LL*m);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[f][j] = __ldcg(r + j);
    }
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      if (f >= nfield) break;
#pragma unroll
      for (int j = 0; j < 4; ++j) { sub[2*j][f] = v[f][j].x; sub[2*j + 1][f] = v[f][j].y; }
    }
  };
  for (int m = st; m < M; m += snt) {
    load_sub(m, nf);
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      if (f >= nf) break;
      const double s = ((sub[0][f] + sub[1][f]) + (sub[2][f] + sub[3][f])) +
        ((sub[4][f] + sub[5][f]) + (sub[6][f] + sub[7][f]));
      (f == 0 ? f0 : f == 1 ? f1 : f == 2 ? f2 : f3)[m] = s;
    }
  }
  sbar();
  CEDR_CLK(13, clk_me);
  if (st < 32) {
    // The tree over the micro-roots: one warp, level by level.
    const int lane = st;
    for (int l = 0; l < T.nlev; ++l) {
      const int je = ts.lvl[l + 1];
      for (int j = ts.lvl[l] + lane; j < je; j += 32) {
        const int k0 = ts.kid0[j], k1 = ts.kid1[j], me = M + j;
        f0[me] = f0[k0] + f0[k1];
        f1[me] = f1[k0] + f1[k1];
        f2[me] = f2[k0] + f2[k1];
        if (sum4) f3[me] = f3[k0] + f3[k1];
      }
      __syncwarp();
    }
    const int root = T.ni ? M + T.ni - 1 : 0;
    if (caas) {
      if (lane == 0) {
        const double clip_sum = f1[root], term_sum = f3[root];
        const double mm = term_sum - clip_sum;
        double mode = 0, fac = 0;
        if (mm < 0) {
          fac = clip_sum - f0[root];
          if (fac > 0) { fac = mm/fac; mode = -1; }
        } else if (mm > 0) {
          fac = f2[root] - clip_sum;
          if (fac > 0) { fac = mm/fac; mode = 1; }
        }
        a.scal[2*t] = mode;
        a.scal[2*t + 1] = fac;
      }
    } else {
      // root_compute: the mass to distribute is sum(Qm_prev) if conserving, else sum(Qm).
      if (lane == 0 && ! sum4) f3[root] = f1[root];
      __syncwarp();
      for (int l = T.nlev - 1; l >= 0; --l) {
        const int je = ts.lvl[l + 1];
        for (int j = ts.lvl[l] + lane; j < je; j += 32) {
          const int k0 = ts.kid0[j], k1 = ts.kid1[j], me = M + j;
          const dev::NodeWQ c = ts.mwq[j];
          double x0, x1;
          if (prefer)
            dev::solve_bounded_lean<true>(c, 0.0, T.mrh + j, f0[me], f1[me], f2[me], f3[me],
                                          f0[k0], f1[k0], f2[k0], f0[k1], f1[k1], f2[k1], x0, x1);
          else
            dev::solve_bounded_lean<false>(c, 0.0, T.mrh + j, f0[me], f1[me], f2[me], f3[me],
                                           f0[k0], f1[k0], f2[k0], f0[k1], f1[k1], f2[k1], x0, x1);
          f3[k0] = x0;
          f3[k1] = x1;
        }
        __syncwarp();
      }
    }
  }
  sbar();
  CEDR_CLK(14, clk_me);
  if ( ! caas) {
    // The three levels under every micro-root, in registers (depths S-3 .. S-1 of a block).
    for (int m = st; m < M; m += snt) {
      if ( ! keep) load_sub(m, 3);
      double s2[4][3], s1[2][3];
#pragma unroll
      for (int f = 0; f < 3; ++f) {
#pragma unroll
        for (int j = 0; j < 4; ++j) s2[j][f] = sub[2*j][f] + sub[2*j + 1][f];
        s1[0][f] = s2[0][f] + s2[1][f];
        s1[1][f] = s2[2][f] + s2[3][f];
      }
      const double s0[3] = {f0[m], f1[m], f2[m]};
      const int cb = T.micro_c[m], h = T.micro_h[m];
      const dev::NodeWQ* const wq = a.wq + (cb - h);     // the block's constants, heap order
      const dev::NodeRh* const rh = a.rh + (cb - h);
      // The seven nodes' constants, all loads in flight before the first solve.
      dev::NodeWQ c[7];
      c[0] = wq[h];
#pragma unroll
      for (int j = 0; j < 2; ++j) c[1 + j] = wq[2*h + 1 + j];
#pragma unroll
      for (int j = 0; j < 4; ++j) c[3 + j] = wq[4*h + 3 + j];
      auto solve = [&] (const dev::NodeWQ& cc, const int hh, const double* nd, const double b,
                        const double* k0, const double* k1, double& x0, double& x1) {
        if (prefer)
          dev::solve_bounded_lean<true>(cc, 0.0, rh + hh, nd[0], nd[1], nd[2], b, k0[0], k0[1],
                                        k0[2], k1[0], k1[1], k1[2], x0, x1);
        else
          dev::solve_bounded_lean<false>(cc, 0.0, rh + hh, nd[0], nd[1], nd[2], b, k0[0], k0[1],
                                         k0[2], k1[0], k1[1], k1[2], x0, x1);
      };
      double x1v[2], x2v[4], x3v[8];
      solve(c[0], h, s0, f3[m], s1[0], s1[1], x1v[0], x1v[1]);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        solve(c[1 + j], 2*h + 1 + j, s1[j], x1v[j], s2[2*j], s2[2*j + 1], x2v[2*j], x2v[2*j + 1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double k0[3] = {sub[2*j][0], sub[2*j][1], sub[2*j][2]};
        const double k1[3] = {sub[2*j + 1][0], sub[2*j + 1][1], sub[2*j + 1][2]};
        solve(c[3 + j], 4*h + 3 + j, s2[j], x2v[j], k0, k1, x3v[2*j], x3v[2*j + 1]);
      }
      double* const o = a.sol + static_cast<long long>(t)*a.sol_ld + 8LL*m;
#pragma unroll
      for (int j = 0; j < 8; j += 2)
        __stcg(reinterpret_cast<double2*>(o + j), make_double2(x3v[j], x3v[j + 1]));
    }
    sbar();
  }
  CEDR_CLK(15, clk_me);
  if (st == 0) {
    __threadfence();
    st_release(a.flag + k, 1u);
  }
  CEDR_CLK(12, clk_me);
}

// ------------------------------------------------------------------ the kernel

constexpr int block_threads (const int NP, const int SW) { return 32*(1 + SW + 5*NP); }

template <int CLS, int NP, int SW>
__global__ void __launch_bounds__(32*(1 + SW + 5*NP), 1)
run_kernel (const Args a) {
  static_assert(CLS == CLS_ST || CLS == CLS_CST || CLS == CLS_CAAS, "ring classes");
  static_assert(NP >= 1 && NP <= kMaxPipes && SW >= 1 && SW <= 4, "ring roles");
  constexpr bool caas = CLS == CLS_CAAS;
  extern __shared__ __align__(16) unsigned char smraw[];

  const PieceDev P = a.pieces[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.S, npn = a.npn, TB = a.TB, plen = a.plen;
  const int NU = a.nuslots, ND = a.ndslots;
  // CAAS stages Qm_prev only if some tracer conserves (cedr_caas.cpp:86-100).
  const bool has_prev = caas ? a.caas_rows == 4 : CLS == CLS_CST;
  const int nrows = has_prev ? 4 : 3;
  // CAAS always carries four sums (its fourth is Qm itself where a tracer does not conserve).
  const bool sum4 = caas || has_prev;
  const int W = TB*npn;                 // depth-7 entries of a unit (<= 128)
  const int Wtop = W >> (7 - S);        // sub-root entries of a unit
  const int nsubmax = npn >> (7 - S);
  const int nhp = 1 << (7 - S);         // depth-7 nodes (= heap slots, one unused) per sub-block
  const int U = (a.ntr + TB - 1)/TB;
  const int shift = P.leaf0 & 1, src0 = P.leaf0 - shift;
  const unsigned rowbytes = 8u*static_cast<unsigned>(((P.leaf0 + P.nl + 1) & ~1) - src0);

  const SmemLayout lay = smem_layout(TB, plen, NU, ND, NP, a.top.M, a.top.ni, a.top.nlev, npn,
                                     a.npairs_max, caas);
  double* const uslots = reinterpret_cast<double*>(smraw + lay.uslots);
  double* const dslots = reinterpret_cast<double*>(smraw + lay.dslots);
  dev::NodeWQ* const tconst = reinterpret_cast<dev::NodeWQ*>(smraw + lay.tconst);
  dev::NodeWQ* const pconst = reinterpret_cast<dev::NodeWQ*>(smraw + lay.pconst);
  dev::NodeWQ* const smwq = reinterpret_cast<dev::NodeWQ*>(smraw + lay.mwq);
  int* const tgi = reinterpret_cast<int*>(smraw + lay.tgi);
  int* const mlvl = reinterpret_cast<int*>(smraw + lay.mlvl);
  unsigned short* const mkid = reinterpret_cast<unsigned short*>(smraw + lay.mkid);
  unsigned short* const tcidx = reinterpret_cast<unsigned short*>(smraw + lay.tcidx);
  uint64_t* const ufull = reinterpret_cast<uint64_t*>(smraw + lay.bars);   // UP rows landed
  uint64_t* const uread = ufull + NU;     // L's UP is done with the slot (128 arrivals)
  uint64_t* const arr = uread + NU;       // T has counted the unit's arrivals (1)
  uint64_t* const dassign = arr + NU;     // DOWN slot given to a unit, flags up (1)
  uint64_t* const dfull = dassign + ND;   // DOWN rows landed
  uint64_t* const topd = dfull + ND;      // T's DOWN done: x7 in the slot (32)
  uint64_t* const done = topd + ND;       // L's DOWN done: results in the slot (128)
  int* const ltask = reinterpret_cast<int*>(smraw + lay.ltask);
  // Window over ktab: entry k at k % kWin, filled by the P warp ahead of every use.
  const int2* const kwin = reinterpret_cast<const int2*>(smraw + lay.kwin);
  auto kent = [&] (const int k) { return kwin[k & (kWin - 1)]; };
  auto uslot = [&] (const int s) { return uslots + static_cast<size_t>(s)*lay.uslot_doubles; };
  auto dslot = [&] (const int s) { return dslots + static_cast<size_t>(s)*lay.dslot_doubles; };
  auto dslot_x7 = [&] (const int s) { return dslot(s) + TB*3*plen; };
  auto unit_ntr = [&] (const int u) { return min(TB, a.ntr - u*TB); };
  auto n7slot = [&] (const int u) {
    return a.n7ring + (static_cast<size_t>(u % a.n7len)*gridDim.x + blockIdx.x)*3*kGroup;
  };

  // ---- one-time setup
  if (tid == 0) {
    for (int s = 0; s < NU; ++s) {
      mbar_init(&ufull[s], 1);
      mbar_init(&uread[s], kGroup);
      mbar_init(&arr[s], 1);
    }
    for (int s = 0; s < ND; ++s) {
      mbar_init(&dassign[s], 1);
      mbar_init(&dfull[s], 1);
      mbar_init(&topd[s], 32);
      mbar_init(&done[s], kGroup);
    }
    mbar_fence_init();
  }
  for (int j = tid; j < a.top.ni; j += blockDim.x) {
    mkid[j] = static_cast<unsigned short>(a.top.kid0[j]);
    mkid[a.top.ni + j] = static_cast<unsigned short>(a.top.kid1[j]);
    if ( ! caas) smwq[j] = a.top.mwq[j];
  }
  for (int j = tid; j <= a.top.nlev; j += blockDim.x) mlvl[j] = a.top.lvlptr[j];
  if ( ! caas) {
    for (int i = tid; i < npn; i += blockDim.x) {
      const int g = i < P.nd7 ? a.topc[P.top_off + i] : -1;
      tgi[i] = g >= 0 ? g : 0;
      dev::NodeWQ one;
      one.w0 = one.w1 = one.q0 = one.q1 = 1.0;
      tconst[i] = g >= 0 ? a.wq[g] : one;
    }
    for (int r = tid; r < P.npairs; r += blockDim.x)
      pconst[r] = a.wq[a.pairtab[P.pair_off + r].y];
    // Node z + i of the sums above depth 7 (level of z entries, z = Wtop 2^l) -> its
    // constants: entry i is (tracer j, sub-block s', position p), i = (j nsubmax + s') 2^l + p.
    for (int node = Wtop + tid; node < W; node += blockDim.x) {
      int l = 0;
      while ((Wtop << (l + 1)) <= node) ++l;
      const int i = node - (Wtop << l);
      const int js = i >> l, p = i & ((1 << l) - 1);
      tcidx[node] = static_cast<unsigned short>((js % nsubmax)*nhp + (1 << l) - 1 + p);
    }
  }
  __syncthreads();

  // =================================================================== P warp
  if (warp == 0) {
    // All 32 lanes run the loop (its decisions are warp-uniform); lane j moves tracer j of
    // a unit.
    int ul = 0, dl = 0, nstored = 0;
    int us[kMaxPipes];
#pragma unroll
    for (int p = 0; p < kMaxPipes; ++p) us[p] = p;
    // freed[s]: how many units of DOWN slot s have been stored (mod 256); unit u may be
    // re-loaded into it once that equals u / ND.
    unsigned char* const freed = reinterpret_cast<unsigned char*>(ltask + 2*kMaxPipes + 2);
    if (lane == 0) for (int s = 0; s < ND; ++s) freed[s] = 0;
    __syncwarp();
    int pending_slot = -1;
    unsigned flagv = 0;
    int flag_unit = -1;
    // ktab window: entries [0, kfill) are in shared memory (those of stored units may be
    // overwritten); `pref` holds the next 32, loaded a round early so that the global
    // latency is never waited for.
    int2* const kw = reinterpret_cast<int2*>(smraw + lay.kwin);
    int kfill = 0;
    int2 pref = make_int2(0, 0);
    if (lane < a.ntr) pref = a.ktab[lane];
    Watchdog wd;
    CEDR_CLK_BEGIN;
    const bool clk_me = lane == 0;
    while (nstored < U) {
      bool did = false;
      CEDR_CLK(23, clk_me);
      {
        // Keep the window a little ahead of the UP pass, never over entries still in use
        // (every unit below min(us[]) has been stored).
        int ulo = us[0];
#pragma unroll
        for (int p = 1; p < NP; ++p) ulo = min(ulo, us[p]);
        if (kfill < a.ntr && kfill < (ul + 2)*TB + 64 && kfill + 32 <= ulo*TB + kWin) {
          kw[(kfill + lane) & (kWin - 1)] = pref;
          kfill += 32;
          if (kfill + lane < a.ntr) pref = a.ktab[kfill + lane];
          __syncwarp();
        }
      }
      // Stores first: they free DOWN slots.
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int u = us[p];
        if (u >= U) continue;
        const int s = u % ND;
        if ( ! mbar_test(&done[s], (u/ND) & 1)) continue;
        const int nt = unit_ntr(u);
        CEDR_RING_TRACE(u, 7);
        if (lane < nt) {
          const int2 ke = kent(u*TB + lane);
          double* const o = caas ?
            const_cast<double*>(a.in) + (static_cast<long long>(ke.y & 0x3fffffff) + 1)*a.in_ld + P.leaf0 :
            a.out + static_cast<long long>(ke.x)*a.out_ld + P.leaf0;
          const double* const x = dslot(s) + (lane*3 + 1)*plen + shift;
          const int q0 = shift, nint = (P.nl - q0) & ~1;
          if (nint) tma_store(o + q0, x + q0, 8u*static_cast<unsigned>(nint));
          tma_store_commit();
          if (q0) o[0] = x[0];
          if (q0 + nint < P.nl) o[P.nl - 1] = x[P.nl - 1];
        }
        // The slot is free once the bulk store has READ it. Waiting for that here would
        // stall the loads behind the store's start-up latency; instead wait for all but the
        // newest store (long done) and free the slot of the store before this one.
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        if (lane == 0 && pending_slot >= 0)
          freed[pending_slot] = static_cast<unsigned char>(freed[pending_slot] + 1);
        pending_slot = s;
        __syncwarp();
        us[p] = u + NP;
        ++nstored;
        did = true;
        CEDR_CLK(20, clk_me);
      }
      // The DOWN pass of the next unit: its tracers' flags are up and a slot is free.
      if (dl < U) {
        const int s = dl % ND, nt = unit_ntr(dl);
        // The flags were read a round ago (flagv); a poll that is not on the critical path
        // costs nothing. No acquire here: this warp only moves the rows; the T warps and
        // the L groups acquire the flags before they read what the S warps wrote.
        if (flag_unit != dl) {
          flagv = lane < nt ? ld_relaxed(a.flag + dl*TB + lane) : 1u;
          flag_unit = dl;
        }
        const bool up_all = __all_sync(0xffffffffu, flagv != 0);
        if (up_all && freed[s] == static_cast<unsigned char>(dl/ND)) {
          CEDR_RING_TRACE(dl, 4);
          if (lane == 0) {
            mbar_arrive(&dassign[s]);
            mbar_expect_tx(&dfull[s], static_cast<unsigned>(nt*3)*rowbytes);
          }
          __syncwarp();
          // lane = 4 (tracer of the unit) + row
          for (int j = lane >> 2; j < nt; j += 8) {
            const int f = lane & 3;
            if (f < 3)
              tma_load(dslot(s) + (j*3 + f)*plen,
                       a.in + (static_cast<long long>(kent(dl*TB + j).y & 0x3fffffff) + f)*a.in_ld + src0,
                       rowbytes, &dfull[s]);
          }
          ++dl;
          did = true;
          // Next unit's flags: in flight until the next round.
          if (dl < U) {
            flagv = lane < unit_ntr(dl) ? ld_relaxed(a.flag + dl*TB + lane) : 1u;
            flag_unit = dl;
          }
          CEDR_CLK(21, clk_me);
        } else if ( ! up_all) {
          flag_unit = -1;    // poll again next round
        }
      }
      // The UP pass of the next unit, at most maxlag units ahead of the DOWN pass (the
      // leaves in between wait in L2).
      if (ul < U && ul < dl + a.maxlag && min((ul + 1)*TB, a.ntr) <= kfill) {
        const int s = ul % NU, r = ul/NU;
        if (r == 0 || mbar_test(&arr[s], (r - 1) & 1)) {
          const int nt = unit_ntr(ul);
          CEDR_RING_TRACE(ul, 0);
          if (lane == 0) {
            mbar_expect_tx(&ufull[s], static_cast<unsigned>(nt*nrows)*rowbytes);
          }
          __syncwarp();
          for (int j = lane >> 2; j < nt; j += 8) {
            const int f = lane & 3;
            if (f < nrows)
              tma_load(uslot(s) + (j*4 + f)*plen,
                       a.in + (static_cast<long long>(kent(ul*TB + j).y & 0x3fffffff) + f)*a.in_ld + src0,
                       rowbytes, &ufull[s]);
          }
          ++ul;
          did = true;
          CEDR_CLK(22, clk_me);
        }
      }
      if (did) { wd.spins = 0; continue; }
      if (wd.expired(a)) break;
      __nanosleep(100);   // (polling warps must not take the working warps' issue slots)
    }
    tma_store_wait_all();
    return;
  }

  // =================================================================== S warps
  if (warp <= SW) {
    const int st = tid - 32, snt = 32*SW;
    TopSmem ts;
    ts.mt = reinterpret_cast<double*>(smraw + lay.mt);
    ts.mwq = smwq;
    ts.kid0 = mkid;
    ts.kid1 = mkid + a.top.ni;
    ts.lvl = mlvl;
    const unsigned need = gridDim.x;
    for (int k = blockIdx.x; k < a.ntr; k += gridDim.x) {
      bool ok = true;
      if (st == 0) {
        Watchdog wd;
        while (ld_relaxed(a.cnt + k) < need) {
          if (wd.expired(a)) { ok = false; break; }
          __nanosleep(60);
        }
        fence_acquire();
        if (SW > 1) ltask[2*kMaxPipes] = ok;   // (past the L groups' entries)
      }
      if (SW > 1) {
        bar_sync_dyn(15, snt);
        ok = ltask[2*kMaxPipes] != 0;
        bar_sync_dyn(15, snt);
      } else {
        ok = __shfl_sync(0xffffffffu, ok, 0);
      }
      if ( ! ok) break;
      if (a.trace && st == 0) a.trace[static_cast<size_t>(gridDim.x)*U*8 + 2*k] = global_ns();
      serve_top<CLS>(a, ts, sum4, k, st, snt);
      if (a.trace && st == 0) a.trace[static_cast<size_t>(gridDim.x)*U*8 + 2*k + 1] = global_ns();
    }
    return;
  }

  // =================================================================== T warps
  if (warp < 1 + SW + NP) {
    const int p = warp - 1 - SW;
    double* const d7s = reinterpret_cast<double*>(smraw + lay.td7) + p*3*kGroup;
    double* const tsum = reinterpret_cast<double*>(smraw + lay.tsum) + p*3*kGroup;
    double* const tx = reinterpret_cast<double*>(smraw + lay.tx) + p*kGroup;
    const bool prefer = a.prefer_mass_con != 0;
    int ta = p, td = p;
    Watchdog wd;
    CEDR_CLK_BEGIN;
    const bool clk_me = lane == 0 && p == 0;
    for (;;) {
      CEDR_CLK(18, clk_me);
      // DOWN of the next unit once P has given it a slot (its tracers' flags are up).
      if ( ! caas && td < U && mbar_test(&dassign[td % ND], (td/ND) & 1)) {
        const int s = td % ND, nt = unit_ntr(td);
        if (lane < nt) ld_acquire(a.flag + td*TB + lane);   // (acquire what the S warps wrote)
        double* const x7 = dslot_x7(s);
        double* const xtop = (Wtop == W) ? x7 : tx + Wtop;
        // Depth-7 sums of the UP pass (L2), the sub-roots' solved masses.
        {
          const double* const g7 = n7slot(td);
          for (int i = lane; i < 3*kGroup; i += 32) d7s[i] = __ldcg(g7 + i);
        }
        for (int i = lane; i < Wtop; i += 32) {
          const int j = i/nsubmax, sb = i - j*nsubmax;
          // (Unused entries -- a short piece, the last unit's missing tracers -- solve the
          // all-zero problem: its quick exit.)
          xtop[i] = (j < nt && sb < P.nsub) ?
            __ldcg(a.sol + static_cast<long long>(kent(td*TB + j).x)*a.sol_ld + P.sub0 + sb) : 0.0;
        }
        __syncwarp();
        CEDR_CLK(9, clk_me);
        // Sums of the levels above depth 7: node z + i of the level with z entries is
        // entries 2i, 2i + 1 of the level below.
        for (int z = W >> 1; z >= Wtop; z >>= 1) {
#pragma unroll
          for (int f = 0; f < 3; ++f) {
            const double* const src = (2*z == W) ? d7s + f*kGroup : tsum + f*kGroup + 2*z;
            for (int i = lane; i < z; i += 32) {
              const double2 v = reinterpret_cast<const double2*>(src)[i];
              tsum[f*kGroup + z + i] = v.x + v.y;
            }
          }
          __syncwarp();
        }
        CEDR_CLK(10, clk_me);
        for (int z = Wtop; z < W; z <<= 1) {
          for (int i = lane; i < z; i += 32) {
            const int node = z + i;
            double k0[3], k1[3], nd[3];
#pragma unroll
            for (int f = 0; f < 3; ++f) {
              const double* const src = (2*z == W) ? d7s + f*kGroup : tsum + f*kGroup + 2*z;
              const double2 v = reinterpret_cast<const double2*>(src)[i];
              k0[f] = v.x; k1[f] = v.y;
              nd[f] = tsum[f*kGroup + node];
            }
            const int ci = tcidx[node];
            const dev::NodeWQ c = tconst[ci];
            double x0, x1;
            if (prefer)
              dev::solve_bounded_lean<true>(c, 0.0, a.rh + tgi[ci], nd[0], nd[1], nd[2], tx[node],
                                            k0[0], k0[1], k0[2], k1[0], k1[1], k1[2], x0, x1);
            else
              dev::solve_bounded_lean<false>(c, 0.0, a.rh + tgi[ci], nd[0], nd[1], nd[2], tx[node],
                                             k0[0], k0[1], k0[2], k1[0], k1[1], k1[2], x0, x1);
            double* const xo = (2*z == W) ? x7 : tx + 2*z;
            reinterpret_cast<double2*>(xo)[i] = make_double2(x0, x1);
          }
          __syncwarp();
        }
        mbar_arrive(&topd[s]);
        CEDR_CLK(11, clk_me);
        CEDR_RING_TRACE(td, 5);
        td += NP;
        wd.spins = 0;
        continue;
      }
      // Arrivals of the next unit whose UP the L group has finished.
      if (ta < U && mbar_test(&uread[ta % NU], (ta/NU) & 1)) {
        const int s = ta % NU, nt = unit_ntr(ta);
        if (lane < nt) {
          __threadfence();
          red_release_add(a.cnt + ta*TB + lane, 1u);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&arr[s]);
        CEDR_CLK(8, clk_me);
        CEDR_RING_TRACE(ta, 3);
        ta += NP;
        wd.spins = 0;
        continue;
      }
      if (ta >= U && (caas || td >= U)) break;
      if (wd.expired(a)) break;
      __nanosleep(100);
    }
    return;
  }

  // =================================================================== L groups
  const int p = (warp - (1 + SW + NP)) >> 2;
  const int gt = tid - 32*(1 + SW + NP + 4*p), gw = gt >> 5;
  const int barid = 1 + p;
  const int j_of = gt/npn, n_of = gt - j_of*npn;
  const bool node_ok = gt < W && n_of < P.nd7;
  double* const d9x = reinterpret_cast<double*>(smraw + lay.d9x) + p*4*kGroup;
  double* const lscal = reinterpret_cast<double*>(smraw + lay.lscal) + p*2*kMaxTB;
  const bool prefer = a.prefer_mass_con != 0;

  int off[4] = {0, 0, 0, 0};
  bool pr[4] = {false, false, false, false};
  dev::NodeWQ c7, c8a, c8b;
  int g7 = 0, g8 = 0;
  c7.w0 = c7.w1 = c7.q0 = c7.q1 = 0;
  c8a = c7; c8b = c7;
  if (node_ok) {
    const ushort4 e = a.d7tab[P.d7_off + n_of];
    off[0] = (e.x & 0x7fff) + shift; off[1] = (e.y & 0x7fff) + shift;
    off[2] = (e.z & 0x7fff) + shift; off[3] = (e.w & 0x7fff) + shift;
    pr[0] = (e.x >> 15) != 0; pr[1] = (e.y >> 15) != 0;
    pr[2] = (e.z >> 15) != 0; pr[3] = (e.w >> 15) != 0;
    if ( ! caas) {
      const int2 ci = a.d7c[P.d7_off + n_of];
      g7 = ci.x; g8 = ci.y;
      c7 = a.wq[g7]; c8a = a.wq[g8]; c8b = a.wq[g8 + 1];
    }
  }
  auto solve = [&] (const dev::NodeWQ& c, const int gi, const double* nd, const double bm,
                    const double* k0, const double* k1, double& x0, double& x1) {
    if (prefer)
      dev::solve_bounded_lean<true>(c, 0.0, a.rh + gi, nd[0], nd[1], nd[2], bm, k0[0], k0[1],
                                    k0[2], k1[0], k1[1], k1[2], x0, x1);
    else
      dev::solve_bounded_lean<false>(c, 0.0, a.rh + gi, nd[0], nd[1], nd[2], bm, k0[0], k0[1],
                                     k0[2], k1[0], k1[1], k1[2], x0, x1);
  };

  int up = p, dn = p, it = 0;
  CEDR_CLK_BEGIN;
  const bool clk_me = gt == 0 && p == 0;
  for (;;) {
    CEDR_CLK(7, clk_me);
    // ---- what next: the DOWN of the oldest unit if its rows (and, QLT, its depth-7
    // masses) are there, else the UP of the next unit if it has landed.
    if (gw == 0) {
      int task = 0;
      Watchdog wd;
      for (;;) {
        if (dn >= U) break;
        {
          const int sd = dn % ND;
          const unsigned par = (dn/ND) & 1;
          if (mbar_test(&dfull[sd], par) && (caas || mbar_test(&topd[sd], par))) {
            if (caas) {
              const int nt = unit_ntr(dn);
              if (lane < nt) {
                ld_acquire(a.flag + dn*TB + lane);   // (acquire what the S warps wrote)
                const int t = kent(dn*TB + lane).x;
                lscal[2*lane] = __ldcg(a.scal + 2*t);
                lscal[2*lane + 1] = __ldcg(a.scal + 2*t + 1);
              }
            }
            task = 1;
            break;
          }
        }
        if (up < U && mbar_test(&ufull[up % NU], (up/NU) & 1)) { task = 2; break; }
        const bool ex = wd.expired(a);
        if (__any_sync(0xffffffffu, ex)) break;
        __nanosleep(60);
      }
      if (lane == 0) ltask[2*p + (it & 1)] = task;
    }
    CEDR_CLK(0, clk_me);
    bar_sync_dyn(barid, kGroup);
    CEDR_CLK(1, clk_me);
    const int task = ltask[2*p + (it & 1)];
    ++it;
    if (task == 0) break;

    if (task == 2) {
      // ------------------------------------------------------------------ UP(up)
      const int u = up, s = u % NU, nt = unit_ntr(u);
      up += NP;
      mbar_test(&ufull[s], (u/NU) & 1);   // (acquire by every thread; it has completed)
      if (gw == 0) CEDR_RING_TRACE(u, 1);
      const bool act = node_ok && j_of < nt;
      double r[4] = {0, 0, 0, 0};    // this thread's depth-7 node: min, Qm | clip, max, prev | term
      if (act) {
        const double* const r0 = uslot(s) + j_of*4*plen;
        bool conserve = true;
        if (caas) conserve = (kent(u*TB + j_of).y >> 30) != 0;
        double n[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          double v[2][4];
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            if (jj == 1 && ! pr[k]) break;
            const int o = off[k] + jj;
            if (caas) {
              const double lo = r0[o], q = r0[plen + o], hi = r0[2*plen + o];
              const double term = (has_prev && conserve) ? r0[3*plen + o] : q;
              const double clip = dev::rmin(hi, dev::rmax(lo, q));
              v[jj][0] = 0.0 + lo; v[jj][1] = 0.0 + clip; v[jj][2] = 0.0 + hi;
              v[jj][3] = 0.0 + term;
            } else {
              v[jj][0] = r0[o]; v[jj][1] = r0[plen + o]; v[jj][2] = r0[2*plen + o];
              if (has_prev) v[jj][3] = r0[3*plen + o];
            }
          }
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            if (f == 3 && ! sum4) continue;
            n[k][f] = pr[k] ? v[0][f] + v[1][f] : v[0][f];
          }
        }
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          if (f == 3 && ! sum4) continue;
          r[f] = (n[0][f] + n[1][f]) + (n[2][f] + n[3][f]);
        }
      }
      CEDR_CLK(2, clk_me);
      // QLT: the depth-7 sums for the T warp's DOWN (unused entries: zero, see there).
      if ( ! caas) {
        double* const g7o = n7slot(u);
#pragma unroll
        for (int f = 0; f < 3; ++f) __stcg(g7o + f*kGroup + gt, r[f]);
      }
      // Up to the sub-roots inside the sub-block's lanes (kids 2i, 2i + 1: left + right).
      for (int Lv = 1; Lv < nhp; Lv <<= 1) {
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          if (f == 3 && ! sum4) continue;
          r[f] = r[f] + __shfl_down_sync(0xffffffffu, r[f], Lv);
        }
      }
      if (act && (n_of & (nhp - 1)) == 0) {
        double* const rec = a.rec + static_cast<long long>(kent(u*TB + j_of).x)*4*a.rec_ld +
          P.sub0 + (n_of >> (7 - S));
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          if (f == 3 && ! sum4) continue;
          rec[f*a.rec_ld] = r[f];
        }
      }
      // The slot has been read and this thread's records are written: the T warp counts the
      // unit's arrivals once all 128 threads are here (its gpu-scope fence then publishes
      // their global stores).
      CEDR_CLK(3, clk_me);
      mbar_arrive(&uread[s]);
      CEDR_CLK(4, clk_me);
      if (gw == 0) CEDR_RING_TRACE(u, 2);
      continue;
    }

    // -------------------------------------------------------------------- DOWN(dn)
    const int u = dn, s = u % ND, nt = unit_ntr(u);
    dn += NP;
    double* const rows = dslot(s);
    mbar_test(&dfull[s], (u/ND) & 1);
    if (caas) {
      // CAAS::finish_locally, cedr_caas.cpp:211-253, on the clipped values
      // (reduce_locally stores the clip in place, :177).
      for (int j = 0; j < nt; ++j) {
        const double mode = lscal[2*j], fac = lscal[2*j + 1];
        double* const d = rows + j*3*plen + shift;
        for (int k = gt; k < P.nl; k += kGroup) {
          const double lo = d[k], hi = d[2*plen + k];
          double q = dev::rmin(hi, dev::rmax(lo, d[plen + k]));
          if (mode < 0) {
            q += fac*(q - lo);
            q = dev::rmax(lo, q);
          } else if (mode > 0) {
            q += fac*(hi - q);
            q = dev::rmin(hi, q);
          }
          d[plen + k] = q;
        }
      }
      CEDR_CLK(5, clk_me);
      fence_proxy_async();
      mbar_arrive(&done[s]);
      CEDR_CLK(6, clk_me);
      if (gw == 0) CEDR_RING_TRACE(u, 6);
      continue;
    }
    mbar_test(&topd[s], (u/ND) & 1);
    const bool act = node_ok && j_of < nt;
    if (act) {
      double* const r0 = rows + j_of*3*plen;
      double* const xout = r0 + plen;               // solved leaves replace the Qm row
      double n9[4][3], n8[2][3], n7[3];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double* const rr = r0 + off[q];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          const double v0 = rr[f*plen];
          n9[q][f] = v0;
          if (pr[q]) n9[q][f] = v0 + rr[f*plen + 1];
        }
      }
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        n8[0][f] = n9[0][f] + n9[1][f];
        n8[1][f] = n9[2][f] + n9[3][f];
        n7[f] = n8[0][f] + n8[1][f];
      }
      const double x7 = dslot_x7(s)[gt];
      double x8[2];
      solve(c7, g7, n7, x7, n8[0], n8[1], x8[0], x8[1]);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        double x9[2];
        solve(hf ? c8b : c8a, g8 + hf, n8[hf], x8[hf], n9[2*hf], n9[2*hf + 1], x9[0], x9[1]);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int q = 2*hf + jj;
          if (pr[q]) d9x[q*kGroup + gt] = x9[jj];
          else xout[off[q]] = x9[jj];
        }
      }
    }
    CEDR_CLK(5, clk_me);
    bar_sync_dyn(barid, kGroup);
    CEDR_CLK(16, clk_me);
    // The unit's depth-9 pairs, densely over the group.
    {
      const int np = P.npairs, tot = nt*np;
      for (int q = gt; q < tot; q += kGroup) {
        const int j = (TB == 1) ? 0 : q/np, r = q - j*np;
        const uint2 e = a.pairtab[P.pair_off + r];
        const int o = (e.x & 0xffff) + shift, qq = (e.x >> 16) & 3, nn = e.x >> 18;
        double* const r0 = rows + j*3*plen;
        double k0[3], k1[3], nd[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          k0[f] = r0[f*plen + o];
          k1[f] = r0[f*plen + o + 1];
          nd[f] = k0[f] + k1[f];
        }
        double x0, x1;
        solve(pconst[r], static_cast<int>(e.y), nd, d9x[qq*kGroup + j*npn + nn], k0, k1, x0, x1);
        r0[plen + o] = x0;
        r0[plen + o + 1] = x1;
      }
    }
    CEDR_CLK(17, clk_me);
    fence_proxy_async();
    mbar_arrive(&done[s]);
    CEDR_CLK(6, clk_me);
    if (gw == 0) CEDR_RING_TRACE(u, 6);
  }
}

// Constants of the tree over the micro-roots, gathered after the rhom sweep: entry j comes
// from the tier-0 fast-order arrays (src >= 0) or from the tier-1 block's NodeConst
// (src = -1 - index).
__global__ void __launch_bounds__(256)
gather_kernel (const int* src, const int n, const dev::NodeWQ* fwq, const dev::NodeRh* frh,
               const dev::NodeConst* nc, dev::NodeWQ* mwq, dev::NodeRh* mrh) {
  const int j = blockIdx.x*blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int s = src[j];
  if (s >= 0) {
    mwq[j] = fwq[s];
    mrh[j] = frh[s];
  } else {
    const dev::NodeConst c = nc[-1 - s];
    dev::NodeWQ w;
    w.w0 = c.w0; w.w1 = c.w1; w.q0 = c.q0; w.q1 = c.q1;
    mwq[j] = w;
    dev::NodeRh r;
    r.rh0 = c.rh0; r.rh1 = c.rh1;
    mrh[j] = r;
  }
}

} // namespace ring
} // namespace cedr_b200

#endif
