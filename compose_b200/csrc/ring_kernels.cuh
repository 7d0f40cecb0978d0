// Persistent single-read run() kernel ("ring" kernel) for plans whose tier-0 blocks all have
// the fast shape (fast_kernels.cuh: recursive bisection, 513..1024 leaves, perfect to depth
// 9) under one tier-1 block: every cubed-sphere config of BASELINE.json.
//
// One cooperative launch does the whole of QLT::run (cedr_qlt.cpp:618-640) for a problem
// class, or the whole of CAAS::run (cedr_caas.cpp:258-270), and the leaf data cross HBM
// exactly once in each direction (32 B in, 8 B out per cell x tracer):
//
//   * The leaves are cut into one PIECE per CTA: a run of consecutive depth-S subtrees of
//     the tier-0 blocks ("sub-blocks"; S is chosen so that there are about 7 per CTA, which
//     balances the CTAs to ~1%). A CTA owns its piece for the whole launch, so everything
//     that depends on the tree only -- leaf offsets, node constants -- is loaded once.
//   * A UNIT is the piece x a batch of TB tracers (TB = 1 at ne120; small pieces batch
//     several tracers so that a 128-thread group still has one depth-7 node per thread).
//     The units flow through a ring of shared-memory slots, worked on by specialised warps:
//       P  (1 warp)       TMA bulk loads of a unit's rows into a free slot; TMA bulk stores
//                         of finished units;
//       L  (4 warps x N)  UP: micro-subtree sums in registers, depth-7 sums into the slot;
//       T  (1 warp x N)   UP: sums up to the sub-roots, records to global memory, one
//                         arrival per tracer on a global counter;
//       S  (1-2 warps)    the CTA that owns tracer t (round robin) waits for all CTAs'
//                         arrivals and sweeps everything above the sub-roots
//                         (l2r_combine_kid_data, root_compute, r2l_solve_qp of the top of the
//                         tree; CAAS: the four global sums and the redistribution scalars),
//                         then raises the tracer's flag;
//       T                 DOWN: waits for the flag, solves from the sub-roots to depth 7;
//       L                 DOWN: solves depths 7..9 in registers, the pairs densely over the
//                         group, results into the slot (CAAS: the clip-and-redistribute
//                         pass, cedr_caas.cpp:211-253).
//     A unit stays in shared memory between its UP and its DOWN; the ring is deep enough
//     that the grid-wide hand-off (a few microseconds) is hidden behind the next units.
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
using fast::tma_store_wait_read;
using fast::smem_u32;

constexpr int kGroup = 128;      // threads of an L group = depth-7 entries of a unit
constexpr int kMaxPipes = 4;
constexpr int kMaxSlots = 24;
constexpr int kMaxTB = 32;

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
  const int* tracers;       // tracer ids of the class
  int ntr;
  int S;                    // sub-root depth within a block, 3..7
  int npn;                  // depth-7 nodes of the largest piece
  int npairs_max;           // pairs of the piece with the most
  int TB;                   // tracers per unit: TB npn <= 128
  int plen;                 // doubles per staged row (even, >= piece leaves + 2)
  int nslots, npslots;      // ring depth; Qm_prev ring depth
  int prefer_mass_con;
  int caas_rows;            // CAAS: 3 or 4 rows per tracer
  unsigned* cnt;            // [ntr] arrivals, zero before the launch
  unsigned* flag;           // [ntr] zero before the launch
  int* status;              // nonzero: a wait gave up (results invalid)
  unsigned long long spin_limit;   // nanoseconds a wait may poll without progress
  // Debug (CEDR_B200_RING_TRACE): globaltimer stamps, [(cta U + unit) 8 + stage] and, for
  // the S warps, [gridDim U 8 + 2 k + {0, 1}]; null in production.
  unsigned long long* trace;
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
    if ((++spins & 127u) != 0) return false;
    if (*reinterpret_cast<volatile int*>(a.status)) return true;
    const unsigned long long now = global_ns();
    if (spins == 128u) { t0 = now; return false; }
    if (now - t0 > a.spin_limit) { atomicExch(a.status, 3); return true; }
    return false;
  }
};

#define CEDR_RING_TRACE(u, stage) do { if (a.trace && lane == 0) \
    a.trace[(static_cast<size_t>(blockIdx.x)*U + (u))*8 + (stage)] = global_ns(); } while (0)

// Shared-memory carve-up, identical on the host (sizes) and the device (pointers).
struct SmemLayout {
  size_t slots, prevring, tsum, tx, d9x, lscal, mt, tconst, tgi, pconst, tcidx, bars, ltask,
    total;
  int slot_doubles;
};

__host__ __device__ inline size_t align16 (size_t x) {
  return (x + 15) & ~static_cast<size_t>(15);
}

__host__ __device__ inline SmemLayout
smem_layout (const int TB, const int plen, const int nslots, const int npslots,
             const int npipes, const int M, const int npn, const int npairs_max,
             const bool caas) {
  SmemLayout l;
  size_t o = 0;
  // per slot: rows [TB][3][plen], d7 [4][128] (depth-7 sums: min, Qm, max, prev), x7 [128]
  l.slot_doubles = TB*3*plen + 4*kGroup + kGroup;
  l.slots = o; o += sizeof(double)*static_cast<size_t>(nslots)*l.slot_doubles;
  l.prevring = o; o += sizeof(double)*static_cast<size_t>(npslots)*TB*plen;
  l.tsum = o; o += sizeof(double)*npipes*4*kGroup;     // node z + i, kids 2(z + i), +1
  l.tx = o; o += caas ? 0 : sizeof(double)*npipes*kGroup;
  l.d9x = o; o += caas ? 0 : sizeof(double)*npipes*4*kGroup;
  l.lscal = o; o += sizeof(double)*npipes*2*kMaxTB;
  l.mt = o; o += sizeof(double)*4*2*static_cast<size_t>(M);
  o = align16(o);
  l.tconst = o; o += caas ? 0 : sizeof(dev::NodeWQ)*static_cast<size_t>(npn);
  l.pconst = o; o += caas ? 0 : sizeof(dev::NodeWQ)*static_cast<size_t>(npairs_max);
  l.tgi = o; o += caas ? 0 : sizeof(int)*static_cast<size_t>(npn);
  l.tcidx = o; o += sizeof(unsigned short)*kGroup;
  o = align16(o);
  l.bars = o; o += sizeof(uint64_t)*(4*static_cast<size_t>(nslots) + npslots);
  l.ltask = o; o += sizeof(int)*(2*kMaxPipes + 2) + kMaxSlots;   // + P's stored_rounds
  l.total = align16(o);
  return l;
}

// ------------------------------------------------------------------ the S warps
//
// Everything above the sub-roots for tracer index k (class-local), once every CTA's records
// have arrived. QLT: l2r_combine_kid_data (cedr_qlt.cpp:339-430), root_compute (:441-476)
// and r2l_solve_qp (:490-604) for those nodes; CAAS: the four tree-ordered global sums
// (cedr_caas.cpp:129-209 with cedr_bfb_tree_allreduce.cpp:86-124) and the scalars of
// finish_locally (:211-227). st / snt: this thread's index among the S threads / their number.
template <int CLS>
__device__ __forceinline__ void
serve_top (const Args& a, double* const mt, const bool has_prev, const int k, const int st,
           const int snt) {
  constexpr bool caas = CLS == CLS_CAAS;
  const TopArgs& T = a.top;
  const int t = a.tracers[k];
  const int M = T.M;
  double* const f0 = mt;
  double* const f1 = f0 + 2*M;
  double* const f2 = f1 + 2*M;
  double* const f3 = f2 + 2*M;
  const bool prefer = a.prefer_mass_con != 0;
  const double* const rbase = a.rec + static_cast<long long>(t)*4*a.rec_ld;
  auto sbar = [&] () { if (snt > 32) bar_sync_dyn(15, snt); else __syncwarp(); };

  // The 8 sub-roots of micro-root m (written by other SMs: bypass L1).
  auto load8 = [&] (const int m, const int f, double (&v)[8]) {
    const double* const r = rbase + f*a.rec_ld + 8LL*m;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const double2 d = __ldcg(reinterpret_cast<const double2*>(r + j));
      v[j] = d.x; v[j + 1] = d.y;
    }
  };
  for (int m = st; m < M; m += snt) {
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      if (f == 3 && ! has_prev) continue;
      double v[8];
      load8(m, f, v);
      const double s = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
      (f == 0 ? f0 : f == 1 ? f1 : f == 2 ? f2 : f3)[m] = s;
    }
  }
  sbar();
  if (st < 32) {
    // The tree over the micro-roots: one warp, level by level.
    const int lane = st;
    for (int l = 0; l < T.nlev; ++l) {
      const int je = T.lvlptr[l + 1];
      for (int j = T.lvlptr[l] + lane; j < je; j += 32) {
        const int k0 = T.kid0[j], k1 = T.kid1[j], me = M + j;
        f0[me] = f0[k0] + f0[k1];
        f1[me] = f1[k0] + f1[k1];
        f2[me] = f2[k0] + f2[k1];
        if (has_prev) f3[me] = f3[k0] + f3[k1];
      }
      __syncwarp();
    }
    const int root = T.ni ? M + T.ni - 1 : 0;
    if (caas) {
      if (lane == 0) {
        const double clip_sum = f1[root], term_sum = has_prev ? f3[root] : f1[root];
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
      if (lane == 0 && ! has_prev) f3[root] = f1[root];
      __syncwarp();
      for (int l = T.nlev - 1; l >= 0; --l) {
        const int je = T.lvlptr[l + 1];
        for (int j = T.lvlptr[l] + lane; j < je; j += 32) {
          const int k0 = T.kid0[j], k1 = T.kid1[j], me = M + j;
          const dev::NodeWQ c = T.mwq[j];
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
  if ( ! caas) {
    // The three levels under every micro-root, in registers (depths S-3 .. S-1 of a block).
    for (int m = st; m < M; m += snt) {
      double sub[8][3], s2[4][3], s1[2][3];
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        double v[8];
        load8(m, f, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) sub[j][f] = v[j];
#pragma unroll
        for (int j = 0; j < 4; ++j) s2[j][f] = v[2*j] + v[2*j + 1];
        s1[0][f] = s2[0][f] + s2[1][f];
        s1[1][f] = s2[2][f] + s2[3][f];
      }
      const double s0[3] = {f0[m], f1[m], f2[m]};
      const int cb = T.micro_c[m], h = T.micro_h[m];
      const dev::NodeWQ* const wq = a.wq + (cb - h);     // the block's constants, heap order
      const dev::NodeRh* const rh = a.rh + (cb - h);
      auto solve = [&] (const int hh, const double* nd, const double b, const double* k0,
                        const double* k1, double& x0, double& x1) {
        const dev::NodeWQ c = wq[hh];
        if (prefer)
          dev::solve_bounded_lean<true>(c, 0.0, rh + hh, nd[0], nd[1], nd[2], b, k0[0], k0[1],
                                        k0[2], k1[0], k1[1], k1[2], x0, x1);
        else
          dev::solve_bounded_lean<false>(c, 0.0, rh + hh, nd[0], nd[1], nd[2], b, k0[0], k0[1],
                                         k0[2], k1[0], k1[1], k1[2], x0, x1);
      };
      double x1v[2], x2v[4], x3v[8];
      solve(h, s0, f3[m], s1[0], s1[1], x1v[0], x1v[1]);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        solve(2*h + 1 + j, s1[j], x1v[j], s2[2*j], s2[2*j + 1], x2v[2*j], x2v[2*j + 1]);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        solve(4*h + 3 + j, s2[j], x2v[j], sub[2*j], sub[2*j + 1], x3v[2*j], x3v[2*j + 1]);
      double* const o = a.sol + static_cast<long long>(t)*a.sol_ld + 8LL*m;
#pragma unroll
      for (int j = 0; j < 8; j += 2)
        __stcg(reinterpret_cast<double2*>(o + j), make_double2(x3v[j], x3v[j + 1]));
    }
    sbar();
  }
  if (st == 0) {
    __threadfence();
    st_release(a.flag + k, 1u);
  }
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
  const int NSLOT = a.nslots, NPS = a.npslots;
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

  const SmemLayout lay = smem_layout(TB, plen, NSLOT, NPS, NP, a.top.M, npn, a.npairs_max, caas);
  double* const slots = reinterpret_cast<double*>(smraw + lay.slots);
  const int slot_doubles = lay.slot_doubles;
  double* const prevring = reinterpret_cast<double*>(smraw + lay.prevring);
  dev::NodeWQ* const tconst = reinterpret_cast<dev::NodeWQ*>(smraw + lay.tconst);
  dev::NodeWQ* const pconst = reinterpret_cast<dev::NodeWQ*>(smraw + lay.pconst);
  int* const tgi = reinterpret_cast<int*>(smraw + lay.tgi);
  unsigned short* const tcidx = reinterpret_cast<unsigned short*>(smraw + lay.tcidx);
  uint64_t* const full = reinterpret_cast<uint64_t*>(smraw + lay.bars);
  uint64_t* const upd = full + NSLOT;
  uint64_t* const topd = upd + NSLOT;
  uint64_t* const done = topd + NSLOT;
  uint64_t* const pempty = done + NSLOT;
  int* const ltask = reinterpret_cast<int*>(smraw + lay.ltask);
  auto slot_rows = [&] (const int s) { return slots + static_cast<size_t>(s)*slot_doubles; };
  auto slot_d7 = [&] (const int s) { return slot_rows(s) + TB*3*plen; };
  auto slot_x7 = [&] (const int s) { return slot_d7(s) + 4*kGroup; };
  auto unit_ntr = [&] (const int u) { return min(TB, a.ntr - u*TB); };

  // ---- one-time setup
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&upd[s], kGroup);
      mbar_init(&topd[s], 32);
      mbar_init(&done[s], kGroup);
    }
    for (int s = 0; s < NPS; ++s) mbar_init(&pempty[s], kGroup);
    mbar_fence_init();
  }
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
    int ul = 0, nstored = 0;
    int us[kMaxPipes];
#pragma unroll
    for (int p = 0; p < kMaxPipes; ++p) us[p] = p;
    // stored_rounds[s]: how many units of slot s have been stored (mod 256); unit u may be
    // loaded once that equals u / NSLOT.
    unsigned char* const stored_rounds = reinterpret_cast<unsigned char*>(ltask + 2*kMaxPipes + 2);
    if (lane == 0) for (int s = 0; s < NSLOT; ++s) stored_rounds[s] = 0;
    __syncwarp();
    int pending_slot = -1;
    Watchdog wd;
    while (nstored < U) {
      bool did = false;
      // Stores first: they free slots.
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int u = us[p];
        if (u >= U) continue;
        const int s = u % NSLOT;
        if ( ! mbar_test(&done[s], (u/NSLOT) & 1)) continue;
        const int nt = unit_ntr(u);
        CEDR_RING_TRACE(u, 7);
        if (lane < nt) {
          const int t = a.tracers[u*TB + lane];
          double* const o = caas ?
            const_cast<double*>(a.in) + (static_cast<long long>(a.trcr_row[t]) + 1)*a.in_ld + P.leaf0 :
            a.out + static_cast<long long>(t)*a.out_ld + P.leaf0;
          const double* const x = slot_rows(s) + (lane*3 + 1)*plen + shift;
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
          stored_rounds[pending_slot] = static_cast<unsigned char>(stored_rounds[pending_slot] + 1);
        pending_slot = s;
        __syncwarp();
        us[p] = u + NP;
        ++nstored;
        did = true;
      }
      if (ul < U) {
        const int s = ul % NSLOT, r = ul/NSLOT;
        const int ps = ul % NPS, pr = ul/NPS;
        const bool slot_free = stored_rounds[s] == static_cast<unsigned char>(r);
        if (slot_free && ( ! has_prev || pr == 0 || mbar_test(&pempty[ps], (pr - 1) & 1))) {
          const int nt = unit_ntr(ul);
          CEDR_RING_TRACE(ul, 0);
          if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(&full[s], static_cast<unsigned>(nt*nrows)*rowbytes);
          }
          __syncwarp();
          if (lane < nt) {
            double* const rows = slot_rows(s);
            double* const prow = prevring + static_cast<size_t>(ps)*TB*plen;
            const int t = a.tracers[ul*TB + lane];
            const double* src = a.in + static_cast<long long>(a.trcr_row[t])*a.in_ld + src0;
#pragma unroll
            for (int f = 0; f < 3; ++f)
              tma_load(rows + (lane*3 + f)*plen, src + f*a.in_ld, rowbytes, &full[s]);
            if (has_prev) tma_load(prow + lane*plen, src + 3*a.in_ld, rowbytes, &full[s]);
          }
          ++ul;
          did = true;
        }
      }
      if (did) { wd.spins = 0; continue; }
      if (wd.expired(a)) break;
    }
    tma_store_wait_all();
    return;
  }

  // =================================================================== S warps
  if (warp <= SW) {
    const int st = tid - 32, snt = 32*SW;
    double* const mt = reinterpret_cast<double*>(smraw + lay.mt);
    const unsigned need = gridDim.x;
    for (int k = blockIdx.x; k < a.ntr; k += gridDim.x) {
      bool ok = true;
      if (st == 0) {
        Watchdog wd;
        while (ld_acquire(a.cnt + k) < need) {
          if (wd.expired(a)) { ok = false; break; }
          __nanosleep(40);
        }
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
      serve_top<CLS>(a, mt, sum4, k, st, snt);
      if (a.trace && st == 0) a.trace[static_cast<size_t>(gridDim.x)*U*8 + 2*k + 1] = global_ns();
    }
    return;
  }

  // =================================================================== T warps
  if (warp < 1 + SW + NP) {
    const int p = warp - 1 - SW;
    double* const ts = reinterpret_cast<double*>(smraw + lay.tsum) + p*4*kGroup;
    double* const tx = reinterpret_cast<double*>(smraw + lay.tx) + p*kGroup;
    const bool prefer = a.prefer_mass_con != 0;
    int tu = p, td = p;
    Watchdog wd;
    // Sums of the levels above depth 7 of the unit in slot s, nf fields: node z + i of the
    // level with z entries is entries 2i, 2i + 1 of the level below (depth 7: the slot's
    // d7 array). With `records`, the sub-roots' sums also go to global memory.
    auto sums = [&] (const int s, const int nf, const int u, const bool records) {
      const double* const d7 = slot_d7(s);
      const int nt = unit_ntr(u);
      for (int z = W >> 1; z >= Wtop; z >>= 1) {
        for (int f = 0; f < nf; ++f) {
          const double* const src = (2*z == W) ? d7 + f*kGroup : ts + f*kGroup + 2*z;
          for (int i = lane; i < z; i += 32) {
            const double2 v = reinterpret_cast<const double2*>(src)[i];
            const double sum = v.x + v.y;
            ts[f*kGroup + z + i] = sum;
            if (records && z == Wtop) {
              const int j = i/nsubmax, sb = i - j*nsubmax;
              if (j < nt && sb < P.nsub)
                a.rec[(static_cast<long long>(a.tracers[u*TB + j])*4 + f)*a.rec_ld + P.sub0 + sb] = sum;
            }
          }
        }
        __syncwarp();
      }
      if (records && Wtop == W) {
        // S = 7: the depth-7 nodes are the sub-roots.
        for (int f = 0; f < nf; ++f)
          for (int i = lane; i < W; i += 32) {
            const int j = i/nsubmax, sb = i - j*nsubmax;
            if (j < nt && sb < P.nsub)
              a.rec[(static_cast<long long>(a.tracers[u*TB + j])*4 + f)*a.rec_ld + P.sub0 + sb] =
                d7[f*kGroup + i];
          }
      }
    };
    for (;;) {
      // DOWN of the oldest unit whose tracers' flags are all up.
      if ( ! caas && td < tu) {
        const int nt = unit_ntr(td);
        const bool ok = lane >= nt || ld_acquire(a.flag + td*TB + lane) != 0;
        if (__all_sync(0xffffffffu, ok)) {
          const int s = td % NSLOT;
          CEDR_RING_TRACE(td, 4);
          const double* const d7 = slot_d7(s);
          double* const x7 = slot_x7(s);
          double* const xtop = (Wtop == W) ? x7 : tx + Wtop;
          for (int i = lane; i < Wtop; i += 32) {
            const int j = i/nsubmax, sb = i - j*nsubmax;
            // (Unused entries -- a short piece, the last unit's missing tracers -- solve
            // the all-zero problem: its quick exit.)
            xtop[i] = (j < nt && sb < P.nsub) ?
              __ldcg(a.sol + static_cast<long long>(a.tracers[td*TB + j])*a.sol_ld + P.sub0 + sb) : 0.0;
          }
          sums(s, 3, td, false);
          for (int z = Wtop; z < W; z <<= 1) {
            for (int i = lane; i < z; i += 32) {
              const int node = z + i;
              double k0[3], k1[3], nd[3];
#pragma unroll
              for (int f = 0; f < 3; ++f) {
                const double* const src = (2*z == W) ? d7 + f*kGroup : ts + f*kGroup + 2*z;
                const double2 v = reinterpret_cast<const double2*>(src)[i];
                k0[f] = v.x; k1[f] = v.y;
                nd[f] = ts[f*kGroup + node];
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
          CEDR_RING_TRACE(td, 5);
          td += NP;
          wd.spins = 0;
          continue;
        }
      }
      if (tu < U && mbar_test(&upd[tu % NSLOT], (tu/NSLOT) & 1)) {
        const int s = tu % NSLOT, nt = unit_ntr(tu);
        sums(s, sum4 ? 4 : 3, tu, true);
        __syncwarp();
        if (lane < nt) {
          __threadfence();
          red_release_add(a.cnt + tu*TB + lane, 1u);
        }
        CEDR_RING_TRACE(tu, 3);
        tu += NP;
        wd.spins = 0;
        continue;
      }
      if (tu >= U && (caas || td >= U)) break;
      if (wd.expired(a)) break;
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
  for (;;) {
    // ---- what next: the DOWN of the oldest unit if its masses (CAAS: scalars) are there,
    // else the UP of the next unit if it has landed.
    if (gw == 0) {
      int task = 0;
      Watchdog wd;
      for (;;) {
        if (dn >= U) break;
        if (dn < up) {
          bool rd;
          if (caas) {
            const int nt = unit_ntr(dn);
            const bool ok = lane >= nt || ld_acquire(a.flag + dn*TB + lane) != 0;
            rd = __all_sync(0xffffffffu, ok);
            if (rd && lane < nt) {
              const int t = a.tracers[dn*TB + lane];
              lscal[2*lane] = __ldcg(a.scal + 2*t);
              lscal[2*lane + 1] = __ldcg(a.scal + 2*t + 1);
            }
          } else {
            rd = mbar_test(&topd[dn % NSLOT], (dn/NSLOT) & 1);
          }
          if (rd) { task = 1; break; }
        }
        if (up < U && mbar_test(&full[up % NSLOT], (up/NSLOT) & 1)) { task = 2; break; }
        const bool ex = wd.expired(a);
        if (__any_sync(0xffffffffu, ex)) break;
      }
      if (lane == 0) ltask[2*p + (it & 1)] = task;
    }
    bar_sync_dyn(barid, kGroup);
    const int task = ltask[2*p + (it & 1)];
    ++it;
    if (task == 0) break;

    if (task == 2) {
      // ------------------------------------------------------------------ UP(up)
      const int u = up, s = u % NSLOT, ps = u % NPS;
      up += NP;
      mbar_test(&full[s], (u/NSLOT) & 1);   // (acquire by every thread; it has completed)
      if (gw == 0) CEDR_RING_TRACE(u, 1);
      const bool act = node_ok && j_of < unit_ntr(u);
      if (act) {
        const double* const r0 = slot_rows(s) + j_of*3*plen;
        const double* const rp = prevring + (static_cast<size_t>(ps)*TB + j_of)*plen;
        bool conserve = true;
        if (caas) conserve = (a.trcr_prob[a.tracers[u*TB + j_of]] & 1) != 0;
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
              const double term = (has_prev && conserve) ? rp[o] : q;
              const double clip = dev::rmin(hi, dev::rmax(lo, q));
              v[jj][0] = 0.0 + lo; v[jj][1] = 0.0 + clip; v[jj][2] = 0.0 + hi;
              v[jj][3] = 0.0 + term;
            } else {
              v[jj][0] = r0[o]; v[jj][1] = r0[plen + o]; v[jj][2] = r0[2*plen + o];
              if (has_prev) v[jj][3] = rp[o];
            }
          }
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            if (f == 3 && ! has_prev && ! caas) continue;
            n[k][f] = pr[k] ? v[0][f] + v[1][f] : v[0][f];
          }
        }
        double* const d7 = slot_d7(s);
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          if (f == 3 && ! has_prev && ! caas) continue;
          d7[f*kGroup + gt] = (n[0][f] + n[1][f]) + (n[2][f] + n[3][f]);
        }
      } else {
        // Unused entries (a short piece, the last unit's missing tracers): keep them tame.
        double* const d7 = slot_d7(s);
#pragma unroll
        for (int f = 0; f < 4; ++f) d7[f*kGroup + gt] = 0.0;
      }
      mbar_arrive(&upd[s]);
      if (has_prev) mbar_arrive(&pempty[ps]);
      if (gw == 0) CEDR_RING_TRACE(u, 2);
      continue;
    }

    // -------------------------------------------------------------------- DOWN(dn)
    const int u = dn, s = u % NSLOT, nt = unit_ntr(u);
    dn += NP;
    double* const rows = slot_rows(s);
    if (caas) {
      if (gw == 0) CEDR_RING_TRACE(u, 4);
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
      fence_proxy_async();
      mbar_arrive(&done[s]);
      if (gw == 0) CEDR_RING_TRACE(u, 6);
      continue;
    }
    mbar_test(&topd[s], (u/NSLOT) & 1);
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
      const double x7 = slot_x7(s)[gt];
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
    bar_sync_dyn(barid, kGroup);
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
    fence_proxy_async();
    mbar_arrive(&done[s]);
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
