// Fast path of the tier-0 sweeps for blocks whose shape is the recursive bisection
// n -> (floor(n/2), n - floor(n/2)) of cedr_tree.cpp:391-413 with 512 < n <= 1024
// leaves (what make_tree_over_1d_mesh yields below any node of that size; all the
// BASELINE.json cubed-sphere configs cut into such blocks: 675 leaves at ne30/ne120,
// 768 at ne256).
//
// Structure of such a block: a PERFECT binary tree down to depth 9 (512 nodes), where
// each depth-9 node is either one leaf or a pair of leaves (n - 512 pairs). One CTA of
// 128 threads sweeps one block for a group of tracers, one tracer at a time:
//   - a tracer's leaf rows are staged into shared memory by TMA bulk copies
//     (cp.async.bulk + mbarrier), double-buffered so tracer i+1 streams in while tracer
//     i is being swept;
//   - a thread owns one depth-7 node (4 depth-9 nodes, 4..8 leaves): its micro-subtree
//     is summed and solved in registers. In the up-sweep thread `tid` owns node `tid`
//     (the levels above go through shuffles); in the down-sweep the nodes are dealt to
//     the threads in a per-shape order that minimises shared-memory bank conflicts
//     (FastArgs::perm, tree_plan.cpp build_down_tables);
//   - the levels above the depth-7 nodes go through warp shuffles (sums) and, in the
//     down-sweep, through a dedicated top warp working a tracer ahead of the leaf warps;
//   - solved leaf masses are staged in shared memory and leave by a TMA bulk store.
// The node arithmetic is the same device code as the generic path (node_solve.cuh), in
// the same tree order, so results are bit-identical to it and to the reference.
#ifndef CEDR_B200_FAST_KERNELS_CUH
#define CEDR_B200_FAST_KERNELS_CUH

#include <cstdint>

#include "kernels.cuh"

namespace cedr_b200 {
namespace fast {

constexpr int kThreads = 128;     // = number of depth-7 nodes of a block
constexpr int kDepthLeafParents = 9;
constexpr int kD9 = 512;          // depth-9 nodes per block
constexpr int kMinLeaves = 513, kMaxLeaves = 1024;

// Positions of a block's internal nodes in the fast const order:
//   [0, 511): heap order (depth d, position p) -> 2^d - 1 + p, for d = 0..8
//   [511, 511 + npairs): the depth-9 pairs, by increasing p
constexpr int kHeapNodes = 511;

// w and q of node_solve.cuh, 32 B, so two LDS.128 / LDG.128 fetch them.
typedef FastWQ NodeWQ;
typedef FastRh NodeRh;

__device__ __forceinline__ unsigned smem_u32 (const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init (uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init () {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx (uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait (uint64_t* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n"
                 " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                 " selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while ( ! ok);
}
// TMA bulk copy shared -> global (16-byte aligned addresses and size), bulk-group tracked.
__device__ __forceinline__ void tma_store (void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit () {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// Wait until the committed bulk stores have finished READING shared memory.
__device__ __forceinline__ void tma_store_wait_read () {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// TMA bulk copy global -> shared (16-byte aligned addresses and size).
__device__ __forceinline__ void tma_load (void* dst, const void* src, unsigned bytes,
                                          uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
               "[%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct FastArgs {
  const BlockDev* blocks;
  int nblocks;
  const unsigned short* dtab;   // per shape: 512 depth-9 entries, off | (pair << 15)
  const unsigned short* ptab;   // per shape: pair list (depth-9 positions)
  // Down-sweep work assignment (tree_plan.cpp build_down_tables), per shape: perm[thread] =
  // depth-7 node of a leaf thread; pent = [first entry of warp 0..3, end | pair entries],
  // entry = leaf offset | d9x slot << 11 | pair's const rank << 20.
  const unsigned short* perm;
  const unsigned* pent;
  const NodeWQ* wq;             // per block: kHeapNodes + npairs entries, fast order
  const NodeRh* rh;
  const double* in;             // tier-0 rows (the CDR's own buffer)
  long long in_ld;
  // Where the tracers' leaf rows are (the CDR's own buffer or arrays bound by the caller):
  // rowaddr[4 t + role], role = 0 min, 1 Qm, 2 max, 3 prev; null where the class has none.
  // (A table rather than the RowMap's four base pointers: the thread that issues the TMA
  // copies then carries one pointer, which keeps down2_kernel at its register budget.)
  const double* const* rowaddr;
  const int* trcr_row;
  const int* trcr_prob;
  double* rec_out;              // UP: [(4 t + f) rec_ld + block]
  long long rec_ld;
  const double* sol_in;         // DOWN: [t sol_in_ld + block]
  long long sol_in_ld;
  double* out;                  // DOWN: [t out_ld + leaf]
  long long out_ld;
  const int* tracers;
  int ntr;
  int group;                    // tracers per CTA
  int sbuf;                     // doubles per staged row (even, >= max_nl + 2)
  int prefer_mass_con;
  const double* qglob;
  unsigned long long* phase_clk;  // debug: per-phase clock sums (CEDR_B200_PHASE_CLOCKS)
  // Depth-7 sums (min, Qm, max) of every block x tracer, written by up_kernel for
  // down2_kernel's top warp: [(t nblocks + block)*384 + f*128 + depth-7 node].
  double* n7buf;
  const double* rq;             // per block, fast order: RN(1/(q0 + q1)) or 0 (node_solve.cuh)
  // split = S > 0: the block's depth-S nodes (2^S "sub-roots"), not its root, are the
  // leaves of the tier above: rec_out / sol_in hold 2^S entries per block, at
  // [.. + gidx 2^S + j], and depths 0..S-1 of the block are swept by the tier above. This
  // takes the narrowest levels of the dependent chain out of the block kernels. S is 0, 2
  // or 3.
  int split;
  // CAAS: rows allotted per tracer in `in` (4 if any tracer conserves, else 3: there is no
  // Qm_prev row to stage, cedr_caas.cpp:86-100).
  int caas_rows;
};

#ifdef CEDR_B200_PHASE_CLOCKS
# define CEDR_PHASE_DECL long long pc_t0 = clock64(); unsigned long long pc_acc[8] = {0}
# define CEDR_PHASE(k) do { const long long pc_t1 = clock64(); pc_acc[k] += pc_t1 - pc_t0; pc_t0 = pc_t1; } while (0)
# define CEDR_PHASE_FLUSH(base) do { if (a.phase_clk) for (int pc_i = 0; pc_i < 8; ++pc_i) atomicAdd(a.phase_clk + (base) + pc_i, pc_acc[pc_i]); } while (0)
#else
# define CEDR_PHASE_DECL do {} while (0)
# define CEDR_PHASE(k) do {} while (0)
# define CEDR_PHASE_FLUSH(base) do {} while (0)
#endif

template <int CLS> struct Rows {
  // Rows of a tracer in the caller-facing buffer (cedr_qlt_inl.hpp:21-58) and where
  // (min, Qm, max, prev) sit among them.
  static constexpr bool caas = CLS == CLS_CAAS;
  static constexpr bool nonneg = CLS == CLS_NN || CLS == CLS_CNN;
  static constexpr bool consistent_only = CLS == CLS_T || CLS == CLS_CT;
  static constexpr bool has_prev = CLS == CLS_CST || CLS == CLS_CT || CLS == CLS_CNN;
  static constexpr int r_min = nonneg ? -1 : 0;
  static constexpr int r_qm = nonneg ? 0 : 1;
  static constexpr int r_max = nonneg ? -1 : 2;
  static constexpr int r_prev = nonneg ? 1 : 3;
};

// ------------------------------------------------------------------------ UP
//
// Block-root record of QLT::l2r_combine_kid_data (cedr_qlt.cpp:339-430) for a
// fast-path block; for CLS_CAAS the four tree-ordered sums of CAAS::reduce_locally
// + BfbTreeAllReducer (cedr_caas.cpp:129-201, cedr_bfb_tree_allreduce.cpp:86-124).
template <int CLS>
__global__ void __launch_bounds__(kThreads)
up_kernel (const FastArgs a) {
  typedef Rows<CLS> R;
  constexpr bool bounds = ! R::nonneg;
  constexpr bool prev = R::has_prev || R::caas;
  constexpr int nrows = (bounds ? 3 : 1) + (prev ? 1 : 0);
  extern __shared__ __align__(16) unsigned char smraw[];
  double* const stage = reinterpret_cast<double*>(smraw);          // [2][nrows][sbuf]
  uint64_t* const mbar = reinterpret_cast<uint64_t*>(stage + 2*nrows*a.sbuf);
  double* const wroot = reinterpret_cast<double*>(mbar + 2);       // [2][4 warps][4]

  const int b = blockIdx.x % a.nblocks, grp = blockIdx.x / a.nblocks;
  const BlockDev B = a.blocks[b];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int src0 = B.leaf0 & ~1, shift = B.leaf0 - src0;
  const unsigned bytes = 8u*static_cast<unsigned>(((B.leaf0 + B.nl + 1) & ~1) - src0);
  const int g0 = grp*a.group;
  const int gn = min(a.group, a.ntr - g0);

  // This thread's depth-7 node -- one of its own warp's 32, in the order that spreads the
  // half-warps' leaf offsets over the shared-memory banks (FastArgs::perm, second table) --
  // and, for the shuffle tree above depth 7, the lane that holds node 32 warp + lane.
  const unsigned short* const pup = a.perm + B.fperm_up_off;
  const int node = pup[tid], src_lane = pup[128 + tid] & 31;
  // The node's four depth-9 nodes: leaf offset and whether it is a pair.
  const ushort4 e = reinterpret_cast<const ushort4*>(a.dtab + B.ftab_off)[node];
  const int o0 = (e.x & 0x7fff) + shift, o1 = (e.y & 0x7fff) + shift,
    o2 = (e.z & 0x7fff) + shift, o3 = (e.w & 0x7fff) + shift;
  const bool p0 = e.x >> 15, p1 = e.y >> 15, p2 = e.z >> 15, p3 = e.w >> 15;

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  // Rows to stage: a CAAS without conserving tracers has no Qm_prev row.
  const int nstage = (R::caas && a.caas_rows == 3) ? 3 : nrows;
  auto issue = [&] (const int i) {
    const int t = a.tracers[g0 + i];
    double* dst = stage + (i & 1)*nrows*a.sbuf;
    // (One load of the tracer's row index: the asm statements below are memory barriers
    // to the compiler.)
    const double* const* const ra = a.rowaddr + 4*t;
    mbar_expect_tx(&mbar[i & 1], nstage*bytes);
#pragma unroll
    for (int f = 0; f < nrows; ++f)
      if (f < nstage) {
        // Staged row f holds role (min, Qm, max, prev)[f]; the nonnegative classes stage
        // Qm[, prev] only.
        const int role = bounds ? f : (f == 0 ? 1 : 3);
        tma_load(dst + f*a.sbuf, ra[role] + src0, bytes, &mbar[i & 1]);
      }
  };
  if (tid == 0) {
    issue(0);
    if (gn > 1) issue(1);
  }

  for (int i = 0; i < gn; ++i) {
    const int t = a.tracers[g0 + i];
    mbar_wait(&mbar[i & 1], (i >> 1) & 1);
    const double* const s = stage + (i & 1)*nrows*a.sbuf;
    // Leaf values of one depth-9 node -> its record (one leaf, or leaf + leaf).
    double r[4];   // (min, Qm|clip, max, prev|term) of this thread's depth-7 node
    {
      double n[4][4];
      const int off[4] = {o0, o1, o2, o3};
      const bool pr[4] = {p0, p1, p2, p3};
      bool conserve = true;
      if (R::caas) conserve = a.trcr_prob[t] & 1;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        double v[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int o = off[k] + j;
          if (j == 1 && ! pr[k]) break;
          if (R::caas) {
            const double lo = s[o], q = s[a.sbuf + o], hi = s[2*a.sbuf + o];
            const double term = conserve ? s[3*a.sbuf + o] : q;
            const double clip = dev::rmin(hi, dev::rmax(lo, q));
            v[j][0] = 0.0 + lo; v[j][1] = 0.0 + clip; v[j][2] = 0.0 + hi;
            v[j][3] = 0.0 + term;
          } else {
            if (bounds) { v[j][0] = s[R::r_min*a.sbuf + o]; v[j][2] = s[R::r_max*a.sbuf + o]; }
            v[j][1] = s[R::r_qm*a.sbuf + o];
            if (prev) v[j][3] = s[R::r_prev*a.sbuf + o];
          }
        }
        if (pr[k]) {
          if (bounds) {
            n[k][0] = R::consistent_only ? dev::rmin(v[0][0], v[1][0]) : v[0][0] + v[1][0];
            n[k][2] = R::consistent_only ? dev::rmax(v[0][2], v[1][2]) : v[0][2] + v[1][2];
          }
          n[k][1] = v[0][1] + v[1][1];
          if (prev) n[k][3] = v[0][3] + v[1][3];
        } else {
          if (bounds) { n[k][0] = v[0][0]; n[k][2] = v[0][2]; }
          n[k][1] = v[0][1];
          if (prev) n[k][3] = v[0][3];
        }
      }
      // depth 8, then depth 7, in tree order (left + right).
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        if ((f == 0 || f == 2) && ! bounds) continue;
        if (f == 3 && ! prev) continue;
        if (R::consistent_only && f == 0)
          r[f] = dev::rmin(dev::rmin(n[0][f], n[1][f]), dev::rmin(n[2][f], n[3][f]));
        else if (R::consistent_only && f == 2)
          r[f] = dev::rmax(dev::rmax(n[0][f], n[1][f]), dev::rmax(n[2][f], n[3][f]));
        else
          r[f] = (n[0][f] + n[1][f]) + (n[2][f] + n[3][f]);
      }
    }
    // Back into lane order: lane l takes the record of node 32 warp + l.
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      if ((f == 0 || f == 2) && ! bounds) continue;
      if (f == 3 && ! prev) continue;
      r[f] = __shfl_sync(0xffffffffu, r[f], src_lane);
    }
    if ( ! R::caas && a.n7buf) {
      // Depth-7 sums for the down-sweep's top warp (the one-field classes need Qm only).
      double* const n7 = a.n7buf + (static_cast<long long>(t)*a.nblocks + b)*384;
#pragma unroll
      for (int f = 0; f < 3; ++f)
        if (f == 1 || (CLS == CLS_ST || CLS == CLS_CST)) n7[f*128 + tid] = r[f];
    }
    // Depths 6..2 inside the warp: lane l (l % 2^(L+1) == 0) takes left + right.
#pragma unroll
    for (int L = 0; L < 5; ++L) {
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        if ((f == 0 || f == 2) && ! bounds) continue;
        if (f == 3 && ! prev) continue;
        const double o = __shfl_down_sync(0xffffffffu, r[f], 1 << L);
        if (R::consistent_only && f == 0) r[f] = dev::rmin(r[f], o);
        else if (R::consistent_only && f == 2) r[f] = dev::rmax(r[f], o);
        else r[f] = r[f] + o;
      }
      // After level L the lanes with lane % 2^(L+1) == 0 hold depth 6-L nodes; depth 3
      // after L = 3 (two per warp), depth 2 after L = 4 (one per warp).
      if (a.split && L == 6 - a.split && (lane & ((2 << L) - 1)) == 0) {
        const int E = 1 << a.split;
        const int j = (warp << (a.split - 2)) + (lane >> (L + 1));
        double* rec = a.rec_out + static_cast<long long>(t)*4*a.rec_ld +
          static_cast<long long>(B.gidx)*E + j;
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          if ((f == 0 || f == 2) && ! bounds) continue;
          if (f == 3 && ! prev) continue;
          rec[f*a.rec_ld] = r[f];
        }
      }
    }
    if (a.split) {
      // The shuffle tree above was also run past the sub-root depth; lanes that held a
      // sub-root at its level saved it (see `sub` below).
      __syncthreads();   // all threads are past their reads of this stage
      if (tid == 0 && i + 2 < gn) issue(i + 2);
      continue;
    }
    double* const wr = wroot + (i & 1)*16;
    if (lane == 0) {
#pragma unroll
      for (int f = 0; f < 4; ++f) wr[warp*4 + f] = r[f];
    }
    __syncthreads();
    if (tid == 0) {
      // Depths 1 and 0, then the record for the next tier.
      double* rec = a.rec_out + static_cast<long long>(t)*4*a.rec_ld + B.gidx;
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        if ((f == 0 || f == 2) && ! bounds) continue;
        if (f == 3 && ! prev) continue;
        double v;
        if (R::consistent_only && f == 0)
          v = dev::rmin(dev::rmin(wr[f], wr[4 + f]), dev::rmin(wr[8 + f], wr[12 + f]));
        else if (R::consistent_only && f == 2)
          v = dev::rmax(dev::rmax(wr[f], wr[4 + f]), dev::rmax(wr[8 + f], wr[12 + f]));
        else
          v = (wr[f] + wr[4 + f]) + (wr[8 + f] + wr[12 + f]);
        rec[f*a.rec_ld] = v;
      }
      // All threads are past their reads of this stage (the barrier above).
      if (i + 2 < gn) issue(i + 2);
    }
  }
}

// ----------------------------------------------------------------------- MID
//
// Multi-rank runs with split = 3: the ranks exchange BLOCK-ROOT records, so depths 0..2 of
// a rank's own blocks are handled here, one thread per (own block, tracer), either side of
// the replicated tier-1 sweep. The 8 sub-root records of a block x tracer are 8 consecutive
// doubles per field in a.rec_out (written by up_kernel); sums pairwise in tree order.
struct MidSums { double s2[4][3], s1[2][3], s0[3]; };

__device__ __forceinline__ void mid_load (const FastArgs& a, const int t, const int gidx,
                                          double (&sub)[8][3], MidSums& m) {
  const double* const r = a.rec_out + static_cast<long long>(t)*4*a.rec_ld +
    static_cast<long long>(gidx)*8;
#pragma unroll
  for (int f = 0; f < 3; ++f) {
    const double* const rf = r + f*a.rec_ld;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const double2 v = __ldcg(reinterpret_cast<const double2*>(rf + j));
      sub[j][f] = v.x; sub[j + 1][f] = v.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) m.s2[j][f] = sub[2*j][f] + sub[2*j + 1][f];
    m.s1[0][f] = m.s2[0][f] + m.s2[1][f];
    m.s1[1][f] = m.s2[2][f] + m.s2[3][f];
    m.s0[f] = m.s1[0][f] + m.s1[1][f];
  }
}

__device__ __forceinline__ void
mid_solve (const dev::NodeWQ& c, const double rq, const dev::NodeRh* rh, const bool prefer,
           const double pmin, const double pqm, const double pmax, const double b,
           const double lo0, const double y0, const double hi0, const double lo1,
           const double y1, const double hi1, double& x0, double& x1) {
  if (prefer)
    dev::solve_bounded_lean<true>(c, rq, rh, pmin, pqm, pmax, b, lo0, y0, hi0, lo1, y1, hi1,
                                  x0, x1);
  else
    dev::solve_bounded_lean<false>(c, rq, rh, pmin, pqm, pmax, b, lo0, y0, hi0, lo1, y1, hi1,
                                   x0, x1);
}

// Block-root records (QLT::l2r_combine_kid_data, cedr_qlt.cpp:339-430, depths 2..0) for the
// tier above: rec1[(4 t + f) ld1 + gidx].
__global__ void __launch_bounds__(256)
mid_up_kernel (const FastArgs a, double* rec1, const long long ld1, const bool has_prev) {
  const long long n = static_cast<long long>(a.nblocks)*a.ntr;
  for (long long k = blockIdx.x*static_cast<long long>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<long long>(gridDim.x)*blockDim.x) {
    const int b = static_cast<int>(k % a.nblocks), t = a.tracers[k / a.nblocks];
    const int gidx = a.blocks[b].gidx;
    double sub[8][3];
    MidSums m;
    mid_load(a, t, gidx, sub, m);
    double* const o = rec1 + static_cast<long long>(t)*4*ld1 + gidx;
#pragma unroll
    for (int f = 0; f < 3; ++f) o[f*ld1] = m.s0[f];
    if (has_prev) {
      const double* const rp = a.rec_out + (static_cast<long long>(t)*4 + 3)*a.rec_ld +
        static_cast<long long>(gidx)*8;
      double p[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) p[j] = __ldcg(rp + j);
      o[3*ld1] = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
    }
  }
}

// Node problems of depths 0..2 (QLT::r2l_solve_qp, cedr_qlt.cpp:490-604): the block root's
// solved mass sol1[t ld1 + gidx] -> the 8 sub-root masses a.sol_in[t sol_in_ld + 8 gidx + j]
// that down2_kernel starts from.
__global__ void __launch_bounds__(256)
mid_down_kernel (const FastArgs a, const double* sol1, const long long ld1, const bool) {
  const long long n = static_cast<long long>(a.nblocks)*a.ntr;
  const bool prefer = a.prefer_mass_con != 0;
  for (long long k = blockIdx.x*static_cast<long long>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<long long>(gridDim.x)*blockDim.x) {
    const int b = static_cast<int>(k % a.nblocks), t = a.tracers[k / a.nblocks];
    const BlockDev B = a.blocks[b];
    const dev::NodeWQ* const wq = a.wq + B.fbase;
    const dev::NodeRh* const rh = a.rh + B.fbase;
    double sub[8][3];
    MidSums m;
    mid_load(a, t, B.gidx, sub, m);
    const double x0 = __ldcg(sol1 + static_cast<long long>(t)*ld1 + B.gidx);
    double x1[2], x2[4], x3[8];
    mid_solve(wq[0], 0.0, rh, prefer, m.s0[0], m.s0[1], m.s0[2], x0,
                               m.s1[0][0], m.s1[0][1], m.s1[0][2],
                               m.s1[1][0], m.s1[1][1], m.s1[1][2], x1[0], x1[1]);
#pragma unroll
    for (int j = 0; j < 2; ++j)
      mid_solve(wq[1 + j], 0.0, rh + 1 + j, prefer, m.s1[j][0], m.s1[j][1],
                                 m.s1[j][2], x1[j], m.s2[2*j][0], m.s2[2*j][1], m.s2[2*j][2],
                                 m.s2[2*j + 1][0], m.s2[2*j + 1][1], m.s2[2*j + 1][2],
                                 x2[2*j], x2[2*j + 1]);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      mid_solve(wq[3 + j], 0.0, rh + 3 + j, prefer, m.s2[j][0], m.s2[j][1],
                                 m.s2[j][2], x2[j], sub[2*j][0], sub[2*j][1], sub[2*j][2],
                                 sub[2*j + 1][0], sub[2*j + 1][1], sub[2*j + 1][2],
                                 x3[2*j], x3[2*j + 1]);
    double* const o = const_cast<double*>(a.sol_in) + static_cast<long long>(t)*a.sol_in_ld +
      static_cast<long long>(B.gidx)*8;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = x3[j];
  }
}

// ---------------------------------------------------------------------- DOWN
//
// QLT::r2l_solve_qp (cedr_qlt.cpp:490-604) over a fast-path block for the
// shape-preserving classes (st, cst), organised so that no warp ever waits for the serial
// part of the block. The block's node problems split into
//   - the block top, depths 0..6 (127 nodes): a chain of 7 dependent levels, at most 64
//     problems wide -- one TOP warp walks it;
//   - the micro-subtrees, depths 7..9 (~80% of the nodes): 128 independent columns, one
//     per thread of the four LEAF warps, dense and barrier-free.
// The two run as a pipeline over the CTA's group of tracers. The top warp needs only the
// 128 depth-7 sums of a tracer, which the up-sweep kernel left in `n7buf` (3 KB per
// block x tracer): it TMA-loads them, sums depths 6..0 with vector loads and warp
// shuffles, solves depths 0..6 and publishes the 128 depth-7 masses (T(i)). The leaf
// warps TMA-stage the tracer's (min, Qm, max) rows two tracers deep, re-sum their
// micro-subtrees, wait for T(i) and solve depths 7..9 (C(i)). Hand-over is by named
// barriers: BAR_T "T(i) published", BAR_C "C(i) has read its masses" (the top warp may
// reuse the buffers of tracer i for tracer i+2).
constexpr int kLeafThreads = 128;
constexpr int kDown2Threads = 160;

// Named barriers with immediate ids (a register id makes ptxas reserve all 16).
template <int ID, int COUNT> __device__ __forceinline__ void bar_sync_i () {
  asm volatile("bar.sync %0, %1;" :: "n"(ID), "n"(COUNT) : "memory");
}
template <int ID, int COUNT> __device__ __forceinline__ void bar_arrive_i () {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" :: "n"(ID), "n"(COUNT) : "memory");
}
// id = BASE + parity
template <int BASE, int COUNT> __device__ __forceinline__ void bar_sync (const int parity) {
  if (parity) bar_sync_i<BASE + 1, COUNT>(); else bar_sync_i<BASE, COUNT>();
}
template <int BASE, int COUNT> __device__ __forceinline__ void bar_arrive (const int parity) {
  if (parity) bar_arrive_i<BASE + 1, COUNT>(); else bar_arrive_i<BASE, COUNT>();
}

// Shared memory of down2_kernel, in doubles: stage[2][3][sbuf] leaf rows, n7s[2][3][128]
// depth-7 sums, un[2][3][128] sums of heap nodes 0..126, xs[2][256] solved masses of heap
// nodes 0..254, d9x[512] solved masses of the depth-9 pairs; then topc[128] (NodeWQ) and
// 4 mbarriers.
inline size_t down2_smem_bytes (const int sbuf) {
  return sizeof(double)*(6*static_cast<size_t>(sbuf) + 2*3*128 + 2*3*128 + 2*256 + kD9 + 128) +
    128*sizeof(dev::NodeWQ) + 32;
}

template <int CLS>
__global__ void __launch_bounds__(kDown2Threads, 3)
down2_kernel (const FastArgs a) {
  static_assert(CLS == CLS_ST || CLS == CLS_CST, "fast down-sweep: st / cst only");
  extern __shared__ __align__(16) unsigned char smraw[];
  const int sbuf = a.sbuf;
  double* const stage = reinterpret_cast<double*>(smraw);            // [2][3][sbuf]
  double* const n7s = stage + 6*sbuf;                                // [2][3][128]
  double* const un = n7s + 2*3*128;                                  // [2][3][128]
  double* const xs = un + 2*3*128;                                   // [2][256]
  double* const d9x = xs + 2*256;                                    // [4][128]
  double* const toprq = d9x + kD9;                                   // [128]
  dev::NodeWQ* const topc = reinterpret_cast<dev::NodeWQ*>(toprq + 128);  // [128]
  uint64_t* const mbar = reinterpret_cast<uint64_t*>(topc + 128);    // [2] rows, [2] n7
  constexpr int BAR_T = 1, BAR_C = 3, BAR_LEAF = 5;   // + parity for T and C

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
  // (Making `prefer` a template argument halves the code but ptxas then spills twice as
  // much: measured 8.0 vs 6.65 ms at ne120.)
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

  if (tid < kHeapNodes/4) { topc[tid] = wq[tid]; toprq[tid] = rqv[tid]; }
  if (tid == 0) {
    for (int q = 0; q < 4; ++q) mbar_init(&mbar[q], 1);
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 4) {
    // ------------------------------------------------------------- TOP warp
    CEDR_PHASE_DECL;
    auto issue_n7 = [&] (const int i) {
      const int t = a.tracers[g0 + i];
      const double* src = a.n7buf + (static_cast<long long>(t)*a.nblocks + b)*384;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&mbar[2 + (i & 1)], 384*8);
      tma_load(n7s + (i & 1)*384, src, 384*8, &mbar[2 + (i & 1)]);
    };
    if (lane == 0) {
      issue_n7(0);
      if (gn > 1) issue_n7(1);
    }
    for (int i = 0; i < gn; ++i) {
      const int t = a.tracers[g0 + i];
      const double* const s7 = n7s + (i & 1)*384;
      double* const u = un + (i & 1)*384;
      double* const x = xs + (i & 1)*256;
      // The block root's mass, from the tier above (issued before the waits).
      const int S = a.split, E = 1 << S;
      double xroot = 0;
      if (lane < E)
        xroot = __ldcg(a.sol_in + static_cast<long long>(t)*a.sol_in_ld +
                       static_cast<long long>(B.gidx)*E + lane);
      CEDR_PHASE(0);
      mbar_wait(&mbar[2 + (i & 1)], (i >> 1) & 1);
      if (i >= 2) bar_sync<BAR_C, kDown2Threads>(i & 1);   // C(i-2) is done with x, u
      CEDR_PHASE(1);
      // Sums of depths 6..0. Lane l holds depth-7 nodes 4l..4l+3: depth 6 and 5 locally,
      // depths 4..0 by shuffles (heap node h has kids 2h+1, 2h+2; left + right).
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        const double2 v01 = reinterpret_cast<const double2*>(s7 + f*128)[2*lane];
        const double2 v23 = reinterpret_cast<const double2*>(s7 + f*128)[2*lane + 1];
        const double d6a = v01.x + v01.y, d6b = v23.x + v23.y;
        u[f*128 + 63 + 2*lane] = d6a;
        u[f*128 + 64 + 2*lane] = d6b;
        double r = d6a + d6b;
        u[f*128 + 31 + lane] = r;
#pragma unroll
        for (int Lv = 0; Lv < 5; ++Lv) {
          if (Lv > 4 - S) break;     // depths above the sub-roots belong to the tier above
          r = r + __shfl_down_sync(0xffffffffu, r, 1 << Lv);
          if ((lane & ((2 << Lv) - 1)) == 0) u[f*128 + (16 >> Lv) - 1 + (lane >> (Lv + 1))] = r;
        }
      }
      if (lane < E) x[E - 1 + lane] = xroot;
      __syncwarp();
      CEDR_PHASE(2);
      // Node problems of depths S..6 (the kids of depth 6 are the depth-7 sums).
      for (int dd = S; dd <= 6; ++dd) {
        for (int p = lane; p < (1 << dd); p += 32) {
          const int h = (1 << dd) - 1 + p;
          const double nd[3] = {u[h], u[128 + h], u[256 + h]};
          double k0[3], k1[3];
          if (dd < 6) {
#pragma unroll
            for (int f = 0; f < 3; ++f) { k0[f] = u[f*128 + 2*h + 1]; k1[f] = u[f*128 + 2*h + 2]; }
          } else {
#pragma unroll
            for (int f = 0; f < 3; ++f) {
              const double2 v = reinterpret_cast<const double2*>(s7 + f*128)[p];
              k0[f] = v.x; k1[f] = v.y;
            }
          }
          double x0, x1;
          solve(topc[h], toprq[h], h, nd, x[h], k0, k1, x0, x1);
          x[2*h + 1] = x0;
          x[2*h + 2] = x1;
        }
        __syncwarp();
      }
      CEDR_PHASE(3);
      if (lane == 0 && i + 2 < gn) issue_n7(i + 2);
      bar_arrive<BAR_T, kDown2Threads>(i & 1);   // T(i): x[127..254] are solved
    }
    if (lane == 0) CEDR_PHASE_FLUSH(8);
    return;
  }

  // ---------------------------------------------------------------- LEAF warps
  // This thread's depth-7 node: dealt out so that the leaf offsets of a half-warp's 16
  // nodes are (nearly) distinct mod 16, the 8-byte shared-memory banks.
  const int node = a.perm[B.fperm_off + tid];
  const ushort4 e = reinterpret_cast<const ushort4*>(dtab)[node];
  const int off[4] = {(e.x & 0x7fff) + shift, (e.y & 0x7fff) + shift,
                      (e.z & 0x7fff) + shift, (e.w & 0x7fff) + shift};
  const bool pr[4] = {(e.x >> 15) != 0, (e.y >> 15) != 0, (e.z >> 15) != 0,
                      (e.w >> 15) != 0};
  // This warp's pairs: the depth-9 pairs below its threads' nodes, pent[ps .. pe). The
  // solved mass of depth-9 position q of thread t's node goes through d9x[q*128 + t].
  const unsigned* const pent = a.pent + B.fpent_off;
  const int ps = pent[warp], pe = pent[warp + 1];
  const dev::NodeWQ c7 = wq[127 + node], c8a = wq[255 + 2*node], c8b = wq[256 + 2*node];
#ifdef CEDR_B200_FASTDIV
  const double rq7 = rqv[127 + node], rq8a = rqv[255 + 2*node], rq8b = rqv[256 + 2*node];
#else
  const double rq7 = 0, rq8a = 0, rq8b = 0;
#endif

  auto issue = [&] (const int i) {
    const int t = a.tracers[g0 + i];
    double* dst = stage + (i & 1)*3*sbuf;
    const double* const* const ra = a.rowaddr + 4*t;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&mbar[i & 1], 3*bytes);
#pragma unroll
    for (int f = 0; f < 3; ++f) tma_load(dst + f*sbuf, ra[f] + src0, bytes, &mbar[i & 1]);
  };
  if (tid == 0) {
    issue(0);
    if (gn > 1) issue(1);
  }

  CEDR_PHASE_DECL;
  for (int k = 0; k < gn; ++k) {
    // ---- C(k): the micro-subtrees of tracer k.
    const int t = a.tracers[g0 + k];
    double* const s = stage + (k & 1)*3*sbuf;
    double* const xout = s + sbuf;                 // solved leaves replace the Qm row
    const double* const x = xs + (k & 1)*256;
    CEDR_PHASE(0);
    mbar_wait(&mbar[k & 1], (k >> 1) & 1);
    CEDR_PHASE(1);
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
    CEDR_PHASE(2);
    // Refill the other stage with tracer k+1: the bulk store of tracer k-1 (issued at the
    // end of the last iteration) has read it by now, so this wait costs nothing, where
    // waiting right after the store held thread 0's warp -- and through the next leaf
    // barrier the whole CTA -- for the store's start-up latency.
    if (tid == 0 && k >= 1) {
      tma_store_wait_read();
      if (k + 1 < gn) issue(k + 1);
    }
    bar_sync<BAR_T, kDown2Threads>(k & 1);       // T(k) published
    const double x7 = x[127 + node];
    bar_arrive<BAR_C, kDown2Threads>(k & 1);     // xs / un of tracer k may be reused
    CEDR_PHASE(3);
    double x8[2];
    solve(c7, rq7, 127 + node, n7, x7, n8[0], n8[1], x8[0], x8[1]);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      double x9[2];
      solve(hf ? c8b : c8a, hf ? rq8b : rq8a, 255 + 2*node + hf, n8[hf], x8[hf], n9[2*hf],
            n9[2*hf + 1], x9[0], x9[1]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int q = 2*hf + j;
        if (pr[q]) d9x[q*128 + tid] = x9[j];
        else xout[off[q]] = x9[j];
      }
    }
    __syncwarp();
    CEDR_PHASE(4);
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
    // One rolled loop over this lane's pairs, constants fetched as needed (L1 hits): holding
    // the first two pairs' constants in registers across the depth-7/8 solves, with the two
    // solves unrolled, cost more in spills and code size than the loads (measured: -5%).
#pragma unroll 1
    for (int j = ps + lane; j < pe; j += 32) {
      const unsigned pe_j = pent[j];
      const int r = pe_j >> 20;
      solve_pair(wq[kHeapNodes + r], rqv[kHeapNodes + r], r, (pe_j & 0x7ff) + shift,
                 (pe_j >> 11) & 0x1ff);
    }
    CEDR_PHASE(5);
    // Write-back: the solved leaves sit in block order in xout; one TMA bulk store moves
    // the 16-byte aligned interior, thread 0 stores the (at most two) edge elements. No
    // other thread waits: the stage is refilled only after the store has read it.
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
      // (The stage is refilled in the next iteration, once the store has read it.)
    }
    CEDR_PHASE(6);
  }
  if (tid == 0) tma_store_wait_read();
  if (tid == 0) CEDR_PHASE_FLUSH(0);
}

// ------------------------------------------------------- DOWN, one-field classes
//
// QLT::r2l_solve_qp for the classes whose node problems need only the Qm sums: consistent
// only (t, ct: the bounds are the tracer's global q_min, q_max times the node's rhom,
// cedr_qlt_inl.hpp:175-188) and nonnegative (nn, cnn: bounds [0, b], :188-197). Same
// pipeline as down2_kernel -- a top warp walks depths 0..6 a tracer ahead of four leaf warps
// that own the depth-7 micro-subtrees -- over one staged row per tracer instead of three.
inline size_t down1_smem_bytes (const int sbuf) {
  return sizeof(double)*(2*static_cast<size_t>(sbuf) + 2*128 + 2*128 + 2*256 + kD9) +
    128*(sizeof(dev::NodeWQ) + sizeof(dev::NodeRh)) + 32;
}

template <int CLS>
__device__ __forceinline__ void
solve_one_field (const dev::NodeWQ& c, const dev::NodeRh& r, const bool prefer,
                 const double qmin, const double qmax, const double ym, const double bm,
                 const double y0, const double y1, double& x0, double& x1) {
  dev::NodeConst nc;
  nc.w0 = c.w0; nc.w1 = c.w1; nc.q0 = c.q0; nc.q1 = c.q1; nc.rh0 = r.rh0; nc.rh1 = r.rh1;
  if (Rows<CLS>::nonneg) {
    dev::solve_node_nonneg(nc, bm, y0, y1, x0, x1);
  } else {
    const double rh = r.rh0 + r.rh1;
    dev::solve_node_bounded(nc, prefer, qmin*rh, ym, qmax*rh, bm, qmin*r.rh0, y0, qmax*r.rh0,
                            qmin*r.rh1, y1, qmax*r.rh1, x0, x1);
  }
}

template <int CLS>
__global__ void __launch_bounds__(kDown2Threads)
down1_kernel (const FastArgs a) {
  typedef Rows<CLS> R;
  static_assert(R::nonneg || R::consistent_only, "one-field down-sweep: t, ct, nn, cnn");
  extern __shared__ __align__(16) unsigned char smraw[];
  const int sbuf = a.sbuf;
  double* const stage = reinterpret_cast<double*>(smraw);            // [2][sbuf]
  double* const n7s = stage + 2*sbuf;                                // [2][128]
  double* const un = n7s + 2*128;                                    // [2][128]
  double* const xs = un + 2*128;                                     // [2][256]
  double* const d9x = xs + 2*256;                                    // [4][128]
  dev::NodeWQ* const topc = reinterpret_cast<dev::NodeWQ*>(d9x + kD9);    // [128]
  dev::NodeRh* const toprh = reinterpret_cast<dev::NodeRh*>(topc + 128);  // [128]
  uint64_t* const mbar = reinterpret_cast<uint64_t*>(toprh + 128);   // [2] rows, [2] n7
  constexpr int BAR_T = 1, BAR_C = 3, BAR_LEAF = 5;

  const int b = blockIdx.x % a.nblocks, grp = blockIdx.x / a.nblocks;
  const BlockDev B = a.blocks[b];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int src0 = B.leaf0 & ~1, shift = B.leaf0 - src0;
  const unsigned bytes = 8u*static_cast<unsigned>(((B.leaf0 + B.nl + 1) & ~1) - src0);
  const int g0 = grp*a.group;
  const int gn = min(a.group, a.ntr - g0);
  const dev::NodeWQ* const wq = a.wq + B.fbase;
  const dev::NodeRh* const rh = a.rh + B.fbase;
  const bool prefer = a.prefer_mass_con != 0;

  if (tid < kHeapNodes/4) { topc[tid] = wq[tid]; toprh[tid] = rh[tid]; }
  if (tid == 0) {
    for (int q = 0; q < 4; ++q) mbar_init(&mbar[q], 1);
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 4) {
    // ------------------------------------------------------------- TOP warp
    auto issue_n7 = [&] (const int i) {
      const int t = a.tracers[g0 + i];
      const double* src = a.n7buf + (static_cast<long long>(t)*a.nblocks + b)*384 + 128;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&mbar[2 + (i & 1)], 128*8);
      tma_load(n7s + (i & 1)*128, src, 128*8, &mbar[2 + (i & 1)]);
    };
    if (lane == 0) {
      issue_n7(0);
      if (gn > 1) issue_n7(1);
    }
    for (int i = 0; i < gn; ++i) {
      const int t = a.tracers[g0 + i];
      const double* const s7 = n7s + (i & 1)*128;
      double* const u = un + (i & 1)*128;
      double* const x = xs + (i & 1)*256;
      const int S = a.split, E = 1 << S;
      double xroot = 0, qmin = 0, qmax = 0;
      if (lane < E)
        xroot = __ldcg(a.sol_in + static_cast<long long>(t)*a.sol_in_ld +
                       static_cast<long long>(B.gidx)*E + lane);
      if (R::consistent_only) { qmin = __ldcg(a.qglob + 2*t); qmax = __ldcg(a.qglob + 2*t + 1); }
      mbar_wait(&mbar[2 + (i & 1)], (i >> 1) & 1);
      if (i >= 2) bar_sync<BAR_C, kDown2Threads>(i & 1);   // C(i-2) is done with x, u
      {
        // Sums of depths 6..0 (heap node h has kids 2h+1, 2h+2; left + right).
        const double2 v01 = reinterpret_cast<const double2*>(s7)[2*lane];
        const double2 v23 = reinterpret_cast<const double2*>(s7)[2*lane + 1];
        const double d6a = v01.x + v01.y, d6b = v23.x + v23.y;
        u[63 + 2*lane] = d6a;
        u[64 + 2*lane] = d6b;
        double r = d6a + d6b;
        u[31 + lane] = r;
#pragma unroll
        for (int Lv = 0; Lv < 5; ++Lv) {
          if (Lv > 4 - S) break;
          r = r + __shfl_down_sync(0xffffffffu, r, 1 << Lv);
          if ((lane & ((2 << Lv) - 1)) == 0) u[(16 >> Lv) - 1 + (lane >> (Lv + 1))] = r;
        }
      }
      if (lane < E) x[E - 1 + lane] = xroot;
      __syncwarp();
      for (int dd = S; dd <= 6; ++dd) {
        for (int p = lane; p < (1 << dd); p += 32) {
          const int h = (1 << dd) - 1 + p;
          double y0, y1;
          if (dd < 6) { y0 = u[2*h + 1]; y1 = u[2*h + 2]; }
          else { const double2 v = reinterpret_cast<const double2*>(s7)[p]; y0 = v.x; y1 = v.y; }
          double x0, x1;
          solve_one_field<CLS>(topc[h], toprh[h], prefer, qmin, qmax, u[h], x[h], y0, y1, x0, x1);
          x[2*h + 1] = x0;
          x[2*h + 2] = x1;
        }
        __syncwarp();
      }
      if (lane == 0 && i + 2 < gn) issue_n7(i + 2);
      bar_arrive<BAR_T, kDown2Threads>(i & 1);   // T(i): x[127..254] are solved
    }
    return;
  }

  // ---------------------------------------------------------------- LEAF warps
  const int node = a.perm[B.fperm_off + tid];
  const ushort4 e = reinterpret_cast<const ushort4*>(a.dtab + B.ftab_off)[node];
  const int off[4] = {(e.x & 0x7fff) + shift, (e.y & 0x7fff) + shift,
                      (e.z & 0x7fff) + shift, (e.w & 0x7fff) + shift};
  const bool pr[4] = {(e.x >> 15) != 0, (e.y >> 15) != 0, (e.z >> 15) != 0,
                      (e.w >> 15) != 0};
  const unsigned* const pent = a.pent + B.fpent_off;
  const int ps = pent[warp], pe = pent[warp + 1];
  const dev::NodeWQ c7 = wq[127 + node], c8a = wq[255 + 2*node], c8b = wq[256 + 2*node];
  const dev::NodeRh h7 = rh[127 + node], h8a = rh[255 + 2*node], h8b = rh[256 + 2*node];

  auto issue = [&] (const int i) {
    const int t = a.tracers[g0 + i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&mbar[i & 1], bytes);
    tma_load(stage + (i & 1)*sbuf, a.rowaddr[4*t + 1] + src0, bytes, &mbar[i & 1]);
  };
  if (tid == 0) {
    issue(0);
    if (gn > 1) issue(1);
  }

  for (int k = 0; k < gn; ++k) {
    const int t = a.tracers[g0 + k];
    double* const s = stage + (k & 1)*sbuf;      // solved leaves replace the Qm row
    const double* const x = xs + (k & 1)*256;
    double qmin = 0, qmax = 0;
    if (R::consistent_only) { qmin = __ldcg(a.qglob + 2*t); qmax = __ldcg(a.qglob + 2*t + 1); }
    mbar_wait(&mbar[k & 1], (k >> 1) & 1);
    double n9[4], n8[2], n7;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      n9[q] = s[off[q]];
      if (pr[q]) n9[q] = n9[q] + s[off[q] + 1];
    }
    n8[0] = n9[0] + n9[1];
    n8[1] = n9[2] + n9[3];
    n7 = n8[0] + n8[1];
    if (tid == 0 && k >= 1) {
      tma_store_wait_read();
      if (k + 1 < gn) issue(k + 1);
    }
    bar_sync<BAR_T, kDown2Threads>(k & 1);       // T(k) published
    const double x7 = x[127 + node];
    bar_arrive<BAR_C, kDown2Threads>(k & 1);     // xs / un of tracer k may be reused
    double x8[2];
    solve_one_field<CLS>(c7, h7, prefer, qmin, qmax, n7, x7, n8[0], n8[1], x8[0], x8[1]);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      double x9[2];
      solve_one_field<CLS>(hf ? c8b : c8a, hf ? h8b : h8a, prefer, qmin, qmax, n8[hf], x8[hf],
                           n9[2*hf], n9[2*hf + 1], x9[0], x9[1]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int q = 2*hf + j;
        if (pr[q]) d9x[q*128 + tid] = x9[j];
        else s[off[q]] = x9[j];
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int j = ps + lane; j < pe; j += 32) {
      const unsigned pe_j = pent[j];
      const int r = pe_j >> 20, o = (pe_j & 0x7ff) + shift;
      const double y0 = s[o], y1 = s[o + 1];
      double x0, x1;
      solve_one_field<CLS>(wq[kHeapNodes + r], rh[kHeapNodes + r], prefer, qmin, qmax, y0 + y1,
                           d9x[(pe_j >> 11) & 0x1ff], y0, y1, x0, x1);
      s[o] = x0;
      s[o + 1] = x1;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    bar_sync_i<BAR_LEAF, kLeafThreads>();
    if (tid == 0) {
      double* const o = a.out + static_cast<long long>(t)*a.out_ld + B.leaf0;
      const int q0 = B.leaf0 & 1;
      const int nint = (B.nl - q0) & ~1;
      if (nint) tma_store(o + q0, s + shift + q0, 8u*static_cast<unsigned>(nint));
      tma_store_commit();
      if (q0) o[0] = s[shift];
      if (q0 + nint < B.nl) o[B.nl - 1] = s[shift + B.nl - 1];
    }
  }
  if (tid == 0) tma_store_wait_read();
}

} // namespace fast
} // namespace cedr_b200

#endif
