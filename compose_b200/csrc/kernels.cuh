// Sweep kernels of the B200 CEDR hot path (generic-shape version).
//
// One CTA sweeps one BLOCK (a subtree of <= max_block_leaves leaves, see
// tree_plan.h) for one tracer, entirely in shared memory:
//   UP    leaves -> block-root record            (QLT::l2r_combine_kid_data,
//                                                 cedr_qlt.cpp:339-430)
//   TOP   UP + root_compute (cedr_qlt.cpp:441-476) + the block's down-sweep
//   DOWN  recompute the block's up-sweep, then solve every node problem from the
//         block root's solved mass down to the leaves (QLT::r2l_solve_qp,
//         cedr_qlt.cpp:490-604)
// The reference runs one launch per tree level over a slot-major AoS buffer
// (cedr_qlt.cpp:349-387, 538-566). Here leaf data are SoA with the cell index
// fastest, so a block's leaves are one contiguous, coalesced segment per field,
// and the whole subtree lives in shared memory between its up- and down-sweep.
#ifndef CEDR_B200_KERNELS_CUH
#define CEDR_B200_KERNELS_CUH

#include <cstdint>

#include "node_solve.cuh"

namespace cedr_b200 {

struct BlockDev {
  int leaf0, nl, ni, nlev;
  int lvlptr_off, kid_off, ibase;
  // Fast-path tables (fast_kernels.cuh); ftab_off < 0 if the shape is not fast.
  int ftab_off, fpair_off, fpos_off, npairs, fbase, fperm_off, fpent_off, fperm_up_off;
  // Index of this block among ALL blocks of its tier (= its leaf number in the next
  // tier). Equals its position in the device block list except in tier 0 of a
  // multi-rank run, where the list holds only the blocks this rank owns.
  int gidx;
};

// The same node constants, split and in the fast path's node order.
typedef dev::NodeWQ FastWQ;
typedef dev::NodeRh FastRh;

// Tracer classes: the six canonical QLT problem classes in the reference's
// order (cedr_qlt_inl.hpp:101-108), plus CAAS.
enum { CLS_ST = 0, CLS_CST = 1, CLS_T = 2, CLS_CT = 3, CLS_NN = 4, CLS_CNN = 5,
       CLS_CAAS = 6,
       // One scalar field of a BfbTreeAllReducer (cedr_bfb_tree_allreduce.cpp:78-159):
       // leaves as given, every node d = 0; d += kid0; d += kid1.
       CLS_BFB = 7, NCLS = 8 };
enum { MODE_UP = 0, MODE_TOP = 1, MODE_DOWN = 2 };

// Where a tracer's leaf rows live: row `role` (0 min, 1 Qm, 2 max, 3 prev) of tracer t is
// p[role] + trow[t]*ld. For the CDR's own buffer p[role] = in + (row offset of the role in
// the tracer's class)*ld and trow = trcr_row; for arrays bound by the caller
// (cedr_b200_bind_arrays) p[role] is the caller's array, trow[t] = t and ld its leading
// dimension. p[role] is null for roles the class does not have.
struct RowMap {
  const double* p[4];
  long long ld;
  const int* trow;
  __host__ __device__ const double* row (const int role, const int t) const {
    return p[role] + static_cast<long long>(trow[t])*ld;
  }
};

struct SweepArgs {
  const BlockDev* blocks;
  int nblocks;
  const int* lvlptr;
  const int* kid0;
  const int* kid1;
  const dev::NodeConst* nc;
  // Leaf input: tier 0 reads the caller-facing rows (row r of cell c at
  // in[r*in_ld + c], first row of tracer t = trcr_row[t]); higher tiers read the
  // records written by the tier below, field f of tracer t at
  // in[(4*t + f)*in_ld + leaf].
  const double* in;
  long long in_ld;
  int tier0;
  RowMap rows;          // tier 0
  const int* trcr_row;
  const int* trcr_prob;
  // UP: block-root records for the next tier, [(4*t + f)*rec_ld + block].
  double* rec_out;
  long long rec_ld;
  // DOWN: the next tier's solved masses, [t*sol_in_ld + block].
  const double* sol_in;
  long long sol_in_ld;
  // TOP/DOWN: solved leaf masses, [t*out_ld + leaf] (tier 0: the r2l buffer).
  double* out;
  long long out_ld;
  const int* tracers;   // tracer ids of this launch's class
  int ntr;
  int prefer_mass_con;
  double* qglob;        // [2*t], [2*t+1]: global q_min, q_max (consistent-only)
  double* caas_scal;    // [2*t]: mode (-1, 0, +1), [2*t+1]: fac
};

// Sweep of one block for one tracer by the whole CTA (any blockDim); `sm` holds
// 4*(nl + ni) doubles. Contains __syncthreads: every thread of the CTA must call it.
template <int CLS, int MODE>
__device__ __forceinline__ void
sweep_block (const SweepArgs& a, const int b, const int t, double* const sm,
             const dev::NodeConst* const nc_block = nullptr, double* const rh_solo = nullptr,
             dev::NodeConst* const nc_solo = nullptr, const BlockDev* const block_solo = nullptr,
             const int* const tabs_solo = nullptr) {
  // solo_kernel's extras. rh_solo / nc_solo: also form the block's rhom sums (in the same
  // level loop as the tracer's sums) and its node constants (one pass over all nodes).
  // block_solo: the block descriptor, passed as a kernel argument. tabs_solo: the block's
  // tree tables staged in shared memory by the caller, [kid0 (ni) | kid1 (ni) | lvlptr].
  constexpr bool caas = CLS == CLS_CAAS;
  constexpr bool bfb = CLS == CLS_BFB;
  constexpr bool nonneg = CLS == CLS_NN || CLS == CLS_CNN || bfb;   // one word: row 0
  constexpr bool consistent_only = CLS == CLS_T || CLS == CLS_CT;
  constexpr bool has_prev = CLS == CLS_CST || CLS == CLS_CT || CLS == CLS_CNN || caas;
  // Which fields this mode needs in shared memory. The down-sweep of the
  // bounded classes needs (min, Qm, max); consistent-only and nonnegative
  // classes need only Qm (their bounds are global q * rhom, resp. [0, b]).
  constexpr bool need_bounds = ! nonneg && (MODE != MODE_DOWN || ! consistent_only);
  constexpr bool need_prev = has_prev && MODE != MODE_DOWN && ! bfb;

  const BlockDev B = block_solo ? *block_solo : a.blocks[b];
  const int nn = B.nl + B.ni;
  const int tid = threadIdx.x, nth = blockDim.x;
  double* const f0 = sm;
  double* const f1 = f0 + nn;
  double* const f2 = f1 + nn;
  double* const f3 = f2 + nn; // Qm_prev sums on the way up, solved masses down

  { // Leaves.
    const double* p0, * p1, * p2, * p3;
    if (a.tier0) {
      p1 = a.rows.row(1, t) + B.leaf0;
      p3 = (need_prev || (caas && a.rows.p[3])) ? a.rows.row(3, t) + B.leaf0 : nullptr;
      p0 = p2 = nullptr;
      if ( ! nonneg) { p0 = a.rows.row(0, t) + B.leaf0; p2 = a.rows.row(2, t) + B.leaf0; }
    } else {
      const double* base = a.in + (long long) t*4*a.in_ld + B.leaf0;
      p0 = base; p1 = base + a.in_ld; p2 = base + 2*a.in_ld; p3 = base + 3*a.in_ld;
    }
    if (rh_solo)
      for (int i = tid; i < B.nl; i += nth) rh_solo[i] = a.in[B.leaf0 + i];   // row 0: rhom
    if (caas && a.tier0) {
      // CAAS::reduce_locally, cedr_caas.cpp:129-201 through calc_Qm_scalars
      // (cedr_caas_inl.hpp:44-57): clip, and the four summands. Each enters the
      // reduction as 0 + value (accumulator start, cedr_caas.cpp:145-151).
      const bool conserve = a.trcr_prob[t] & 1;
      // (Four elements per thread and pass: all their loads are in flight before the first
      // is used; a rolled loop pays one memory round trip per element.)
      for (int i0 = tid; i0 < B.nl; i0 += 4*nth) {
        double lo[4], q[4], hi[4], term[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + k*nth;
          if (i < B.nl) {
            lo[k] = __ldcg(p0 + i); q[k] = __ldcg(p1 + i); hi[k] = __ldcg(p2 + i);
            term[k] = conserve ? __ldcg(p3 + i) : q[k];
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + k*nth;
          if (i < B.nl) {
            const double clip = dev::rmin(hi[k], dev::rmax(lo[k], q[k]));
            f0[i] = 0.0 + lo[k];
            f1[i] = 0.0 + clip;
            f2[i] = 0.0 + hi[k];
            f3[i] = 0.0 + term[k];
          }
        }
      }
    } else {
      // Read-once data; for tiers >= 1 it may have been written by other SMs of the
      // same (fused) launch, so bypass L1.
      for (int i0 = tid; i0 < B.nl; i0 += 4*nth) {
        double v0[4], v1[4], v2[4], v3[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + k*nth;
          if (i < B.nl) {
            if (need_bounds) { v0[k] = __ldcg(p0 + i); v2[k] = __ldcg(p2 + i); }
            v1[k] = __ldcg(p1 + i);
            if (need_prev) v3[k] = __ldcg(p3 + i);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + k*nth;
          if (i < B.nl) {
            if (need_bounds) { f0[i] = v0[k]; f2[i] = v2[k]; }
            f1[i] = v1[k];
            if (need_prev) f3[i] = v3[k];
          }
        }
      }
    }
  }
  __syncthreads();

  // Up-sweep inside the block, level by level (kids before parents).
  const int* const lvlptr = tabs_solo ? tabs_solo + 2*B.ni : a.lvlptr + B.lvlptr_off;
  const int* const kid0 = tabs_solo ? tabs_solo : a.kid0 + B.kid_off;
  const int* const kid1 = tabs_solo ? tabs_solo + B.ni : a.kid1 + B.kid_off;
  for (int l = 0; l < B.nlev; ++l) {
    const int je = lvlptr[l+1];
    for (int j = lvlptr[l] + tid; j < je; j += nth) {
      const int k0 = kid0[j], k1 = kid1[j], me = B.nl + j;
      if (rh_solo) rh_solo[me] = rh_solo[k0] + rh_solo[k1];
      if (bfb) {
        f1[me] = (0.0 + f1[k0]) + f1[k1];
      } else if (caas) {
        // BfbTreeAllReducer: d = 0; d += kid0; d += kid1
        // (cedr_bfb_tree_allreduce.cpp:115-124).
        f0[me] = (0.0 + f0[k0]) + f0[k1];
        f1[me] = (0.0 + f1[k0]) + f1[k1];
        f2[me] = (0.0 + f2[k0]) + f2[k1];
        f3[me] = (0.0 + f3[k0]) + f3[k1];
      } else {
        if (need_bounds) {
          if (consistent_only) {
            f0[me] = dev::rmin(f0[k0], f0[k1]);
            f2[me] = dev::rmax(f2[k0], f2[k1]);
          } else {
            f0[me] = f0[k0] + f0[k1];
            f2[me] = f2[k0] + f2[k1];
          }
        }
        f1[me] = f1[k0] + f1[k1];
        if (need_prev) f3[me] = f3[k0] + f3[k1];
      }
    }
    __syncthreads();
  }
  const int root = B.ni ? nn - 1 : 0;
  if (nc_solo) {
    // Node constants (what rhom_kernel leaves in global memory for larger problems); the
    // barriers below order them before the down-sweep.
    for (int j = tid; j < B.ni; j += nth) {
      const double rh0 = rh_solo[kid0[j]], rh1 = rh_solo[kid1[j]];
      dev::NodeConst c;
      c.w0 = 1/rh0;
      c.w1 = 1/rh1;
      c.q0 = 1/c.w0;
      c.q1 = 1/c.w1;
      c.rh0 = rh0;
      c.rh1 = rh1;
      nc_solo[j] = c;
    }
  }

  if (MODE == MODE_UP) {
    if (tid == 0) {
      double* r = a.rec_out + (long long) t*4*a.rec_ld + B.gidx;
      if (need_bounds) { r[0] = f0[root]; r[2*a.rec_ld] = f2[root]; }
      r[a.rec_ld] = f1[root];
      if (need_prev) r[3*a.rec_ld] = f3[root];
    }
    return;
  }

  if (bfb) {
    // MODE_TOP: the reduced field is at the root.
    if (tid == 0) a.qglob[t] = f1[root];
    return;
  }
  if (caas) {
    // MODE_TOP: the four global sums are at the root. CAAS::finish_locally,
    // cedr_caas.cpp:211-253, scalar part.
    if (tid == 0) {
      const double clip_sum = f1[root], term_sum = f3[root];
      const double m = term_sum - clip_sum;
      double mode = 0, fac = 0;
      if (m < 0) {
        fac = clip_sum - f0[root];
        if (fac > 0) { fac = m/fac; mode = -1; }
      } else if (m > 0) {
        fac = f2[root] - clip_sum;
        if (fac > 0) { fac = m/fac; mode = 1; }
      }
      a.caas_scal[2*t] = mode;
      a.caas_scal[2*t+1] = fac;
    }
    return;
  }

  double qmin = 0, qmax = 0;
  if (MODE == MODE_TOP) {
    // root_compute, cedr_qlt.cpp:441-476.
    if (consistent_only) {
      qmin = f0[root];
      qmax = f2[root];
      if (tid == 0) { a.qglob[2*t] = qmin; a.qglob[2*t+1] = qmax; }
    }
    __syncthreads();
    if (tid == 0 && ! has_prev) f3[root] = f1[root];
  } else {
    if (consistent_only) { qmin = a.qglob[2*t]; qmax = a.qglob[2*t+1]; }
    if (tid == 0) f3[root] = a.sol_in[(long long) t*a.sol_in_ld + B.gidx];
  }
  __syncthreads();

  // Down-sweep: parents before kids.
  const dev::NodeConst* const nc = nc_block ? nc_block : a.nc + B.ibase;
  const bool prefer = a.prefer_mass_con != 0;
  for (int l = B.nlev - 1; l >= 0; --l) {
    const int je = lvlptr[l+1];
    for (int j = lvlptr[l] + tid; j < je; j += nth) {
      const int k0 = kid0[j], k1 = kid1[j], me = B.nl + j;
      const dev::NodeConst& c = nc[j];
      const double bm = f3[me];
      double x0, x1;
      if (nonneg) {
        dev::solve_node_nonneg(c, bm, f1[k0], f1[k1], x0, x1);
      } else if (consistent_only) {
        // cedr_qlt_inl.hpp:181-188: q bounds scaled by each node's rhom.
        const double rh0 = c.rh0, rh1 = c.rh1, rh = rh0 + rh1;
        dev::solve_node_bounded(c, prefer, qmin*rh, f1[me], qmax*rh, bm,
                                qmin*rh0, f1[k0], qmax*rh0,
                                qmin*rh1, f1[k1], qmax*rh1, x0, x1);
      } else {
        dev::solve_node_bounded(c, prefer, f0[me], f1[me], f2[me], bm,
                                f0[k0], f1[k0], f2[k0], f0[k1], f1[k1], f2[k1],
                                x0, x1);
      }
      f3[k0] = x0;
      f3[k1] = x1;
    }
    __syncthreads();
  }
  double* const o = a.out + (long long) t*a.out_ld + B.leaf0;
  for (int i = tid; i < B.nl; i += nth) o[i] = f3[i];
}

template <int CLS, int MODE>
__global__ void __launch_bounds__(256)
sweep_kernel (const SweepArgs a) {
  extern __shared__ double sm[];
  sweep_block<CLS, MODE>(a, blockIdx.x % a.nblocks, a.tracers[blockIdx.x / a.nblocks], sm);
}

// A whole run() in ONE launch for problems that are a single small block (the whole tree
// in one tier: cedr_test_1d_transport's 111 cells, the randomized unit test's trees): one
// CTA per tracer forms the rhom sums and node constants of the block in shared memory
// (what rhom_kernel leaves in global memory for larger problems; the same operations),
// sweeps the block up and down (QLT), or forms the four sums and adjusts the block's
// cells (CAAS). These problems are launch-latency bound: one launch instead of two.
// Shared memory: 4 (nl + ni) doubles for the sweep, then nl + ni doubles of rhom sums,
// then ni node constants.
template <int CLS>
__global__ void __launch_bounds__(256)
solo_kernel (const SweepArgs a, const BlockDev B) {
  extern __shared__ double sm[];
  const int t = a.tracers[blockIdx.x];
  const int nn = B.nl + B.ni, tid = threadIdx.x, nth = blockDim.x;
  double* const rh = sm + 4*nn;
  dev::NodeConst* const nc = reinterpret_cast<dev::NodeConst*>(rh + nn + (nn & 1));
  // The tree tables go to shared memory in one round trip, together with the leaf loads
  // (from global memory every level of the sweep would wait for its own).
  int* const tabs = reinterpret_cast<int*>(nc + B.ni);
  for (int i = tid; i < B.ni; i += nth) {
    tabs[i] = a.kid0[B.kid_off + i];
    tabs[B.ni + i] = a.kid1[B.kid_off + i];
  }
  for (int i = tid; i <= B.nlev; i += nth) tabs[2*B.ni + i] = a.lvlptr[B.lvlptr_off + i];
  if (CLS == CLS_CAAS) {
    sweep_block<CLS_CAAS, MODE_TOP>(a, 0, t, sm, nullptr, nullptr, nullptr, &B, tabs);
    __syncthreads();
    // CAAS::finish_locally (cedr_caas.cpp:211-253) on this tracer's cells.
    const double mode = a.caas_scal[2*t], fac = a.caas_scal[2*t+1];
    const double* const rlo = a.rows.row(0, t) + B.leaf0;
    const double* const rhi = a.rows.row(2, t) + B.leaf0;
    double* const rq = const_cast<double*>(a.rows.row(1, t)) + B.leaf0;
    for (int i = tid; i < B.nl; i += nth) {
      const double lo = rlo[i], hi = rhi[i];
      double q = dev::rmin(hi, dev::rmax(lo, rq[i]));
      if (mode < 0) { q += fac*(q - lo); q = dev::rmax(lo, q); }
      else if (mode > 0) { q += fac*(hi - q); q = dev::rmin(hi, q); }
      rq[i] = q;
    }
    return;
  }
  sweep_block<CLS, MODE_TOP>(a, 0, t, sm, nc, rh, nc, &B, tabs);
}

// rhom sweep: word 0 of every slot in the reference (cedr_qlt.cpp:356-360),
// summed kid0 + kid1 up the tree, plus the per-node constants of
// node_solve.cuh. One CTA per block.
struct RhomArgs {
  const BlockDev* blocks;
  int nblocks;
  const int* lvlptr;
  const int* kid0;
  const int* kid1;
  const double* in;    // this tier's leaf rhom
  double* root_out;    // next tier's leaf rhom, [block] (may be null at the top)
  dev::NodeConst* nc;
  const unsigned short* fpos;  // fast-path const positions (may be null)
  FastWQ* fwq;
  FastRh* frh;
  double* frq;                 // RN(1/(q0 + q1)) or 0, see node_solve.cuh div_by_qmass
  // rhom of the block's 2^split depth-`split` nodes, [gidx 2^split + j] (fast shapes; may
  // be null): the leaf rhom of the expanded tier above (FastArgs::split).
  double* sub_out;
  int split;
};

__global__ void __launch_bounds__(256)
rhom_kernel (const RhomArgs a) {
  extern __shared__ double sm[];
  const BlockDev B = a.blocks[blockIdx.x];
  const int nn = B.nl + B.ni;
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int i = tid; i < B.nl; i += nth) sm[i] = a.in[B.leaf0 + i];
  __syncthreads();
  const int* const lvlptr = a.lvlptr + B.lvlptr_off;
  const int* const kid0 = a.kid0 + B.kid_off;
  const int* const kid1 = a.kid1 + B.kid_off;
  for (int l = 0; l < B.nlev; ++l) {
    const int je = lvlptr[l+1];
    for (int j = lvlptr[l] + tid; j < je; j += nth) {
      const double rh0 = sm[kid0[j]], rh1 = sm[kid1[j]];
      sm[B.nl + j] = rh0 + rh1;
      dev::NodeConst c;
      c.w0 = 1/rh0;
      c.w1 = 1/rh1;
      c.q0 = 1/c.w0;
      c.q1 = 1/c.w1;
      c.rh0 = rh0;
      c.rh1 = rh1;
      a.nc[B.ibase + j] = c;
      if (a.fpos && B.ftab_off >= 0) {
        const int pos = B.fbase + a.fpos[B.fpos_off + j];
        FastWQ wq;
        wq.w0 = c.w0; wq.w1 = c.w1; wq.q0 = c.q0; wq.q1 = c.q1;
        a.fwq[pos] = wq;
        FastRh r;
        r.rh0 = rh0; r.rh1 = rh1;
        a.frh[pos] = r;
        a.frq[pos] = dev::reciprocal_for_div(c.q0 + c.q1);
        // Fast positions of depths < 9 are heap indices: depth d holds [2^d - 1, 2^(d+1) - 1).
        const int E = 1 << a.split, hp = a.fpos[B.fpos_off + j];
        if (a.sub_out && a.split && hp >= E - 1 && hp < 2*E - 1)
          a.sub_out[static_cast<long long>(B.gidx)*E + hp - (E - 1)] = rh0 + rh1;
      }
    }
    __syncthreads();
  }
  if (tid == 0 && a.root_out) a.root_out[B.gidx] = sm[B.ni ? nn - 1 : 0];
}

// Multi-rank exchange (replaces the per-level messages of cedr_qlt.cpp:327-337, 432-439
// and CAAS's MPI_Allreduce, cedr_caas.cpp:203-209): a rank's message carries, for each of
// its (at most nown_max) tier-0 blocks, the global block index, the rhom of the block root
// and the 4 nt words of its record. Layout: WORD-major, message[w*nown_max + j] for block
// j -- word 0 the index (-1 pads), word 1 + e*(4 nt + 1) + i the i-th value of sub-root e --
// so that consecutive threads move consecutive blocks: a rank's blocks are consecutive
// leaves of the tier above, which makes both the gather from / scatter to the record rows
// and the message accesses coalesced. After the exchange every rank scatters all entries
// into its (replicated) tier-1 leaf arrays.
__device__ __forceinline__ double
exchange_word (const BlockDev* blocks, const int nown, const int j, const long long w,
               const long long per, const int E, const double* rhom1, const double* rec,
               const long long rec_ld) {
  if (j >= nown) return w == 0 ? -1.0 : 0.0;
  const long long g = blocks[j].gidx;
  if (w == 0) return static_cast<double>(g);
  const long long e = (w - 1)/per, i = (w - 1) % per, leaf = g*E + e;
  return i == 0 ? (rhom1 ? rhom1[leaf] : 0.0) : rec[(i - 1)*rec_ld + leaf];
}

__global__ void __launch_bounds__(256)
pack_kernel (const BlockDev* blocks, const int nown, const int nown_max, const int nt,
             const int E, const double* rhom1, const double* rec, const long long rec_ld,
             double* send) {
  const long long per = 4LL*nt + 1, nw = 1 + E*per, n = nw*nown_max;
  for (long long k = blockIdx.x*(long long) blockDim.x + threadIdx.x; k < n;
       k += (long long) gridDim.x*blockDim.x) {
    const int j = (int) (k % nown_max);
    send[k] = exchange_word(blocks, nown, j, k / nown_max, per, E, rhom1, rec, rec_ld);
  }
}

// The same message, stored straight into every rank's receive buffer over peer-mapped
// memory (NVLink): rank `me` owns slot `me` of each peer's rank-major buffer. Replaces
// pack + all-gather; p2p_barrier_kernel then publishes and awaits the epoch. blockIdx.y is
// the peer: the gather is repeated per peer (L2 hits) so that the stores to the eight peers
// go out in parallel rather than one after the other from each thread.
struct PeerPtrs { double* recv[16]; unsigned long long* flags[16]; };

__global__ void __launch_bounds__(256)
pack_p2p_kernel (const BlockDev* blocks, const int nown, const int nown_max, const int nt,
                 const int E, const double* rhom1, const double* rec, const long long rec_ld,
                 const PeerPtrs peers, const int me, const int nranks,
                 const long long rank_stride) {
  const long long per = 4LL*nt + 1, nw = 1 + E*per, n = nw*nown_max;
  double* const dst = peers.recv[blockIdx.y] + me*rank_stride;
  for (long long k = blockIdx.x*(long long) blockDim.x + threadIdx.x; k < n;
       k += (long long) gridDim.x*blockDim.x) {
    const int j = (int) (k % nown_max);
    dst[k] = exchange_word(blocks, nown, j, k / nown_max, per, E, rhom1, rec, rec_ld);
  }
  // (No fence here: the kernel boundary orders these stores before p2p_barrier_kernel,
  // whose system-scope release publishes them together with the epoch flag.)
}

// Epoch barrier across the ranks' GPUs: publish "rank me has delivered epoch e" into every
// peer's flag array, then wait until every rank has delivered to this one. One warp; lane
// r talks to rank r. The ranks are different devices, each running its own stream, so the
// wait cannot starve the writer (unlike two spinning kernels on ONE device).
// The epoch is counted on the device (*epoch_ctr, one more per barrier) so that a captured
// CUDA graph of run() can be replayed: the host only has to know its parity.
__global__ void p2p_barrier_kernel (const PeerPtrs peers, unsigned long long* my_flags,
                                    const int me, const int nranks,
                                    unsigned long long* epoch_ctr, int* status,
                                    const unsigned long long timeout_ns) {
  const int r = threadIdx.x;
  const unsigned long long epoch = *epoch_ctr + 1;
  __syncwarp();
  if (r == 0) *epoch_ctr = epoch;
  __threadfence_system();
  if (r < nranks) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(peers.flags[r] + me), "l"(epoch)
                 : "memory");
    unsigned long long v = 0, t0 = 0;
    unsigned spins = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(my_flags + r)
                   : "memory");
      if (v >= epoch) break;
      __nanosleep(100);
      // A rank that is late (first step, I/O, load imbalance) is waited for; one that is
      // gone must not hang the device: give up after timeout_ns (CEDR_B200_P2P_TIMEOUT_MS,
      // default 30 s) and let poison_kernel mark this run's results.
      if ((++spins & 1023u) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > timeout_ns) { atomicExch(status, 2); break; }
      }
    }
  }
  __threadfence_system();
}

// After a run() whose exchange gave up: the results were formed from stale or partial peer
// data -- overwrite them with NaN so that no caller mistakes them for an answer (the error
// itself is reported by cedr_b200_synchronize). Costs one load per thread otherwise.
__global__ void __launch_bounds__(256)
poison_kernel (const int* status, double* out, const long long ld, const int ncells,
               const int nrows, const int row0, const int row_stride) {
  if (*status == 0) return;
  const long long n = static_cast<long long>(ncells)*nrows;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (long long k = blockIdx.x*static_cast<long long>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<long long>(gridDim.x)*blockDim.x)
    out[(row0 + (k/ncells)*row_stride)*ld + k % ncells] = nan;
}

__global__ void __launch_bounds__(256)
unpack_kernel (const double* recv, const int nranks, const int nown_max, const int nt,
               const int E, double* rhom1, double* rec, const long long rec_ld,
               const long long rank_stride) {
  const long long per = 4LL*nt + 1, nw = 1 + E*per, nper = nw*nown_max;
  const long long n = nper*nranks;
  for (long long k = blockIdx.x*(long long) blockDim.x + threadIdx.x; k < n;
       k += (long long) gridDim.x*blockDim.x) {
    const long long r = k / nper, kk = k % nper;
    const double* const msg = recv + r*rank_stride;
    const long long w = kk / nown_max;
    const int j = (int) (kk % nown_max);
    const long long g = (long long) msg[j];
    if (g < 0 || w == 0) continue;
    const long long e = (w - 1)/per, i = (w - 1) % per, leaf = g*E + e;
    if (i == 0) { if (rhom1) rhom1[leaf] = msg[kk]; }
    else rec[(i - 1)*rec_ld + leaf] = msg[kk];
  }
}

// General cell -> rank maps: where a rank's cells are not whole blocks of the
// tree plan -- the reference's pseudorandom test map rank = (ci + ci/nranks) % nranks,
// cedr_tree.cpp:366-375, cuts every block -- every rank gathers every rank's rows, sweeps
// the WHOLE tree as the one-rank path does, and keeps its own cells (replicated mode).
// A rank's message: row-major [nrows][nlcl_max], zero padded.
__global__ void __launch_bounds__(256)
repl_pack_kernel (const double* in, const long long ld, const int nlcl, const int nrows,
                  const int nlcl_max, double* send) {
  const long long n = static_cast<long long>(nrows)*nlcl_max;
  for (long long k = blockIdx.x*static_cast<long long>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<long long>(gridDim.x)*blockDim.x) {
    const long long row = k/nlcl_max;
    const int i = static_cast<int>(k % nlcl_max);
    send[k] = i < nlcl ? in[row*ld + i] : 0.0;
  }
}

// Cell i of rank r is leaf pos[pos_off[r] + i] of the whole tree (DFS order).
__global__ void __launch_bounds__(256)
repl_unpack_kernel (const double* recv, const int nranks, const int nrows, const int nlcl_max,
                    const int* pos, const int* pos_off, double* whole_in,
                    const long long whole_ld) {
  const long long per = static_cast<long long>(nrows)*nlcl_max, n = per*nranks;
  for (long long k = blockIdx.x*static_cast<long long>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<long long>(gridDim.x)*blockDim.x) {
    const int r = static_cast<int>(k/per);
    const long long kk = k % per, row = kk/nlcl_max;
    const int i = static_cast<int>(kk % nlcl_max);
    if (i < pos_off[r + 1] - pos_off[r]) whole_in[row*whole_ld + pos[pos_off[r] + i]] = recv[k];
  }
}

__global__ void __launch_bounds__(256)
repl_keep_own_kernel (const double* whole_out, const long long whole_ld, const int* pos_me,
                      const int nlcl, const int nt, double* out, const long long ld) {
  const long long n = static_cast<long long>(nlcl)*nt;
  for (long long k = blockIdx.x*static_cast<long long>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<long long>(gridDim.x)*blockDim.x) {
    const long long t = k/nlcl;
    const int i = static_cast<int>(k % nlcl);
    out[t*ld + i] = whole_out[t*whole_ld + pos_me[i]];
  }
}

// CAAS::finish_locally, cedr_caas.cpp:211-253, per-cell part; also writes the
// clipped value the reference stores in place during reduce_locally
// (cedr_caas.cpp:177).
__global__ void __launch_bounds__(256)
caas_adjust_kernel (const RowMap rows, const int ncells, const double* scal, const int nt) {
  const long long n = (long long) ncells*nt;
  for (long long k = blockIdx.x*(long long) blockDim.x + threadIdx.x; k < n;
       k += (long long) gridDim.x*blockDim.x) {
    const int t = (int) (k / ncells), i = (int) (k % ncells);
    double* const rq = const_cast<double*>(rows.row(1, t)) + i;
    const double lo = rows.row(0, t)[i], hi = rows.row(2, t)[i];
    double q = dev::rmin(hi, dev::rmax(lo, rq[0]));
    const double mode = scal[2*t], fac = scal[2*t+1];
    if (mode < 0) {
      q += fac*(q - lo);
      q = dev::rmax(lo, q);
    } else if (mode > 0) {
      q += fac*(hi - q);
      q = dev::rmin(hi, q);
    }
    rq[0] = q;
  }
}

// CAAS with the reference's own host summation order (CEDR_B200_CAAS_SUM_SEQUENTIAL):
// reduce_locally on a host backend runs each tracer's cells i = 0..n-1 through one
// accumulator per sum (team size 1, cedr_caas.cpp:171-199, cedr_kokkos.hpp:118), so the
// four sums are ((0 + v0) + v1) + ... One thread per tracer walks its rows in that order
// (a compatibility mode for one rank: bit parity with the reference's default CAAS, not
// speed), then forms the redistribution scalars of finish_locally (cedr_caas.cpp:211-253).
__global__ void __launch_bounds__(128)
caas_seq_sums_kernel (const RowMap rows, const int ncells, const int* trcr_prob, const int nt,
                      double* scal) {
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const double* const rlo = rows.row(0, t);
  const double* const rq = rows.row(1, t);
  const double* const rhi = rows.row(2, t);
  const bool conserve = trcr_prob[t] & 1;
  const double* const rp = conserve ? rows.row(3, t) : rq;
  double clip_sum = 0, term_sum = 0, min_sum = 0, max_sum = 0;
  for (int i = 0; i < ncells; ++i) {
    const double lo = rlo[i], q = rq[i], hi = rhi[i];
    const double term = rp[i];
    clip_sum += dev::rmin(hi, dev::rmax(lo, q));
    term_sum += term;
    min_sum += lo;
    max_sum += hi;
  }
  const double m = term_sum - clip_sum;
  double mode = 0, fac = 0;
  if (m < 0) {
    fac = clip_sum - min_sum;
    if (fac > 0) { fac = m/fac; mode = -1; }
  } else if (m > 0) {
    fac = max_sum - clip_sum;
    if (fac > 0) { fac = m/fac; mode = 1; }
  }
  scal[2*t] = mode;
  scal[2*t+1] = fac;
}

// CAAS with the caller's UserAllReducer (CEDR_B200_CAAS_SUM_USER): reduce_locally's
// user-reducer branch (cedr_caas.cpp:140-168). One thread per (tracer k, block bi of
// n_accum cells): clips the block's cells in place and accumulates, each sum starting at
// 0 and adding cell after cell as the reference's lambdas do, the four partial sums
//   send[nlocal*k + bi] (clip), [nlocal*(nt + k) + bi] (term),
//   send[nlocal*(2 nt + k) + bi] (min), [nlocal*(3 nt + k) + bi] (max).
__global__ void __launch_bounds__(256)
caas_user_partials_kernel (const RowMap rows, const int nlocal, const int n_accum,
                           const int* trcr_prob, const int nt, double* send) {
  const long long n = static_cast<long long>(nlocal)*nt;
  for (long long j = blockIdx.x*static_cast<long long>(blockDim.x) + threadIdx.x; j < n;
       j += static_cast<long long>(gridDim.x)*blockDim.x) {
    const int k = static_cast<int>(j / nlocal), bi = static_cast<int>(j % nlocal);
    const double* const rlo = rows.row(0, k);
    double* const rq = const_cast<double*>(rows.row(1, k));
    const double* const rhi = rows.row(2, k);
    const bool conserve = trcr_prob[k] & 1;
    const double* const rp = conserve ? rows.row(3, k) : rq;
    double a_clip = 0, a_term = 0, a_min = 0, a_max = 0;
    for (int ai = 0; ai < n_accum; ++ai) {
      const long long i = static_cast<long long>(n_accum)*bi + ai;
      const double lo = rlo[i], q = rq[i], hi = rhi[i];
      const double term = rp[i];
      const double clip = dev::rmin(hi, dev::rmax(lo, q));
      rq[i] = clip;
      a_clip += clip;
      a_term += term;
      a_min += lo;
      a_max += hi;
    }
    const long long nl = nlocal;
    send[nl*k + bi] = a_clip;
    send[nl*(nt + k) + bi] = a_term;
    send[nl*(2LL*nt + k) + bi] = a_min;
    send[nl*(3LL*nt + k) + bi] = a_max;
  }
}

// The scalar part of finish_locally (cedr_caas.cpp:211-227) from the reducer's output
// recv = [sum clip (nt) | sum term (nt) | sum min (nt) | sum max (nt)].
__global__ void __launch_bounds__(128)
caas_scal_from_recv_kernel (const double* recv, const int nt, double* scal) {
  const int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const double clip_sum = recv[t], term_sum = recv[nt + t];
  const double m = term_sum - clip_sum;
  double mode = 0, fac = 0;
  if (m < 0) {
    fac = clip_sum - recv[2*nt + t];
    if (fac > 0) { fac = m/fac; mode = -1; }
  } else if (m > 0) {
    fac = recv[3*nt + t] - clip_sum;
    if (fac > 0) { fac = m/fac; mode = 1; }
  }
  scal[2*t] = mode;
  scal[2*t + 1] = fac;
}

// Bulk DeviceOp::set_Qm (cedr_qlt_inl.hpp:21-58, cedr_caas_inl.hpp:21-34) from
// SoA caller arrays a[t*lda + lci].
// BfbTreeAllReducer leaf fill: send is (nfield fastest, nlocal) unless transpose, then
// (nlocal fastest, nfield) (cedr_bfb_tree_allreduce.cpp:92-103); row j of `in` is field j.
__global__ void __launch_bounds__(256)
bfb_fill_kernel (double* in, const long long ld, const int nlocal, const int nfield,
                 const int transpose, const double* send) {
  const long long n = (long long) nlocal*nfield;
  for (long long k = blockIdx.x*(long long) blockDim.x + threadIdx.x; k < n;
       k += (long long) gridDim.x*blockDim.x) {
    const int j = (int) (k / nlocal), id = (int) (k % nlocal);
    in[(long long) j*ld + id] = transpose ? send[(long long) nlocal*j + id]
                                          : send[(long long) id*nfield + j];
  }
}

__global__ void __launch_bounds__(256)
set_qm_bulk_kernel (double* in, const long long ld, const int ncells,
                    const int* trcr_row, const int* trcr_prob, const int t0,
                    const int nt, const long long lda, const double* qm,
                    const double* qm_min, const double* qm_max, const double* qm_prev,
                    const int caas_store_prev) {
  const long long n = (long long) ncells*nt;
  for (long long k = blockIdx.x*(long long) blockDim.x + threadIdx.x; k < n;
       k += (long long) gridDim.x*blockDim.x) {
    const int tl = (int) (k / ncells), i = (int) (k % ncells), t = t0 + tl;
    const long long s = (long long) tl*lda + i;
    const int pt = trcr_prob[t];
    double* bd = in + (long long) trcr_row[t]*ld + i;
    int next;
    if (pt & 2) {
      bd[0] = qm_min[s]; bd[ld] = qm[s]; bd[2*ld] = qm_max[s]; next = 3;
    } else if (pt & 4) {
      const double rhom = in[i];
      bd[0] = qm_min[s]/rhom; bd[ld] = qm[s]; bd[2*ld] = qm_max[s]/rhom; next = 3;
    } else {
      bd[0] = qm[s]; next = 1;
    }
    if ((pt & 1) || caas_store_prev)
      bd[next*ld] = qm_prev ? qm_prev[s] : __longlong_as_double(0x7ff0000000000000LL);
  }
}

__global__ void __launch_bounds__(256)
get_qm_bulk_kernel (const double* src, const long long ld, const int ncells,
                    const int* trcr_row, const int is_caas, const int t0,
                    const int nt, const long long lda, double* qm) {
  const long long n = (long long) ncells*nt;
  for (long long k = blockIdx.x*(long long) blockDim.x + threadIdx.x; k < n;
       k += (long long) gridDim.x*blockDim.x) {
    const int tl = (int) (k / ncells), i = (int) (k % ncells), t = t0 + tl;
    qm[(long long) tl*lda + i] =
      is_caas ? src[((long long) trcr_row[t] + 1)*ld + i] : src[(long long) t*ld + i];
  }
}

// ---- the 1-D transport harness on the device (cedr_test_1d_transport.cpp:136-255): one
// semi-Lagrangian step = periodic cubic interpolation at the departure points
// (interp::cubic_interp_periodic, :38-85, with get_cubic, :24-36), the caller side of
// run_cdr (:170-189: bounds from the domain of dependence, set_Qm) -- fused in one kernel
// -- then CDR::run, then get_Qm / area (t1d_get_kernel). Tracer 0 of the CDR.
struct T1dArgs {
  int ncells;
  const double* xcp;    // [ncells + 1] cell centres, the last one periodic image of the first
  const double* area;   // [ncells]
  const int* tgt_i;     // [ncells + 1] interval of each (wrapped) departure point
  const double* tgt_xp; // [ncells + 1] the wrapped departure point
  const int* lci;       // [ncells] cell -> local cell index of the CDR
  double* in;           // the CDR's rows
  long long ld;
  int row;              // first row of tracer 0
  int layout;           // 0: (min, Qm, max, prev) rows; 1: nonnegative (Qm, prev)
  const double* out;    // results: QLT out row 0 / CAAS Qm row
};

// Point j of one step (j = ncells is the periodic image: interpolated, not set).
__device__ __forceinline__ void
t1d_interp_set_point (const T1dArgs& a, const double* const y, double* const yi, const int j) {
  const int nc = a.ncells;
  const double* const x = a.xcp;
  const int i = a.tgt_i[j], ip1 = i + 1;
  auto slope = [&] (const int k) { return (y[k + 1] - y[k])/(x[k + 1] - x[k]); };
  const double smid = slope(i);
  double s1, s2;
  if (i == 0) {
    const double w = (x[nc] - x[nc - 1])/((x[1] - x[0]) + (x[nc] - x[nc - 1]));
    s1 = (1 - w)*slope(nc - 1) + w*smid;
  } else {
    const double w = (x[i] - x[i - 1])/(x[ip1] - x[i - 1]);
    s1 = (1 - w)*slope(i - 1) + w*smid;
  }
  if (i == nc - 1) {
    const double w = (x[ip1] - x[i])/((x[ip1] - x[i]) + (x[1] - x[0]));
    s2 = (1 - w)*smid + w*slope(0);
  } else {
    const double w = (x[ip1] - x[i])/(x[i + 2] - x[i]);
    s2 = (1 - w)*smid + w*slope(ip1);
  }
  const double dx = x[ip1] - x[i], dx2 = dx*dx, dx3 = dx2*dx, den = -dx3;
  const double c2 = s1, c3 = y[i];
  const double b1 = y[ip1] - dx*c2 - c3, b2 = s2 - c2;
  const double c0 = (2.0*b1 - dx*b2)/den, c1 = (-3.0*dx*b1 + dx2*b2)/den;
  const double xij = a.tgt_xp[j] - x[i];
  const double v = (((c0*xij + c1)*xij) + c2)*xij + c3;
  yi[j] = v;
  if (j == nc) return;
  // run_cdr: bounds over the four cells of the domain of dependence.
  double mn = y[(i - 1 + nc) % nc], mx = mn;
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    const double u = y[(i - 1 + k + nc) % nc];
    mn = u < mn ? u : mn;
    mx = mx < u ? u : mx;
  }
  const double ar = a.area[j];
  double* const r = a.in + static_cast<long long>(a.row)*a.ld + a.lci[j];
  if (a.layout == 0) {
    r[0] = mn*ar; r[a.ld] = v*ar; r[2*a.ld] = mx*ar; r[3*a.ld] = y[j]*ar;
  } else {
    r[0] = v*ar; r[a.ld] = y[j]*ar;
  }
}

__global__ void __launch_bounds__(128)
t1d_interp_set_kernel (const T1dArgs a, const double* __restrict__ y, double* __restrict__ yi) {
  const int j = blockIdx.x*blockDim.x + threadIdx.x;
  if (j > a.ncells) return;
  t1d_interp_set_point(a, y, yi, j);
}

__global__ void __launch_bounds__(128)
t1d_get_kernel (const T1dArgs a, double* __restrict__ yi) {
  const int j = blockIdx.x*blockDim.x + threadIdx.x;
  if (j > a.ncells) return;
  const int c = j == a.ncells ? 0 : j;
  yi[j] = a.out[a.lci[c]]/a.area[c];
}

// The whole cycle in ONE launch, for problems solo_kernel takes (a single block of at most
// 256 leaves) with one tracer: a persistent CTA keeps the two time levels of y in shared
// memory and runs, per step, the same three phases -- t1d_interp_set_point for every point,
// solo_kernel's body (rhom sums and node constants, the block's sweeps, or CAAS's sums and
// adjustment), the read-back of t1d_get_kernel -- separated by block barriers instead of
// kernel boundaries. The many-tiny-calls pattern is launch-latency bound on a GPU; this
// form pays the latency once per cycle. Same operations on the same operands: the bits of
// the three-launch form. Shared memory: 2 x (ncells + 1) doubles (padded to even), then
// solo_kernel's layout.
template <int CLS>
__global__ void __launch_bounds__(256)
t1d_cycle_kernel (const SweepArgs a, const BlockDev B, const T1dArgs ta,
                  const double* const y_in, double* const y_out, const int nsteps) {
  extern __shared__ double sm_all[];
  const int ncp = ta.ncells + 1, ncp2 = ncp + (ncp & 1);
  double* const ys0 = sm_all;
  double* const ys1 = ys0 + ncp2;
  double* const sm = ys1 + ncp2;
  const int t = a.tracers[0];
  const int nn = B.nl + B.ni, tid = threadIdx.x, nth = blockDim.x;
  double* const rh = sm + 4*nn;
  dev::NodeConst* const nc = reinterpret_cast<dev::NodeConst*>(rh + nn + (nn & 1));
  int* const tabs = reinterpret_cast<int*>(nc + B.ni);
  for (int i = tid; i < B.ni; i += nth) {
    tabs[i] = a.kid0[B.kid_off + i];
    tabs[B.ni + i] = a.kid1[B.kid_off + i];
  }
  for (int i = tid; i <= B.nlev; i += nth) tabs[2*B.ni + i] = a.lvlptr[B.lvlptr_off + i];
  for (int j = tid; j < ncp; j += nth) ys0[j] = y_in[j];
  __syncthreads();
  for (int s = 0; s < nsteps; ++s) {
    const double* const y = (s & 1) ? ys1 : ys0;
    double* const yi = (s & 1) ? ys0 : ys1;
    for (int j = tid; j < ncp; j += nth) t1d_interp_set_point(ta, y, yi, j);
    __syncthreads();   // the rows in global memory, written and read by this CTA alone
    if (CLS == CLS_CAAS) {
      sweep_block<CLS_CAAS, MODE_TOP>(a, 0, t, sm, nullptr, nullptr, nullptr, &B, tabs);
      __syncthreads();
      // CAAS::finish_locally (cedr_caas.cpp:211-253), as in solo_kernel.
      const double mode = a.caas_scal[2*t], fac = a.caas_scal[2*t+1];
      const double* const rlo = a.rows.row(0, t) + B.leaf0;
      const double* const rhi = a.rows.row(2, t) + B.leaf0;
      double* const rq = const_cast<double*>(a.rows.row(1, t)) + B.leaf0;
      for (int i = tid; i < B.nl; i += nth) {
        const double lo = rlo[i], hi = rhi[i];
        double q = dev::rmin(hi, dev::rmax(lo, rq[i]));
        if (mode < 0) { q += fac*(q - lo); q = dev::rmax(lo, q); }
        else if (mode > 0) { q += fac*(hi - q); q = dev::rmin(hi, q); }
        rq[i] = q;
      }
    } else {
      // rhom does not change over the cycle: its sums and the node constants are formed in
      // the first step and stay in shared memory.
      sweep_block<CLS, MODE_TOP>(a, 0, t, sm, nc, s == 0 ? rh : nullptr,
                                 s == 0 ? nc : nullptr, &B, tabs);
    }
    __syncthreads();
    // get_Qm: QLT's solved leaf masses are still in shared memory (sweep_block's f3, what
    // it has just stored to `out`); CAAS's are in its Qm row.
    const double* const res = CLS == CLS_CAAS ? ta.out : sm + 3*nn - B.leaf0;
    for (int j = tid; j < ncp; j += nth) {
      const int c = j == ta.ncells ? 0 : j;
      yi[j] = res[ta.lci[c]]/ta.area[c];
    }
    __syncthreads();
  }
  const double* const yf = (nsteps & 1) ? ys1 : ys0;
  for (int j = tid; j < ncp; j += nth) y_out[j] = yf[j];
}

// Synthetic workload of SURVEY.md 8(d): splitmix64, U = (z >> 11) * 2^-53.
__device__ __forceinline__ double splitmix_u (const unsigned long long seed,
                                              const unsigned long long k) {
  unsigned long long z = seed + (k + 1ull)*0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30))*0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27))*0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double) (z >> 11)*0x1.0p-53;
}

__global__ void __launch_bounds__(256)
fill_headline_kernel (const int ncells, const int nt, const long long lda,
                      const unsigned long long seed, double* rhom, double* qm_min,
                      double* qm, double* qm_max, double* qm_prev, const int cell0,
                      const int nlcl) {
  // Cells [cell0, cell0 + nlcl) of the ncells-cell workload.
  const long long n = (long long) nlcl*nt;
  for (long long kk = blockIdx.x*(long long) blockDim.x + threadIdx.x; kk < n;
       kk += (long long) gridDim.x*blockDim.x) {
    const int t = (int) (kk / nlcl), il = (int) (kk % nlcl), i = cell0 + il;
    const long long k = (long long) t*ncells + i;
    const double rh = 0.5*(1 + splitmix_u(seed, i));
    if (t == 0) rhom[il] = rh;
    const unsigned long long p = (unsigned long long) ncells + 4ull*k;
    const double q_min = 0.1*splitmix_u(seed, p);
    const double q_max = q_min + splitmix_u(seed, p + 1);
    const double q = q_min + (q_max - q_min)*(1.4*splitmix_u(seed, p + 2) - 0.2);
    const double q_prev = q_min + (q_max - q_min)*splitmix_u(seed, p + 3);
    const long long s = (long long) t*lda + il;
    qm_min[s] = q_min*rh;
    qm_max[s] = q_max*rh;
    qm[s] = q*rh;
    qm_prev[s] = q_prev*rh;
  }
}

} // namespace cedr_b200

#endif
