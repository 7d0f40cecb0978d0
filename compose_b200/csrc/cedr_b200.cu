// C ABI of the B200-native CEDR hot path (see include/cedr_b200.h) and the host
// orchestration of run(): which kernels are launched, in which order, on which
// buffers. Host code is plain C++; all arithmetic on tracer data happens in the
// CUDA kernels of kernels.cuh. There is no CPU fallback.
#include "cedr_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "kernels.cuh"
#include "fast_kernels.cuh"
#include "ring_kernels.cuh"
#include "transposed_kernels.cuh"
#include "cluster_caas.cuh"
#include "cedr_b200_local.hpp"
#include "tree_plan.h"

namespace {

using namespace cedr_b200;

thread_local std::string g_err;

// Same message shape as the reference's cedr_throw_if (cedr_util.hpp:70-77).
#define cedr_b200_throw_if(condition, message) do {                     \
    if (condition) {                                                    \
      std::stringstream _ss_;                                           \
      _ss_ << __FILE__ << ":" << __LINE__ << ": The condition:\n"       \
           << #condition "\nled to the exception\n" << message << "\n"; \
      throw std::logic_error(_ss_.str());                               \
    }                                                                   \
  } while (0)

#define CUDA_CHECK(call) do {                                           \
    const cudaError_t _e_ = (call);                                     \
    if (_e_ != cudaSuccess) {                                           \
      std::stringstream _ss_;                                           \
      _ss_ << __FILE__ << ":" << __LINE__ << ": CUDA error in " #call ": " \
           << cudaGetErrorString(_e_);                                  \
      throw std::runtime_error(_ss_.str());                             \
    }                                                                   \
  } while (0)

template <typename F> int guarded (F&& f) {
  try {
    f();
    return 0;
  } catch (const std::logic_error& e) {
    g_err = e.what();
    return 1;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 2;
  }
}

template <typename T> struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf () {}
  DevBuf (const DevBuf&) = delete;
  DevBuf& operator= (const DevBuf&) = delete;
  ~DevBuf () { release(); }
  void release () { if (p) cudaFree(p); p = nullptr; n = 0; }
  void alloc (size_t n_) {
    release();
    n = n_;
    if (n) CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&p), n*sizeof(T)));
  }
  void upload (const std::vector<T>& h) {
    alloc(h.size());
    if (n) CUDA_CHECK(cudaMemcpy(p, h.data(), n*sizeof(T), cudaMemcpyHostToDevice));
  }
};

long long round_up (long long x, long long m) { return (x + m - 1)/m*m; }

// QLT::MetaData::get_problem_type_idx, cedr_qlt.cpp:85-96.
int qlt_class_of (int mask) {
  enum { c = 1, s = 2, t = 4, n = 8 };
  switch (mask) {
  case s: case s | t: return CLS_ST;
  case c | s: case c | s | t: return CLS_CST;
  case t: return CLS_T;
  case c | t: return CLS_CT;
  case n: return CLS_NN;
  case c | n: return CLS_CNN;
  default: return -1;
  }
}
// cedr_qlt_inl.hpp:101-108
int qlt_canonical_type (int cls) {
  static const int pt[] = {2 | 4, 1 | 2 | 4, 4, 1 | 4, 8, 1 | 8};
  return pt[cls];
}
// cedr_qlt_inl.hpp:109-117
int qlt_l2r_words (int cls) {
  static const int w[] = {3, 4, 3, 4, 1, 2};
  return w[cls];
}

const int kThreads = 256;

} // namespace

struct cedr_b200_cdr {
  bool is_caas = false;
  bool is_bfb = false;          // a BfbTreeAllReducer: "tracers" are the fields
  int rank = 0, nranks = 1;
  bool prefer_mass_con = false;
  int caas_sum_mode = CEDR_B200_CAAS_SUM_TREE;
  bool caas_need_conserve = false;
  // Caller arrays bound in place of the `in` / `out` rows (cedr_b200_bind_arrays).
  struct Bound {
    bool on = false;
    long long lda = 0;
    const double* qm_min = nullptr;
    double* qm = nullptr;
    const double* qm_max = nullptr;
    const double* qm_prev = nullptr;
    double* qm_out = nullptr;
  } bound;
  DevBuf<int> d_ident;          // 0, 1, 2, ...: tracer -> row of a bound array
  DevBuf<const double*> d_rowaddr;   // [4 nt]: row_map() per tracer and role, for the fast kernels
  // CAAS::UserAllReducer (CEDR_B200_CAAS_SUM_USER)
  cedr_b200_user_reducer_fn user_reducer = nullptr;
  void* user_reducer_ctx = nullptr;
  int user_naccum = 1;
  double* usend = nullptr;      // (nlocal, 4 nt), then recv (4 nt): in buf2 if the caller set it
  double* urecv = nullptr;

  // Tree (QLT: the caller's; CAAS: bisection over the cells, for the ordered sums).
  int ncells = 0;           // global
  int nlcl = 0;             // owned by this rank (== ncells on one rank)
  std::vector<int> tree_kids;
  std::vector<int64_t> tree_cellidx;
  std::vector<int> tree_rank;
  int tree_root = 0;
  int max_block_leaves = 1024;
  Plan plan;
  std::unordered_map<int64_t,int> gci2lci;
  std::vector<int> leaf_lci;   // DFS leaf index -> local cell index (-1: not owned)
  std::vector<int> own_blocks; // tier-0 blocks this rank owns (all of them on one rank)
  int nown_max = 0;            // blocks per rank in the exchange message (padded)
  std::string partition_error; // why this rank's cells are not a subtree partition
  // Replicated mode for cell -> rank maps that cut blocks (kernels.cuh repl_*): a one-rank
  // CDR over the whole tree, fed by an all-gather of every rank's rows.
  bool repl = false;
  std::unique_ptr<cedr_b200_cdr> whole;
  int nlcl_max = 0;
  DevBuf<int> d_pos, d_pos_off;
  int pos_me_off = 0;
  int64_t caas_cell0 = 0;

  // Tracers.
  bool declaring = true;
  std::vector<int> trcr_prob;  // canonical type (QLT) / declared type (CAAS)
  std::vector<int> trcr_cls;
  std::vector<int> trcr_row;
  int nrows = 0;               // rows of the `in` buffer, incl. the rhom row
  std::vector<int> cls_tracers[NCLS];

  // Buffers.
  long long ld = 0;
  bool finished = false;
  bool user_buffers = false;
  double* in = nullptr;
  double* out = nullptr;
  DevBuf<double> in_own, out_own, usend_own;
  DevBuf<int> d_trcr_row, d_trcr_prob, d_cls_tracers[NCLS];
  DevBuf<int> d_lvlptr, d_kid0, d_kid1;
  DevBuf<unsigned short> d_dtab, d_ptab, d_fpos, d_perm;
  DevBuf<unsigned> d_pent;
  DevBuf<FastWQ> d_fwq;
  DevBuf<FastRh> d_frh;
  DevBuf<double> d_frq;
  bool fast_enabled = true;   // cedr_b200_set_fast_path
  BlockDev solo_block;        // tier 0's only block, when run() is the solo kernel
  bool fast_ok = false;       // plan + buffers allow the fast tier-0 kernels
  DevBuf<unsigned long long> d_phase_clk;   // debug (CEDR_B200_PHASE_CLOCKS builds)
  DevBuf<double> d_n7;        // depth-7 sums per own block x tracer (fast path)
  DevBuf<double> d_x7;        // depth-7 masses per own block x tracer (transposed_kernels.cuh)
  // Expanded tier above the fast blocks (FastArgs::split): its leaves are the 2^split
  // depth-`split` nodes of every tier-0 block; one block, swept by the generic kernels.
  int split = 0;
  bool x_tier = false;        // the expanded tier exists (one rank); else the split is local
  int x_nl = 0, x_ni = 0;
  long long x_ld = 0;
  DevBuf<BlockDev> d_xblock;
  DevBuf<dev::NodeConst> d_xnc;
  DevBuf<double> d_xrhom, d_xrec, d_xsol;
  // CAAS::run as one cluster kernel (cluster_caas.cuh): a tracer on chip, one pass over HBM.
  int cluster_mode = 0;       // cedr_b200_set_cluster_caas: 0 off (default: slower today), 1 on
  bool cluster_ok = false;
  ccaas::Args cluster_args;
  int cluster_ng = 0;
  size_t cluster_smem = 0;
  // Persistent single-read kernel (ring_kernels.cuh), the default run() where it applies.
  bool ring_enabled = false;  // cedr_b200_set_ring (opt-in: slower than the multi-launch path today)
  bool ring_ok = false;
  struct Ring {
    int grid = 0, S = 0, npn = 0, npairs_max = 0, TB = 1, plen = 0, nuslots = 0, ndslots = 0;
    int maxlag = 0, n7len = 0;
    int np = 3, sw = 2;
    int M = 0, mni = 0, mnlev = 0;
    long long ld = 0;         // padded sub-root count
    size_t smem = 0;
    DevBuf<ring::PieceDev> pieces;
    DevBuf<ushort4> d7tab;
    DevBuf<int2> d7c;
    DevBuf<uint2> pairtab;
    DevBuf<int> topc, lvlptr, kid0, kid1, micro_c, micro_h, msrc;
    DevBuf<dev::NodeWQ> mwq;
    DevBuf<dev::NodeRh> mrh;
    DevBuf<double> rec, sol, n7ring;
    DevBuf<int2> ktab[NCLS];
    DevBuf<unsigned> sync;    // [2 nt]: arrivals, flags
    DevBuf<unsigned long long> trace;   // debug (CEDR_B200_RING_TRACE)
    size_t trace_n = 0;
  } ring;
  DevBuf<int> d_status;
  std::vector<DevBuf<BlockDev> > d_blocks;   // per tier
  DevBuf<dev::NodeConst> d_nc;
  std::vector<DevBuf<double> > d_rhom_tier;  // leaf rhom of tiers >= 1
  std::vector<DevBuf<double> > d_rec, d_sol; // records / solved masses, tiers >= 1
  std::vector<long long> tier_ld;            // padded leaf counts per tier
  DevBuf<double> d_qglob, d_caas_scal;

  cudaStream_t stream = 0;
  // run_qlt computes the node constants (rhom sweeps) beside the up-sweep, which does not
  // read them: a stream of our own, forked from and joined to `stream` by events.
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cedr_b200_allgather_fn allgather = nullptr;
  void* allgather_ctx = nullptr;
  DevBuf<double> xsend_own, xrecv_own;   // exchange message / gathered messages
  double* xsend = nullptr;
  double* xrecv = nullptr;
  // Peer-to-peer exchange (cedr_b200_p2p_*): one allocation [64 epoch flags | receive
  // buffer, parity 0 | receive buffer, parity 1], mapped by every peer through CUDA IPC.
  DevBuf<double> p2p_arena;
  std::vector<void*> p2p_peer;           // peer r's arena in this process's address space
  bool p2p_on = false;
  unsigned long long p2p_epoch = 0;   // host mirror of the device counter (parity of buffers)
  // run() as a replayed CUDA graph (cedr_b200_set_graph): one per exchange-buffer parity.
  struct RunGraph { cudaGraphExec_t exec = nullptr; int launches = 0; };
  RunGraph graph[2];
  int graph_mode = -1;         // -1: where it pays (multi-rank p2p; short one-rank runs), 0: never, 1: whenever possible
  int graph_plain_runs = 0;    // plain run() calls since the last reset (the first ones stay plain)
  cudaStream_t cap_stream = nullptr;   // capture happens here (the caller's may be stream 0)
  void graph_reset () {
    for (RunGraph& g : graph) {
      if (g.exec) cudaGraphExecDestroy(g.exec);
      g.exec = nullptr;
    }
    graph_plain_runs = 0;
  }
  int last_launches = 0;

  // Optional per-launch timing (cedr_b200_set_profiling).
  bool profiling = false;
  struct Timed { cudaEvent_t e0, e1; int tag, tier; };
  std::vector<Timed> timed;
  size_t ntimed = 0;
  ~cedr_b200_cdr () {
    graph_reset();
    if (cap_stream) cudaStreamDestroy(cap_stream);
    if (side_stream) cudaStreamDestroy(side_stream);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    for (auto& t : timed) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
    for (size_t r = 0; r < p2p_peer.size(); ++r)
      if (p2p_peer[r] && static_cast<int>(r) != rank) cudaIpcCloseMemHandle(p2p_peer[r]);
  }
};

namespace {
// RAII bracket around one launch when profiling is on.
struct LaunchTimer {
  cedr_b200_cdr& c;
  cedr_b200_cdr::Timed* t = nullptr;
  LaunchTimer (cedr_b200_cdr& c_, int tag, int tier) : c(c_) {
    if (! c.profiling) return;
    if (c.ntimed == c.timed.size()) {
      cedr_b200_cdr::Timed n;
      cudaEventCreate(&n.e0);
      cudaEventCreate(&n.e1);
      c.timed.push_back(n);
    }
    t = &c.timed[c.ntimed++];
    t->tag = tag;
    t->tier = tier;
    cudaEventRecord(t->e0, c.stream);
  }
  ~LaunchTimer () { if (t) cudaEventRecord(t->e1, c.stream); }
};
}

namespace {

int env_int (const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}

size_t sweep_smem_bytes (const cedr_b200_cdr& c, int tier) {
  return sizeof(double)*4*static_cast<size_t>(2*c.plan.tiers[tier].max_nl);
}

// Raise (never lower) the dynamic shared memory limit of sweep_kernel<CLS, MODE>.
// (cudaFuncSetAttribute is per device: the cache is keyed by the current device.)
int current_device () {
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  return dev;
}

template <int CLS, int MODE> void sweep_smem_configure (const size_t smem) {
  static std::unordered_map<int, size_t> configured;
  size_t& have = configured.emplace(current_device(), 48*1024).first->second;
  if (smem > have) {
    CUDA_CHECK(cudaFuncSetAttribute(sweep_kernel<CLS, MODE>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
    have = smem;
  }
}

template <int CLS, int MODE>
void launch_sweep (cedr_b200_cdr& c, int tier, const SweepArgs& a) {
  if (a.ntr == 0 || a.nblocks == 0) return;
  const size_t smem = sweep_smem_bytes(c, tier);
  sweep_smem_configure<CLS, MODE>(smem);
  const long long grid = static_cast<long long>(a.nblocks)*a.ntr;
  cedr_b200_throw_if(grid > 0x7fffffffLL, "grid too large");
  LaunchTimer lt(c, MODE == MODE_UP ? CEDR_B200_TAG_UP : MODE == MODE_TOP ?
                 CEDR_B200_TAG_TOP : CEDR_B200_TAG_DOWN, tier);
  // One thread per leaf of the largest block (the widest level), 64..256: small blocks --
  // the tier above the tier-0 blocks has 128 leaves at ne120 -- get more CTAs per SM and
  // cheaper barriers.
  int threads = std::max(64, std::min(kThreads, (c.plan.tiers[tier].max_nl + 31)/32*32));
  if (const int e = env_int("CEDR_B200_SWEEP_THREADS", 0)) threads = std::min(threads, e);
  sweep_kernel<CLS, MODE><<<static_cast<unsigned>(grid), threads, smem, c.stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  ++c.last_launches;
}

template <int CLS>
void launch_sweep_mode (cedr_b200_cdr& c, int tier, int mode, const SweepArgs& a) {
  switch (mode) {
  case MODE_UP: launch_sweep<CLS, MODE_UP>(c, tier, a); break;
  case MODE_TOP: launch_sweep<CLS, MODE_TOP>(c, tier, a); break;
  case MODE_DOWN: launch_sweep<CLS, MODE_DOWN>(c, tier, a); break;
  }
}

void launch_sweep_any (cedr_b200_cdr& c, int cls, int tier, int mode, const SweepArgs& a) {
  switch (cls) {
  case CLS_ST: launch_sweep_mode<CLS_ST>(c, tier, mode, a); break;
  case CLS_CST: launch_sweep_mode<CLS_CST>(c, tier, mode, a); break;
  case CLS_T: launch_sweep_mode<CLS_T>(c, tier, mode, a); break;
  case CLS_CT: launch_sweep_mode<CLS_CT>(c, tier, mode, a); break;
  case CLS_NN: launch_sweep_mode<CLS_NN>(c, tier, mode, a); break;
  case CLS_CNN: launch_sweep_mode<CLS_CNN>(c, tier, mode, a); break;
  case CLS_CAAS:
    if (mode == MODE_UP) launch_sweep<CLS_CAAS, MODE_UP>(c, tier, a);
    else launch_sweep<CLS_CAAS, MODE_TOP>(c, tier, a);
    break;
  case CLS_BFB:
    if (mode == MODE_UP) launch_sweep<CLS_BFB, MODE_UP>(c, tier, a);
    else launch_sweep<CLS_BFB, MODE_TOP>(c, tier, a);
    break;
  }
}

// Blocks in the device list of a tier: tier 0 holds only this rank's blocks.
int nblocks_dev (const cedr_b200_cdr& c, int tier) {
  return tier == 0 ? static_cast<int>(c.own_blocks.size()) :
    static_cast<int>(c.plan.tiers[tier].blocks.size());
}

// Where the leaf rows of class `cls` are: the CDR's own buffer (row offsets of the class,
// cedr_qlt_inl.hpp:21-58, cedr_caas_inl.hpp:21-34) or the caller's bound arrays.
RowMap row_map (const cedr_b200_cdr& c, int cls) {
  RowMap rm;
  std::memset(&rm, 0, sizeof(rm));
  if (c.bound.on) {
    rm.p[0] = c.bound.qm_min;
    rm.p[1] = c.bound.qm;
    rm.p[2] = c.bound.qm_max;
    rm.p[3] = c.bound.qm_prev;
    rm.ld = c.bound.lda;
    rm.trow = c.d_ident.p;
    return rm;
  }
  rm.ld = c.ld;
  rm.trow = c.d_trcr_row.p;
  if (cls == CLS_NN || cls == CLS_CNN || cls == CLS_BFB) {
    rm.p[1] = c.in;
    if (cls == CLS_CNN) rm.p[3] = c.in + c.ld;
  } else {
    for (int f = 0; f < 3; ++f) rm.p[f] = c.in + f*c.ld;
    const bool prev = cls == CLS_CAAS ? c.caas_need_conserve : (cls == CLS_CST || cls == CLS_CT);
    if (prev) rm.p[3] = c.in + 3*c.ld;
  }
  return rm;
}

// The fast kernels' table of row addresses (FastArgs::rowaddr), rebuilt whenever the rows
// move (finish_setup, bind_arrays). Stream-ordered before the next run().
void upload_rowaddr (cedr_b200_cdr& c) {
  const size_t nt = c.trcr_prob.size();
  if (nt == 0) return;
  std::vector<const double*> h(4*nt, nullptr);
  for (size_t t = 0; t < nt; ++t) {
    const RowMap rm = row_map(c, c.trcr_cls[t]);
    const long long o = c.bound.on ? static_cast<long long>(t)*rm.ld :
      static_cast<long long>(c.trcr_row[t])*rm.ld;
    for (int f = 0; f < 4; ++f)
      if (rm.p[f]) h[4*t + f] = rm.p[f] + o;
  }
  if (c.d_rowaddr.n != h.size()) c.d_rowaddr.alloc(h.size());
  CUDA_CHECK(cudaMemcpyAsync(c.d_rowaddr.p, h.data(), h.size()*sizeof(const double*),
                             cudaMemcpyHostToDevice, c.stream));
}

SweepArgs base_args (cedr_b200_cdr& c, int cls, int tier) {
  SweepArgs a;
  std::memset(&a, 0, sizeof(a));
  a.blocks = c.d_blocks[tier].p;
  a.nblocks = nblocks_dev(c, tier);
  a.lvlptr = c.d_lvlptr.p;
  a.kid0 = c.d_kid0.p;
  a.kid1 = c.d_kid1.p;
  a.nc = c.d_nc.p;
  a.tier0 = tier == 0;
  if (tier == 0) {
    a.in = c.in;
    a.in_ld = c.ld;
    a.rows = row_map(c, cls);
    a.out = c.bound.on && ! c.is_caas ? c.bound.qm_out : c.out;
    a.out_ld = c.bound.on && ! c.is_caas ? c.bound.lda : c.ld;
  } else {
    a.in = c.d_rec[tier].p;
    a.in_ld = c.tier_ld[tier];
    a.out = c.d_sol[tier].p;
    a.out_ld = c.tier_ld[tier];
  }
  a.trcr_row = c.d_trcr_row.p;
  a.trcr_prob = c.d_trcr_prob.p;
  const int ntiers = static_cast<int>(c.plan.tiers.size());
  if (tier + 1 < ntiers) {
    a.rec_out = c.d_rec[tier+1].p;
    a.rec_ld = c.tier_ld[tier+1];
    a.sol_in = c.d_sol[tier+1].p;
    a.sol_in_ld = c.tier_ld[tier+1];
  }
  a.tracers = c.d_cls_tracers[cls].p;
  a.ntr = static_cast<int>(c.cls_tracers[cls].size());
  a.prefer_mass_con = c.prefer_mass_con;
  a.qglob = c.d_qglob.p;
  a.caas_scal = c.d_caas_scal.p;
  return a;
}

// ---- fast tier-0 kernels (fast_kernels.cuh)

bool fast_class (int cls, int mode) {
  if (std::getenv("CEDR_B200_FAST_ST_ONLY") && cls != CLS_ST && cls != CLS_CST &&
      cls != CLS_CAAS)
    return false;
  if (mode == MODE_UP) return cls <= CLS_CAAS;
  return cls < CLS_CAAS;
}

// Classes the mid kernels (local split of a multi-rank run) cover.
bool three_field_class (int cls) { return cls == CLS_ST || cls == CLS_CST; }

fast::FastArgs fast_args (cedr_b200_cdr& c, int cls) {
  fast::FastArgs a;
  std::memset(&a, 0, sizeof(a));
  a.blocks = c.d_blocks[0].p;
  a.nblocks = nblocks_dev(c, 0);
  a.dtab = c.d_dtab.p;
  a.ptab = c.d_ptab.p;
  a.perm = c.d_perm.p;
  a.pent = c.d_pent.p;
  a.wq = c.d_fwq.p;
  a.rh = c.d_frh.p;
  a.in = c.in;
  a.in_ld = c.ld;
  a.rowaddr = c.d_rowaddr.p;
  a.trcr_row = c.d_trcr_row.p;
  a.trcr_prob = c.d_trcr_prob.p;
  a.rec_out = c.d_rec[1].p;
  a.rec_ld = c.tier_ld[1];
  a.sol_in = c.d_sol[1].p;
  a.sol_in_ld = c.tier_ld[1];
  a.out = c.bound.on && ! c.is_caas ? c.bound.qm_out : c.out;
  a.out_ld = c.bound.on && ! c.is_caas ? c.bound.lda : c.ld;
  a.tracers = c.d_cls_tracers[cls].p;
  a.ntr = static_cast<int>(c.cls_tracers[cls].size());
  a.sbuf = (c.plan.tiers[0].max_nl + 2 + 1) & ~1;
  a.prefer_mass_con = c.prefer_mass_con;
  a.qglob = c.d_qglob.p;
  // Tracers per CTA: enough CTAs for several waves, enough tracers per CTA to amortise
  // the per-CTA setup and keep the TMA double buffer busy.
  const long long work = static_cast<long long>(a.nblocks)*a.ntr;
  a.group = static_cast<int>(std::max<long long>(1, std::min<long long>(32, work/(148*5*4))));
  if (const char* e = std::getenv("CEDR_B200_GROUP")) a.group = std::max(1, std::atoi(e));
  a.n7buf = c.d_n7.p;
  a.rq = c.d_frq.p;
  a.caas_rows = c.caas_need_conserve ? 4 : 3;
  if (c.split && cls != CLS_CAAS && (c.x_tier || three_field_class(cls))) {
    a.split = c.split;
    a.rec_out = c.d_xrec.p;
    a.rec_ld = c.x_ld;
    a.sol_in = c.d_xsol.p;
    a.sol_in_ld = c.x_ld;
  }
#ifdef CEDR_B200_PHASE_CLOCKS
  if ( ! c.d_phase_clk.p) {
    c.d_phase_clk.alloc(16);
    CUDA_CHECK(cudaMemset(c.d_phase_clk.p, 0, 16*sizeof(unsigned long long)));
  }
  a.phase_clk = c.d_phase_clk.p;
#endif
  return a;
}

template <typename K, typename... Extra>
void launch_fast (cedr_b200_cdr& c, K kernel, const fast::FastArgs& a, size_t smem, int tag,
                  int threads = fast::kThreads, Extra... extra) {
  if (a.ntr == 0) return;
  // Raise (never lower) the kernel's dynamic shared memory limit, once per size.
  static std::map<std::pair<int, const void*>, size_t> configured;
  size_t& have = configured[std::make_pair(current_device(),
                                           reinterpret_cast<const void*>(kernel))];
  if (smem > have) {
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(std::max<size_t>(smem, 48*1024))));
    have = smem;
  }
  const long long grid = static_cast<long long>(a.nblocks)*((a.ntr + a.group - 1)/a.group);
  cedr_b200_throw_if(grid > 0x7fffffffLL, "grid too large");
  LaunchTimer lt(c, tag, 0);
  kernel<<<static_cast<unsigned>(grid), threads, smem, c.stream>>>(a, extra...);
  CUDA_CHECK(cudaGetLastError());
  ++c.last_launches;
}

void launch_fast_up (cedr_b200_cdr& c, int cls) {
  const fast::FastArgs a = fast_args(c, cls);
  const int nrows = (cls == CLS_ST || cls == CLS_T) ? 3 : cls == CLS_NN ? 1 : cls == CLS_CNN ? 2 : 4;
  const size_t smem = sizeof(double)*2*nrows*a.sbuf + 16 + sizeof(double)*32;
  switch (cls) {
  case CLS_ST: launch_fast(c, fast::up_kernel<CLS_ST>, a, smem, CEDR_B200_TAG_UP); break;
  case CLS_CST: launch_fast(c, fast::up_kernel<CLS_CST>, a, smem, CEDR_B200_TAG_UP); break;
  case CLS_T: launch_fast(c, fast::up_kernel<CLS_T>, a, smem, CEDR_B200_TAG_UP); break;
  case CLS_CT: launch_fast(c, fast::up_kernel<CLS_CT>, a, smem, CEDR_B200_TAG_UP); break;
  case CLS_NN: launch_fast(c, fast::up_kernel<CLS_NN>, a, smem, CEDR_B200_TAG_UP); break;
  case CLS_CNN: launch_fast(c, fast::up_kernel<CLS_CNN>, a, smem, CEDR_B200_TAG_UP); break;
  case CLS_CAAS: launch_fast(c, fast::up_kernel<CLS_CAAS>, a, smem, CEDR_B200_TAG_UP); break;
  }
}

// The transposed down-sweep (transposed_kernels.cuh) takes the three-field classes when the
// block tops come from the tier above (split 2 or 3) and a class has enough tracers to fill
// the lanes of a warp.
bool transposed_down_ok (const cedr_b200_cdr& c, const fast::FastArgs& a) {
  return env_int("CEDR_B200_TRANSPOSED", 1) && (a.split == 2 || a.split == 3) &&
    a.ntr >= env_int("CEDR_B200_TRANSPOSED_MIN", 4) && c.d_x7.p;
}

template <int CLS, bool PREFER>
void launch_mid_cls (cedr_b200_cdr& c, const fast::FastArgs& a, const fast::TArgs& ta) {
  const unsigned ngroups = static_cast<unsigned>((a.ntr + fast::kTLanes - 1)/fast::kTLanes);
  cedr_b200_throw_if(ngroups > 65535u, "grid too large");
  LaunchTimer lt(c, CEDR_B200_TAG_MID, 0);
  fast::midT_kernel<CLS, PREFER><<<dim3(4u*a.nblocks, ngroups), 128, 0, c.stream>>>(a, ta);
  CUDA_CHECK(cudaGetLastError());
  ++c.last_launches;
}

void launch_transposed_down (cedr_b200_cdr& c, int cls, const fast::FastArgs& a) {
  if (a.ntr == 0) return;
  fast::TArgs ta;
  ta.x7 = c.d_x7.p;
  const bool prefer = a.prefer_mass_con != 0;
  const size_t smem = fast::down3_smem_bytes(a.sbuf);
  if (cls == CLS_ST) {
    if (prefer) launch_mid_cls<CLS_ST, true>(c, a, ta);
    else launch_mid_cls<CLS_ST, false>(c, a, ta);
    launch_fast(c, fast::down3_kernel<CLS_ST>, a, smem, CEDR_B200_TAG_DOWN,
                fast::kLeafThreads, ta);
  } else {
    if (prefer) launch_mid_cls<CLS_CST, true>(c, a, ta);
    else launch_mid_cls<CLS_CST, false>(c, a, ta);
    launch_fast(c, fast::down3_kernel<CLS_CST>, a, smem, CEDR_B200_TAG_DOWN,
                fast::kLeafThreads, ta);
  }
}

void launch_fast_down (cedr_b200_cdr& c, int cls) {
  const fast::FastArgs a = fast_args(c, cls);
  if ( ! three_field_class(cls)) {
    const size_t smem1 = fast::down1_smem_bytes(a.sbuf);
    switch (cls) {
    case CLS_T: launch_fast(c, fast::down1_kernel<CLS_T>, a, smem1, CEDR_B200_TAG_DOWN,
                            fast::kDown2Threads); break;
    case CLS_CT: launch_fast(c, fast::down1_kernel<CLS_CT>, a, smem1, CEDR_B200_TAG_DOWN,
                             fast::kDown2Threads); break;
    case CLS_NN: launch_fast(c, fast::down1_kernel<CLS_NN>, a, smem1, CEDR_B200_TAG_DOWN,
                             fast::kDown2Threads); break;
    case CLS_CNN: launch_fast(c, fast::down1_kernel<CLS_CNN>, a, smem1, CEDR_B200_TAG_DOWN,
                              fast::kDown2Threads); break;
    }
    return;
  }
  if (transposed_down_ok(c, a)) {
    launch_transposed_down(c, cls, a);
    return;
  }
  size_t smem = fast::down2_smem_bytes(a.sbuf);
  if (const char* e = std::getenv("CEDR_B200_SMEM_PAD")) smem += std::atoi(e);  // occupancy experiments
  if (cls == CLS_ST)
    launch_fast(c, fast::down2_kernel<CLS_ST>, a, smem, CEDR_B200_TAG_DOWN,
                fast::kDown2Threads);
  else
    launch_fast(c, fast::down2_kernel<CLS_CST>, a, smem, CEDR_B200_TAG_DOWN,
                fast::kDown2Threads);
}

// ---- persistent single-read kernel (ring_kernels.cuh)

bool ring_class (int cls) { return cls == CLS_ST || cls == CLS_CST || cls == CLS_CAAS; }

const void* ring_kernel_ptr (int cls, int np, int sw) {
#define CEDR_RK(C, N, W) reinterpret_cast<const void*>(ring::run_kernel<C, N, W>)
#define CEDR_RK_CLS(C)                                                  \
  (np == 2 ? (sw == 2 ? CEDR_RK(C, 2, 2) : CEDR_RK(C, 2, 4))            \
           : (sw == 2 ? CEDR_RK(C, 3, 2) : CEDR_RK(C, 3, 4)))
  switch (cls) {
  case CLS_ST: return CEDR_RK_CLS(CLS_ST);
  case CLS_CST: return CEDR_RK_CLS(CLS_CST);
  default: return CEDR_RK_CLS(CLS_CAAS);
  }
#undef CEDR_RK_CLS
#undef CEDR_RK
}


// ---- cluster-resident CAAS (cluster_caas.cuh)

const void* cluster_kernel_ptr (int ng) {
  switch (ng) {
  case 1: return reinterpret_cast<const void*>(ccaas::run_kernel<1>);
  case 2: return reinterpret_cast<const void*>(ccaas::run_kernel<2>);
  default: return reinterpret_cast<const void*>(ccaas::run_kernel<4>);
  }
}

void cluster_launch_config (const cedr_b200_cdr& c, int nclusters, cudaLaunchConfig_t& cfg,
                            cudaLaunchAttribute* at) {
  std::memset(&cfg, 0, sizeof(cfg));
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = c.cluster_args.cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(static_cast<unsigned>(nclusters)*c.cluster_args.cs);
  cfg.blockDim = dim3(128u*c.cluster_ng);
  cfg.dynamicSmemBytes = c.cluster_smem;
  cfg.stream = c.stream;
  cfg.attrs = at;
  cfg.numAttrs = 1;
}

// Decide whether CAAS::run can be the cluster kernel; called from finish_setup. A CTA takes
// whole tier-0 blocks (at most 2 per 128-thread group, 4 groups), a cluster at most 16 CTAs,
// and five row slots of a CTA's slice must fit its shared memory.
void cluster_setup (cedr_b200_cdr& c) {
  c.cluster_ok = false;
  if ( ! c.is_caas || c.is_bfb || ! c.fast_ok || c.nranks > 1) return;
  if (c.cluster_mode <= 0 && ! env_int("CEDR_B200_CLUSTER_CAAS", 0)) return;
  if (c.caas_sum_mode != CEDR_B200_CAAS_SUM_TREE) return;
  if (c.plan.tiers.size() != 2 || c.plan.tiers[1].blocks.size() != 1) return;
  const int nt = static_cast<int>(c.trcr_prob.size());
  if (nt == 0) return;
  const std::vector<Block>& blocks = c.plan.tiers[0].blocks;
  const int nb = static_cast<int>(blocks.size());
  for (int b = 0; b + 1 < nb; ++b)
    if (blocks[b + 1].leaf0 != blocks[b].leaf0 + blocks[b].nl ||
        c.leaf_lci[blocks[b + 1].leaf0] != c.leaf_lci[blocks[b].leaf0] + blocks[b].nl) return;
  if (c.leaf_lci[blocks[0].leaf0] != blocks[0].leaf0) return;
  int dev = 0, smem_max = 0, cluster_launch = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&cluster_launch, cudaDevAttrClusterLaunch, dev));
  if ( ! cluster_launch) return;
  const Block& tb = c.plan.tiers[1].blocks[0];
  const Shape& ts = c.plan.shapes[tb.shape];
  ccaas::Args& a = c.cluster_args;
  std::memset(&a, 0, sizeof(a));
  a.top.nl = ts.nl;
  a.top.ni = ts.ni;
  a.top.nlev = ts.nlev;
  a.top.lvlptr_off = ts.dev_lvlptr_off;
  a.top.kid_off = ts.dev_kid_off;
  // As many whole blocks per CTA as its shared memory holds (fewer, fatter CTAs: a cluster
  // of one needs no exchange at all), up to 2 per 128-thread group x 4 groups.
  const int max_cs = env_int("CEDR_B200_CLUSTER_SIZE", 16);
  int bpc = std::min(nb, 4*ccaas::kMaxBlocksPerGroup);
  for (; bpc >= 1; --bpc) {
    int slice = 0;
    for (int b = 0; b < nb; b += bpc) {
      const int e = std::min(nb, b + bpc) - 1;
      slice = std::max(slice, blocks[e].leaf0 + blocks[e].nl - blocks[b].leaf0);
    }
    a.cap = (slice + 2 + 1) & ~1;
    c.cluster_ng = bpc >= 3 ? 4 : bpc;     // bpc 1 -> 1 group, 2 -> 2, 3..8 -> 4
    c.cluster_smem = ccaas::smem_bytes(a, c.cluster_ng);
    if (c.cluster_smem <= static_cast<size_t>(smem_max)) break;
  }
  if (bpc < 1 || (nb + bpc - 1)/bpc > max_cs) return;
  a.bpc = bpc;
  a.cs = (nb + bpc - 1)/bpc;
  a.nblocks = nb;
  a.ntr = nt;
  a.need_prev = c.caas_need_conserve;
  const void* const k = cluster_kernel_ptr(c.cluster_ng);
  CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(c.cluster_smem)));
  if (a.cs > 8)
    CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  cluster_launch_config(c, 1, cfg, at);
  int maxc = 0;
  if (cudaOccupancyMaxActiveClusters(&maxc, k, &cfg) != cudaSuccess || maxc < 1) {
    cudaGetLastError();
    return;
  }
  a.nclusters = std::min(maxc, nt);
  if (const char* e = std::getenv("CEDR_B200_CLUSTERS")) a.nclusters = std::max(1, std::atoi(e));
  c.cluster_ok = true;
}

void launch_cluster_caas (cedr_b200_cdr& c) {
  ccaas::Args a = c.cluster_args;
  a.blocks = c.d_blocks[0].p;
  a.dtab = c.d_dtab.p;
  a.perm = c.d_perm.p;
  a.rowaddr = c.d_rowaddr.p;
  a.trcr_prob = c.d_trcr_prob.p;
  a.lvlptr = c.d_lvlptr.p;
  a.kid0 = c.d_kid0.p;
  a.kid1 = c.d_kid1.p;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  cluster_launch_config(c, a.nclusters, cfg, at);
  void* args[1] = {&a};
  LaunchTimer lt(c, CEDR_B200_TAG_FUSED, 0);
  CUDA_CHECK(cudaLaunchKernelExC(&cfg, cluster_kernel_ptr(c.cluster_ng), args));
  ++c.last_launches;
}

// Decide whether run() can be the ring kernel and build its tables; called from
// finish_setup. One piece per CTA: consecutive depth-S subtrees of the tier-0 blocks.
void ring_setup (cedr_b200_cdr& c) {
  c.ring_ok = false;
  cedr_b200_cdr::Ring& R = c.ring;
  if ( ! c.ring_enabled || ! c.fast_ok || c.is_bfb || std::getenv("CEDR_B200_NO_RING")) return;
  if (c.nranks > 1) return;
  if (c.is_caas && c.caas_sum_mode != CEDR_B200_CAAS_SUM_TREE) return;
  if (c.plan.tiers.size() != 2 || c.plan.tiers[1].blocks.size() != 1) return;
  const int nt = static_cast<int>(c.trcr_prob.size());
  if (nt == 0) return;
  if ( ! c.is_caas && c.cls_tracers[CLS_ST].empty() && c.cls_tracers[CLS_CST].empty()) return;
  int dev = 0, coop = 0, nsm = 0, smem_max = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if ( ! coop) return;
  nsm = env_int("CEDR_B200_RING_GRID", nsm);
  const std::vector<Block>& blocks = c.plan.tiers[0].blocks;
  const int nb = static_cast<int>(blocks.size());
  // The leaves of consecutive blocks must be consecutive local cells.
  for (int b = 0; b + 1 < nb; ++b)
    if (c.leaf_lci[blocks[b + 1].leaf0] != c.leaf_lci[blocks[b].leaf0] + blocks[b].nl) return;
  // S: the shallowest cut that keeps the SMs busy (sub-blocks per CTA x CTAs against the
  // sub-blocks there are), with at most 128 depth-7 nodes per piece. Deeper cuts cost more
  // records and a larger tree above the sub-roots.
  int S = -1, G = 0, per = 0;
  {
    double eff[8] = {0}, best = 0;
    for (int s = 3; s <= 7; ++s) {
      const long long n1 = static_cast<long long>(nb) << s;
      const int g = static_cast<int>(std::min<long long>(nsm, n1));
      const int pr = static_cast<int>((n1 + g - 1)/g);
      if ((pr << (7 - s)) > ring::kGroup) continue;
      eff[s] = static_cast<double>(n1)/(static_cast<double>(pr)*nsm);
      best = std::max(best, eff[s]);
    }
    for (int s = 3; s <= 7 && S < 0; ++s)
      if (eff[s] > 0 && eff[s] >= best - 0.03) {
        const long long n1 = static_cast<long long>(nb) << s;
        S = s;
        G = static_cast<int>(std::min<long long>(nsm, n1));
        per = static_cast<int>((n1 + G - 1)/G);
      }
  }
  if (const char* e = std::getenv("CEDR_B200_RING_S")) {
    const int s = std::atoi(e);
    const long long n1 = static_cast<long long>(nb) << s;
    if (s >= 3 && s <= 7) {
      S = s; G = static_cast<int>(std::min<long long>(nsm, n1));
      per = static_cast<int>((n1 + G - 1)/G);
      if ((per << (7 - s)) > ring::kGroup) return;
    }
  }
  if (S < 0) return;
  const int N1 = nb << S, nhp = 1 << (7 - S);
  R.S = S;
  R.grid = G;
  R.npn = per*nhp;
  R.ld = round_up(N1, 16);
  R.np = std::max(2, std::min(3, env_int("CEDR_B200_RING_NP", c.is_caas ? 3 : 2)));
  R.sw = env_int("CEDR_B200_RING_SW", 2) >= 4 ? 4 : 2;

  // ---- pieces
  std::vector<ring::PieceDev> pieces(G);
  std::vector<ushort4> d7tab;
  std::vector<int2> d7c;
  std::vector<uint2> pairtab;
  std::vector<int> topc;
  int maxnl = 0;
  R.npairs_max = 1;
  for (int g = 0, sub = 0; g < G; ++g) {
    // The first N1 % G pieces take `per` sub-blocks, the others per - 1 (all `per` if even).
    const int rem = N1 % G;
    const int ns = (rem == 0 || g < rem) ? per : per - 1;
    ring::PieceDev P;
    std::memset(&P, 0, sizeof(P));
    P.nsub = ns;
    P.sub0 = sub;
    P.nd7 = ns*nhp;
    P.d7_off = static_cast<int>(d7tab.size());
    P.pair_off = static_cast<int>(pairtab.size());
    P.top_off = static_cast<int>(topc.size());
    int leaf0 = -1, leaf_end = -1;
    for (int i = 0; i < ns; ++i, ++sub) {
      const int b = sub >> S, ps = sub & ((1 << S) - 1);
      const Block& blk = blocks[b];
      const Shape& sh = c.plan.shapes[blk.shape];
      const int lci0 = c.leaf_lci[blk.leaf0];
      for (int k = 0; k < nhp; ++k) {
        const int n7 = ps*nhp + k;            // depth-7 node of the block
        const int nloc = i*nhp + k;           // ... of the piece
        unsigned short e[4];
        for (int q = 0; q < 4; ++q) {
          const unsigned short de = sh.dtab[4*n7 + q];
          const int o = lci0 + (de & 0x7fff);
          const bool pair = (de >> 15) != 0;
          if (leaf0 < 0) leaf0 = o;
          leaf_end = o + (pair ? 2 : 1);
          const int rel = o - leaf0;
          if (rel > 0x7ffe) return;           // does not fit the 15-bit offsets
          e[q] = static_cast<unsigned short>(rel | (pair ? 0x8000 : 0));
          if (pair) {
            const int r = static_cast<int>(std::lower_bound(sh.ptab.begin(), sh.ptab.end(),
                                                            4*n7 + q) - sh.ptab.begin());
            uint2 pe;
            pe.x = static_cast<unsigned>(rel) | (static_cast<unsigned>(q) << 16) |
              (static_cast<unsigned>(nloc) << 18);
            pe.y = static_cast<unsigned>(blk.ibase + 511 + r);
            pairtab.push_back(pe);
          }
        }
        d7tab.push_back(make_ushort4(e[0], e[1], e[2], e[3]));
        d7c.push_back(make_int2(blk.ibase + 127 + n7, blk.ibase + 255 + 2*n7));
      }
      // Constants of the sub-block's nodes of depths S..6: heap order within the sub-block.
      for (int h = 0; h < nhp; ++h) {
        if (h == nhp - 1) { topc.push_back(-1); break; }
        int l = 0;
        while ((2 << l) - 1 <= h) ++l;
        const int pp = h - ((1 << l) - 1);
        topc.push_back(blk.ibase + (1 << (S + l)) - 1 + ps*(1 << l) + pp);
      }
    }
    P.leaf0 = leaf0;
    P.nl = leaf_end - leaf0;
    P.npairs = static_cast<int>(pairtab.size()) - P.pair_off;
    maxnl = std::max(maxnl, P.nl);
    R.npairs_max = std::max(R.npairs_max, P.npairs);
    pieces[g] = P;
  }
  R.plen = (maxnl + 2 + 1) & ~1;

  // ---- the tree over the micro-roots is sized first (it sits in shared memory)
  const int M = nb << (S - 3);
  const int mni_est = ((1 << (S - 3)) - 1)*nb + c.plan.shapes[c.plan.tiers[1].blocks[0].shape].ni;
  const int mnlev_est = (S - 3) + c.plan.shapes[c.plan.tiers[1].blocks[0].shape].nlev;
  if (M + mni_est > 0xffff) return;

  // ---- tracers per unit and ring depths, from the shared-memory budget
  int TB = std::max(1, std::min(std::min(ring::kMaxTB, nt), ring::kGroup/R.npn));
  TB = std::max(1, std::min(TB, env_int("CEDR_B200_RING_TB", TB)));
  const size_t budget = static_cast<size_t>(smem_max) - 1024;
  for (;; TB = std::max(1, TB/2)) {
    const ring::SmemLayout l0 = ring::smem_layout(TB, R.plen, 0, 0, R.np, M, mni_est, mnlev_est,
                                                  R.npn, R.npairs_max, c.is_caas);
    const size_t per_u = sizeof(double)*l0.uslot_doubles + 3*sizeof(uint64_t);
    const size_t per_d = sizeof(double)*l0.dslot_doubles + 4*sizeof(uint64_t);
    // The UP ring holds a unit from its load to its arrival, the DOWN ring from its re-load
    // through the solves to its store: about 1 : 2.
    int nu = 0, nd = 0;
    if (l0.total < budget) {
      const size_t avail = budget - l0.total;
      nu = static_cast<int>(std::max<size_t>(2, avail/3/per_u));
      nu = std::min(nu, env_int("CEDR_B200_RING_USLOTS", std::min(nu, 6)));
      nd = avail > nu*per_u ? static_cast<int>((avail - nu*per_u)/per_d) : 0;
      nd = std::min(nd, ring::kMaxSlots);
      nd = std::min(nd, env_int("CEDR_B200_RING_DSLOTS", nd));
      // (whatever the DOWN ring's cap left over goes to the UP ring)
      if (avail > nd*per_d) nu = std::max(nu, std::min<int>(env_int("CEDR_B200_RING_USLOTS", ring::kMaxSlots),
                                                       std::min<int>(ring::kMaxSlots, (avail - nd*per_d)/per_u)));
    }
    // A slot always serves the same pipe (slot = unit % depth): depths are multiples of
    // the pipe count, so that every waiter sees the phases of its barriers in order.
    nu = nu/R.np*R.np;
    nd = nd/R.np*R.np;
    if (nu >= R.np && nd >= R.np) {
      R.TB = TB; R.nuslots = nu; R.ndslots = nd;
      R.smem = ring::smem_layout(TB, R.plen, nu, nd, R.np, M, mni_est, mnlev_est, R.npn,
                                 R.npairs_max, c.is_caas).total;
      break;
    }
    if (TB == 1) return;
  }
  // The UP pass may run ahead of the DOWN pass by what L2 holds comfortably (the leaves of
  // the units in between are re-read from there).
  {
    const double unit_bytes = 32.0*c.nlcl*R.TB;
    const double l2_window = 1e6*env_int("CEDR_B200_RING_L2_MB", 48);
    R.maxlag = static_cast<int>(std::max(4.0, std::min(64.0, l2_window/unit_bytes)));
    R.maxlag = env_int("CEDR_B200_RING_MAXLAG", R.maxlag);
    // (the kernel's window over the tracer table spans the units in flight)
    R.maxlag = std::max(1, std::min(R.maxlag, (ring::kWin - 96)/R.TB - R.nuslots - R.ndslots - 4));
    R.n7len = R.maxlag + R.ndslots + 1;
  }

  // ---- the tree over the micro-roots: the blocks' nodes above depth S-3 (perfect, Ep
  // micro-roots per block), then the tier-1 block with the block roots as its leaves.
  {
    const int Sp = S - 3, Ep = 1 << Sp, nx = (Ep - 1)*nb;
    const Block& b1 = c.plan.tiers[1].blocks[0];
    const Shape& s1 = c.plan.shapes[b1.shape];
    if (s1.nl != nb) return;
    std::vector<int> hstart(Sp + 2, 0);       // first internal index of height h
    for (int h = 1; h <= Sp; ++h) hstart[h + 1] = hstart[h] + nb*(1 << (Sp - h));
    auto xnode = [&] (int b, int d, int p) { return M + hstart[Sp - d] + b*(1 << d) + p; };
    const int ni = nx + s1.ni;
    if (ni != mni_est) return;
    std::vector<int> kid0(ni), kid1(ni), lvlptr(1, 0), msrc(ni);
    for (int h = 1; h <= Sp; ++h) {
      const int d = Sp - h;
      for (int b = 0; b < nb; ++b)
        for (int p = 0; p < (1 << d); ++p) {
          const int me = xnode(b, d, p) - M;
          if (d == Sp - 1) { kid0[me] = b*Ep + 2*p; kid1[me] = b*Ep + 2*p + 1; }
          else { kid0[me] = xnode(b, d + 1, 2*p); kid1[me] = xnode(b, d + 1, 2*p + 1); }
          msrc[me] = blocks[b].ibase + (1 << d) - 1 + p;
        }
      lvlptr.push_back(hstart[h + 1]);
    }
    auto map1 = [&] (int id) {
      return id < nb ? (Sp > 0 ? xnode(id, 0, 0) : id) : M + nx + (id - nb);
    };
    for (int j = 0; j < s1.ni; ++j) {
      kid0[nx + j] = map1(s1.kid0[j]);
      kid1[nx + j] = map1(s1.kid1[j]);
      msrc[nx + j] = -1 - (b1.ibase + j);
    }
    for (int l = 1; l <= s1.nlev; ++l) lvlptr.push_back(nx + s1.lvlptr[l]);
    std::vector<int> micro_c(M), micro_h(M);
    for (int m = 0; m < M; ++m) {
      const int b = m >> Sp, pm = m & (Ep - 1);
      micro_h[m] = Ep - 1 + pm;
      micro_c[m] = blocks[b].ibase + micro_h[m];
    }
    R.M = M;
    R.mni = ni;
    R.mnlev = static_cast<int>(lvlptr.size()) - 1;
    R.lvlptr.upload(lvlptr);
    R.kid0.upload(kid0);
    R.kid1.upload(kid1);
    R.msrc.upload(msrc);
    R.micro_c.upload(micro_c);
    R.micro_h.upload(micro_h);
    R.mwq.alloc(std::max(1, ni));
    R.mrh.alloc(std::max(1, ni));
  }
  R.pieces.upload(pieces);
  R.d7tab.upload(d7tab);
  R.d7c.upload(d7c);
  if (pairtab.empty()) pairtab.push_back(make_uint2(0, 0));
  R.pairtab.upload(pairtab);
  R.topc.upload(topc);
  R.rec.alloc(static_cast<size_t>(4)*nt*R.ld);
  R.sol.alloc(static_cast<size_t>(nt)*R.ld);
  if ( ! c.is_caas) {
    R.n7ring.alloc(static_cast<size_t>(R.n7len)*R.grid*3*ring::kGroup);
    CUDA_CHECK(cudaMemsetAsync(R.n7ring.p, 0, R.n7ring.n*sizeof(double), c.stream));
  }
  R.sync.alloc(2*static_cast<size_t>(nt));
  for (int cls : {CLS_ST, CLS_CST, CLS_CAAS}) {
    std::vector<int2> kt;
    for (int t : c.cls_tracers[cls])
      kt.push_back(make_int2(t, c.trcr_row[t] | ((c.trcr_prob[t] & 1) << 30)));
    R.ktab[cls].upload(kt);
  }
  if ( ! c.d_status.p) {
    c.d_status.alloc(1);
    CUDA_CHECK(cudaMemsetAsync(c.d_status.p, 0, sizeof(int), c.stream));
  }
  // All CTAs must be co-resident (they wait on each other): one per SM.
  for (int cls : {CLS_ST, CLS_CST, CLS_CAAS}) {
    if (c.is_caas != (cls == CLS_CAAS)) continue;
    const void* fn = ring_kernel_ptr(cls, R.np, R.sw);
    CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(R.smem)));
    int per_sm = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
      &per_sm, fn, ring::block_threads(R.np, R.sw), R.smem));
    int nsm_real = 0;
    CUDA_CHECK(cudaDeviceGetAttribute(&nsm_real, cudaDevAttrMultiProcessorCount, dev));
    if (per_sm*nsm_real < R.grid) return;
  }
  c.ring_ok = true;
}

// Constants of the nodes above the micro-roots, after the rhom sweep of this run().
void ring_gather (cedr_b200_cdr& c) {
  cedr_b200_cdr::Ring& R = c.ring;
  if (R.mni == 0) return;
  LaunchTimer lt(c, CEDR_B200_TAG_RHOM, 1);
  ring::gather_kernel<<<(R.mni + 255)/256, 256, 0, c.stream>>>(
    R.msrc.p, R.mni, c.d_fwq.p, c.d_frh.p, c.d_nc.p, R.mwq.p, R.mrh.p);
  CUDA_CHECK(cudaGetLastError());
  ++c.last_launches;
}

void launch_ring (cedr_b200_cdr& c, int cls) {
  const int ntr = static_cast<int>(c.cls_tracers[cls].size());
  if (ntr == 0) return;
  cedr_b200_cdr::Ring& R = c.ring;
  ring::Args a;
  std::memset(&a, 0, sizeof(a));
  a.pieces = R.pieces.p;
  a.d7tab = R.d7tab.p;
  a.d7c = R.d7c.p;
  a.pairtab = R.pairtab.p;
  a.topc = R.topc.p;
  a.wq = c.d_fwq.p;
  a.rh = c.d_frh.p;
  a.in = c.in;
  a.in_ld = c.ld;
  a.trcr_row = c.d_trcr_row.p;
  a.trcr_prob = c.d_trcr_prob.p;
  a.out = c.out;
  a.out_ld = c.ld;
  a.rec = R.rec.p;
  a.rec_ld = R.ld;
  a.sol = R.sol.p;
  a.sol_ld = R.ld;
  a.scal = c.d_caas_scal.p;
  a.tracers = c.d_cls_tracers[cls].p;
  a.ktab = R.ktab[cls].p;
  a.ntr = ntr;
  a.S = R.S;
  a.npn = R.npn;
  a.npairs_max = R.npairs_max;
  a.TB = R.TB;
  a.plen = R.plen;
  a.nuslots = R.nuslots;
  a.ndslots = R.ndslots;
  a.maxlag = R.maxlag;
  a.n7ring = R.n7ring.p;
  a.n7len = R.n7len;
  a.prefer_mass_con = c.prefer_mass_con;
  a.caas_rows = c.caas_need_conserve ? 4 : 3;
  a.cnt = R.sync.p;
  a.flag = R.sync.p + ntr;
  a.status = c.d_status.p;
  a.spin_limit = static_cast<unsigned long long>(env_int("CEDR_B200_RING_TIMEOUT_MS", 4000))*1000000ull;
  a.top.M = R.M;
  a.top.ni = R.mni;
  a.top.nlev = R.mnlev;
  a.top.lvlptr = R.lvlptr.p;
  a.top.kid0 = R.kid0.p;
  a.top.kid1 = R.kid1.p;
  a.top.mwq = R.mwq.p;
  a.top.mrh = R.mrh.p;
  a.top.micro_c = R.micro_c.p;
  a.top.micro_h = R.micro_h.p;
  CUDA_CHECK(cudaMemsetAsync(R.sync.p, 0, 2*static_cast<size_t>(ntr)*sizeof(unsigned),
                             c.stream));
  if (std::getenv("CEDR_B200_RING_TRACE")) {
    const size_t U = (static_cast<size_t>(ntr) + R.TB - 1)/R.TB;
    R.trace_n = static_cast<size_t>(R.grid)*U*8 + 2*static_cast<size_t>(ntr) + 64;
    if (R.trace.n < R.trace_n) R.trace.alloc(R.trace_n);
    CUDA_CHECK(cudaMemsetAsync(R.trace.p, 0, R.trace_n*sizeof(unsigned long long), c.stream));
    a.trace = R.trace.p;
    a.clk = R.trace.p + R.trace_n - 64;
  }
  void* params[] = {&a};
  const void* fn = ring_kernel_ptr(cls, R.np, R.sw);
  LaunchTimer lt(c, CEDR_B200_TAG_FUSED, 0);
  CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(R.grid), dim3(ring::block_threads(R.np, R.sw)),
                                         params, R.smem, c.stream));
  ++c.last_launches;
}

// rhom sweep of tiers [k0, k1).
void run_rhom (cedr_b200_cdr& c, int k0, int k1) {
  const int ntiers = static_cast<int>(c.plan.tiers.size());
  for (int k = k0; k < k1; ++k) {
    RhomArgs a;
    a.blocks = c.d_blocks[k].p;
    a.nblocks = nblocks_dev(c, k);
    a.lvlptr = c.d_lvlptr.p;
    a.kid0 = c.d_kid0.p;
    a.kid1 = c.d_kid1.p;
    a.in = k == 0 ? c.in : c.d_rhom_tier[k].p;
    a.root_out = k + 1 < ntiers ? c.d_rhom_tier[k+1].p : nullptr;
    a.nc = c.d_nc.p;
    a.fpos = (k == 0 && c.fast_ok) ? c.d_fpos.p : nullptr;
    a.fwq = c.d_fwq.p;
    a.frh = c.d_frh.p;
    a.frq = c.d_frq.p;
    a.sub_out = (k == 0 && c.split) ? c.d_xrhom.p : nullptr;
    a.split = c.split;
    const size_t smem = sizeof(double)*2*static_cast<size_t>(c.plan.tiers[k].max_nl);
    LaunchTimer lt(c, CEDR_B200_TAG_RHOM, k);
    rhom_kernel<<<a.nblocks, kThreads, smem, c.stream>>>(a);
    CUDA_CHECK(cudaGetLastError());
    ++c.last_launches;
  }
}

// ---- expanded tier above the fast blocks (FastArgs::split)

// rhom sums and node constants of the expanded tier, from the sub-root rhom the tier-0
// rhom sweep left in d_xrhom.
void run_rhom_x (cedr_b200_cdr& c) {
  if ( ! c.split || ! c.x_tier) return;
  RhomArgs a;
  std::memset(&a, 0, sizeof(a));
  a.blocks = c.d_xblock.p;
  a.nblocks = 1;
  a.lvlptr = c.d_lvlptr.p;
  a.kid0 = c.d_kid0.p;
  a.kid1 = c.d_kid1.p;
  a.in = c.d_xrhom.p;
  a.nc = c.d_xnc.p;
  const size_t smem = sizeof(double)*2*static_cast<size_t>(c.x_nl);
  LaunchTimer lt(c, CEDR_B200_TAG_RHOM, 1);
  rhom_kernel<<<1, kThreads, smem, c.stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  ++c.last_launches;
}

// Sweep of the expanded tier for a fast class: up over the sub-roots, root_compute, down.
template <int CLS> void launch_top_x_cls (cedr_b200_cdr& c, int cls) {
  SweepArgs a;
  std::memset(&a, 0, sizeof(a));
  a.blocks = c.d_xblock.p;
  a.nblocks = 1;
  a.lvlptr = c.d_lvlptr.p;
  a.kid0 = c.d_kid0.p;
  a.kid1 = c.d_kid1.p;
  a.nc = c.d_xnc.p;
  a.in = c.d_xrec.p;
  a.in_ld = c.x_ld;
  a.out = c.d_xsol.p;
  a.out_ld = c.x_ld;
  a.trcr_row = c.d_trcr_row.p;
  a.trcr_prob = c.d_trcr_prob.p;
  a.tracers = c.d_cls_tracers[cls].p;
  a.ntr = static_cast<int>(c.cls_tracers[cls].size());
  a.prefer_mass_con = c.prefer_mass_con;
  a.qglob = c.d_qglob.p;
  a.caas_scal = c.d_caas_scal.p;
  const size_t smem = sizeof(double)*4*static_cast<size_t>(c.x_nl + c.x_ni);
  sweep_smem_configure<CLS, MODE_TOP>(smem);
  LaunchTimer lt(c, CEDR_B200_TAG_TOP, 1);
  sweep_kernel<CLS, MODE_TOP><<<a.ntr, kThreads, smem, c.stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  ++c.last_launches;
}

void launch_top_x (cedr_b200_cdr& c, int cls) {
  switch (cls) {
  case CLS_ST: launch_top_x_cls<CLS_ST>(c, cls); break;
  case CLS_CST: launch_top_x_cls<CLS_CST>(c, cls); break;
  case CLS_T: launch_top_x_cls<CLS_T>(c, cls); break;
  case CLS_CT: launch_top_x_cls<CLS_CT>(c, cls); break;
  case CLS_NN: launch_top_x_cls<CLS_NN>(c, cls); break;
  case CLS_CNN: launch_top_x_cls<CLS_CNN>(c, cls); break;
  }
}

// Build the expanded tier: every leaf b of the tier-1 block (a tier-0 block root) becomes
// a perfect subtree of depth S whose 2^S leaves are block b's depth-S nodes.
void build_split (cedr_b200_cdr& c) {
  c.split = 0;
  if ( ! c.fast_ok || c.is_caas || std::getenv("CEDR_B200_NO_SPLIT")) return;
  if (c.plan.tiers.size() != 2 || c.plan.tiers[1].blocks.size() != 1) return;
  const int nl = c.plan.tiers[1].nleaves;
  // Multi-rank runs exchange block roots (the expanded tier would multiply the message by
  // 2^S and its replicated sweep does not shrink with the rank count): the split is local,
  // depths 0..S-1 of the OWN blocks are summed / solved by the mid kernels either side of
  // the replicated tier-1 sweep (run_qlt).
  if (c.nranks > 1) {
    if (std::getenv("CEDR_B200_NO_MULTI_SPLIT")) return;
    c.split = 3;
    c.x_tier = false;
    c.x_ld = round_up(static_cast<long long>(8)*nl, 16);
    const size_t nt = std::max<size_t>(1, c.trcr_prob.size());
    c.d_xrhom.alloc(c.x_ld);
    c.d_xrec.alloc(4*nt*c.x_ld);
    c.d_xsol.alloc(nt*c.x_ld);
    return;
  }
  const int S = 8*nl <= 2048 ? 3 : 4*nl <= 2048 ? 2 : 0;
  if (S == 0) return;
  const Shape& s1 = c.plan.shapes[c.plan.tiers[1].blocks[0].shape];
  const int E = 1 << S, nx = (E - 1)*nl;     // expansion nodes
  Shape x;
  x.nl = E*nl;
  x.ni = nx + s1.ni;
  x.nlev = s1.nlev + S;
  // Heights 1..S: expansion nodes of depth d = S - h, per block b and position p.
  std::vector<int> hstart(S + 2, 0);   // first internal index of height h
  for (int h = 1; h <= S; ++h) hstart[h + 1] = hstart[h] + nl*(1 << (S - h));
  auto xnode = [&] (int b, int d, int p) { return x.nl + hstart[S - d] + b*(1 << d) + p; };
  x.kid0.resize(x.ni);
  x.kid1.resize(x.ni);
  x.lvlptr.assign(1, 0);
  for (int h = 1; h <= S; ++h) {
    const int d = S - h;
    for (int b = 0; b < nl; ++b)
      for (int p = 0; p < (1 << d); ++p) {
        const int me = xnode(b, d, p) - x.nl;
        if (d == S - 1) { x.kid0[me] = b*E + 2*p; x.kid1[me] = b*E + 2*p + 1; }
        else { x.kid0[me] = xnode(b, d + 1, 2*p); x.kid1[me] = xnode(b, d + 1, 2*p + 1); }
      }
    x.lvlptr.push_back(hstart[h + 1]);
  }
  // The tier-1 block's own internal nodes, with its leaves replaced by the block roots.
  auto map1 = [&] (int id) { return id < nl ? xnode(id, 0, 0) : x.nl + nx + (id - nl); };
  for (int j = 0; j < s1.ni; ++j) {
    x.kid0[nx + j] = map1(s1.kid0[j]);
    x.kid1[nx + j] = map1(s1.kid1[j]);
  }
  for (int l = 1; l <= s1.nlev; ++l) x.lvlptr.push_back(nx + s1.lvlptr[l]);
  if (s1.ni == 0) {
    // One tier-0 block: its depth-0 expansion node is the root (nothing to append).
  }
  x.dev_lvlptr_off = static_cast<int>(c.plan.dev_lvlptr.size());
  x.dev_kid_off = static_cast<int>(c.plan.dev_kid0.size());
  c.plan.dev_lvlptr.insert(c.plan.dev_lvlptr.end(), x.lvlptr.begin(), x.lvlptr.end());
  c.plan.dev_kid0.insert(c.plan.dev_kid0.end(), x.kid0.begin(), x.kid0.end());
  c.plan.dev_kid1.insert(c.plan.dev_kid1.end(), x.kid1.begin(), x.kid1.end());
  BlockDev bd;
  std::memset(&bd, 0, sizeof(bd));
  bd.leaf0 = 0;
  bd.nl = x.nl;
  bd.ni = x.ni;
  bd.nlev = x.nlev;
  bd.lvlptr_off = x.dev_lvlptr_off;
  bd.kid_off = x.dev_kid_off;
  bd.ibase = 0;
  bd.ftab_off = -1;
  bd.gidx = 0;
  c.d_xblock.upload(std::vector<BlockDev>(1, bd));
  c.split = S;
  c.x_tier = true;
  c.x_nl = x.nl;
  c.x_ni = x.ni;
  c.x_ld = round_up(x.nl, 16);
  const size_t nt = std::max<size_t>(1, c.trcr_prob.size());
  c.d_xnc.alloc(x.ni);
  c.d_xrhom.alloc(c.x_ld);
  c.d_xrec.alloc(4*nt*c.x_ld);
  c.d_xsol.alloc(nt*c.x_ld);
}

void launch_up (cedr_b200_cdr& c, int cls, int k) {
  if (k == 0 && c.fast_ok && fast_class(cls, MODE_UP)) launch_fast_up(c, cls);
  else launch_sweep_any(c, cls, k, MODE_UP, base_args(c, cls, k));
}

void launch_down (cedr_b200_cdr& c, int cls, int k) {
  if (k == 0 && c.fast_ok && fast_class(cls, MODE_DOWN)) launch_fast_down(c, cls);
  else launch_sweep_any(c, cls, k, MODE_DOWN, base_args(c, cls, k));
}

// Doubles per rank in the exchange message: per owned block its index, its root's rhom
// and the 4 nt record words. (Multi-rank runs exchange block roots: split == 0.)
size_t exchange_count (const cedr_b200_cdr& c) {
  if (c.repl) return static_cast<size_t>(c.nrows)*c.nlcl_max;
  return static_cast<size_t>(c.nown_max)*(1 + (4*c.trcr_prob.size() + 1));
}

int grid_for (long long n) {
  return static_cast<int>(std::max<long long>(1, std::min<long long>((n + kThreads - 1)/kThreads,
                                                                   148*32)));
}

// This rank's tier-0 block roots -> the exchange message.
void exchange_pack (cedr_b200_cdr& c, bool with_rhom) {
  LaunchTimer lt(c, CEDR_B200_TAG_EXCHANGE, 0);
  pack_kernel<<<grid_for(exchange_count(c)), kThreads, 0, c.stream>>>(
    c.d_blocks[0].p, nblocks_dev(c, 0), c.nown_max, static_cast<int>(c.trcr_prob.size()), 1,
    with_rhom ? c.d_rhom_tier[1].p : nullptr, c.d_rec[1].p, c.tier_ld[1], c.xsend);
  CUDA_CHECK(cudaGetLastError());
  ++c.last_launches;
}

// Every rank's message -> the (replicated) tier-1 leaves.
void exchange_unpack (cedr_b200_cdr& c, bool with_rhom) {
  LaunchTimer lt(c, CEDR_B200_TAG_EXCHANGE, 1);
  unpack_kernel<<<grid_for(exchange_count(c)*c.nranks), kThreads, 0, c.stream>>>(
    c.xrecv, c.nranks, c.nown_max, static_cast<int>(c.trcr_prob.size()), 1,
    with_rhom ? c.d_rhom_tier[1].p : nullptr, c.d_rec[1].p, c.tier_ld[1],
    static_cast<long long>(exchange_count(c)));
  CUDA_CHECK(cudaGetLastError());
  ++c.last_launches;
}

// Word of the arena's 64-word header that counts this rank's exchanges (the flags written by
// the peers are words 0..nranks-1, nranks <= 16).
constexpr int kP2pEpochWord = 48;

size_t p2p_arena_doubles (const cedr_b200_cdr& c) {
  return 64 + 2*exchange_count(c)*c.nranks;
}

// pack + all-gather in one step over peer memory, then the epoch barrier.
void exchange_p2p (cedr_b200_cdr& c, bool with_rhom) {
  ++c.p2p_epoch;
  const size_t cnt = exchange_count(c);
  const int parity = static_cast<int>(c.p2p_epoch & 1);
  PeerPtrs pp;
  std::memset(&pp, 0, sizeof(pp));
  for (int r = 0; r < c.nranks; ++r) {
    double* base = static_cast<double*>(c.p2p_peer[r]);
    pp.flags[r] = reinterpret_cast<unsigned long long*>(base);
    pp.recv[r] = base + 64 + parity*cnt*c.nranks;
  }
  {
    LaunchTimer lt(c, CEDR_B200_TAG_EXCHANGE, 0);
    pack_p2p_kernel<<<dim3(grid_for(cnt), c.nranks), kThreads, 0, c.stream>>>(
      c.d_blocks[0].p, nblocks_dev(c, 0), c.nown_max, static_cast<int>(c.trcr_prob.size()), 1,
      with_rhom ? c.d_rhom_tier[1].p : nullptr, c.d_rec[1].p, c.tier_ld[1], pp, c.rank,
      c.nranks, static_cast<long long>(cnt));
    CUDA_CHECK(cudaGetLastError());
    ++c.last_launches;
  }
  {
    LaunchTimer lt(c, CEDR_B200_TAG_EXCHANGE, 2);
    p2p_barrier_kernel<<<1, 32, 0, c.stream>>>(
      pp, reinterpret_cast<unsigned long long*>(c.p2p_arena.p), c.rank, c.nranks,
      reinterpret_cast<unsigned long long*>(c.p2p_arena.p) + kP2pEpochWord, c.d_status.p,
      static_cast<unsigned long long>(env_int("CEDR_B200_P2P_TIMEOUT_MS", 30000))*1000000ull);
    CUDA_CHECK(cudaGetLastError());
    ++c.last_launches;
  }
  c.xrecv = c.p2p_arena.p + 64 + parity*cnt*c.nranks;
}

void exchange_allgather (cedr_b200_cdr& c) {
  cedr_b200_throw_if( ! c.allgather, "nranks > 1 but no all-gather was set "
                     "(cedr_b200_set_allgather)");
  const int e = c.allgather(c.allgather_ctx, c.xsend, c.xrecv, exchange_count(c), c.stream);
  cedr_b200_throw_if(e != 0, "the all-gather callback failed with code " << e);
}

// Depths 0..S-1 of this rank's own fast blocks, either side of the replicated tier-1 sweep
// (multi-rank, local split): sub-root records -> block-root records, and block-root
// masses -> sub-root masses.
void launch_mid (cedr_b200_cdr& c, int cls, bool down) {
  const fast::FastArgs a = fast_args(c, cls);
  if (a.ntr == 0 || a.nblocks == 0) return;
  const long long n = static_cast<long long>(a.nblocks)*a.ntr;
  LaunchTimer lt(c, down ? CEDR_B200_TAG_DOWN : CEDR_B200_TAG_UP, 1);
  if (down)
    fast::mid_down_kernel<<<grid_for(n), kThreads, 0, c.stream>>>(
      a, c.d_sol[1].p, c.tier_ld[1], cls == CLS_CST);
  else
    fast::mid_up_kernel<<<grid_for(n), kThreads, 0, c.stream>>>(
      a, c.d_rec[1].p, c.tier_ld[1], cls == CLS_CST);
  CUDA_CHECK(cudaGetLastError());
  ++c.last_launches;
}

// One small block is the whole problem: run() is a single launch (solo_kernel).
constexpr int kSoloMaxLeaves = 256;

bool solo_ok (const cedr_b200_cdr& c) {
  return c.nranks == 1 && c.plan.tiers.size() == 1 && c.plan.tiers[0].blocks.size() == 1 &&
    c.plan.tiers[0].max_nl <= kSoloMaxLeaves && ! std::getenv("CEDR_B200_NO_SOLO") &&
    (! c.is_caas || c.caas_sum_mode == CEDR_B200_CAAS_SUM_TREE);
}

template <int CLS> void launch_solo_cls (cedr_b200_cdr& c, int cls) {
  const SweepArgs a = base_args(c, cls, 0);
  if (a.ntr == 0) return;
  const size_t nn = 2*static_cast<size_t>(c.plan.tiers[0].max_nl);
  // 4 nn sweep fields, nn rhom sums, ni < nn/2 node constants, 2 ni + nlev + 1 table ints.
  const size_t smem = sizeof(double)*(5*nn + 2) + sizeof(dev::NodeConst)*(nn/2) +
    sizeof(int)*(3*(nn/2) + 2);
  LaunchTimer lt(c, CEDR_B200_TAG_TOP, 0);
  solo_kernel<CLS><<<a.ntr, kThreads, smem, c.stream>>>(a, c.solo_block);
  CUDA_CHECK(cudaGetLastError());
  ++c.last_launches;
}

void launch_solo (cedr_b200_cdr& c, int cls) {
  switch (cls) {
  case CLS_ST: launch_solo_cls<CLS_ST>(c, cls); break;
  case CLS_CST: launch_solo_cls<CLS_CST>(c, cls); break;
  case CLS_T: launch_solo_cls<CLS_T>(c, cls); break;
  case CLS_CT: launch_solo_cls<CLS_CT>(c, cls); break;
  case CLS_NN: launch_solo_cls<CLS_NN>(c, cls); break;
  case CLS_CNN: launch_solo_cls<CLS_CNN>(c, cls); break;
  case CLS_CAAS: launch_solo_cls<CLS_CAAS>(c, cls); break;
  }
}

// The side stream picks up after everything queued on the CDR's stream so far.
void side_stream_fork (cedr_b200_cdr& c) {
  if ( ! c.side_stream) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&c.side_stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming));
  }
  CUDA_CHECK(cudaEventRecord(c.ev_fork, c.stream));
  CUDA_CHECK(cudaStreamWaitEvent(c.side_stream, c.ev_fork, 0));
}

// QLT::run, cedr_qlt.cpp:618-640. phase < 0: everything (one rank: no exchange);
// phase 0: up to the exchange message; phase 1: from the gathered messages on.
void run_qlt (cedr_b200_cdr& c, int phase) {
  const int ntiers = static_cast<int>(c.plan.tiers.size());
  const int top = ntiers - 1;
  const bool multi = c.nranks > 1;
  // Classes on the fast kernels hand their blocks' sub-roots to the expanded tier.
  auto via_x = [&] (int cls) {
    return c.split && c.x_tier && c.fast_ok && fast_class(cls, MODE_DOWN) &&
      ! (c.ring_ok && ring_class(cls) && ! c.bound.on);
  };
  if (solo_ok(c)) {
    for (int cls = 0; cls < CLS_CAAS; ++cls)
      if ( ! c.cls_tracers[cls].empty()) launch_solo(c, cls);
    return;
  }
  // Multi-rank: the fast classes' blocks are split locally (build_split).
  auto local_split = [&] (int cls) {
    return c.split && ! c.x_tier && c.fast_ok && three_field_class(cls);
  };
  // One rank, every class on the fast kernels: the rhom sweeps run on the side stream
  // while the first class's up-sweep runs on the CDR's (joined before its tier above).
  bool rhom_aside = ! multi && phase < 0 && ! c.profiling && ! c.ring_ok &&
    ! std::getenv("CEDR_B200_NO_SIDE_STREAM");
  for (int cls = 0; cls < CLS_CAAS && rhom_aside; ++cls)
    if ( ! c.cls_tracers[cls].empty() && ! via_x(cls)) rhom_aside = false;
  if (rhom_aside) {
    side_stream_fork(c);
    const cudaStream_t main_stream = c.stream;
    c.stream = c.side_stream;
    try {
      run_rhom(c, 0, ntiers);
      run_rhom_x(c);
    } catch (...) { c.stream = main_stream; throw; }
    c.stream = main_stream;
    CUDA_CHECK(cudaEventRecord(c.ev_join, c.side_stream));
  }
  // Several ranks: the tier-0 rhom sweep (the block roots' rhom goes into the exchange
  // message) runs beside the ranks' own up-sweeps in the same way.
  bool rhom_aside_multi = multi && phase <= 0 && ! c.profiling && c.fast_ok &&
    ! std::getenv("CEDR_B200_NO_SIDE_STREAM");
  if (rhom_aside_multi) {
    side_stream_fork(c);
    const cudaStream_t main_stream = c.stream;
    c.stream = c.side_stream;
    try { run_rhom(c, 0, 1); } catch (...) { c.stream = main_stream; throw; }
    c.stream = main_stream;
    CUDA_CHECK(cudaEventRecord(c.ev_join, c.side_stream));
  }
  if (phase <= 0 && ! rhom_aside && ! rhom_aside_multi) {
    run_rhom(c, 0, multi ? 1 : ntiers);
    if ( ! multi) run_rhom_x(c);
  }
  if (phase <= 0) {
    if (c.ring_ok) ring_gather(c);
    if (multi) {
      for (int cls = 0; cls < CLS_CAAS; ++cls)
        if ( ! c.cls_tracers[cls].empty()) {
          launch_up(c, cls, 0);
          if (local_split(cls)) launch_mid(c, cls, false);
        }
      if (rhom_aside_multi) CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.ev_join, 0));
      if (c.p2p_on && phase < 0) exchange_p2p(c, true); else exchange_pack(c, true);
    }
  }
  if (multi && phase < 0 && ! c.p2p_on) exchange_allgather(c);
  if (phase != 0) {
    if (multi) {
      exchange_unpack(c, true);
      run_rhom(c, 1, ntiers);
    }
    for (int cls = 0; cls < CLS_CAAS; ++cls) {
      if (c.cls_tracers[cls].empty()) continue;
      if (c.ring_ok && ring_class(cls) && ! c.bound.on) { launch_ring(c, cls); continue; }
      if (via_x(cls)) {
        if ( ! multi) launch_up(c, cls, 0);
        if (rhom_aside) {
          CUDA_CHECK(cudaStreamWaitEvent(c.stream, c.ev_join, 0));
          rhom_aside = false;
        }
        launch_top_x(c, cls);
        launch_down(c, cls, 0);
        continue;
      }
      for (int k = multi ? 1 : 0; k < top; ++k) launch_up(c, cls, k);
      launch_sweep_any(c, cls, top, MODE_TOP, base_args(c, cls, top));
      for (int k = top - 1; k >= 0; --k) {
        if (k == 0 && multi && local_split(cls)) launch_mid(c, cls, true);
        launch_down(c, cls, k);
      }
    }
  }
}

// CAAS::run, cedr_caas.cpp:258-270; phases as for run_qlt.
void run_caas (cedr_b200_cdr& c, int phase) {
  if (c.caas_sum_mode == CEDR_B200_CAAS_SUM_USER) {
    // CAAS::run with a UserAllReducer, cedr_caas.cpp:258-270.
    cedr_b200_throw_if( ! c.user_reducer, "CEDR_B200_CAAS_SUM_USER but no reducer was set "
                       "(cedr_b200_caas_set_user_reducer)");
    if (phase == 1) return;
    const int nt = static_cast<int>(c.trcr_prob.size());
    const int nlocal = c.nlcl/c.user_naccum;
    {
      LaunchTimer lt(c, CEDR_B200_TAG_UP, 0);
      caas_user_partials_kernel<<<grid_for(static_cast<long long>(nlocal)*nt), kThreads, 0,
                                  c.stream>>>(row_map(c, CLS_CAAS), nlocal, c.user_naccum,
                                              c.d_trcr_prob.p, nt, c.usend);
      CUDA_CHECK(cudaGetLastError());
      ++c.last_launches;
    }
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    const int e = c.user_reducer(c.user_reducer_ctx, c.usend, c.urecv, nlocal, 4*nt, c.stream);
    cedr_b200_throw_if(e != 0, "the UserAllReducer failed with code " << e);
    {
      LaunchTimer lt(c, CEDR_B200_TAG_TOP, 0);
      caas_scal_from_recv_kernel<<<(nt + 127)/128, 128, 0, c.stream>>>(c.urecv, nt,
                                                                     c.d_caas_scal.p);
      CUDA_CHECK(cudaGetLastError());
      ++c.last_launches;
    }
    LaunchTimer lt(c, CEDR_B200_TAG_CAAS_ADJUST, 0);
    caas_adjust_kernel<<<grid_for(static_cast<long long>(c.nlcl)*nt), kThreads, 0, c.stream>>>(
      row_map(c, CLS_CAAS), c.nlcl, c.d_caas_scal.p, nt);
    CUDA_CHECK(cudaGetLastError());
    ++c.last_launches;
    return;
  }
  const bool multi = c.nranks > 1;
  if (c.caas_sum_mode == CEDR_B200_CAAS_SUM_SEQUENTIAL) {
    cedr_b200_throw_if(multi, "CEDR_B200_CAAS_SUM_SEQUENTIAL is a one-rank mode (the order of "
                       "an MPI_Allreduce is unspecified, cedr_caas.cpp:203-209)");
    if (phase == 1) return;
    const int nt = static_cast<int>(c.trcr_prob.size());
    {
      LaunchTimer lt(c, CEDR_B200_TAG_UP, 0);
      caas_seq_sums_kernel<<<(nt + 127)/128, 128, 0, c.stream>>>(
        row_map(c, CLS_CAAS), c.nlcl, c.d_trcr_prob.p, nt, c.d_caas_scal.p);
      CUDA_CHECK(cudaGetLastError());
      ++c.last_launches;
    }
    const long long n = static_cast<long long>(c.nlcl)*nt;
    LaunchTimer lt(c, CEDR_B200_TAG_CAAS_ADJUST, 0);
    caas_adjust_kernel<<<grid_for(n), kThreads, 0, c.stream>>>(row_map(c, CLS_CAAS), c.nlcl,
                                                             c.d_caas_scal.p, nt);
    CUDA_CHECK(cudaGetLastError());
    ++c.last_launches;
    return;
  }
  if (c.ring_ok && ! c.bound.on) { launch_ring(c, CLS_CAAS); return; }
  if (solo_ok(c)) { launch_solo(c, CLS_CAAS); return; }
  if (c.cluster_ok) { launch_cluster_caas(c); return; }
  const int ntiers = static_cast<int>(c.plan.tiers.size());
  const int top = ntiers - 1;
  if (phase <= 0 && multi) {
    launch_up(c, CLS_CAAS, 0);
    if (c.p2p_on && phase < 0) exchange_p2p(c, false); else exchange_pack(c, false);
  }
  if (multi && phase < 0 && ! c.p2p_on) exchange_allgather(c);
  if (phase == 0) return;
  if (multi) exchange_unpack(c, false);
  for (int k = multi ? 1 : 0; k < top; ++k) launch_up(c, CLS_CAAS, k);
  launch_sweep_any(c, CLS_CAAS, top, MODE_TOP, base_args(c, CLS_CAAS, top));
  const int nt = static_cast<int>(c.trcr_prob.size());
  const long long n = static_cast<long long>(c.nlcl)*nt;
  const int grid = static_cast<int>(std::min<long long>((n + kThreads - 1)/kThreads, 148*32));
  LaunchTimer lt(c, CEDR_B200_TAG_CAAS_ADJUST, 0);
  caas_adjust_kernel<<<grid, kThreads, 0, c.stream>>>(row_map(c, CLS_CAAS), c.nlcl,
                                                      c.d_caas_scal.p, nt);
  CUDA_CHECK(cudaGetLastError());
  ++c.last_launches;
}

void build_plan (cedr_b200_cdr& c) {
  c.plan.build(c.ncells, static_cast<int>(c.tree_cellidx.size()), c.tree_root,
               c.tree_kids.data(), c.tree_cellidx.data(),
               c.tree_rank.empty() ? nullptr : c.tree_rank.data(), c.max_block_leaves);
  // Local cell indices: this rank's leaves in DFS order (QLT::init_ordinals,
  // cedr_qlt.cpp:220-227; the reference numbers a rank's leaf slots the same way).
  c.gci2lci.clear();
  c.leaf_lci.assign(c.ncells, -1);
  c.nlcl = 0;
  for (int i = 0; i < c.ncells; ++i)
    if (c.plan.leaf_rank[i] == c.rank) {
      c.leaf_lci[i] = c.nlcl;
      c.gci2lci[c.plan.lci2gci[i]] = c.nlcl;
      ++c.nlcl;
    }
  // Subtree partition: a rank owns whole tier-0 blocks; everything above
  // tier 0 is replicated on every rank.
  c.own_blocks.clear();
  c.partition_error.clear();
  const std::vector<Block>& blocks = c.plan.tiers[0].blocks;
  for (size_t b = 0; b < blocks.size(); ++b) {
    if (c.nranks == 1 || blocks[b].owner == c.rank) {
      c.own_blocks.push_back(static_cast<int>(b));
      continue;
    }
    if (blocks[b].owner == -1 && c.partition_error.empty())
      for (int i = blocks[b].leaf0; i < blocks[b].leaf0 + blocks[b].nl; ++i)
        if (c.plan.leaf_rank[i] == c.rank) {
          // Reported by end_tracer_declarations: max_block_leaves may still change.
          std::stringstream ss;
          ss << "with nranks > 1 the cells of a rank must form whole blocks of the tree "
            "plan (subtree partition); block " << b << " of " << blocks[b].nl
             << " leaves mixes ranks. General cell->rank maps are not supported yet.";
          c.partition_error = ss.str();
          break;
        }
  }
  if (c.nranks > 1) {
    c.nown_max = static_cast<int>((blocks.size() + c.nranks - 1)/c.nranks);
    if ( ! c.is_caas) {
      // QLT and the reducer know every leaf's rank: size the message for the rank that
      // owns the most blocks, so any assignment of whole blocks to ranks works.
      std::vector<int> cnt(c.nranks, 0);
      for (size_t b = 0; b < blocks.size(); ++b)
        if (blocks[b].owner >= 0 && blocks[b].owner < c.nranks) ++cnt[blocks[b].owner];
      c.nown_max = std::max(1, *std::max_element(cnt.begin(), cnt.end()));
    }
    if (static_cast<int>(c.own_blocks.size()) > c.nown_max && c.partition_error.empty()) {
      std::stringstream ss;
      ss << "this rank owns " << c.own_blocks.size() << " blocks, more than "
        "ceil(#blocks/#ranks) = " << c.nown_max << ": uneven partitions are not supported yet";
      c.partition_error = ss.str();
    }
  }
}

// BfbTreeAllReducer::allreduce after the leaf fill: tree-ordered sums of every field up
// to the root (one all-gather of block roots when nranks > 1); results in d_qglob[field].
void run_bfb (cedr_b200_cdr& c, int phase) {
  const bool multi = c.nranks > 1;
  const int ntiers = static_cast<int>(c.plan.tiers.size());
  const int top = ntiers - 1;
  if (phase <= 0 && multi) {
    launch_up(c, CLS_BFB, 0);
    if (c.p2p_on && phase < 0) exchange_p2p(c, false); else exchange_pack(c, false);
  }
  if (multi && phase < 0 && ! c.p2p_on) exchange_allgather(c);
  if (phase == 0) return;
  if (multi) exchange_unpack(c, false);
  for (int k = multi ? 1 : 0; k < top; ++k) launch_up(c, CLS_BFB, k);
  launch_sweep_any(c, CLS_BFB, top, MODE_TOP, base_args(c, CLS_BFB, top));
}

void run_repl (cedr_b200_cdr& c, int phase) {
  cedr_b200_cdr& w = *c.whole;
  const int nt = static_cast<int>(c.trcr_prob.size());
  if (phase <= 0) {
    LaunchTimer lt(c, CEDR_B200_TAG_EXCHANGE, 0);
    repl_pack_kernel<<<grid_for(exchange_count(c)), kThreads, 0, c.stream>>>(
      c.in, c.ld, c.nlcl, c.nrows, c.nlcl_max, c.xsend);
    CUDA_CHECK(cudaGetLastError());
    ++c.last_launches;
  }
  if (phase < 0) exchange_allgather(c);
  if (phase == 0) return;
  {
    LaunchTimer lt(c, CEDR_B200_TAG_EXCHANGE, 1);
    repl_unpack_kernel<<<grid_for(exchange_count(c)*c.nranks), kThreads, 0, c.stream>>>(
      c.xrecv, c.nranks, c.nrows, c.nlcl_max, c.d_pos.p, c.d_pos_off.p, w.in, w.ld);
    CUDA_CHECK(cudaGetLastError());
    ++c.last_launches;
  }
  w.stream = c.stream;
  w.last_launches = 0;
  run_qlt(w, -1);
  c.last_launches += w.last_launches;
  repl_keep_own_kernel<<<grid_for(static_cast<long long>(c.nlcl)*nt), kThreads, 0, c.stream>>>(
    w.out, w.ld, c.d_pos.p + c.pos_me_off, c.nlcl, nt, c.out, c.ld);
  CUDA_CHECK(cudaGetLastError());
  ++c.last_launches;
}

void run_any (cedr_b200_cdr& c, int phase) {
  if (c.repl) { run_repl(c, phase); return; }
  if (c.is_bfb) run_bfb(c, phase);
  else if (c.is_caas) run_caas(c, phase);
  else run_qlt(c, phase);
  if (c.p2p_on && phase != 0 && ! c.is_bfb && ! c.bound.on) {
    // (see poison_kernel) CAAS results are its Qm rows, QLT's the out rows.
    const int nt = static_cast<int>(c.trcr_prob.size());
    if (c.is_caas)
      poison_kernel<<<148, kThreads, 0, c.stream>>>(c.d_status.p, c.in, c.ld, c.nlcl, nt, 2,
                                                    c.caas_need_conserve ? 4 : 3);
    else
      poison_kernel<<<148, kThreads, 0, c.stream>>>(c.d_status.p, c.out, c.ld, c.nlcl, nt, 0, 1);
    CUDA_CHECK(cudaGetLastError());
    ++c.last_launches;
  }
}

// run() as a CUDA graph: the launches of run_any are captured once per exchange-buffer
// parity and replayed, which takes the per-launch host cost and the gaps between the dozen
// small kernels of a multi-rank run() off the critical path (at 8 GPUs they were ~10 % of the
// step). Only for runs that are pure stream work: no profiling events, no host callbacks
// (all-gather hook, UserAllReducer), no cooperative ring kernel.
bool graph_eligible (const cedr_b200_cdr& c) {
  if (c.graph_mode == 0 || std::getenv("CEDR_B200_NO_GRAPH")) return false;
  if (c.profiling || c.repl || c.is_bfb || c.ring_ok || solo_ok(c)) return false;
  if (c.is_caas && c.caas_sum_mode != CEDR_B200_CAAS_SUM_TREE) return false;
  if (c.nranks > 1 && ! c.p2p_on) return false;
  if (c.graph_mode == 1 || c.nranks > 1) return true;
  // Auto, one rank: short runs, where the gaps between run()'s launches are a visible share
  // (ne30 x 2,880 tracers: 0.338 -> 0.322 ms; x 320: 0.095 -> 0.080 ms; ne120 x 1,280:
  // no difference, profiles/r02c_graph_ab.txt).
  return static_cast<long long>(c.nlcl)*static_cast<long long>(c.trcr_prob.size()) <=
    static_cast<long long>(env_int("CEDR_B200_GRAPH_AUTO_MAX", 32000000));
}

void run_graphed (cedr_b200_cdr& c) {
  if ( ! graph_eligible(c)) { run_any(c, -1); return; }
  // The first run stays plain: it sets kernel attributes, which is not stream work.
  if (c.graph_plain_runs < 1) { ++c.graph_plain_runs; run_any(c, -1); return; }
  const int parity = c.nranks > 1 ? static_cast<int>((c.p2p_epoch + 1) & 1) : 0;
  cedr_b200_cdr::RunGraph& g = c.graph[parity];
  if ( ! g.exec) {
    // Captured on a stream of our own (the caller's may be the legacy default stream,
    // which cannot capture) and replayed on the caller's.
    if ( ! c.cap_stream)
      CUDA_CHECK(cudaStreamCreateWithFlags(&c.cap_stream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    const cudaStream_t user = c.stream;
    CUDA_CHECK(cudaStreamBeginCapture(c.cap_stream, cudaStreamCaptureModeThreadLocal));
    c.stream = c.cap_stream;
    try {
      run_any(c, -1);      // advances the host epoch mirror itself
    } catch (...) {
      c.stream = user;
      cudaStreamEndCapture(c.cap_stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    c.stream = user;
    CUDA_CHECK(cudaStreamEndCapture(c.cap_stream, &graph));
    const cudaError_t e = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    CUDA_CHECK(e);
    g.launches = c.last_launches;
  } else if (c.nranks > 1) {
    ++c.p2p_epoch;
    c.xrecv = c.p2p_arena.p + 64 + parity*exchange_count(c)*c.nranks;
  }
  CUDA_CHECK(cudaGraphLaunch(g.exec, c.stream));
  c.last_launches = g.launches;
}

void get_buffers_sizes (cedr_b200_cdr& c, size_t& b1, size_t& b2) {
  cedr_b200_throw_if(c.declaring, "end_tracer_declarations must be called first.");
  b1 = static_cast<size_t>(c.nrows)*c.ld;
  b2 = c.is_caas ? 0 : c.trcr_prob.size()*static_cast<size_t>(c.ld);
  // CAAS with a UserAllReducer: send (nlocal, 4 nt) then recv (4 nt), cedr_caas.cpp:75-90.
  if (c.is_caas && c.caas_sum_mode == CEDR_B200_CAAS_SUM_USER)
    b2 = 4*c.trcr_prob.size()*(static_cast<size_t>(c.nlcl/c.user_naccum) + 1);
}

void finish_setup (cedr_b200_cdr& c) {
  cedr_b200_throw_if(c.declaring, "end_tracer_declarations must be called first.");
  if (c.finished) return;
  size_t b1, b2;
  get_buffers_sizes(c, b1, b2);
  if (! c.user_buffers) {
    c.in_own.alloc(b1);
    c.out_own.alloc(b2);
    c.in = c.in_own.p;
    c.out = c.out_own.p;
    // Rows are padded to ld; keep the padding defined.
    CUDA_CHECK(cudaMemsetAsync(c.in, 0, b1*sizeof(double), c.stream));
    if (b2) CUDA_CHECK(cudaMemsetAsync(c.out, 0, b2*sizeof(double), c.stream));
  }
  if (c.is_caas && c.caas_sum_mode == CEDR_B200_CAAS_SUM_USER) {
    cedr_b200_throw_if( ! c.user_reducer, "CEDR_B200_CAAS_SUM_USER but no reducer was set "
                       "(cedr_b200_caas_set_user_reducer)");
    if (c.user_buffers && c.out) c.usend = c.out;
    else { c.usend_own.alloc(b2); c.usend = c.usend_own.p; }
    c.urecv = c.usend + 4*c.trcr_prob.size()*static_cast<size_t>(c.nlcl/c.user_naccum);
  }
  if (c.is_caas) c.out = c.in;
  const int nt = static_cast<int>(c.trcr_prob.size());
  c.d_trcr_row.upload(c.trcr_row);
  c.d_trcr_prob.upload(c.trcr_prob);
  if (c.repl) {
    // Only the caller-facing buffers, the exchange buffers and the cell -> leaf table live
    // here; the sweeps belong to the whole-tree CDR.
    std::vector<int> pos, off(c.nranks + 1, 0);
    for (int r = 0; r < c.nranks; ++r) {
      for (int i = 0; i < c.ncells; ++i)
        if (c.plan.leaf_rank[i] == r) pos.push_back(i);
      off[r + 1] = static_cast<int>(pos.size());
    }
    c.pos_me_off = off[c.rank];
    c.d_pos.upload(pos);
    c.d_pos_off.upload(off);
    if ( ! c.xsend) {
      c.xsend_own.alloc(exchange_count(c));
      c.xrecv_own.alloc(exchange_count(c)*c.nranks);
      c.xsend = c.xsend_own.p;
      c.xrecv = c.xrecv_own.p;
    }
    c.whole->stream = c.stream;
    finish_setup(*c.whole);
    c.finished = true;
    return;
  }
  for (int k = 0; k < NCLS; ++k) c.d_cls_tracers[k].upload(c.cls_tracers[k]);
  c.fast_ok = c.fast_enabled && c.plan.tier0_fast &&
    reinterpret_cast<uintptr_t>(c.in) % 16 == 0 &&
    (c.is_caas || reinterpret_cast<uintptr_t>(c.out) % 16 == 0) &&
    ! std::getenv("CEDR_B200_NO_FAST");
  build_split(c);
  c.d_lvlptr.upload(c.plan.dev_lvlptr);
  c.d_kid0.upload(c.plan.dev_kid0);
  c.d_kid1.upload(c.plan.dev_kid1);
  const int ntiers = static_cast<int>(c.plan.tiers.size());
  c.d_blocks = std::vector<DevBuf<BlockDev> >(ntiers);
  c.d_rhom_tier = std::vector<DevBuf<double> >(ntiers);
  c.d_rec = std::vector<DevBuf<double> >(ntiers);
  c.d_sol = std::vector<DevBuf<double> >(ntiers);
  c.tier_ld.assign(ntiers, 0);
  for (int k = 0; k < ntiers; ++k) {
    const Tier& tier = c.plan.tiers[k];
    // Tier 0: only this rank's blocks, their leaves as local cell indices.
    std::vector<int> list;
    if (k == 0) list = c.own_blocks;
    else for (size_t b = 0; b < tier.blocks.size(); ++b) list.push_back(static_cast<int>(b));
    std::vector<BlockDev> hb(list.size());
    for (size_t b = 0; b < hb.size(); ++b) {
      const Block& blk = tier.blocks[list[b]];
      const Shape& sh = c.plan.shapes[blk.shape];
      hb[b].gidx = list[b];
      hb[b].leaf0 = k == 0 ? c.leaf_lci[blk.leaf0] : blk.leaf0;
      hb[b].nl = blk.nl;
      hb[b].ni = sh.ni;
      hb[b].nlev = sh.nlev;
      hb[b].lvlptr_off = sh.dev_lvlptr_off;
      hb[b].kid_off = sh.dev_kid_off;
      hb[b].ibase = blk.ibase;
      hb[b].ftab_off = sh.fast ? sh.dev_dtab_off : -1;
      hb[b].fpair_off = sh.dev_ptab_off;
      hb[b].fperm_off = sh.dev_perm_off;
      hb[b].fpent_off = sh.dev_pent_off;
      hb[b].fperm_up_off = sh.dev_perm_up_off;
      hb[b].fpos_off = sh.dev_fpos_off;
      hb[b].npairs = sh.fast ? static_cast<int>(sh.ptab.size()) : 0;
      hb[b].fbase = blk.ibase;
    }
    if (hb.empty()) hb.resize(1);
    if (k == 0) c.solo_block = hb[0];
    c.d_blocks[k].upload(hb);
    c.tier_ld[k] = k == 0 ? c.ld : round_up(tier.nleaves, 16);
    if (k > 0) {
      c.d_rhom_tier[k].alloc(c.tier_ld[k]);
      c.d_rec[k].alloc(static_cast<size_t>(4)*nt*c.tier_ld[k]);
      c.d_sol[k].alloc(static_cast<size_t>(nt)*c.tier_ld[k]);
    }
  }
  c.d_nc.alloc(std::max(1, c.plan.ninternal));
  c.d_dtab.upload(c.plan.dev_dtab);
  c.d_ptab.upload(c.plan.dev_ptab);
  c.d_perm.upload(c.plan.dev_perm);
  c.d_pent.upload(c.plan.dev_pent);
  c.d_fpos.upload(c.plan.dev_fpos);
  if (c.fast_ok) {
    c.d_fwq.alloc(std::max(1, c.plan.ninternal));
    c.d_frh.alloc(std::max(1, c.plan.ninternal));
    c.d_frq.alloc(std::max(1, c.plan.ninternal));
    if ( ! c.is_caas) {
      c.d_n7.alloc(static_cast<size_t>(384)*std::max<size_t>(1, c.own_blocks.size())*
                   std::max(1, nt));
      c.d_x7.alloc(static_cast<size_t>(128)*std::max<size_t>(1, c.own_blocks.size())*
                   std::max(1, nt));
    }
  }
  c.d_qglob.alloc(2*static_cast<size_t>(nt));
  c.d_caas_scal.alloc(2*static_cast<size_t>(nt));
  if (c.nranks > 1 && ! c.d_status.p) {
    c.d_status.alloc(1);
    CUDA_CHECK(cudaMemsetAsync(c.d_status.p, 0, sizeof(int), c.stream));
  }
  if (c.nranks > 1 && ! c.xsend) {
    c.xsend_own.alloc(exchange_count(c));
    c.xrecv_own.alloc(exchange_count(c)*c.nranks);
    c.xsend = c.xsend_own.p;
    c.xrecv = c.xrecv_own.p;
  }
  ring_setup(c);
  cluster_setup(c);
  upload_rowaddr(c);
  c.finished = true;
}

void fill_1d_tree (cedr_b200_cdr& c, int ncells, bool imbalanced) {
  cedr_b200_throw_if(c.nranks < 1 || c.nranks > ncells, "#GIDs < #ranks is not supported.");
  make_bisection_tree(ncells, imbalanced, c.tree_kids, c.tree_cellidx);
  c.tree_root = 0;
  c.tree_rank.assign(c.tree_cellidx.size(), 0);
  if (c.nranks > 1) {
    // oned::Mesh::rank, contiguous decomposition (cedr_tree.cpp:366-369).
    for (size_t i = 0; i < c.tree_cellidx.size(); ++i)
      if (c.tree_cellidx[i] >= 0)
        c.tree_rank[i] = std::min<int64_t>(c.nranks - 1,
                                           c.tree_cellidx[i]/(ncells/c.nranks));
  }
}

void require_device () {
  int n = 0;
  const cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    throw std::runtime_error("cedr_b200: no usable CUDA device (there is no CPU fallback)");
}

} // namespace

namespace {
// One thread per element, the element's values in registers (cedr_b200_local.hpp).
template <int N>
__global__ void __launch_bounds__(128)
local_solve_kernel (const int method, const int nprob, const int n, const double* w,
                    const double* a, const double* b, const double* xlo, const double* xhi,
                    const double* y, double* x, int* info, const long long sv,
                    const long long sp, const int max_its, const bool clip) {
  namespace L = cedr::local;
  const int p = blockIdx.x*blockDim.x + threadIdx.x;
  if (p >= nprob) return;
  const long long o = p*sp;
  const bool nonneg = method == CEDR_B200_LOCAL_NONNEG_LS || method == CEDR_B200_LOCAL_NONNEG_CAAS;
  const double bp = b[p];
  if (nonneg && bp < 0) { info[p] = -1; return; }    // x untouched, as in the reference
  L::Registers<N> v;
  v.gather(n, w ? w + o : nullptr, a ? a + o : nullptr, bp, nonneg ? nullptr : xlo + o,
           nonneg ? nullptr : xhi + o, y + o, sv);
  int r = 0;
  switch (method) {
  case CEDR_B200_LOCAL_QP: r = L::solve_qp(v, max_its); break;
  case CEDR_B200_LOCAL_CAAS: L::clip_and_spread(v, clip); break;
  case CEDR_B200_LOCAL_NONNEG_LS: r = L::solve_nonneg(v, L::Method::least_squares); break;
  case CEDR_B200_LOCAL_NONNEG_CAAS: r = L::solve_nonneg(v, L::Method::caas); break;
  case CEDR_B200_LOCAL_QP_2D: r = L::solve_qp2(v, clip, true); break;
  }
  v.scatter(x + o, sv);
  info[p] = r;
}
} // namespace

extern "C" {

const char* cedr_b200_last_error (void) { return g_err.c_str(); }

int cedr_b200_version (void) { return 100; }

int cedr_b200_device_available (void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
}

int cedr_b200_qlt_create (cedr_b200_cdr** out, int ncells, int nnodes, int root,
                          const int* kids, const int64_t* cellidx, const int* node_rank,
                          int prefer, int rank, int nranks) {
  return guarded([&] {
    cedr_b200_throw_if(! out, "null output pointer");
    require_device();
    std::unique_ptr<cedr_b200_cdr> c(new cedr_b200_cdr);
    c->rank = rank;
    c->nranks = nranks;
    c->prefer_mass_con = prefer != 0;
    c->ncells = ncells;
    cedr_b200_throw_if(nnodes < 1 || ! kids || ! cellidx, "bad tree arrays");
    c->tree_kids.assign(kids, kids + 2*static_cast<size_t>(nnodes));
    c->tree_cellidx.assign(cellidx, cellidx + nnodes);
    if (node_rank) c->tree_rank.assign(node_rank, node_rank + nnodes);
    c->tree_root = root;
    build_plan(*c);
    cedr_b200_throw_if(c->nlcl == 0, "QLT does not support 0 cells on a rank.");
    *out = c.release();
  });
}

int cedr_b200_qlt_create_1d (cedr_b200_cdr** out, int ncells, int imbalanced,
                             int prefer, int rank, int nranks) {
  return guarded([&] {
    cedr_b200_throw_if(! out, "null output pointer");
    require_device();
    cedr_b200_throw_if(nranks > ncells, "#GIDs < #ranks is not supported.");
    std::unique_ptr<cedr_b200_cdr> c(new cedr_b200_cdr);
    c->rank = rank;
    c->nranks = nranks;
    c->prefer_mass_con = prefer != 0;
    c->ncells = ncells;
    fill_1d_tree(*c, ncells, imbalanced != 0);
    build_plan(*c);
    cedr_b200_throw_if(c->nlcl == 0, "QLT does not support 0 cells on a rank.");
    *out = c.release();
  });
}

int cedr_b200_caas_create (cedr_b200_cdr** out, int nlclcells, int sum_mode,
                           int64_t cell0, int64_t ncells_global, int rank, int nranks) {
  return guarded([&] {
    cedr_b200_throw_if(! out, "null output pointer");
    require_device();
    cedr_b200_throw_if(nlclcells == 0, "CAAS does not support 0 cells on a rank.");
    if (sum_mode == CEDR_B200_CAAS_SUM_USER) {
      // The reducer owns the cross-rank sum: any cells, no exchange of ours.
      cell0 = 0;
      ncells_global = nlclcells;
      rank = 0;
      nranks = 1;
    }
    cedr_b200_throw_if(nranks == 1 && (cell0 != 0 || ncells_global != nlclcells),
                       "one rank: cell0 must be 0 and ncells_global == nlclcells");
    cedr_b200_throw_if(cell0 < 0 || cell0 + nlclcells > ncells_global ||
                       ncells_global > 0x7fffffffLL, "bad cell range");
    std::unique_ptr<cedr_b200_cdr> c(new cedr_b200_cdr);
    c->is_caas = true;
    c->rank = rank;
    c->nranks = nranks;
    c->caas_sum_mode = sum_mode;
    c->ncells = static_cast<int>(ncells_global);
    c->caas_cell0 = cell0;
    // The tree-ordered sums run over a bisection tree of ALL cells; this rank's cells
    // [cell0, cell0 + nlclcells) must be whole blocks of it. Other ranks' cells only
    // need to be "not mine" (-2).
    make_bisection_tree(c->ncells, false, c->tree_kids, c->tree_cellidx);
    c->tree_root = 0;
    c->tree_rank.assign(c->tree_cellidx.size(), 0);
    for (size_t i = 0; i < c->tree_cellidx.size(); ++i) {
      const int64_t ci = c->tree_cellidx[i];
      if (ci >= 0) c->tree_rank[i] = (ci >= cell0 && ci < cell0 + nlclcells) ? rank : -2;
    }
    build_plan(*c);
    cedr_b200_throw_if(c->nlcl != nlclcells, "internal: owned cell count mismatch");
    *out = c.release();
  });
}

// ---- BfbTreeAllReducer (cedr_bfb_tree_allreduce.hpp:15-55)

namespace {
void bfb_finish (cedr_b200_cdr& c, int nfield) {
  cedr_b200_throw_if(nfield < 1, "nfield must be >= 1");
  c.is_bfb = true;
  c.fast_enabled = false;      // the generic sweeps carry the one-word records
  c.ring_enabled = false;
  for (int j = 0; j < nfield; ++j) {
    c.trcr_prob.push_back(0);
    c.trcr_cls.push_back(CLS_BFB);
    c.trcr_row.push_back(j);
    c.cls_tracers[CLS_BFB].push_back(j);
  }
  c.nrows = nfield;
  c.ld = round_up(std::max(1, c.nlcl), 16);
  c.declaring = false;
  cedr_b200_throw_if( ! c.partition_error.empty(), c.partition_error);
  cedr_b200_throw_if(c.nranks > 1 && c.plan.tiers.size() < 2,
                     "with nranks > 1 the tree plan needs >= 2 tiers");
}
}

int cedr_b200_bfb_create (cedr_b200_cdr** out, int nleaf, int nnodes, int root,
                          const int* kids, const int64_t* cellidx, const int* node_rank,
                          int nfield, int max_block_leaves, int rank, int nranks) {
  return guarded([&] {
    cedr_b200_throw_if( ! out, "null output pointer");
    require_device();
    std::unique_ptr<cedr_b200_cdr> c(new cedr_b200_cdr);
    c->rank = rank;
    c->nranks = nranks;
    c->ncells = nleaf;
    if (max_block_leaves > 0) c->max_block_leaves = max_block_leaves;
    if (kids) {
      cedr_b200_throw_if(nnodes < 1 || ! cellidx, "bad tree arrays");
      c->tree_kids.assign(kids, kids + 2*static_cast<size_t>(nnodes));
      c->tree_cellidx.assign(cellidx, cellidx + nnodes);
      if (node_rank) c->tree_rank.assign(node_rank, node_rank + nnodes);
      c->tree_root = root;
    } else {
      fill_1d_tree(*c, nleaf, false);
    }
    build_plan(*c);
    cedr_b200_throw_if(c->nlcl == 0, "BfbTreeAllReducer: this rank owns no leaf");
    bfb_finish(*c, nfield);
    finish_setup(*c);
    *out = c.release();
  });
}

int cedr_b200_bfb_allreduce (cedr_b200_cdr* c, const double* send, double* recv,
                             int transpose, int phase) {
  return guarded([&] {
    cedr_b200_throw_if( ! c->is_bfb, "not a BfbTreeAllReducer");
    cedr_b200_throw_if(phase < -1 || phase > 1, "phase must be -1 (all), 0 or 1");
    const int nf = static_cast<int>(c->trcr_prob.size());
    if (phase <= 0) {
      c->last_launches = 0;
      c->ntimed = 0;
      bfb_fill_kernel<<<grid_for(static_cast<long long>(c->nlcl)*nf), kThreads, 0,
                        c->stream>>>(c->in, c->ld, c->nlcl, nf, transpose, send);
      CUDA_CHECK(cudaGetLastError());
      ++c->last_launches;
    }
    if (c->nranks == 1) { if (phase <= 0) run_bfb(*c, -1); }
    else run_bfb(*c, phase);
    if (phase != 0)
      CUDA_CHECK(cudaMemcpyAsync(recv, c->d_qglob.p, sizeof(double)*nf,
                                 cudaMemcpyDeviceToDevice, c->stream));
  });
}

int cedr_b200_caas_set_user_reducer (cedr_b200_cdr* c, cedr_b200_user_reducer_fn fn, void* ctx,
                                     int n_accum_in_place) {
  return guarded([&] {
    cedr_b200_throw_if( ! c->is_caas || c->caas_sum_mode != CEDR_B200_CAAS_SUM_USER,
                       "set_user_reducer needs a CAAS created with CEDR_B200_CAAS_SUM_USER");
    cedr_b200_throw_if( ! c->declaring, "set_user_reducer must precede end_tracer_declarations");
    cedr_b200_throw_if( ! fn, "null reducer");
    cedr_b200_throw_if(n_accum_in_place < 1 || c->nlcl % n_accum_in_place != 0,
                       "n_accum_in_place must be >= 1 and divide nlclcells");
    c->user_reducer = fn;
    c->user_reducer_ctx = ctx;
    c->user_naccum = n_accum_in_place;
  });
}

int cedr_b200_local_solve (int method, int nprob, int n, const double* w, const double* a,
                           const double* b, const double* xlo, const double* xhi,
                           const double* y, double* x, int* info, int64_t sv, int64_t sp,
                           int max_its, int clip, void* stream) {
  return guarded([&] {
    require_device();
    cedr_b200_throw_if(method < CEDR_B200_LOCAL_QP || method > CEDR_B200_LOCAL_QP_2D,
                       "unknown local method");
    cedr_b200_throw_if(n < 1 || n > 16, "local solvers: 1 <= n <= 16 (cedr_local_inl.hpp:310)");
    cedr_b200_throw_if(method == CEDR_B200_LOCAL_QP_2D && n != 2, "solve_1eq_bc_qp_2d: n is 2");
    const bool nonneg = method == CEDR_B200_LOCAL_NONNEG_LS || method == CEDR_B200_LOCAL_NONNEG_CAAS;
    cedr_b200_throw_if( ! b || ! y || ! x || ! info || ( ! nonneg && ( ! xlo || ! xhi)),
                       "null array");
    if (nprob <= 0) return;
    if (max_its <= 0) max_its = 100;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = (nprob + 127)/128;
#define CEDR_LOCAL_LAUNCH(N) local_solve_kernel<N><<<grid, 128, 0, st>>>(                  \
      method, nprob, n, w, a, b, xlo, xhi, y, x, info, sv, sp, max_its, clip != 0)
    if (n <= 2) CEDR_LOCAL_LAUNCH(2);
    else if (n <= 4) CEDR_LOCAL_LAUNCH(4);
    else if (n <= 8) CEDR_LOCAL_LAUNCH(8);
    else CEDR_LOCAL_LAUNCH(16);
#undef CEDR_LOCAL_LAUNCH
    CUDA_CHECK(cudaGetLastError());
  });
}

int cedr_b200_destroy (cedr_b200_cdr* c) {
  return guarded([&] { delete c; });
}

int cedr_b200_set_max_block_leaves (cedr_b200_cdr* c, int m) {
  return guarded([&] {
    cedr_b200_throw_if(c->finished, "set_max_block_leaves must precede finish_setup");
    cedr_b200_throw_if(m < 2 || m > 2048, "max_block_leaves must be in [2, 2048]");
    c->max_block_leaves = m;
    build_plan(*c);
  });
}

int cedr_b200_declare_tracer (cedr_b200_cdr* c, int problem_type, int rhomidx) {
  return guarded([&] {
    cedr_b200_throw_if(! c->declaring, "end_tracer_declarations was already called; "
                       "it is an error to call declare_tracer now.");
    cedr_b200_throw_if(rhomidx > 0, "rhomidx > 0 is not supported yet.");
    if (c->is_caas) {
      cedr_b200_throw_if(! (problem_type & CEDR_B200_SHAPEPRESERVE),
                         "CAAS does not support ! shapepreserve yet.");
      c->trcr_prob.push_back(problem_type);
      c->trcr_cls.push_back(CLS_CAAS);
      if (problem_type & CEDR_B200_CONSERVE) c->caas_need_conserve = true;
    } else {
      const int cls = qlt_class_of(problem_type);
      cedr_b200_throw_if(cls < 0, "Invalid problem type.");
      c->trcr_prob.push_back(qlt_canonical_type(cls));
      c->trcr_cls.push_back(cls);
    }
  });
}

namespace {
void end_tracer_declarations (cedr_b200_cdr& c) {
  cedr_b200_throw_if(! c.declaring, "end_tracer_declarations was already called.");
  if (c.is_caas)
    cedr_b200_throw_if(c.trcr_prob.size() == 0, "#tracers is 0.");
  const int nt = static_cast<int>(c.trcr_prob.size());
  c.trcr_row.resize(nt);
  int row = 1; // row 0 is rhom
  for (int k = 0; k < NCLS; ++k) c.cls_tracers[k].clear();
  for (int t = 0; t < nt; ++t) {
    c.trcr_row[t] = row;
    const int cls = c.trcr_cls[t];
    row += c.is_caas ? (c.caas_need_conserve ? 4 : 3) : qlt_l2r_words(cls);
    c.cls_tracers[cls].push_back(t);
  }
  c.nrows = row;
  c.ld = round_up(std::max(1, c.nlcl), 16);
  // QLT with a cell -> rank map that cuts blocks of the plan (every rank sees the same
  // plan, so all take this branch together): replicated mode.
  bool cut = false;
  if (c.nranks > 1 && ! c.is_caas && ! c.is_bfb) {
    cut = c.plan.tiers.size() < 2;
    for (const Block& b : c.plan.tiers[0].blocks) cut |= b.owner < 0;
  }
  if (cut) {
    c.repl = true;
    std::vector<int> cnt(c.nranks, 0);
    for (int i = 0; i < c.ncells; ++i) {
      const int r = c.plan.leaf_rank[i];
      cedr_b200_throw_if(r < 0 || r >= c.nranks, "leaf " << i << " has rank " << r);
      ++cnt[r];
    }
    c.nlcl_max = *std::max_element(cnt.begin(), cnt.end());
    c.whole.reset(new cedr_b200_cdr);
    cedr_b200_cdr& w = *c.whole;
    w.prefer_mass_con = c.prefer_mass_con;
    w.ncells = c.ncells;
    w.tree_kids = c.tree_kids;
    w.tree_cellidx = c.tree_cellidx;
    w.tree_root = c.tree_root;
    w.max_block_leaves = c.max_block_leaves;
    w.fast_enabled = c.fast_enabled;
    build_plan(w);
    w.trcr_prob = c.trcr_prob;
    w.trcr_cls = c.trcr_cls;
    end_tracer_declarations(w);
    c.declaring = false;
    return;
  }
  cedr_b200_throw_if( ! c.partition_error.empty(), c.partition_error);
  cedr_b200_throw_if(c.nranks > 1 && c.plan.tiers.size() < 2,
                     "with nranks > 1 the tree plan needs >= 2 tiers "
                     "(lower max_block_leaves for tiny trees)");
  c.declaring = false;
}
} // namespace

int cedr_b200_end_tracer_declarations (cedr_b200_cdr* c) {
  return guarded([&] { end_tracer_declarations(*c); });
}

int cedr_b200_get_buffers_sizes (cedr_b200_cdr* c, size_t* b1, size_t* b2) {
  return guarded([&] { get_buffers_sizes(*c, *b1, *b2); });
}

int cedr_b200_set_buffers (cedr_b200_cdr* c, double* b1, double* b2) {
  return guarded([&] {
    cedr_b200_throw_if(c->declaring, "end_tracer_declarations must be called first.");
    cedr_b200_throw_if(c->finished, "set_buffers must precede finish_setup");
    c->in = b1;
    c->out = b2;
    c->user_buffers = true;
  });
}

int cedr_b200_finish_setup (cedr_b200_cdr* c) {
  return guarded([&] { finish_setup(*c); });
}

int cedr_b200_get_problem_type (const cedr_b200_cdr* c, int tracer_idx, int* type) {
  return guarded([&] {
    cedr_b200_throw_if(tracer_idx < 0 ||
                       tracer_idx >= static_cast<int>(c->trcr_prob.size()),
                       "tracer_idx is out of bounds: " << tracer_idx);
    *type = c->trcr_prob[tracer_idx];
  });
}

int cedr_b200_get_num_tracers (const cedr_b200_cdr* c, int* n) {
  return guarded([&] { *n = static_cast<int>(c->trcr_prob.size()); });
}

int cedr_b200_run (cedr_b200_cdr* c) {
  return guarded([&] {
    cedr_b200_throw_if(! c->finished, "finish_setup must be called before run.");
    c->last_launches = 0;
    c->ntimed = 0;
    run_graphed(*c);
  });
}

int cedr_b200_run_phase (cedr_b200_cdr* c, int phase) {
  return guarded([&] {
    cedr_b200_throw_if( ! c->finished, "finish_setup must be called before run.");
    cedr_b200_throw_if(phase != 0 && phase != 1, "phase must be 0 or 1");
    if (phase == 0) { c->last_launches = 0; c->ntimed = 0; }
    if (c->nranks == 1) { if (phase == 0) run_any(*c, -1); }
    else run_any(*c, phase);
  });
}

int cedr_b200_get_exchange_count (const cedr_b200_cdr* c, size_t* count) {
  return guarded([&] {
    cedr_b200_throw_if(c->declaring, "end_tracer_declarations must be called first.");
    *count = c->nranks > 1 ? exchange_count(*c) : 0;
  });
}

int cedr_b200_set_exchange_buffers (cedr_b200_cdr* c, double* send, double* recv) {
  return guarded([&] {
    c->xsend = send;
    c->xrecv = recv;
    c->graph_reset();
  });
}

int cedr_b200_get_exchange_buffers (const cedr_b200_cdr* c, double** send, double** recv) {
  return guarded([&] {
    cedr_b200_throw_if( ! c->finished, "finish_setup must be called first.");
    *send = c->xsend;
    *recv = c->xrecv;
  });
}

int cedr_b200_print (const cedr_b200_cdr* c, char* buf, size_t bufsize) {
  return guarded([&] {
    std::stringstream ss;
    ss << (c->is_caas ? "CAAS" : "QLT") << " pid " << c->rank << ": ncells " << c->ncells
       << " #levels " << c->plan.nlevels_ref << " #tiers " << c->plan.tiers.size();
    for (size_t k = 0; k < c->plan.tiers.size(); ++k)
      ss << "\n  tier " << k << ": " << c->plan.tiers[k].nleaves << " leaves, "
         << c->plan.tiers[k].blocks.size() << " blocks (max "
         << c->plan.tiers[k].max_nl << " leaves)";
    if (c->repl)
      ss << "\n  cell -> rank map cuts blocks: replicated mode (all-gather of " << c->nrows
         << " rows x " << c->nlcl_max << " cells per rank, whole tree swept on every rank)";
    ss << "\n";
    const std::string s = ss.str();
    if (buf && bufsize) {
      std::strncpy(buf, s.c_str(), bufsize - 1);
      buf[bufsize - 1] = 0;
    }
  });
}

int cedr_b200_nlclcells (const cedr_b200_cdr* c, int* n) {
  return guarded([&] { *n = c->nlcl; });
}

int cedr_b200_get_owned_glblcells (const cedr_b200_cdr* c, int64_t* gcis) {
  return guarded([&] {
    // One rank owns all leaves and lci is the DFS leaf order
    // (cedr_qlt.cpp:243-256).
    int k = 0;
    for (int i = 0; i < c->ncells; ++i)
      if (c->plan.leaf_rank[i] == c->rank) gcis[k++] = c->plan.lci2gci[i];
  });
}

int cedr_b200_gci2lci (const cedr_b200_cdr* c, int64_t gci, int* lci) {
  return guarded([&] {
    const auto it = c->gci2lci.find(gci);
    cedr_b200_throw_if(it == c->gci2lci.end(), "gci " << gci << " not in gci2lci map.");
    *lci = it->second;
  });
}

int cedr_b200_get_device_op (cedr_b200_cdr* c, cedr_b200_device_op* op) {
  return guarded([&] {
    cedr_b200_throw_if(! c->finished, "finish_setup must be called first.");
    op->in = c->in;
    op->out = c->out;
    op->ld = c->ld;
    op->trcr_row = c->d_trcr_row.p;
    op->trcr_prob = c->d_trcr_prob.p;
    op->ntracers = static_cast<int>(c->trcr_prob.size());
    op->nlclcells = c->nlcl;
    op->is_caas = c->is_caas;
    op->reserved = c->caas_need_conserve;
  });
}

int cedr_b200_set_rhom_bulk (cedr_b200_cdr* c, const double* rhom) {
  return guarded([&] {
    cedr_b200_throw_if(! c->finished, "finish_setup must be called first.");
    CUDA_CHECK(cudaMemcpyAsync(c->in, rhom, sizeof(double)*c->nlcl,
                               cudaMemcpyDeviceToDevice, c->stream));
  });
}

int cedr_b200_set_Qm_bulk (cedr_b200_cdr* c, int t0, int nt, int64_t lda,
                           const double* qm, const double* qm_min,
                           const double* qm_max, const double* qm_prev) {
  return guarded([&] {
    cedr_b200_throw_if(! c->finished, "finish_setup must be called first.");
    cedr_b200_throw_if(t0 < 0 || nt < 0 ||
                       t0 + nt > static_cast<int>(c->trcr_prob.size()),
                       "tracer range out of bounds");
    cedr_b200_throw_if(c->bound.on, "arrays are bound (cedr_b200_bind_arrays): write them "
                       "directly instead of set_Qm");
    if (nt == 0) return;
    bool need_prev = c->is_caas && c->caas_need_conserve;
    for (int t = t0; t < t0 + nt; ++t) need_prev |= (c->trcr_prob[t] & 1) != 0;
    cedr_b200_throw_if(need_prev && ! qm_prev && ! c->is_caas,
                       "Qm_prev was not provided to set_Q.");
    const long long n = static_cast<long long>(c->nlcl)*nt;
    const int grid = static_cast<int>(std::min<long long>((n + kThreads - 1)/kThreads,
                                                          148*32));
    set_qm_bulk_kernel<<<grid, kThreads, 0, c->stream>>>(
      c->in, c->ld, c->nlcl, c->d_trcr_row.p, c->d_trcr_prob.p, t0, nt, lda, qm,
      qm_min, qm_max, qm_prev, c->is_caas && c->caas_need_conserve);
    CUDA_CHECK(cudaGetLastError());
  });
}

int cedr_b200_get_Qm_bulk (cedr_b200_cdr* c, int t0, int nt, int64_t lda, double* qm) {
  return guarded([&] {
    cedr_b200_throw_if(! c->finished, "finish_setup must be called first.");
    cedr_b200_throw_if(t0 < 0 || nt < 0 ||
                       t0 + nt > static_cast<int>(c->trcr_prob.size()),
                       "tracer range out of bounds");
    cedr_b200_throw_if(c->bound.on, "arrays are bound (cedr_b200_bind_arrays): the results "
                       "are in the bound output array");
    if (nt == 0) return;
    const long long n = static_cast<long long>(c->nlcl)*nt;
    const int grid = static_cast<int>(std::min<long long>((n + kThreads - 1)/kThreads,
                                                          148*32));
    get_qm_bulk_kernel<<<grid, kThreads, 0, c->stream>>>(
      c->is_caas ? c->in : c->out, c->ld, c->nlcl, c->d_trcr_row.p, c->is_caas, t0,
      nt, lda, qm);
    CUDA_CHECK(cudaGetLastError());
  });
}

int cedr_b200_bind_arrays (cedr_b200_cdr* c, int64_t lda, const double* qm_min, double* qm,
                            const double* qm_max, const double* qm_prev, double* qm_out) {
  return guarded([&] {
    cedr_b200_throw_if( ! c->finished, "finish_setup must be called first.");
    if ( ! qm) {          // unbind
      c->bound = cedr_b200_cdr::Bound();
      upload_rowaddr(*c);
      c->graph_reset();   // the output array is a launch argument of the captured kernels
      return;
    }
    cedr_b200_throw_if(c->is_bfb, "bind_arrays is for QLT and CAAS");
    cedr_b200_throw_if( ! qm_min || ! qm_max, "Qm_min and Qm_max are required");
    cedr_b200_throw_if(lda < c->nlcl, "lda must be >= nlclcells");
    bool need_prev = c->is_caas && c->caas_need_conserve;
    for (size_t t = 0; t < c->trcr_prob.size(); ++t) {
      const int cls = c->trcr_cls[t];
      // Consistent-only tracers store q bounds (Qm bounds / rhom, cedr_qlt_inl.hpp:36-45)
      // and nonnegative ones have no bound rows: those need set_Qm's transformation.
      cedr_b200_throw_if( ! (cls == CLS_ST || cls == CLS_CST || cls == CLS_CAAS),
                         "bind_arrays needs shape-preserving tracers (tracer " << t << ")");
      need_prev |= (c->trcr_prob[t] & 1) != 0;
    }
    cedr_b200_throw_if(need_prev && ! qm_prev, "Qm_prev was not provided to set_Q.");
    cedr_b200_throw_if( ! c->is_caas && ! qm_out, "QLT needs an output array");
    cedr_b200_throw_if( ! c->is_caas && (qm_out == qm || qm_out == qm_min || qm_out == qm_max),
                       "QLT's output must not alias its inputs (the down-sweep re-reads them)");
    if (c->fast_ok) {
      // The fast kernels move rows with 16-byte TMA bulk copies.
      auto al = [] (const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
      cedr_b200_throw_if(lda % 2 != 0 || ! al(qm_min) || ! al(qm) || ! al(qm_max) ||
                         (qm_prev && ! al(qm_prev)) || (qm_out && ! al(qm_out)),
                         "bound arrays must be 16-byte aligned with an even lda");
    }
    if (c->d_ident.n < c->trcr_prob.size()) {
      std::vector<int> id(c->trcr_prob.size());
      for (size_t i = 0; i < id.size(); ++i) id[i] = static_cast<int>(i);
      c->d_ident.upload(id);
    }
    c->bound.on = true;
    c->graph_reset();
    c->bound.lda = lda;
    c->bound.qm_min = qm_min;
    c->bound.qm = qm;
    c->bound.qm_max = qm_max;
    c->bound.qm_prev = qm_prev;
    c->bound.qm_out = qm_out;
    upload_rowaddr(*c);
  });
}

int cedr_b200_transport1d_cycle (cedr_b200_cdr* c, int nsteps, const double* y0_host,
                                 double* yf_host, int use_graph, float* ms_per_step) {
  return guarded([&] {
    cedr_b200_throw_if( ! c->finished, "finish_setup must be called first.");
    cedr_b200_throw_if(c->is_bfb || c->nranks != 1 || c->trcr_prob.empty() || nsteps < 1,
                       "transport1d: one rank, at least one tracer and one step");
    cedr_b200_throw_if(c->bound.on, "transport1d: unbind the arrays first");
    const int n = c->ncells;
    cedr_b200_throw_if(n < 4, "transport1d: at least 4 cells");
    const int cls = c->trcr_cls[0];
    cedr_b200_throw_if(cls == CLS_T || cls == CLS_CT,
                       "transport1d: consistent-only tracers are not driven by this harness");
    // Problem1D::init_mesh, uniform (cedr_test_1d_transport.cpp:143-169), and the static
    // part of cycle / cubic_interp_periodic: departure points, their wrapped images and
    // intervals (:14-18, :46-58, :238-241).
    std::vector<double> xb(n + 1), xcp(n + 1), area(n), xp(n + 1);
    std::vector<int> ti(n + 1), lci(n);
    xb[0] = 0;
    xb[n] = 1;
    for (int i = 1; i < n; ++i) xb[i] = static_cast<double>(i)/n;
    for (int i = 0; i < n; ++i) { xcp[i] = 0.5*(xb[i] + xb[i + 1]); area[i] = xb[i + 1] - xb[i]; }
    xcp[n] = 1 + xcp[0];
    const double xos = -1.0/nsteps;
    for (int j = 0; j <= n; ++j) {
      const double xi = xcp[j] + xos, xl = xcp[0], xr = xcp[n];
      double x = xi;
      if ( ! (x >= xl && x <= xr)) { const double w = xr - xl; x = xi - w*std::floor((xi - xl)/w); }
      xp[j] = x;
      int ip1 = static_cast<int>(std::upper_bound(xcp.begin(), xcp.end(), x) - xcp.begin());
      if (ip1 == 0) ++ip1; else if (ip1 == n + 1) --ip1;
      ti[j] = ip1 - 1;
    }
    for (int i = 0; i < n; ++i) {
      const auto it = c->gci2lci.find(i);
      cedr_b200_throw_if(c->is_caas ? false : it == c->gci2lci.end(),
                         "transport1d: the CDR's cells must be 0..ncells-1");
      lci[i] = c->is_caas ? i : it->second;
    }
    DevBuf<double> d_xcp, d_area, d_xp, d_y[2];
    DevBuf<int> d_ti, d_lci;
    d_xcp.upload(xcp); d_area.upload(area); d_xp.upload(xp); d_ti.upload(ti); d_lci.upload(lci);
    d_y[0].alloc(n + 1); d_y[1].alloc(n + 1);
    cudaStream_t own = nullptr, saved = c->stream;
    CUDA_CHECK(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamSynchronize(saved));
    c->stream = own;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&] () {
      c->stream = saved;
      if (exec) cudaGraphExecDestroy(exec);
      if (graph) cudaGraphDestroy(graph);
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
      cudaStreamDestroy(own);
    };
    try {
      CUDA_CHECK(cudaMemcpyAsync(d_y[0].p, y0_host, sizeof(double)*(n + 1),
                                 cudaMemcpyHostToDevice, own));
      // rhom = the cell areas (cedr_test_1d_transport.cpp:283-289).
      std::vector<double> rh(c->nlcl);
      for (int i = 0; i < n; ++i) rh[lci[i]] = area[i];
      CUDA_CHECK(cudaMemcpyAsync(c->in, rh.data(), sizeof(double)*c->nlcl,
                                 cudaMemcpyHostToDevice, own));
      CUDA_CHECK(cudaStreamSynchronize(own));
      T1dArgs a;
      std::memset(&a, 0, sizeof(a));
      a.ncells = n;
      a.xcp = d_xcp.p; a.area = d_area.p; a.tgt_i = d_ti.p; a.tgt_xp = d_xp.p; a.lci = d_lci.p;
      a.in = c->in;
      a.ld = c->ld;
      a.row = c->trcr_row[0];
      a.layout = (cls == CLS_NN || cls == CLS_CNN) ? 1 : 0;
      cedr_b200_throw_if(a.layout == 0 && cls != CLS_CAAS && cls != CLS_CST,
                         "transport1d: tracer 0 must conserve (Qm_prev is set every step)");
      a.out = c->is_caas ? c->in + (static_cast<long long>(a.row) + 1)*c->ld : c->out;
      const int grid = (n + 1 + 127)/128;
      auto step = [&] (const int from) {
        t1d_interp_set_kernel<<<grid, 128, 0, own>>>(a, d_y[from].p, d_y[1 - from].p);
        CUDA_CHECK(cudaGetLastError());
        run_any(*c, -1);
        t1d_get_kernel<<<grid, 128, 0, own>>>(a, d_y[1 - from].p);
        CUDA_CHECK(cudaGetLastError());
      };
      c->last_launches = 0;
      c->ntimed = 0;
      const bool prof = c->profiling;
      c->profiling = false;
      CUDA_CHECK(cudaEventCreate(&e0));
      CUDA_CHECK(cudaEventCreate(&e1));
      int done = 0;
      if (use_graph == 2) {
        // The whole cycle in one launch (t1d_cycle_kernel).
        cedr_b200_throw_if( ! solo_ok(*c) || c->trcr_prob.size() != 1,
                           "transport1d: the one-launch cycle takes a single block of at "
                           "most 256 cells with one tracer");
        const SweepArgs sa = base_args(*c, cls, 0);
        const size_t nn = 2*static_cast<size_t>(c->plan.tiers[0].max_nl);
        const size_t smem = sizeof(double)*(2*(static_cast<size_t>(n) + 2) + 5*nn + 2) +
          sizeof(dev::NodeConst)*(nn/2) + sizeof(int)*(3*(nn/2) + 2);
        CUDA_CHECK(cudaEventRecord(e0, own));
        double* const yo = d_y[nsteps & 1].p;
        switch (cls) {
        case CLS_CST: t1d_cycle_kernel<CLS_CST><<<1, kThreads, smem, own>>>(
            sa, c->solo_block, a, d_y[0].p, yo, nsteps); break;
        case CLS_NN: t1d_cycle_kernel<CLS_NN><<<1, kThreads, smem, own>>>(
            sa, c->solo_block, a, d_y[0].p, yo, nsteps); break;
        case CLS_CNN: t1d_cycle_kernel<CLS_CNN><<<1, kThreads, smem, own>>>(
            sa, c->solo_block, a, d_y[0].p, yo, nsteps); break;
        default: t1d_cycle_kernel<CLS_CAAS><<<1, kThreads, smem, own>>>(
            sa, c->solo_block, a, d_y[0].p, yo, nsteps); break;
        }
        CUDA_CHECK(cudaGetLastError());
        c->last_launches = 1;
        done = nsteps;
      } else if (use_graph && nsteps >= 2) {
        // Two steps (y0 -> y1 -> y0) per graph launch: launch-bound work, replayed.
        CUDA_CHECK(cudaStreamBeginCapture(own, cudaStreamCaptureModeThreadLocal));
        step(0);
        step(1);
        CUDA_CHECK(cudaStreamEndCapture(own, &graph));
        CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
        CUDA_CHECK(cudaEventRecord(e0, own));
        for (; done + 2 <= nsteps; done += 2) CUDA_CHECK(cudaGraphLaunch(exec, own));
      } else {
        CUDA_CHECK(cudaEventRecord(e0, own));
      }
      for (; done < nsteps; ++done) step(done & 1);
      CUDA_CHECK(cudaEventRecord(e1, own));
      CUDA_CHECK(cudaStreamSynchronize(own));
      c->profiling = prof;
      float ms = 0;
      CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms_per_step) *ms_per_step = ms/nsteps;
      CUDA_CHECK(cudaMemcpy(yf_host, d_y[nsteps & 1].p, sizeof(double)*(n + 1),
                            cudaMemcpyDeviceToHost));
    } catch (...) {
      cleanup();
      throw;
    }
    cleanup();
  });
}

int cedr_b200_set_stream (cedr_b200_cdr* c, void* s) {
  return guarded([&] { c->stream = static_cast<cudaStream_t>(s); });
}

int cedr_b200_synchronize (cedr_b200_cdr* c) {
  return guarded([&] {
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (c->d_status.p) {
      int st = 0;
      CUDA_CHECK(cudaMemcpy(&st, c->d_status.p, sizeof(int), cudaMemcpyDeviceToHost));
      if (st) {
        // Report once; the next run() starts clean.
        CUDA_CHECK(cudaMemset(c->d_status.p, 0, sizeof(int)));
        if (st == 2)
          throw std::runtime_error("cedr_b200: the peer-to-peer exchange gave up waiting for "
                                   "another rank's block roots (results are invalid)");
        throw std::runtime_error("cedr_b200: the persistent run() kernel gave up waiting for "
                                 "a tracer's root (results are invalid)");
      }
    }
  });
}

int cedr_b200_debug_phase_clocks (cedr_b200_cdr* c, unsigned long long* out16) {
  return guarded([&] {
    for (int i = 0; i < 16; ++i) out16[i] = 0;
    if (c->d_phase_clk.p) {
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      CUDA_CHECK(cudaMemcpy(out16, c->d_phase_clk.p, 16*sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost));
      CUDA_CHECK(cudaMemset(c->d_phase_clk.p, 0, 16*sizeof(unsigned long long)));
    }
  });
}

int cedr_b200_set_ring (cedr_b200_cdr* c, int on) {
  return guarded([&] {
    cedr_b200_throw_if(c->finished, "set_ring must precede finish_setup");
    c->ring_enabled = on != 0;
  });
}

int cedr_b200_uses_ring (const cedr_b200_cdr* c, int* on) {
  return guarded([&] { *on = c->ring_ok; });
}

int cedr_b200_set_cluster_caas (cedr_b200_cdr* c, int mode) {
  return guarded([&] {
    cedr_b200_throw_if(c->finished, "set_cluster_caas must precede finish_setup");
    c->cluster_mode = mode;
  });
}

int cedr_b200_uses_cluster_caas (const cedr_b200_cdr* c, int* on) {
  return guarded([&] { *on = c->cluster_ok && ! c->ring_ok && ! solo_ok(*c); });
}

int cedr_b200_ring_info (const cedr_b200_cdr* c, int* info8) {
  return guarded([&] {
    const cedr_b200_cdr::Ring& R = c->ring;
    const int v[8] = {R.grid, R.S, R.npn, R.TB, R.nuslots*100 + R.ndslots, R.np, R.sw,
                      static_cast<int>(R.smem)};
    for (int i = 0; i < 8; ++i) info8[i] = c->ring_ok ? v[i] : 0;
  });
}

int cedr_b200_ring_trace (cedr_b200_cdr* c, unsigned long long* host, size_t cap, size_t* n) {
  return guarded([&] {
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *n = c->ring.trace_n;
    if (host && c->ring.trace.p)
      CUDA_CHECK(cudaMemcpy(host, c->ring.trace.p, std::min(cap, c->ring.trace_n)*sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost));
  });
}

int cedr_b200_set_allgather (cedr_b200_cdr* c, cedr_b200_allgather_fn fn, void* ctx) {
  return guarded([&] { c->allgather = fn; c->allgather_ctx = ctx; });
}

int cedr_b200_p2p_get_handle (cedr_b200_cdr* c, void* handle64) {
  return guarded([&] {
    cedr_b200_throw_if( ! c->finished, "finish_setup must be called first.");
    cedr_b200_throw_if(c->nranks < 2, "peer-to-peer exchange needs nranks > 1");
    cedr_b200_throw_if(c->repl, "peer-to-peer exchange: this cell -> rank map cuts blocks of "
                       "the tree plan (replicated mode uses the all-gather)");
    cedr_b200_throw_if(c->nranks > 16, "peer-to-peer exchange supports up to 16 ranks");
    if ( ! c->p2p_arena.p) {
      c->p2p_arena.alloc(p2p_arena_doubles(*c));
      CUDA_CHECK(cudaMemset(c->p2p_arena.p, 0, p2p_arena_doubles(*c)*sizeof(double)));
      c->p2p_peer.assign(c->nranks, nullptr);
      c->p2p_peer[c->rank] = c->p2p_arena.p;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CUDA_CHECK(cudaIpcGetMemHandle(&h, c->p2p_arena.p));
    std::memcpy(handle64, &h, 64);
  });
}

int cedr_b200_p2p_set_peer (cedr_b200_cdr* c, int peer_rank, const void* handle64) {
  return guarded([&] {
    cedr_b200_throw_if( ! c->p2p_arena.p, "call p2p_get_handle first");
    cedr_b200_throw_if(peer_rank < 0 || peer_rank >= c->nranks, "peer rank out of range");
    if (peer_rank == c->rank) return;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    void* p = nullptr;
    CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->p2p_peer[peer_rank] = p;
  });
}

int cedr_b200_p2p_enable (cedr_b200_cdr* c, int on) {
  return guarded([&] {
    if (on)
      for (int r = 0; r < c->nranks; ++r)
        cedr_b200_throw_if(r >= static_cast<int>(c->p2p_peer.size()) || ! c->p2p_peer[r],
                           "peer " << r << " was not set (cedr_b200_p2p_set_peer)");
    c->p2p_on = on != 0;
    c->graph_reset();
  });
}

int cedr_b200_set_graph (cedr_b200_cdr* c, int mode) {
  return guarded([&] {
    cedr_b200_throw_if(mode < -1 || mode > 1, "graph mode must be -1 (auto), 0 or 1");
    c->graph_mode = mode;
    c->graph_reset();
  });
}

int cedr_b200_uses_graph (const cedr_b200_cdr* c, int* on) {
  return guarded([&] { *on = c->graph[0].exec || c->graph[1].exec; });
}

int cedr_b200_last_run_launches (const cedr_b200_cdr* c, int* n) {
  return guarded([&] { *n = c->last_launches; });
}

int cedr_b200_set_fast_path (cedr_b200_cdr* c, int on) {
  return guarded([&] {
    cedr_b200_throw_if(c->finished, "set_fast_path must precede finish_setup");
    c->fast_enabled = on != 0;
  });
}

int cedr_b200_uses_fast_path (const cedr_b200_cdr* c, int* on) {
  return guarded([&] { *on = c->fast_ok; });
}

int cedr_b200_set_profiling (cedr_b200_cdr* c, int on) {
  return guarded([&] { c->profiling = on != 0; c->ntimed = 0; });
}

int cedr_b200_get_launch_times (cedr_b200_cdr* c, int cap, float* ms, int* tags,
                                int* tiers, int* n) {
  return guarded([&] {
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    const int m = static_cast<int>(std::min<size_t>(c->ntimed, cap));
    for (int i = 0; i < m; ++i) {
      CUDA_CHECK(cudaEventElapsedTime(&ms[i], c->timed[i].e0, c->timed[i].e1));
      tags[i] = c->timed[i].tag;
      tiers[i] = c->timed[i].tier;
    }
    *n = m;
  });
}

int cedr_b200_plan_info (const cedr_b200_cdr* c, int* ntiers, int* nblocks0,
                         int* max_block_leaves, int* nlevels_ref) {
  return guarded([&] {
    if (ntiers) *ntiers = static_cast<int>(c->plan.tiers.size());
    if (nblocks0) *nblocks0 = static_cast<int>(c->plan.tiers[0].blocks.size());
    if (max_block_leaves) *max_block_leaves = c->plan.tiers[0].max_nl;
    if (nlevels_ref) *nlevels_ref = c->plan.nlevels_ref;
  });
}

int cedr_b200_make_1d_tree (int ncells, int imbalanced, int* kids_host,
                            int64_t* cellidx_host) {
  return guarded([&] {
    std::vector<int> kids;
    std::vector<int64_t> cellidx;
    make_bisection_tree(ncells, imbalanced != 0, kids, cellidx);
    std::copy(kids.begin(), kids.end(), kids_host);
    std::copy(cellidx.begin(), cellidx.end(), cellidx_host);
  });
}

int cedr_b200_merge_partial_trees (int nparts, const int* part_nnodes, const int* part_root,
                                   const int* kids, const int64_t* cellidx,
                                   const int* node_rank, int cap_nodes, int* out_nnodes,
                                   int* out_kids, int64_t* out_cellidx, int* out_rank) {
  return guarded([&] {
    cedr_b200_throw_if( ! part_nnodes || ! part_root || ! kids || ! cellidx || ! out_nnodes,
                       "merge_partial_trees: null argument");
    std::vector<int> k, r;
    std::vector<int64_t> ci;
    merge_partial_trees(nparts, part_nnodes, part_root, kids, cellidx, node_rank, k, ci, r);
    *out_nnodes = static_cast<int>(ci.size());
    // Size query: out arrays may be null; otherwise they must hold the merged tree.
    if ( ! out_kids && ! out_cellidx && ! out_rank) return;
    cedr_b200_throw_if(static_cast<int>(ci.size()) > cap_nodes,
                       "merge_partial_trees: output arrays are too small");
    if (out_kids) std::copy(k.begin(), k.end(), out_kids);
    if (out_cellidx) std::copy(ci.begin(), ci.end(), out_cellidx);
    if (out_rank) std::copy(r.begin(), r.end(), out_rank);
  });
}

int cedr_b200_allgather_host (cedr_b200_allgather_fn fn, void* ctx, int nranks,
                              const double* send_host, double* recv_host, size_t count) {
  return guarded([&] {
    cedr_b200_throw_if(nranks < 1 || ! send_host || ! recv_host, "allgather_host: bad argument");
    if (nranks == 1) {
      std::copy(send_host, send_host + count, recv_host);
      return;
    }
    cedr_b200_throw_if( ! fn, "allgather_host: nranks > 1 needs the all-gather hook");
    // The hook gathers device memory on a stream (it is the run()-time exchange's hook):
    // stage through two scratch buffers on the legacy default stream. Setup-time only.
    double* d = nullptr;
    CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d), (1 + static_cast<size_t>(nranks))*count*
                          sizeof(double)));
    int e = 0;
    cudaError_t ce = cudaMemcpy(d, send_host, count*sizeof(double), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) {
      e = fn(ctx, d, d + count, count, nullptr);
      if (e == 0) ce = cudaDeviceSynchronize();
      if (e == 0 && ce == cudaSuccess)
        ce = cudaMemcpy(recv_host, d + count, static_cast<size_t>(nranks)*count*sizeof(double),
                        cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    cedr_b200_throw_if(e != 0, "allgather_host: the all-gather hook failed");
    CUDA_CHECK(ce);
  });
}

int cedr_b200_plan_probe (int ncells, int nnodes, int root, const int* kids,
                          const int64_t* cellidx, int max_block_leaves,
                          int64_t* lci2gci_host, int* ntiers,
                          int* nblocks_per_tier_host, int* nshapes, int* nlevels_ref,
                          int64_t* idsum) {
  return guarded([&] {
    Plan plan;
    plan.build(ncells, nnodes, root, kids, cellidx, nullptr, max_block_leaves);
    if (lci2gci_host) std::copy(plan.lci2gci.begin(), plan.lci2gci.end(), lci2gci_host);
    if (ntiers) *ntiers = static_cast<int>(plan.tiers.size());
    if (nshapes) *nshapes = static_cast<int>(plan.shapes.size());
    if (nlevels_ref) *nlevels_ref = plan.nlevels_ref;
    cedr_b200_throw_if(plan.tiers.size() > 64, "more than 64 tiers");
    // Sum the cell ids through the plan, exactly as the sweep kernels walk it.
    std::vector<int64_t> leaf(plan.lci2gci);
    std::vector<int> internal_used(std::max(1, plan.ninternal), 0);
    for (size_t k = 0; k < plan.tiers.size(); ++k) {
      const Tier& tier = plan.tiers[k];
      if (nblocks_per_tier_host)
        nblocks_per_tier_host[k] = static_cast<int>(tier.blocks.size());
      cedr_b200_throw_if(static_cast<int>(leaf.size()) != tier.nleaves,
                         "tier leaf count mismatch");
      std::vector<int> leaf_used(tier.nleaves, 0);
      std::vector<int64_t> next(tier.blocks.size());
      for (size_t b = 0; b < tier.blocks.size(); ++b) {
        const Block& blk = tier.blocks[b];
        const Shape& sh = plan.shapes[blk.shape];
        cedr_b200_throw_if(sh.nl != blk.nl || sh.ni != blk.nl - 1, "shape mismatch");
        std::vector<int64_t> v(sh.nl + sh.ni);
        std::vector<int> used(sh.nl + sh.ni, 0);
        for (int i = 0; i < sh.nl; ++i) {
          v[i] = leaf[blk.leaf0 + i];
          ++leaf_used[blk.leaf0 + i];
        }
        for (int l = 0; l < sh.nlev; ++l)
          for (int j = sh.lvlptr[l]; j < sh.lvlptr[l+1]; ++j) {
            cedr_b200_throw_if(sh.kid0[j] >= sh.nl + j || sh.kid1[j] >= sh.nl + j,
                               "kid not computed before its parent");
            v[sh.nl + j] = v[sh.kid0[j]] + v[sh.kid1[j]];
            ++used[sh.kid0[j]];
            ++used[sh.kid1[j]];
            ++internal_used[blk.ibase + j];
          }
        for (int i = 0; i + 1 < sh.nl + sh.ni; ++i)
          cedr_b200_throw_if(used[i] != 1, "block node not used exactly once");
        next[b] = v[sh.nl + sh.ni - 1];
      }
      for (int i = 0; i < tier.nleaves; ++i)
        cedr_b200_throw_if(leaf_used[i] != 1, "tier leaf not used exactly once");
      leaf.swap(next);
    }
    cedr_b200_throw_if(leaf.size() != 1, "top tier must have one block");
    // Work-assignment tables of the fast shapes (tree_plan.cpp build_down_tables): every
    // depth-7 node has one leaf thread, every pair one entry, in the warp of its owner.
    for (const Shape& sh : plan.shapes) {
      if ( ! sh.fast) continue;
      std::vector<int> thread_of(128, -1);
      cedr_b200_throw_if(sh.perm.size() != 128, "perm size");
      for (int i = 0; i < 128; ++i) {
        cedr_b200_throw_if(sh.perm[i] >= 128 || thread_of[sh.perm[i]] >= 0,
                           "perm is not a permutation");
        thread_of[sh.perm[i]] = i;
      }
      cedr_b200_throw_if(sh.perm_up.size() != 256, "perm_up size");
      for (int i = 0; i < 128; ++i) {
        const int nd = sh.perm_up[i];
        cedr_b200_throw_if(nd >= 128 || (nd >> 5) != (i >> 5) || sh.perm_up[128 + nd] != i,
                           "perm_up is not a permutation within each warp");
      }
      const int np = static_cast<int>(sh.ptab.size());
      cedr_b200_throw_if(sh.pent.size() != static_cast<size_t>(5 + np) || sh.pent[0] != 5 ||
                         sh.pent[4] != static_cast<unsigned>(5 + np), "pair entry table size");
      std::vector<int> seen(std::max(1, np), 0);
      for (int w = 0; w < 4; ++w) {
        cedr_b200_throw_if(sh.pent[w] > sh.pent[w + 1], "pair entry ranges");
        for (unsigned j = sh.pent[w]; j < sh.pent[w + 1]; ++j) {
          const unsigned e = sh.pent[j], o = e & 0x7ff, slot = (e >> 11) & 0x1ff, r = e >> 20;
          cedr_b200_throw_if(static_cast<int>(r) >= np, "pair rank out of range");
          const int p = sh.ptab[r], owner = slot & 127;
          ++seen[r];
          cedr_b200_throw_if((sh.dtab[p] >> 15) == 0 || o != (sh.dtab[p] & 0x7fffu),
                             "pair entry does not match its depth-9 node");
          cedr_b200_throw_if(static_cast<int>(slot >> 7) != (p & 3) ||
                             sh.perm[owner] != (p >> 2) || (owner >> 5) != w,
                             "pair entry is not in the warp that owns its node");
        }
      }
      for (int r = 0; r < np; ++r)
        cedr_b200_throw_if(seen[r] != 1, "pair not assigned exactly once");
    }
    for (int i = 0; i < plan.ninternal; ++i)
      cedr_b200_throw_if(internal_used[i] != 1, "internal node not used exactly once");
    if (idsum) *idsum = leaf[0];
  });
}

int cedr_b200_partition_probe (int ncells, int imbalanced, int max_block_leaves, int rank,
                               int nranks, int cap, int* nlclcells, int* nown, int* nown_max,
                               int* nblocks_global, int* gidx_host, int* leaf0_host,
                               int* nl_host) {
  return guarded([&] {
    cedr_b200_cdr c;
    c.rank = rank;
    c.nranks = nranks;
    c.ncells = ncells;
    c.max_block_leaves = max_block_leaves;
    cedr_b200_throw_if(nranks > ncells, "#GIDs < #ranks is not supported.");
    fill_1d_tree(c, ncells, imbalanced != 0);
    build_plan(c);
    cedr_b200_throw_if( ! c.partition_error.empty(), c.partition_error);
    if (nlclcells) *nlclcells = c.nlcl;
    if (nown) *nown = static_cast<int>(c.own_blocks.size());
    if (nown_max) *nown_max = c.nown_max;
    if (nblocks_global) *nblocks_global = static_cast<int>(c.plan.tiers[0].blocks.size());
    for (size_t j = 0; j < c.own_blocks.size() && static_cast<int>(j) < cap; ++j) {
      const Block& blk = c.plan.tiers[0].blocks[c.own_blocks[j]];
      if (gidx_host) gidx_host[j] = c.own_blocks[j];
      if (leaf0_host) leaf0_host[j] = blk.leaf0;
      if (nl_host) nl_host[j] = blk.nl;
    }
  });
}

int cedr_b200_fill_headline (int ncells, int nt, int64_t lda, int config_id,
                             double* rhom, double* qm_min, double* qm, double* qm_max,
                             double* qm_prev, void* stream) {
  return cedr_b200_fill_headline_range(ncells, 0, ncells, nt, lda, config_id, rhom, qm_min,
                                       qm, qm_max, qm_prev, stream);
}

int cedr_b200_fill_headline_range (int ncells, int cell0, int nlcl, int nt, int64_t lda,
                                   int config_id, double* rhom, double* qm_min, double* qm,
                                   double* qm_max, double* qm_prev, void* stream) {
  return guarded([&] {
    require_device();
    cedr_b200_throw_if(cell0 < 0 || nlcl < 0 || cell0 + nlcl > ncells, "bad cell range");
    const long long n = static_cast<long long>(nlcl)*nt;
    const int grid = static_cast<int>(std::min<long long>((n + kThreads - 1)/kThreads,
                                                          148*32));
    fill_headline_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      ncells, nt, lda, 0xCED20000ull + static_cast<unsigned long long>(config_id),
      rhom, qm_min, qm, qm_max, qm_prev, cell0, nlcl);
    CUDA_CHECK(cudaGetLastError());
  });
}

} // extern "C"
