// cedr_b200.hpp -- C++ mirror of COMPOSE's cedr::CDR interface over the C ABI of
// cedr_b200.h. Header-only; host code is plain C++11, device code needs nvcc.
//
// Same names, argument meaning, call order and error behaviour as the reference, so a
// HOMME-style caller keeps its code and swaps the include:
//
//   reference (cedr/)                               here
//   cedr_cdr.hpp:16-112   struct CDR, Options,      cedr::CDR, CDR::Options,
//                         DeviceOp                  CDR::DeviceOp
//   cedr.hpp:29-39        ProblemType               cedr::ProblemType
//   cedr_qlt.hpp:26-219   qlt::QLT<ES>              cedr::qlt::QLT<ES>
//   cedr_caas.hpp:15-118  caas::CAAS<ES>            cedr::caas::CAAS<ES>
//   cedr_tree_caller.hpp  tree::Node,               cedr::tree::Node,
//                         make_tree_over_1d_mesh    tree::make_tree_over_1d_mesh
//   cedr_mpi.hpp:17-39    mpi::Parallel,            cedr::mpi::Parallel,
//                         make_parallel             mpi::make_parallel
//   cedr_util.hpp:70-77   cedr_throw_if ->          std::logic_error with the same
//                         std::logic_error          message shape
//
// Differences a caller sees:
//   - `ES` is an opaque tag (there is no Kokkos); every CDR runs on the current CUDA
//     device. mpi::Parallel carries (rank, size) and an all-gather hook instead of an
//     MPI_Comm (the path's one exchange step, see cedr_b200.h).
//   - CDR::DeviceOp is a concrete trivially-copyable struct with __host__ __device__
//     methods (the reference's is an abstract base whose concrete subclasses are
//     copied into Kokkos lambdas, cedr_test_randomized_inl.hpp:27-58): copy it by value
//     into kernels. Host-side calls (cedr_test_1d_transport.cpp:173-187) work when the
//     CDR was built with Memory::managed, which puts its buffers in CUDA managed
//     memory; synchronise (run() does) before touching them from the host.
//   - run() is synchronous like the reference's; run_async() only enqueues.
#ifndef CEDR_B200_HPP
#define CEDR_B200_HPP

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "cedr_b200.h"

#if defined(__CUDACC__)
# define CEDR_B200_HD __host__ __device__ __forceinline__
#else
# define CEDR_B200_HD inline
#endif

// UserAllReducer::operator() takes an MPI_Op (cedr_caas.hpp:39); without an MPI in the build
// a stand-in is declared (CAAS only ever passes MPI_SUM, cedr_caas.cpp:262-266).
#if ! defined(MPI_VERSION) && ! defined(CEDR_B200_HAVE_MPI_OP)
# define CEDR_B200_HAVE_MPI_OP
typedef int MPI_Op;
# ifndef MPI_SUM
static const MPI_Op MPI_SUM = 0;
# endif
#endif

namespace cedr {
typedef int Int;
typedef long Long;
typedef std::size_t Size;
typedef double Real;

// cedr.hpp:29-39
struct ProblemType {
  enum : Int { conserve = 1, shapepreserve = 2, consistent = 4, nonnegative = 8 };
};

namespace impl {
inline void check (const int e) {
  if (e == 0) return;
  // Code 1: a cedr_throw_if-style logic error; otherwise a runtime (CUDA) failure.
  if (e == 1) throw std::logic_error(cedr_b200_last_error());
  throw std::runtime_error(cedr_b200_last_error());
}
} // namespace impl

namespace mpi {
// cedr_mpi.hpp:17-27. One process per GPU; `gather` is the exchange hook
// (cedr_b200_allgather_fn), e.g. a thin wrapper over ncclAllGather.
class Parallel {
  Int rank_, size_;
  cedr_b200_allgather_fn gather_;
  void* ctx_;
public:
  typedef std::shared_ptr<Parallel> Ptr;
  Parallel (Int rank = 0, Int size = 1, cedr_b200_allgather_fn gather = nullptr,
            void* ctx = nullptr) : rank_(rank), size_(size), gather_(gather), ctx_(ctx) {}
  Int size () const { return size_; }
  Int rank () const { return rank_; }
  Int root () const { return 0; }
  bool amroot () const { return rank() == root(); }
  cedr_b200_allgather_fn allgather () const { return gather_; }
  void* allgather_ctx () const { return ctx_; }
};
inline Parallel::Ptr make_parallel (Int rank = 0, Int size = 1,
                                    cedr_b200_allgather_fn gather = nullptr,
                                    void* ctx = nullptr) {
  return std::make_shared<Parallel>(rank, size, gather, ctx);
}
} // namespace mpi

namespace tree {
// cedr_tree_caller.hpp:12-24
struct Node {
  typedef std::shared_ptr<Node> Ptr;
  const Node* parent;
  Int rank;
  Long cellidx;
  Int nkids;
  Node::Ptr kids[2];
  Int reserved;
  Int level;
  Node () : parent(nullptr), rank(-1), cellidx(-1), nkids(0), reserved(-1), level(-1) {}
};

// cedr_tree_caller.hpp:26-29: recursive bisection (cn/2, or cn/3 when imbalanced and
// cn > 2), leaves ranked by the contiguous map of cedr_tree.cpp:366-369.
inline Node::Ptr make_tree_over_1d_mesh (const mpi::Parallel::Ptr& p, const Int& ncells,
                                         const bool imbalanced = false) {
  std::vector<int> kids(2*(2*static_cast<size_t>(ncells) - 1));
  std::vector<int64_t> cellidx(2*static_cast<size_t>(ncells) - 1);
  impl::check(cedr_b200_make_1d_tree(ncells, imbalanced, kids.data(), cellidx.data()));
  std::vector<Node::Ptr> nodes(cellidx.size());
  for (size_t i = 0; i < nodes.size(); ++i) nodes[i] = std::make_shared<Node>();
  const Int nr = p ? p->size() : 1;
  for (size_t i = 0; i < nodes.size(); ++i) {
    Node& n = *nodes[i];
    if (kids[2*i] < 0) {
      n.cellidx = static_cast<Long>(cellidx[i]);
      const Int r = static_cast<Int>(cellidx[i]/(ncells/nr));
      n.rank = r < nr ? r : nr - 1;
    } else {
      n.nkids = 2;
      for (int k = 0; k < 2; ++k) {
        n.kids[k] = nodes[kids[2*i + k]];
        n.kids[k]->parent = &n;
      }
    }
  }
  return nodes[0];
}
} // namespace tree

// Where a CDR's buffers live when the caller does not set_buffers.
enum class Memory { device, managed };

// cedr_cdr.hpp:16-112
struct CDR {
  typedef std::shared_ptr<CDR> Ptr;

  struct Options {
    bool prefer_numerical_mass_conservation_to_numerical_bounds;
    Options () : prefer_numerical_mass_conservation_to_numerical_bounds(false) {}
  };

  // cedr_cdr.hpp:67-102; layout documented at cedr_b200_device_op (cedr_b200.h).
  struct DeviceOp {
    cedr_b200_device_op v;

    CEDR_B200_HD void set_rhom (const Int& lclcellidx, const Int& /*rhomidx*/,
                                const Real& rhom) const {
      v.in[lclcellidx] = rhom;
    }

    CEDR_B200_HD void set_Qm (const Int& lclcellidx, const Int& tracer_idx, const Real& Qm,
                              const Real& Qm_min, const Real& Qm_max,
                              const Real Qm_prev = inf()) const {
      const int pt = v.trcr_prob[tracer_idx];
      Real* bd = v.in + static_cast<int64_t>(v.trcr_row[tracer_idx])*v.ld + lclcellidx;
      int next;
      if (pt & ProblemType::shapepreserve) {
        bd[0] = Qm_min; bd[v.ld] = Qm; bd[2*v.ld] = Qm_max; next = 3;
      } else if (pt & ProblemType::consistent) {
        const Real rhom = v.in[lclcellidx];   // set_rhom precedes set_Qm (cedr_cdr.hpp:80)
        bd[0] = Qm_min/rhom; bd[v.ld] = Qm; bd[2*v.ld] = Qm_max/rhom; next = 3;
      } else {
        bd[0] = Qm; next = 1;
      }
      if ((pt & ProblemType::conserve) || (v.is_caas && v.reserved)) bd[next*v.ld] = Qm_prev;
    }

    CEDR_B200_HD Real get_Qm (const Int& lclcellidx, const Int& tracer_idx) const {
      if (v.is_caas)
        return v.in[(static_cast<int64_t>(v.trcr_row[tracer_idx]) + 1)*v.ld + lclcellidx];
      return v.out[static_cast<int64_t>(tracer_idx)*v.ld + lclcellidx];
    }

    CEDR_B200_HD static Real inf () {
#if defined(__CUDA_ARCH__)
      return __longlong_as_double(0x7ff0000000000000LL);
#else
      return std::numeric_limits<Real>::infinity();
#endif
    }
  };

  CDR (const Options options = Options()) : options_(options), h_(nullptr),
                                            memory_(Memory::device), buf_(nullptr),
                                            host_meta_(nullptr) {}
  CDR (const CDR&) = delete;
  CDR& operator= (const CDR&) = delete;

  virtual ~CDR () {
    if (h_) cedr_b200_destroy(h_);
    if (buf_) cudaFree(buf_);
    if (host_meta_) cudaFree(host_meta_);
  }

  virtual void print (std::ostream& os) const {
    char buf[4096];
    impl::check(cedr_b200_print(h_, buf, sizeof(buf)));
    os << buf;
  }

  const Options& get_options () const { return options_; }

  virtual void declare_tracer (int problem_type, const Int& rhomidx) {
    impl::check(cedr_b200_declare_tracer(h_, problem_type, rhomidx));
  }

  virtual void end_tracer_declarations () {
    impl::check(cedr_b200_end_tracer_declarations(h_));
  }

  virtual void get_buffers_sizes (size_t& buf1, size_t& buf2) {
    impl::check(cedr_b200_get_buffers_sizes(h_, &buf1, &buf2));
  }

  // Device (or managed) memory of at least get_buffers_sizes() Reals each.
  virtual void set_buffers (Real* buf1, Real* buf2) {
    impl::check(cedr_b200_set_buffers(h_, buf1, buf2));
    user_buffers_ = true;
  }

  virtual void finish_setup () {
    if (memory_ == Memory::managed && ! user_buffers_) {
      size_t b1, b2;
      get_buffers_sizes(b1, b2);
      if (cudaMallocManaged(reinterpret_cast<void**>(&buf_), (b1 + b2 + 1)*sizeof(Real)) !=
          cudaSuccess)
        throw std::runtime_error("cedr_b200: cudaMallocManaged failed");
      cudaMemset(buf_, 0, (b1 + b2 + 1)*sizeof(Real));
      impl::check(cedr_b200_set_buffers(h_, buf_, buf_ + b1));
    }
    impl::check(cedr_b200_finish_setup(h_));
    impl::check(cedr_b200_get_device_op(h_, &op_.v));
    if (memory_ == Memory::managed) {
      // The tracer metadata the DeviceOp dereferences must be host-visible too.
      const int nt = get_num_tracers();
      if (cudaMallocManaged(reinterpret_cast<void**>(&host_meta_),
                            2*(nt + 1)*sizeof(int)) != cudaSuccess)
        throw std::runtime_error("cedr_b200: cudaMallocManaged failed");
      cudaMemcpy(host_meta_, op_.v.trcr_row, nt*sizeof(int), cudaMemcpyDeviceToHost);
      cudaMemcpy(host_meta_ + nt, op_.v.trcr_prob, nt*sizeof(int), cudaMemcpyDeviceToHost);
      op_.v.trcr_row = host_meta_;
      op_.v.trcr_prob = host_meta_ + nt;
    }
    impl::check(cedr_b200_synchronize(h_));
  }

  virtual int get_problem_type (const Int& tracer_idx) const {
    int t = 0;
    impl::check(cedr_b200_get_problem_type(h_, tracer_idx, &t));
    return t;
  }

  virtual Int get_num_tracers () const {
    int n = 0;
    impl::check(cedr_b200_get_num_tracers(h_, &n));
    return n;
  }

  virtual const DeviceOp& get_device_op () { return op_; }

  // Synchronous, like the reference's (it fences after every level).
  virtual void run () {
    impl::check(cedr_b200_run(h_));
    impl::check(cedr_b200_synchronize(h_));
  }
  // Enqueue only; results are ready after synchronize() (or stream order).
  void run_async () { impl::check(cedr_b200_run(h_)); }
  void synchronize () { impl::check(cedr_b200_synchronize(h_)); }
  void set_stream (cudaStream_t s) { impl::check(cedr_b200_set_stream(h_, s)); }

  cedr_b200_cdr* c_handle () { return h_; }

protected:
  void adopt (cedr_b200_cdr* h, const mpi::Parallel::Ptr& p, Memory m) {
    h_ = h;
    memory_ = m;
    if (p && p->size() > 1 && p->allgather())
      impl::check(cedr_b200_set_allgather(h_, p->allgather(), p->allgather_ctx()));
  }

  Options options_;
  cedr_b200_cdr* h_;
  Memory memory_;
  Real* buf_;
  int* host_meta_;
  bool user_buffers_ = false;
  DeviceOp op_;
};

// Tag standing in for the reference's Kokkos execution-space template argument.
struct DefaultExecutionSpace {};

namespace qlt {
// cedr_qlt.hpp:26-219
template <typename ES = DefaultExecutionSpace>
class QLT : public CDR {
public:
  typedef QLT<ES> Me;
  typedef std::shared_ptr<Me> Ptr;
  typedef CDR::DeviceOp DeviceOp;

  // ncells and tree refer to the global mesh (cedr_qlt.hpp:127-130).
  QLT (const mpi::Parallel::Ptr& p, const Int& ncells, const tree::Node::Ptr& tree,
       CDR::Options options = Options(), Memory memory = Memory::device)
    : CDR(options) {
    std::vector<int> kids, rank;
    std::vector<int64_t> cellidx;
    flatten(tree.get(), kids, cellidx, rank);
    // tree::Node::level set: this rank passed only its part of the tree
    // (cedr_tree_caller.hpp:20-22).
    if (tree->level >= 0) assemble_partial_trees(p, kids, cellidx, rank);
    cedr_b200_cdr* h = nullptr;
    impl::check(cedr_b200_qlt_create(
                  &h, ncells, static_cast<int>(cellidx.size()), 0, kids.data(), cellidx.data(),
                  rank.data(), options.prefer_numerical_mass_conservation_to_numerical_bounds,
                  p ? p->rank() : 0, p ? p->size() : 1));
    adopt(h, p, memory);
  }

  Int nlclcells () const {
    int n = 0;
    impl::check(cedr_b200_nlclcells(h_, &n));
    return n;
  }

  // gci2lci(gcis[i]) == i (cedr_qlt.hpp:140-144).
  void get_owned_glblcells (std::vector<Long>& gcis) const {
    std::vector<int64_t> g(nlclcells());
    impl::check(cedr_b200_get_owned_glblcells(h_, g.data()));
    gcis.assign(g.begin(), g.end());
  }

  Int gci2lci (const Int& gci) const {
    int lci = 0;
    impl::check(cedr_b200_gci2lci(h_, gci, &lci));
    return lci;
  }

  // Pre-order flattening of the caller's tree::Node graph; node 0 is the root.
  static void flatten_tree (const tree::Node* root, std::vector<int>& kids,
                            std::vector<int64_t>& cellidx, std::vector<int>& rank) {
    flatten(root, kids, cellidx, rank);
  }

  // Replace this rank's flattened partial tree by the union of all ranks' parts: two
  // setup-time all-gathers through Parallel's hook (sizes, then the padded node tables as
  // doubles -- every entry is an integer below 2^53) and cedr_b200_merge_partial_trees.
  static void assemble_partial_trees (const mpi::Parallel::Ptr& p, std::vector<int>& kids,
                                      std::vector<int64_t>& cellidx, std::vector<int>& rank) {
    const int nr = p ? p->size() : 1;
    cedr_b200_allgather_fn fn = p ? p->allgather() : nullptr;
    void* ctx = p ? p->allgather_ctx() : nullptr;
    const size_t n = cellidx.size();
    std::vector<double> cnt(nr), mine(1, static_cast<double>(n));
    impl::check(cedr_b200_allgather_host(fn, ctx, nr, mine.data(), cnt.data(), 1));
    size_t nmax = 0, ntot = 0;
    for (int r = 0; r < nr; ++r) {
      nmax = std::max(nmax, static_cast<size_t>(cnt[r]));
      ntot += static_cast<size_t>(cnt[r]);
    }
    mine.assign(4*nmax, -1.0);
    for (size_t i = 0; i < n; ++i) {
      mine[4*i] = kids[2*i];
      mine[4*i+1] = kids[2*i+1];
      mine[4*i+2] = static_cast<double>(cellidx[i]);
      mine[4*i+3] = rank[i];
    }
    std::vector<double> all(4*nmax*nr);
    impl::check(cedr_b200_allgather_host(fn, ctx, nr, mine.data(), all.data(), 4*nmax));
    std::vector<int> pn(nr), proot(nr, 0), ak(2*ntot), ar(ntot);
    std::vector<int64_t> ac(ntot);
    for (size_t r = 0, o = 0; r < static_cast<size_t>(nr); ++r) {
      pn[r] = static_cast<int>(cnt[r]);
      const double* t = all.data() + 4*nmax*r;
      for (int i = 0; i < pn[r]; ++i, ++o) {
        ak[2*o] = static_cast<int>(t[4*i]);
        ak[2*o+1] = static_cast<int>(t[4*i+1]);
        ac[o] = static_cast<int64_t>(t[4*i+2]);
        ar[o] = static_cast<int>(t[4*i+3]);
      }
    }
    int nn = 0;
    kids.assign(2*ntot, -1);
    cellidx.assign(ntot, -1);
    rank.assign(ntot, 0);
    impl::check(cedr_b200_merge_partial_trees(nr, pn.data(), proot.data(), ak.data(), ac.data(),
                                              ar.data(), static_cast<int>(ntot), &nn,
                                              kids.data(), cellidx.data(), rank.data()));
    kids.resize(2*static_cast<size_t>(nn));
    cellidx.resize(nn);
    rank.resize(nn);
  }

private:
  static void flatten (const tree::Node* root, std::vector<int>& kids,
                       std::vector<int64_t>& cellidx, std::vector<int>& rank) {
    if ( ! root) throw std::logic_error("cedr_b200: null tree");
    std::vector<std::pair<const tree::Node*, int> > stack;   // node, slot to patch
    stack.push_back(std::make_pair(root, -1));
    while ( ! stack.empty()) {
      const tree::Node* n = stack.back().first;
      const int patch = stack.back().second;
      stack.pop_back();
      const int me = static_cast<int>(cellidx.size());
      if (patch >= 0) kids[patch] = me;
      kids.push_back(-1);
      kids.push_back(-1);
      cellidx.push_back(-1);
      rank.push_back(0);
      if (n->nkids == 0) {
        cellidx[me] = n->cellidx;
        rank[me] = n->rank < 0 ? 0 : n->rank;
      } else {
        if (n->nkids != 2)   // cedr_qlt.cpp:355
          throw std::logic_error("cedr_b200: every internal tree node must have 2 kids");
        stack.push_back(std::make_pair(n->kids[1].get(), 2*me + 1));
        stack.push_back(std::make_pair(n->kids[0].get(), 2*me));
      }
    }
  }
};
} // namespace qlt

namespace caas {
// cedr_caas.hpp:15-118
template <typename ES = DefaultExecutionSpace>
class CAAS : public CDR {
public:
  typedef CAAS<ES> Me;
  typedef std::shared_ptr<Me> Ptr;
  typedef CDR::DeviceOp DeviceOp;

  // The caller's own all-reduce (cedr_caas.hpp:27-49): an MPI_Allreduce-like call that CAAS
  // makes once per run() with nlclcells / n_accum_in_place() partial sums per field
  // (cedr_caas.cpp:140-168, 262-266). sendbuf(nlocal, nfld) and rcvbuf(nfld) are DEVICE
  // pointers, as on the reference's GPU builds; the CDR's stream has been synchronised, and
  // the result must be complete (or ordered on the CDR's stream) on return.
  struct UserAllReducer {
    typedef std::shared_ptr<const UserAllReducer> Ptr;
    virtual ~UserAllReducer () {}
    virtual int operator() (const mpi::Parallel& p, Real* sendbuf, Real* rcvbuf, int nlocal,
                            int nfld, MPI_Op op) const = 0;
    virtual int n_accum_in_place () const { return 1; }
  };

  // CAAS(p, nlclcells, r), cedr_caas.hpp:51-52. With a reducer `r` this rank's cells may be
  // any set (the reducer owns the cross-rank sum). Without one the built-in tree-ordered
  // sum runs (what the reference computes when `r` is backed by its BfbTreeAllReducer);
  // on several ranks this rank's cells are then [cell0, cell0 + nlclcells) of
  // ncells_global, in the global cell order the tree-ordered sums run over.
  CAAS (const mpi::Parallel::Ptr& p, const Int nlclcells,
        const typename UserAllReducer::Ptr& r = nullptr,
        Memory memory = Memory::device, const Long cell0 = 0, const Long ncells_global = -1,
        const int sum_mode = CEDR_B200_CAAS_SUM_TREE) : user_reducer_(r) {
    cedr_b200_cdr* h = nullptr;
    impl::check(cedr_b200_caas_create(&h, nlclcells, r ? CEDR_B200_CAAS_SUM_USER : sum_mode,
                                      cell0, ncells_global < 0 ? nlclcells : ncells_global,
                                      p ? p->rank() : 0, p ? p->size() : 1));
    adopt(h, p, memory);
    if (r) {
      par_ = p ? p : mpi::make_parallel();
      impl::check(cedr_b200_caas_set_user_reducer(h, &Me::reduce_trampoline, this,
                                                  r->n_accum_in_place()));
    }
  }

private:
  typename UserAllReducer::Ptr user_reducer_;
  mpi::Parallel::Ptr par_;
  static int reduce_trampoline (void* ctx, double* send, double* recv, int nlocal, int nfld,
                                void* /*stream*/) {
    const Me* me = static_cast<const Me*>(ctx);
    return (*me->user_reducer_)(*me->par_, send, recv, nlocal, nfld, MPI_SUM);
  }
};
} // namespace caas

// cedr_bfb_tree_allreduce.hpp:15-55. Device-resident: send and recv are device pointers
// (the reference takes Kokkos device views and stages through the host). The host-buffer
// calls of the reference are accepted and ignored.
template <typename ES = DefaultExecutionSpace>
struct BfbTreeAllReducer {
  typedef BfbTreeAllReducer<ES> Me;
  typedef std::shared_ptr<Me> Ptr;

  BfbTreeAllReducer (const mpi::Parallel::Ptr& p, const tree::Node::Ptr& tree, const Int nleaf,
                     const Int nfield) : h_(nullptr), nfield_(nfield) {
    std::vector<int> kids, rank;
    std::vector<int64_t> cellidx;
    qlt::QLT<ES>::flatten_tree(tree.get(), kids, cellidx, rank);
    impl::check(cedr_b200_bfb_create(&h_, nleaf, static_cast<int>(cellidx.size()), 0,
                                     kids.data(), cellidx.data(), rank.data(), nfield, 0,
                                     p ? p->rank() : 0, p ? p->size() : 1));
    if (p && p->size() > 1 && p->allgather())
      impl::check(cedr_b200_set_allgather(h_, p->allgather(), p->allgather_ctx()));
  }
  BfbTreeAllReducer (const BfbTreeAllReducer&) = delete;
  BfbTreeAllReducer& operator= (const BfbTreeAllReducer&) = delete;
  ~BfbTreeAllReducer () { if (h_) cedr_b200_destroy(h_); }

  void get_host_buffers_sizes (size_t& buf1, size_t& buf2) { buf1 = buf2 = 0; }
  void set_host_buffers (Real*, Real*) {}
  void finish_setup () {}
  Int get_nfield () const { return nfield_; }

  // recv(nfield); send(nfield, nlocal) with the fastest index first, or
  // send(nlocal, nfield) if transpose (cedr_bfb_tree_allreduce.hpp:37-41). Synchronous.
  void allreduce (const Real* send, Real* recv, const bool transpose = false) const {
    impl::check(cedr_b200_bfb_allreduce(h_, send, recv, transpose, -1));
    impl::check(cedr_b200_synchronize(h_));
  }

private:
  cedr_b200_cdr* h_;
  Int nfield_;
};

} // namespace cedr

#endif
