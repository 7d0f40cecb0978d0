// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// C-callable driver around the UNMODIFIED reference CEDR sources
// (/root/reference/cedr/*.cpp, compiled where they lie by oracle/Makefile with
// the stand-in headers in oracle/shim/). It lets tests/ and bench.py's
// cpu_baseline / --impl reference legs feed the reference's own QLT
// (cedr_qlt.cpp:618 QLT::run) and CAAS (cedr_caas.cpp:258 CAAS::run) identical
// inputs and read back identical-layout outputs.
//
// Array convention for every entry point: SoA, tracer-major, *global cell id*
// fastest: a[t*ncells + gci]; rhom[gci]. The driver translates gci -> the
// reference's local cell index through QLT::get_owned_glblcells
// (cedr_qlt.cpp:243-256).
//
// The reference indexes its buffers with 32-bit ints (cedr_qlt_inl.hpp:17-18),
// so nslots*(1+4*nt) must stay below 2^31; the driver refuses larger problems
// (callers batch tracers -- tracers are independent problems).

#include "cedr_qlt.hpp"
#include "cedr_caas.hpp"
#include "cedr_bfb_tree_allreduce.hpp"
#include "cedr_tree.hpp"

#ifdef _OPENMP
# include <omp.h>
#endif
#include <chrono>
#include <cstdint>
#include <sstream>
#include <string>
#include <vector>

namespace {

using namespace cedr;
typedef qlt::QLT<Kokkos::DefaultHostExecutionSpace> QLTT;
typedef caas::CAAS<Kokkos::DefaultHostExecutionSpace> CAAST;

thread_local std::string g_err;

double now_s () {
  return std::chrono::duration<double>(
    std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Build the caller-side tree (cedr_tree_caller.hpp:12-24) from flat arrays:
// kids[2*i], kids[2*i+1] are node indices or -1 for a leaf; cellidx[i] is the
// global cell of leaf i.
tree::Node::Ptr build_tree (const int* kids, const int64_t* cellidx, int idx,
                            const tree::Node* parent) {
  auto n = std::make_shared<tree::Node>();
  n->parent = parent;
  if (kids[2*idx] < 0) {
    n->nkids = 0;
    n->rank = 0;
    n->cellidx = cellidx[idx];
    return n;
  }
  n->nkids = 2;
  n->kids[0] = build_tree(kids, cellidx, kids[2*idx], n.get());
  n->kids[1] = build_tree(kids, cellidx, kids[2*idx+1], n.get());
  return n;
}

tree::Node::Ptr get_tree (const mpi::Parallel::Ptr& p, int ncells, int tree_kind,
                          int root, const int* kids, const int64_t* cellidx) {
  switch (tree_kind) {
  case 0: return tree::make_tree_over_1d_mesh(p, ncells, false);
  case 1: return tree::make_tree_over_1d_mesh(p, ncells, true);
  case 2: return build_tree(kids, cellidx, root, nullptr);
  }
  throw std::logic_error("ref_driver: bad tree_kind");
}

// CAAS UserAllReducer (cedr_caas.hpp:27-49) that sums in tree order by
// delegating to the reference BfbTreeAllReducer (cedr_bfb_tree_allreduce.cpp:
// 78-159) with transpose=true, which matches CAAS's (nlocal fastest, nfld) send
// layout (cedr_caas.cpp:153-154).
struct TreeOrderedReducer : public CAAST::UserAllReducer {
  BfbTreeAllReducer<Kokkos::DefaultHostExecutionSpace>::Ptr r;
  TreeOrderedReducer (const mpi::Parallel::Ptr& p, const tree::Node::Ptr& t,
                      Int nleaf, Int nfield) {
    r = std::make_shared<BfbTreeAllReducer<Kokkos::DefaultHostExecutionSpace> >(
      p, t, nleaf, nfield);
  }
  int operator() (const mpi::Parallel&, Real* send, Real* recv, int nlocal,
                  int nfld, MPI_Op) const override {
    typedef BfbTreeAllReducer<Kokkos::DefaultHostExecutionSpace> B;
    B::ConstRealList s(send, static_cast<size_t>(nlocal)*nfld);
    B::RealList rv(recv, nfld);
    r->allreduce(s, rv, true);
    return 0;
  }
};

} // namespace

extern "C" {

const char* cedr_ref_last_error () { return g_err.c_str(); }

int cedr_ref_num_threads () {
  return Kokkos::DefaultHostExecutionSpace::concurrency();
}

// torchrun exports OMP_NUM_THREADS=1 to every rank; bench.py's reference arm asks for
// the host's cores explicitly. Returns the thread count in effect afterwards.
int cedr_ref_set_num_threads (int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void) n;
#endif
  return Kokkos::DefaultHostExecutionSpace::concurrency();
}

// Reference QLT on one rank. Returns 0 on success, nonzero + last_error on
// failure. seconds[0..nrep) receives the wall time of each QLT::run() alone
// (inputs are re-set before every repetition because consistent-only tracers
// mutate the l2r buffer, cedr_qlt.cpp:555-559).
int cedr_ref_qlt (int ncells, int tree_kind, int tree_root, const int* tree_kids,
                  const int64_t* tree_cellidx,
                  int nt, const int* ptypes, int prefer_mass_con,
                  const double* rhom, const double* qm_min, const double* qm,
                  const double* qm_max, const double* qm_prev,
                  double* qm_out, int* ptypes_out, int nrep, double* seconds) {
  try {
    auto p = mpi::make_parallel(MPI_COMM_WORLD);
    auto tree = get_tree(p, ncells, tree_kind, tree_root, tree_kids, tree_cellidx);
    CDR::Options opts;
    opts.prefer_numerical_mass_conservation_to_numerical_bounds = prefer_mass_con != 0;
    QLTT q(p, ncells, tree, opts);
    tree = nullptr;
    for (int t = 0; t < nt; ++t) q.declare_tracer(ptypes[t], 0);
    q.end_tracer_declarations();
    {
      size_t b1, b2;
      q.get_buffers_sizes(b1, b2);
      if (b1 >= (size_t(1) << 31) || b2 >= (size_t(1) << 31)) {
        g_err = "ref_driver: problem exceeds the reference's 32-bit buffer indexing; "
                "batch the tracers";
        return 2;
      }
    }
    q.finish_setup();
    if (ptypes_out)
      for (int t = 0; t < nt; ++t) ptypes_out[t] = q.get_problem_type(t);
    std::vector<Long> gcis;
    q.get_owned_glblcells(gcis);
    const Int n = q.nlclcells();
    const auto op = static_cast<const QLTT::DeviceOp&>(q.get_device_op());
    for (Int i = 0; i < n; ++i) op.set_rhom(i, 0, rhom[gcis[i]]);
    for (int rep = 0; rep < (nrep < 1 ? 1 : nrep); ++rep) {
#ifdef _OPENMP
#     pragma omp parallel for
#endif
      for (int t = 0; t < nt; ++t) {
        const size_t os = static_cast<size_t>(t)*ncells;
        const bool conserve = ptypes[t] & ProblemType::conserve;
        for (Int i = 0; i < n; ++i) {
          const size_t k = os + gcis[i];
          if (conserve)
            op.set_Qm(i, t, qm[k], qm_min[k], qm_max[k], qm_prev[k]);
          else
            op.set_Qm(i, t, qm[k], qm_min[k], qm_max[k]);
        }
      }
      const double t0 = now_s();
      q.run();
      const double t1 = now_s();
      if (seconds) seconds[rep] = t1 - t0;
    }
#ifdef _OPENMP
#   pragma omp parallel for
#endif
    for (int t = 0; t < nt; ++t) {
      const size_t os = static_cast<size_t>(t)*ncells;
      for (Int i = 0; i < n; ++i) qm_out[os + gcis[i]] = op.get_Qm(i, t);
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Reference CAAS on one rank. reducer = 0: the reference's own local sums
// (sequential over cells on a host backend) + MPI_Allreduce; reducer = 1: a
// UserAllReducer that delegates to the reference BfbTreeAllReducer over the
// given tree (tree-ordered pairwise sums). Inputs are re-set before every
// repetition because CAAS::run() clips Qm in place (cedr_caas.cpp:177).
int cedr_ref_caas (int ncells, int reducer, int tree_kind, int tree_root,
                   const int* tree_kids, const int64_t* tree_cellidx,
                   int nt, const int* ptypes,
                   const double* rhom, const double* qm_min, const double* qm,
                   const double* qm_max, const double* qm_prev,
                   double* qm_out, int nrep, double* seconds) {
  try {
    auto p = mpi::make_parallel(MPI_COMM_WORLD);
    CAAST::UserAllReducer::Ptr red;
    if (reducer == 1) {
      auto tree = get_tree(p, ncells, tree_kind, tree_root, tree_kids, tree_cellidx);
      red = std::make_shared<TreeOrderedReducer>(p, tree, ncells, 4*nt);
    }
    CAAST c(p, ncells, red);
    for (int t = 0; t < nt; ++t) c.declare_tracer(ptypes[t], 0);
    c.end_tracer_declarations();
    {
      size_t b1, b2;
      c.get_buffers_sizes(b1, b2);
      if (b1 >= (size_t(1) << 31) || b2 >= (size_t(1) << 31)) {
        g_err = "ref_driver: problem exceeds the reference's 32-bit buffer indexing; "
                "batch the tracers";
        return 2;
      }
    }
    c.finish_setup();
    const auto op = static_cast<const CAAST::DeviceOp&>(c.get_device_op());
    for (Int i = 0; i < ncells; ++i) op.set_rhom(i, 0, rhom[i]);
    for (int rep = 0; rep < (nrep < 1 ? 1 : nrep); ++rep) {
#ifdef _OPENMP
#     pragma omp parallel for
#endif
      for (int t = 0; t < nt; ++t) {
        const size_t os = static_cast<size_t>(t)*ncells;
        for (Int i = 0; i < ncells; ++i) {
          const size_t k = os + i;
          op.set_Qm(i, t, qm[k], qm_min[k], qm_max[k], qm_prev[k]);
        }
      }
      const double t0 = now_s();
      c.run();
      const double t1 = now_s();
      if (seconds) seconds[rep] = t1 - t0;
    }
#ifdef _OPENMP
#   pragma omp parallel for
#endif
    for (int t = 0; t < nt; ++t) {
      const size_t os = static_cast<size_t>(t)*ncells;
      for (Int i = 0; i < ncells; ++i) qm_out[os + i] = op.get_Qm(i, t);
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Reference BfbTreeAllReducer::allreduce (cedr_bfb_tree_allreduce.cpp:78-159)
// on one rank: send is (nfield fastest, nleaf) unless transpose, recv is
// (nfield).
int cedr_ref_bfb_allreduce (int nleaf, int tree_kind, int tree_root,
                            const int* tree_kids, const int64_t* tree_cellidx,
                            int nfield, int transpose, const double* send,
                            double* recv) {
  try {
    typedef BfbTreeAllReducer<Kokkos::DefaultHostExecutionSpace> B;
    auto p = mpi::make_parallel(MPI_COMM_WORLD);
    auto tree = get_tree(p, nleaf, tree_kind, tree_root, tree_kids, tree_cellidx);
    B b(p, tree, nleaf, nfield);
    B::ConstRealList s(send, static_cast<size_t>(nleaf)*nfield);
    B::RealList r(recv, nfield);
    b.allreduce(s, r, transpose != 0);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// The reference's leaf numbering for a tree: lci -> gci
// (QLT::get_owned_glblcells, cedr_qlt.cpp:243-256). Used to pin the new code's
// gci2lci against the reference's.
int cedr_ref_qlt_leaf_order (int ncells, int tree_kind, int tree_root,
                             const int* tree_kids, const int64_t* tree_cellidx,
                             int64_t* gcis_out, int* nlevels_out, int* nslots_out) {
  try {
    auto p = mpi::make_parallel(MPI_COMM_WORLD);
    auto tree = get_tree(p, ncells, tree_kind, tree_root, tree_kids, tree_cellidx);
    auto ns = tree::analyze(p, ncells, tree);
    for (const auto& idx : ns->levels[0].nodes) {
      const auto n = ns->node_h(idx);
      gcis_out[n->offset] = n->id;
    }
    if (nlevels_out) *nlevels_out = static_cast<int>(ns->levels.size());
    if (nslots_out) *nslots_out = ns->nslots;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Reference local solvers (cedr_local_inl.hpp) for pinning the restatement.
int cedr_ref_solve_1eq_bc_qp_2d (const double* w, const double* a, double b,
                                 const double* xlo, const double* xhi,
                                 const double* y, double* x, int clip,
                                 int early_exit_on_tol) {
  return local::solve_1eq_bc_qp_2d(w, a, b, xlo, xhi, y, x, clip != 0,
                                   early_exit_on_tol != 0);
}

int cedr_ref_solve_1eq_bc_qp (int n, const double* w, const double* a, double b,
                              const double* xlo, const double* xhi,
                              const double* y, double* x, int max_its) {
  return local::solve_1eq_bc_qp(n, w, a, b, xlo, xhi, y, x, max_its);
}

void cedr_ref_local_caas (int n, const double* a, double b, const double* xlo,
                          const double* xhi, const double* y, double* x, int clip) {
  local::caas(n, a, b, xlo, xhi, y, x, clip != 0);
}

int cedr_ref_solve_1eq_nonneg (int n, const double* a, double b, const double* y,
                               double* x, const double* w, int method) {
  return local::solve_1eq_nonneg(n, a, b, y, x, w,
                                 method ? local::Method::caas
                                        : local::Method::least_squares);
}

void cedr_ref_solve_node_problem (int problem_type, double rhom, const double* pd,
                                  double Qm, double rhom0, const double* k0d,
                                  double* Qm0, double rhom1, const double* k1d,
                                  double* Qm1, int prefer_mass_con) {
  qlt::impl::solve_node_problem(problem_type, rhom, pd, Qm, rhom0, k0d, *Qm0,
                                rhom1, k1d, *Qm1, prefer_mass_con != 0);
}

} // extern "C"
