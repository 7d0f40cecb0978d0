// C++ caller-side test of include/cedr_b200.hpp, written the way the reference's own
// callers use cedr::CDR:
//   (A) device path: the concrete DeviceOp copied by value into kernels that call
//       set_rhom / set_Qm / get_Qm per (cell, tracer), as TestRandomized::run does with
//       Kokkos lambdas (cedr_test_randomized_inl.hpp:12-65);
//   (B) host path: plain host loops over op.set_Qm / cdr.run() / op.get_Qm, as the 1-D
//       transport test does (cedr_test_1d_transport.cpp:170-189), on a CDR built with
//       Memory::managed;
// and checks the results bit-for-bit against the CPU oracle (oracle/cedr_oracle.h --
// test infrastructure, linked only into this test), plus the reference's error behaviour.
#include <cstdio>
#include <cstring>
#include <sstream>
#include <vector>

#include "cedr_b200.hpp"
#include "cedr_b200_local.hpp"
#include "cedr_oracle.h"

using namespace cedr;

template <typename Op>
__global__ void set_all (const Op op, const int n, const int nt, const double* rhom,
                         const double* lo, const double* q, const double* hi,
                         const double* prev, const long long* gcis) {
  // One thread per (cell, tracer); rhom first, as the contract requires -- here by a
  // separate launch (see main).
  const long long k = blockIdx.x*(long long) blockDim.x + threadIdx.x;
  if (k >= (long long) n*nt) return;
  const int t = (int) (k / n), i = (int) (k % n);
  const long long g = gcis[i];
  op.set_Qm(i, t, q[(long long) t*n + g], lo[(long long) t*n + g], hi[(long long) t*n + g],
            prev[(long long) t*n + g]);
}
template <typename Op>
__global__ void set_rhom_all (const Op op, const int n, const double* rhom,
                              const long long* gcis) {
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) op.set_rhom(i, 0, rhom[gcis[i]]);
}
template <typename Op>
__global__ void get_all (const Op op, const int n, const int nt, double* out,
                         const long long* gcis) {
  const long long k = blockIdx.x*(long long) blockDim.x + threadIdx.x;
  if (k >= (long long) n*nt) return;
  const int t = (int) (k / n), i = (int) (k % n);
  out[(long long) t*n + gcis[i]] = op.get_Qm(i, t);
}

// The element-local solvers on the device: one thread per problem, 16-wide slots.
// which: 0 solve_1eq_bc_qp, 1 caas, 2 solve_1eq_nonneg(least_squares), 3 (caas),
// 4 solve_1eq_bc_qp_2d (n = 2).
__global__ void local_kernel (const int nprob, const int which, const int* n, const double* w,
                              const double* a, const double* b, const double* xlo,
                              const double* xhi, const double* y, double* x, int* info) {
  const int p = blockIdx.x*blockDim.x + threadIdx.x;
  if (p >= nprob) return;
  const int o = 16*p;
  namespace L = cedr::local;
  int r = 0;
  switch (which) {
  case 0: r = L::solve_1eq_bc_qp(n[p], w + o, a + o, b[p], xlo + o, xhi + o, y + o, x + o); break;
  case 1: L::caas(n[p], a + o, b[p], xlo + o, xhi + o, y + o, x + o); break;
  case 2: r = L::solve_1eq_nonneg(n[p], a + o, b[p], y + o, x + o, w + o, L::Method::least_squares); break;
  case 3: r = L::solve_1eq_nonneg(n[p], a + o, b[p], y + o, x + o, w + o, L::Method::caas); break;
  case 4: r = L::solve_1eq_bc_qp_2d(w + o, a + o, b[p], xlo + o, xhi + o, y + o, x + o); break;
  }
  info[p] = r;
}

template <typename T> T* to_dev (const std::vector<T>& h) {
  T* d = nullptr;
  cudaMalloc(&d, h.size()*sizeof(T));
  cudaMemcpy(d, h.data(), h.size()*sizeof(T), cudaMemcpyHostToDevice);
  return d;
}

struct Problem {
  int n, nt;
  std::vector<int> pts;
  std::vector<double> rhom, lo, q, hi, prev;
  Problem (int n_, const std::vector<int>& pts_) : n(n_), nt((int) pts_.size()), pts(pts_) {
    rhom.resize(n);
    for (auto* v : {&lo, &q, &hi, &prev}) v->resize((size_t) n*nt);
    oracle_fill_headline(n, 9, 0, nt, rhom.data(), lo.data(), q.data(), hi.data(), prev.data());
  }
};

static int nerr = 0;
#define REQUIRE(c) do { if ( ! (c)) { ++nerr; std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); } } while (0)

static bool same_bits (const std::vector<double>& a, const std::vector<double>& b) {
  return a.size() == b.size() && std::memcmp(a.data(), b.data(), a.size()*sizeof(double)) == 0;
}

static std::vector<double> oracle_qlt (const Problem& p, bool imbalanced, bool prefer) {
  const int nn = 2*p.n - 1;
  std::vector<int> kids(2*nn);
  std::vector<int64_t> cellidx(nn);
  oracle_make_bisection_tree(p.n, imbalanced, kids.data(), cellidx.data());
  std::vector<double> out((size_t) p.n*p.nt);
  REQUIRE(oracle_qlt_run(p.n, nn, 0, kids.data(), cellidx.data(), p.nt, p.pts.data(), prefer,
                         p.rhom.data(), p.lo.data(), p.q.data(), p.hi.data(), p.prev.data(),
                         out.data()) == 0);
  return out;
}

static std::vector<double> oracle_caas (const Problem& p) {
  const int nn = 2*p.n - 1;
  std::vector<int> kids(2*nn);
  std::vector<int64_t> cellidx(nn);
  oracle_make_bisection_tree(p.n, 0, kids.data(), cellidx.data());
  std::vector<double> out((size_t) p.n*p.nt);
  REQUIRE(oracle_caas_run(p.n, 1, nn, 0, kids.data(), cellidx.data(), p.nt, p.pts.data(),
                          p.lo.data(), p.q.data(), p.hi.data(), p.prev.data(), out.data()) == 0);
  return out;
}

// (A) through kernels that copy the DeviceOp by value.
template <typename CDRT>
static std::vector<double> run_device (CDRT& cdr, const Problem& p,
                                       const std::vector<long long>& gcis) {
  for (int t = 0; t < p.nt; ++t) cdr.declare_tracer(p.pts[t], 0);
  cdr.end_tracer_declarations();
  cdr.finish_setup();
  const typename CDRT::DeviceOp op = cdr.get_device_op();
  double* rhom = to_dev(p.rhom), * lo = to_dev(p.lo), * q = to_dev(p.q), * hi = to_dev(p.hi),
    * prev = to_dev(p.prev);
  long long* g = to_dev(gcis);
  double* out = nullptr;
  cudaMalloc(&out, (size_t) p.n*p.nt*sizeof(double));
  const int nb = (int) (((long long) p.n*p.nt + 255)/256);
  set_rhom_all<<<(p.n + 255)/256, 256>>>(op, p.n, rhom, g);
  set_all<<<nb, 256>>>(op, p.n, p.nt, rhom, lo, q, hi, prev, g);
  cdr.run();
  get_all<<<nb, 256>>>(op, p.n, p.nt, out, g);
  std::vector<double> h((size_t) p.n*p.nt);
  cudaMemcpy(h.data(), out, h.size()*sizeof(double), cudaMemcpyDeviceToHost);
  for (void* d : {(void*) rhom, (void*) lo, (void*) q, (void*) hi, (void*) prev, (void*) g,
        (void*) out})
    cudaFree(d);
  return h;
}

// (B) through host calls on a managed-memory CDR.
template <typename CDRT>
static std::vector<double> run_host (CDRT& cdr, const Problem& p,
                                     const std::vector<long long>& gcis) {
  for (int t = 0; t < p.nt; ++t) cdr.declare_tracer(p.pts[t], 0);
  cdr.end_tracer_declarations();
  cdr.finish_setup();
  const CDR::DeviceOp& op = cdr.get_device_op();
  std::vector<double> out((size_t) p.n*p.nt);
  for (int rep = 0; rep < 2; ++rep) {   // CDR::run is repeatable (a time-step loop)
    for (int i = 0; i < p.n; ++i) op.set_rhom(i, 0, p.rhom[gcis[i]]);
    for (int t = 0; t < p.nt; ++t)
      for (int i = 0; i < p.n; ++i) {
        const size_t k = (size_t) t*p.n + gcis[i];
        op.set_Qm(i, t, p.q[k], p.lo[k], p.hi[k], p.prev[k]);
      }
    cdr.run();
    for (int t = 0; t < p.nt; ++t)
      for (int i = 0; i < p.n; ++i) out[(size_t) t*p.n + gcis[i]] = op.get_Qm(i, t);
  }
  return out;
}

int main () {
  typedef qlt::QLT<> QLTT;
  typedef caas::CAAS<> CAAST;
  const int cst = ProblemType::conserve | ProblemType::shapepreserve | ProblemType::consistent,
    st = ProblemType::shapepreserve | ProblemType::consistent,
    ct = ProblemType::conserve | ProblemType::consistent, t_ = ProblemType::consistent;
  mpi::Parallel::Ptr par = mpi::make_parallel();

  for (const int n : {1, 2, 21, 111, 1350, 5400})
    for (const bool imbalanced : {false, true})
      for (const bool prefer : {false, true}) {
        if (n == 5400 && (imbalanced || prefer)) continue;
        const Problem p(n, {cst, st, ct, t_, ProblemType::shapepreserve, cst});
        CDR::Options o;
        o.prefer_numerical_mass_conservation_to_numerical_bounds = prefer;
        const std::vector<double> ref = oracle_qlt(p, imbalanced, prefer);
        tree::Node::Ptr tree = tree::make_tree_over_1d_mesh(par, n, imbalanced);
        std::vector<Long> g;
        {
          QLTT q(par, n, tree, o);
          tree = nullptr;   // the CDR keeps nothing of the caller's tree (cedr_tree.cpp:477)
          q.get_owned_glblcells(g);
          REQUIRE(q.nlclcells() == n && (int) g.size() == n);
          for (int i = 0; i < n; i += 1 + n/7) REQUIRE(q.gci2lci((Int) g[i]) == i);
          const std::vector<long long> gl(g.begin(), g.end());
          REQUIRE(same_bits(run_device(q, p, gl), ref));
          REQUIRE(q.get_num_tracers() == p.nt);
          REQUIRE(q.get_problem_type(4) == st);   // `s` is reported canonically as `st`
        }
        if (n <= 1350) {
          QLTT q(par, n, tree::make_tree_over_1d_mesh(par, n, imbalanced), o, Memory::managed);
          const std::vector<long long> gl(g.begin(), g.end());
          REQUIRE(same_bits(run_host(q, p, gl), ref));
        }
      }

  { // tree::Node::level set (cedr_tree_caller.hpp:20-22): the constructor takes the tree
    // for this rank's PART, gathers the parts through Parallel's hook and merges them.
    // On one rank the part is the whole tree; the multi-part merge is tested on the host
    // (tests/test_host_logic.py, tests/test_multirank_host.py).
    struct SetLevel {
      static int go (tree::Node* nd) {
        int l = 0;
        for (int k = 0; k < nd->nkids; ++k) l = std::max(l, 1 + go(nd->kids[k].get()));
        return nd->level = l;
      }
    };
    for (const int n : {21, 1350}) {
      const Problem p(n, {cst, st, ct, t_, ProblemType::shapepreserve, cst});
      tree::Node::Ptr tree = tree::make_tree_over_1d_mesh(par, n, true);
      SetLevel::go(tree.get());
      QLTT q(par, n, tree, CDR::Options());
      std::vector<Long> g;
      q.get_owned_glblcells(g);
      const std::vector<long long> gl(g.begin(), g.end());
      REQUIRE(same_bits(run_device(q, p, gl), oracle_qlt(p, true, false)));
    }
  }

  for (const int n : {1, 4, 11, 1350, 5400}) {
    const Problem p(n, {cst, ProblemType::shapepreserve, cst, st});
    const std::vector<double> ref = oracle_caas(p);
    std::vector<long long> id(n);
    for (int i = 0; i < n; ++i) id[i] = i;
    { CAAST c(par, n); REQUIRE(same_bits(run_device(c, p, id), ref)); }
    if (n <= 1350) {
      CAAST c(par, n, nullptr, Memory::managed);
      REQUIRE(same_bits(run_host(c, p, id), ref));
    }
  }

  { // A caller-supplied UserAllReducer (cedr_caas.hpp:27-49), here one that delegates to
    // the public BfbTreeAllReducer with transpose = true -- the (nlocal, nfld) layout CAAS
    // sends (cedr_caas.cpp:153-154, cedr_bfb_tree_allreduce.cpp:95-97): must reproduce the
    // built-in tree-ordered sums, i.e. the oracle, bit for bit.
    struct TreeReducer : public CAAST::UserAllReducer {
      mutable std::shared_ptr<BfbTreeAllReducer<> > r;
      mpi::Parallel::Ptr par;
      int ncells, calls = 0;
      TreeReducer (const mpi::Parallel::Ptr& p, int n) : par(p), ncells(n) {}
      int operator() (const mpi::Parallel&, Real* send, Real* recv, int nlocal, int nfld,
                      MPI_Op op) const override {
        if (op != MPI_SUM || nlocal != ncells) return 1;
        if ( ! r)
          r = std::make_shared<BfbTreeAllReducer<> >(
            par, tree::make_tree_over_1d_mesh(par, ncells, false), ncells, nfld);
        r->allreduce(send, recv, true);
        ++const_cast<TreeReducer*>(this)->calls;
        return 0;
      }
    };
    for (const int n : {11, 1350}) {
      const Problem p(n, {cst, ProblemType::shapepreserve, cst, st});
      const std::vector<double> ref = oracle_caas(p);
      std::vector<long long> id(n);
      for (int i = 0; i < n; ++i) id[i] = i;
      auto red = std::make_shared<TreeReducer>(par, n);
      CAAST c(par, n, red);
      REQUIRE(same_bits(run_device(c, p, id), ref));
      REQUIRE(red->calls >= 1);
    }
  }

  { // cedr::local on the device == the oracle's restatement of the reference, bitwise.
    const int np = 4096;
    unsigned long long seed = 12345;
    auto U = [&] () {   // splitmix64 -> [0, 1)
      unsigned long long z = (seed += 0x9E3779B97F4A7C15ull);
      z = (z ^ (z >> 30))*0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27))*0x94D049BB133111EBull;
      return (double) ((z ^ (z >> 31)) >> 11)*0x1.0p-53;
    };
    for (int which = 0; which < 5; ++which) {
      std::vector<int> n(np), info(np), info_ref(np);
      std::vector<double> w(16*np, 1), a(16*np, 1), b(np), xlo(16*np, 0), xhi(16*np, 1),
        y(16*np, 0), x(16*np, 0), x_ref(16*np, 0);
      for (int p = 0; p < np; ++p) {
        n[p] = which == 4 ? 2 : 2 + (int) (U()*14.999);
        double blo = 0, bhi = 0, bmid = 0;
        for (int i = 0; i < n[p]; ++i) {
          const int k = 16*p + i;
          w[k] = 0.1 + U(); a[k] = 0.1 + U();
          xlo[k] = U() - 0.5; xhi[k] = xlo[k] + U();
          y[k] = xlo[k] + (xhi[k] - xlo[k])*(1.6*U() - 0.3);   // some outside the bounds
          if (which == 2 || which == 3) y[k] = U() - 0.2;
          blo += a[k]*xlo[k]; bhi += a[k]*xhi[k];
          bmid += a[k]*(xlo[k] + (xhi[k] - xlo[k])*U());
        }
        // Mostly feasible masses; some at the corners, some infeasible.
        const double u = U();
        b[p] = u < 0.8 ? bmid : u < 0.85 ? blo : u < 0.9 ? bhi : u < 0.95 ? blo - 0.1 : bhi + 0.1;
        if (which == 2 || which == 3) b[p] = u < 0.9 ? U()*n[p] : -0.1;
        const int o = 16*p;
        switch (which) {
        case 0: info_ref[p] = oracle_solve_1eq_bc_qp(n[p], &w[o], &a[o], b[p], &xlo[o], &xhi[o], &y[o], &x_ref[o], 100); break;
        case 1: oracle_local_caas(n[p], &a[o], b[p], &xlo[o], &xhi[o], &y[o], &x_ref[o], 1); info_ref[p] = 0; break;
        case 2: info_ref[p] = oracle_solve_1eq_nonneg(n[p], &a[o], b[p], &y[o], &x_ref[o], &w[o], 0); break;
        case 3: info_ref[p] = oracle_solve_1eq_nonneg(n[p], &a[o], b[p], &y[o], &x_ref[o], &w[o], 1); break;
        case 4: info_ref[p] = oracle_solve_1eq_bc_qp_2d(&w[o], &a[o], b[p], &xlo[o], &xhi[o], &y[o], &x_ref[o], 1, 1); break;
        }
      }
      int* dn = to_dev(n), * dinfo = to_dev(info);
      double* dw = to_dev(w), * da = to_dev(a), * db = to_dev(b), * dlo = to_dev(xlo),
        * dhi = to_dev(xhi), * dy = to_dev(y), * dx = to_dev(x);
      local_kernel<<<(np + 127)/128, 128>>>(np, which, dn, dw, da, db, dlo, dhi, dy, dx, dinfo);
      cudaMemcpy(x.data(), dx, x.size()*sizeof(double), cudaMemcpyDeviceToHost);
      cudaMemcpy(info.data(), dinfo, np*sizeof(int), cudaMemcpyDeviceToHost);
      int nfeas = 0;
      for (int p = 0; p < np; ++p) nfeas += info_ref[p] >= 0;
      REQUIRE(nfeas > np/2);
      REQUIRE(info == info_ref);
      // An infeasible return of solve_1eq_nonneg leaves x untouched in both versions
      // (zeros here), so the whole array compares.
      REQUIRE(same_bits(x, x_ref));
      for (void* d : {(void*) dn, (void*) dinfo, (void*) dw, (void*) da, (void*) db, (void*) dlo,
            (void*) dhi, (void*) dy, (void*) dx})
        cudaFree(d);
    }
  }

  { // BfbTreeAllReducer: device sums in tree order == the oracle's, both layouts.
    const int nleaf = 777, nf = 5;
    std::vector<double> data((size_t) nleaf*nf), ref(nf), got(nf);
    for (size_t i = 0; i < data.size(); ++i) data[i] = std::sin(0.37*i)*(1 + (i % 11));
    const int nn = 2*nleaf - 1;
    std::vector<int> kids(2*nn);
    std::vector<int64_t> cellidx(nn);
    oracle_make_bisection_tree(nleaf, 0, kids.data(), cellidx.data());
    BfbTreeAllReducer<> red(par, tree::make_tree_over_1d_mesh(par, nleaf), nleaf, nf);
    double* d = to_dev(data);
    for (int transpose = 0; transpose < 2; ++transpose) {
      REQUIRE(oracle_bfb_allreduce(nleaf, nn, 0, kids.data(), cellidx.data(), nf, transpose,
                                   data.data(), ref.data()) == 0);
      red.allreduce(d, d, transpose != 0);    // in place, like the reference allows
      cudaMemcpy(got.data(), d, nf*sizeof(double), cudaMemcpyDeviceToHost);
      REQUIRE(same_bits(got, ref));
      cudaMemcpy(d, data.data(), data.size()*sizeof(double), cudaMemcpyHostToDevice);
    }
    cudaFree(d);
  }

  { // Error behaviour: std::logic_error, message shaped like cedr_throw_if's.
    QLTT q(par, 8, tree::make_tree_over_1d_mesh(par, 8));
    q.declare_tracer(cst, 0);
    q.end_tracer_declarations();
    bool threw = false;
    try { q.declare_tracer(cst, 0); }
    catch (const std::logic_error& e) {
      threw = std::strstr(e.what(), "The condition:") && std::strstr(e.what(), "led to the exception");
    }
    REQUIRE(threw);
    threw = false;
    try { CAAST c(par, 4); c.declare_tracer(ct, 0); }
    catch (const std::logic_error&) { threw = true; }
    REQUIRE(threw);
    std::stringstream ss;
    q.finish_setup();
    q.print(ss);
    REQUIRE(ss.str().find("QLT") != std::string::npos);
  }

  std::printf(nerr ? "FAIL (%d)\n" : "PASS\n", nerr);
  return nerr ? 1 : 0;
}
