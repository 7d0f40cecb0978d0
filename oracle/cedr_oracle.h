/* TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
 *
 * Plain-C restatement of COMPOSE/CEDR's property-preservation hot path, used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER
 * for the CUDA implementation. Nothing under compose_b200/ may call into it.
 *
 * Parity status: PINNED. tests/test_oracle_vs_ref.py checks every function here
 * bit-for-bit against the unmodified reference sources compiled by
 * oracle/Makefile into oracle/_ref/ (when /root/reference is present), and
 * tests/test_oracle_golden.py checks it against fixtures under tests/golden/
 * that were generated from that same reference build
 * (tests/golden/make_golden.py).
 *
 * Array convention (same as oracle/ref_driver.cpp): SoA, tracer-major, global
 * cell id fastest: a[t*ncells + gci]; rhom[gci].
 */
#ifndef CEDR_B200_ORACLE_H
#define CEDR_B200_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cedr.hpp:29-39 */
enum {
  ORACLE_PT_CONSERVE = 1,
  ORACLE_PT_SHAPEPRESERVE = 2,
  ORACLE_PT_CONSISTENT = 4,
  ORACLE_PT_NONNEGATIVE = 8
};

/* Recursive-bisection tree of cedr_tree.cpp:391-413 (oned::make_tree) as flat
 * arrays. kids has 2*(2*ncells-1) entries (-1,-1 for a leaf), cellidx one per
 * node (-1 for internal nodes). Returns the root index (always 0) or -1. */
int oracle_make_bisection_tree(int ncells, int imbalanced, int* kids,
                               int64_t* cellidx);

/* Leaf numbering of tree::analyze on one rank (cedr_tree.cpp:55-213): lci ->
 * gci in DFS order; also the number of levels (= tree height + 1). */
int oracle_leaf_order(int nnodes, int root, const int* kids, const int64_t* cellidx,
                      int64_t* lci2gci, int* nlevels);

/* Canonical problem type, QLT::MetaData::get_problem_type(get_problem_type_idx)
 * (cedr_qlt.cpp:85-96, cedr_qlt_inl.hpp:101-108). Returns -1 if invalid. */
int oracle_qlt_canonical_problem_type(int mask);

/* QLT::run (cedr_qlt.cpp:618-640) for nt independent tracers on one rank.
 * Returns 0, or nonzero on invalid input. */
int oracle_qlt_run(int ncells, int nnodes, int root, const int* kids,
                   const int64_t* cellidx, int nt, const int* ptypes,
                   int prefer_mass_con, const double* rhom, const double* qm_min,
                   const double* qm, const double* qm_max, const double* qm_prev,
                   double* qm_out);

/* CAAS::run (cedr_caas.cpp:258-270). reducer 0: sequential local sums
 * (cedr_caas.cpp:171-199 on a host backend); reducer 1: tree-ordered sums as the
 * reference BfbTreeAllReducer computes them (cedr_bfb_tree_allreduce.cpp:
 * 86-124) over the given tree. The tree arguments are ignored for reducer 0. */
int oracle_caas_run(int ncells, int reducer, int nnodes, int root, const int* kids,
                    const int64_t* cellidx, int nt, const int* ptypes,
                    const double* qm_min, const double* qm, const double* qm_max,
                    const double* qm_prev, double* qm_out);

/* BfbTreeAllReducer::allreduce on one rank (cedr_bfb_tree_allreduce.cpp:78-159).
 * send is (nfield fastest, nleaf) unless transpose, then (nleaf fastest,
 * nfield); leaf index = local leaf order (lci). */
int oracle_bfb_allreduce(int nleaf, int nnodes, int root, const int* kids,
                         const int64_t* cellidx, int nfield, int transpose,
                         const double* send, double* recv);

/* cedr_local_inl.hpp */
int oracle_solve_1eq_bc_qp_2d(const double* w, const double* a, double b,
                              const double* xlo, const double* xhi,
                              const double* y, double* x, int clip,
                              int early_exit_on_tol);
int oracle_solve_1eq_bc_qp(int n, const double* w, const double* a, double b,
                           const double* xlo, const double* xhi, const double* y,
                           double* x, int max_its);
void oracle_local_caas(int n, const double* a, double b, const double* xlo,
                       const double* xhi, const double* y, double* x, int clip);
int oracle_solve_1eq_nonneg(int n, const double* a, double b, const double* y,
                            double* x, const double* w, int method);
/* cedr_qlt_inl.hpp:119-203 */
void oracle_solve_node_problem(int problem_type, double rhom, const double* pd,
                               double Qm, double rhom0, const double* k0d,
                               double* Qm0, double rhom1, const double* k1d,
                               double* Qm1, int prefer_mass_con);

/* Synthetic inputs of SURVEY.md section 8(d) for tracers [t0, t1). */
void oracle_fill_headline(int ncells, int config_id, int t0, int t1, double* rhom,
                          double* qm_min, double* qm, double* qm_max, double* qm_prev);

#ifdef __cplusplus
}
#endif
#endif
