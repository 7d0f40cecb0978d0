/* TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
 * See cedr_oracle.h for scope and parity status. Every function cites the
 * reference lines (relative to /root/reference/cedr/) whose arithmetic, in
 * that order of operations, it restates. Compile with -ffp-contract=off.
 */
#include "cedr_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PT_C ORACLE_PT_CONSERVE
#define PT_S ORACLE_PT_SHAPEPRESERVE
#define PT_T ORACLE_PT_CONSISTENT
#define PT_N ORACLE_PT_NONNEGATIVE

static double dmin(double a, double b) { return a < b ? a : b; } /* cedr_kokkos.hpp:136 */
static double dmax(double a, double b) { return a > b ? a : b; } /* cedr_kokkos.hpp:138 */

/* ------------------------------------------------------------------ local */

/* cedr_local_inl.hpp:13-18 */
static double calc_r_tol(double b, const double* a, const double* y, int n) {
  double ab = fabs(b);
  for (int i = 0; i < n; ++i) ab = dmax(ab, fabs(a[i]*y[i]));
  return 1e1*DBL_EPSILON*fabs(ab);
}

/* cedr_local_inl.hpp:23-41 */
static int check_lu(int n, const double* a, double b, const double* xlo,
                    const double* xhi, double r_tol, double* x) {
  double r = -b;
  for (int i = 0; i < n; ++i) {
    x[i] = xlo[i];
    r += a[i]*x[i];
  }
  if (fabs(r) <= r_tol) return 1;
  if (r > 0) return -1;
  r = -b;
  for (int i = 0; i < n; ++i) {
    x[i] = xhi[i];
    r += a[i]*x[i];
  }
  if (fabs(r) <= r_tol) return 1;
  if (r < 0) return -1;
  return 0;
}

/* cedr_local_inl.hpp:43-64 */
static void calc_r(int n, const double* w, const double* a, double b,
                   const double* xlo, const double* xhi, const double* y,
                   double lambda, double* x, double* r_out, double* r_lambda_out) {
  double r = 0, r_lambda = 0;
  for (int i = 0; i < n; ++i) {
    const double q = a[i]/w[i];
    const double x_trial = y[i] + lambda*q;
    double xtmp;
    if (x_trial < (xtmp = xlo[i]))
      x[i] = xtmp;
    else if (x_trial > (xtmp = xhi[i]))
      x[i] = xtmp;
    else {
      x[i] = x_trial;
      r_lambda += a[i]*q;
    }
    r += a[i]*x[i];
  }
  r -= b;
  *r_out = r;
  *r_lambda_out = r_lambda;
}

/* cedr_local_inl.hpp:68-165. Note the reference declares a second `info`
 * inside the early-exit block, so a corner solution found by check_lu (return
 * value 1) does NOT return early; only infeasibility (-1) does, leaving x at
 * xlo or xhi. */
int oracle_solve_1eq_bc_qp_2d(const double* w, const double* a, double b,
                              const double* xlo, const double* xhi,
                              const double* y, double* x, int clip,
                              int early_exit_on_tol) {
  int info;
  if (early_exit_on_tol) {
    const double r_tol = calc_r_tol(b, a, y, 2);
    const int info_inner = check_lu(2, a, b, xlo, xhi, r_tol, x);
    if (info_inner == -1) return info_inner;
  }

  {
    double qmass = 0, dm = b;
    for (int i = 0; i < 2; ++i) {
      const double qi = a[i]/w[i];
      qmass += a[i]*qi;
      dm -= a[i]*y[i];
    }
    const double lambda = dm/qmass;
    int ok = 1;
    for (int i = 0; i < 2; ++i) {
      x[i] = y[i] + lambda*(a[i]/w[i]);
      if (x[i] < xlo[i] || x[i] > xhi[i]) {
        ok = 0;
        break;
      }
    }
    if (ok) return 1;
  }

  double x_base[2];
  for (int i = 0; i < 2; ++i) x_base[i] = 0.5*b/a[i];
  const double x_dir[2] = {-a[1], a[0]};

  double alphas[4];
  alphas[0] = (xlo[1] - x_base[1])/x_dir[1]; /* bottom */
  alphas[1] = (xhi[0] - x_base[0])/x_dir[0]; /* right */
  alphas[2] = (xhi[1] - x_base[1])/x_dir[1]; /* top */
  alphas[3] = (xlo[0] - x_base[0])/x_dir[0]; /* left */

  double mn = alphas[0], mx = mn;
  int imin = 0, imax = 0;
  for (int i = 1; i < 4; ++i) {
    const double alpha = alphas[i];
    if (alpha < mn) { mn = alpha; imin = i; }
    if (alpha > mx) { mx = alpha; imax = i; }
  }
  int ais[2] = {0, 0};
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    if (i != imin && i != imax) {
      ais[cnt++] = i;
      if (cnt == 2) break;
    }

  double objs[2];
  for (int j = 0; j < 2; ++j) {
    const double alpha = alphas[ais[j]];
    double obj = 0;
    for (int i = 0; i < 2; ++i) {
      x[i] = x_base[i] + alpha*x_dir[i];
      const double d = y[i] - x[i];
      obj += w[i]*(d*d);
    }
    objs[j] = obj;
  }

  const int ai = ais[objs[0] <= objs[1] ? 0 : 1];

  info = 1;
  int i0 = 0;
  switch (ai) {
  case 0: case 2:
    x[1] = ai == 0 ? xlo[1] : xhi[1];
    i0 = 1;
    break;
  case 1: case 3:
    x[0] = ai == 1 ? xhi[0] : xlo[0];
    i0 = 0;
    break;
  default: info = -2;
  }
  const int i1 = (i0 + 1) % 2;
  x[i1] = (b - a[i0]*x[i0])/a[i1];
  if (clip) x[i1] = dmin(xhi[i1], dmax(xlo[i1], x[i1]));
  return info;
}

/* cedr_local_inl.hpp:167-270 */
int oracle_solve_1eq_bc_qp(int n, const double* w, const double* a, double b,
                           const double* xlo, const double* xhi, const double* y,
                           double* x, int max_its) {
  const double r_tol = calc_r_tol(b, a, y, n);
  int info = check_lu(n, a, b, xlo, xhi, r_tol, x);
  if (info != 0) return info;

  for (int i = 0; i < n; ++i)
    if (x[i] != y[i]) {
      info = 1;
      x[i] = y[i];
    }

  const double wall_dist = 1e-3;

  double lamlo = 0, lamhi = 0;
  for (int i = 0; i < n; ++i) {
    const double rq = w[i]/a[i];
    const double lamlo_i = rq*(xlo[i] - y[i]);
    const double lamhi_i = rq*(xhi[i] - y[i]);
    if (i == 0) {
      lamlo = lamlo_i;
      lamhi = lamhi_i;
    } else {
      lamlo = dmin(lamlo, lamlo_i);
      lamhi = dmax(lamhi, lamhi_i);
    }
  }
  const double lamlo_feas = lamlo, lamhi_feas = lamhi;
  double lambda = lamlo <= 0 && lamhi >= 0 ? 0 : lamlo;

  int prev_step_bisect = 0;
  int nbisect = 0;
  info = -2;
  for (int iteration = 0; iteration < max_its; ++iteration) {
    double r, r_lambda;
    calc_r(n, w, a, b, xlo, xhi, y, lambda, x, &r, &r_lambda);
    if (fabs(r) <= r_tol) {
      info = 1;
      break;
    }
    if (nbisect > 64) {
      if (lamhi == lamhi_feas || lamlo == lamlo_feas) {
        info = -1;
        break;
      }
      info = 1;
      break;
    }
    if (r > 0)
      lamhi = lambda;
    else
      lamlo = lambda;
    if (r_lambda != 0) {
      lambda -= r/r_lambda;
    } else {
      lambda = lamlo;
    }
    const double D = prev_step_bisect ? 0 : wall_dist*(lamhi - lamlo);
    if (lambda - lamlo < D || lamhi - lambda < D) {
      lambda = 0.5*(lamlo + lamhi);
      ++nbisect;
      prev_step_bisect = 1;
    } else {
      prev_step_bisect = 0;
    }
  }
  return info;
}

/* cedr_local_inl.hpp:272-305 */
void oracle_local_caas(int n, const double* a, double b, const double* xlo,
                       const double* xhi, const double* y, double* x, int clip) {
  double dm = b;
  for (int i = 0; i < n; ++i) {
    x[i] = dmax(xlo[i], dmin(xhi[i], y[i]));
    dm -= a[i]*x[i];
  }
  if (dm == 0) return;
  if (dm > 0) {
    double fac = 0;
    for (int i = 0; i < n; ++i) fac += a[i]*(xhi[i] - x[i]);
    if (fac > 0) {
      fac = dm/fac;
      for (int i = 0; i < n; ++i) x[i] += fac*(xhi[i] - x[i]);
    }
  } else if (dm < 0) {
    double fac = 0;
    for (int i = 0; i < n; ++i) fac += a[i]*(x[i] - xlo[i]);
    if (fac > 0) {
      fac = dm/fac;
      for (int i = 0; i < n; ++i) x[i] += fac*(x[i] - xlo[i]);
    }
  }
  if (clip)
    for (int i = 0; i < n; ++i) x[i] = dmax(xlo[i], dmin(xhi[i], x[i]));
}

/* cedr_local_inl.hpp:307-330 */
int oracle_solve_1eq_nonneg(int n, const double* a, double b, const double* y,
                            double* x, const double* w, int method) {
  if (n > 16) return -3;
  if (b < 0) return -1;
  const double zero[16] = {0};
  double xhi[16];
  for (int i = 0; i < n; ++i) xhi[i] = b/a[i];
  if (method == 1) {
    oracle_local_caas(n, a, b, zero, xhi, y, x, 1);
    return 1;
  }
  if (n == 2) return oracle_solve_1eq_bc_qp_2d(w, a, b, zero, xhi, y, x, 1, 1);
  return oracle_solve_1eq_bc_qp(n, w, a, b, zero, xhi, y, x, 100);
}

/* ------------------------------------------------------- QLT node problem */

/* cedr_qlt_inl.hpp:69-99 */
static void r2l_nl_adjust_bounds(double Qm_bnd[2], const double rhom[2],
                                 double Qm_extra) {
  double q[2];
  for (int i = 0; i < 2; ++i) q[i] = Qm_bnd[i]/rhom[i];
  if (Qm_extra < 0) {
    int i0, i1;
    if (q[0] >= q[1]) { i0 = 0; i1 = 1; } else { i0 = 1; i1 = 0; }
    const double Qm_gap = (q[i1] - q[i0])*rhom[i0];
    if (Qm_gap <= Qm_extra) {
      Qm_bnd[i0] += Qm_extra;
      return;
    }
  } else {
    int i0, i1;
    if (q[0] <= q[1]) { i0 = 0; i1 = 1; } else { i0 = 1; i1 = 0; }
    const double Qm_gap = (q[i1] - q[i0])*rhom[i0];
    if (Qm_gap >= Qm_extra) {
      Qm_bnd[i0] += Qm_extra;
      return;
    }
  }
  {
    const double Qm_tot = Qm_bnd[0] + Qm_bnd[1] + Qm_extra;
    const double rhom_tot = rhom[0] + rhom[1];
    const double q_tot = Qm_tot/rhom_tot;
    for (int i = 0; i < 2; ++i) Qm_bnd[i] = q_tot*rhom[i];
  }
}

/* cedr_qlt_inl.hpp:119-173 */
static void solve_node_problem_generic(double rhom, const double* pd, double Qm,
                                       double rhom0, const double* k0d, double* Qm0,
                                       double rhom1, const double* k1d, double* Qm1,
                                       int prefer_mass_con) {
  (void) rhom;
  double Qm_min_kids[2] = {k0d[0], k1d[0]};
  double Qm_orig_kids[2] = {k0d[1], k1d[1]};
  double Qm_max_kids[2] = {k0d[2], k1d[2]};
  {
    const double Qm_min = pd[0], Qm_max = pd[2];
    const int lo = Qm < Qm_min, hi = Qm > Qm_max;
    if (lo || hi) {
      const double tol = 10*DBL_EPSILON;
      const double discrepancy = lo ? Qm_min - Qm : Qm - Qm_max;
      if (discrepancy > tol*(Qm_max - Qm_min)) {
        const double rhom_kids[2] = {rhom0, rhom1};
        r2l_nl_adjust_bounds(lo ? Qm_min_kids : Qm_max_kids, rhom_kids,
                             Qm - (lo ? Qm_min : Qm_max));
      }
    } else {
      if (Qm == pd[1] &&
          Qm_orig_kids[0] >= Qm_min_kids[0] && Qm_orig_kids[0] <= Qm_max_kids[0] &&
          Qm_orig_kids[1] >= Qm_min_kids[1] && Qm_orig_kids[1] <= Qm_max_kids[1]) {
        *Qm0 = Qm_orig_kids[0];
        *Qm1 = Qm_orig_kids[1];
        return;
      }
    }
  }
  {
    static const double ones[2] = {1, 1};
    const double w[2] = {1/rhom0, 1/rhom1};
    double Qm_kids[2] = {k0d[1], k1d[1]};
    oracle_solve_1eq_bc_qp_2d(w, ones, Qm, Qm_min_kids, Qm_max_kids, Qm_orig_kids,
                              Qm_kids, !prefer_mass_con, !prefer_mass_con);
    *Qm0 = Qm_kids[0];
    *Qm1 = Qm_kids[1];
  }
}

/* cedr_qlt_inl.hpp:175-203 */
void oracle_solve_node_problem(int problem_type, double rhom, const double* pd,
                               double Qm, double rhom0, const double* k0d,
                               double* Qm0, double rhom1, const double* k1d,
                               double* Qm1, int prefer_mass_con) {
  if ((problem_type & PT_T) && !(problem_type & PT_S)) {
    double mpd[3], mk0d[3], mk1d[3];
    mpd[0] = pd[0]*rhom;    mpd[1] = pd[1];   mpd[2] = pd[2]*rhom;
    mk0d[0] = k0d[0]*rhom0; mk0d[1] = k0d[1]; mk0d[2] = k0d[2]*rhom0;
    mk1d[0] = k1d[0]*rhom1; mk1d[1] = k1d[1]; mk1d[2] = k1d[2]*rhom1;
    solve_node_problem_generic(rhom, mpd, Qm, rhom0, mk0d, Qm0, rhom1, mk1d, Qm1,
                               prefer_mass_con);
  } else if (problem_type & PT_N) {
    static const double ones[2] = {1, 1};
    const double w[2] = {1/rhom0, 1/rhom1};
    double Qm_orig_kids[2] = {k0d[0], k1d[0]};
    double Qm_kids[2] = {k0d[0], k1d[0]};
    oracle_solve_1eq_nonneg(2, ones, Qm, Qm_orig_kids, Qm_kids, w, 0);
    *Qm0 = Qm_kids[0];
    *Qm1 = Qm_kids[1];
  } else {
    solve_node_problem_generic(rhom, pd, Qm, rhom0, k0d, Qm0, rhom1, k1d, Qm1,
                               prefer_mass_con);
  }
}

/* ------------------------------------------------------------------- tree */

/* cedr_tree.cpp:391-413: cn0 = cn/2, or cn/3 if imbalanced and cn > 2. Nodes
 * are emitted in pre-order, so the root is node 0. */
static int bisect(int cs, int ce, int imbalanced, int* kids, int64_t* cellidx,
                  int* next) {
  const int me = (*next)++;
  const int cn = ce - cs;
  if (cn == 1) {
    kids[2*me] = kids[2*me+1] = -1;
    cellidx[me] = cs;
    return me;
  }
  const int cn0 = (imbalanced && cn > 2) ? cn/3 : cn/2;
  cellidx[me] = -1;
  const int k0 = bisect(cs, cs + cn0, imbalanced, kids, cellidx, next);
  const int k1 = bisect(cs + cn0, ce, imbalanced, kids, cellidx, next);
  kids[2*me] = k0;
  kids[2*me+1] = k1;
  return me;
}

int oracle_make_bisection_tree(int ncells, int imbalanced, int* kids,
                               int64_t* cellidx) {
  if (ncells < 1) return -1;
  int next = 0;
  return bisect(0, ncells, imbalanced, kids, cellidx, &next);
}

/* Level schedule of tree::analyze on one rank: level = 1 + max(kid levels),
 * leaves at level 0 (cedr_tree.cpp:55-70); nodes enter their level in DFS
 * post-order (cedr_tree.cpp:85), and leaf slots are numbered in that order
 * (cedr_tree.cpp:148-180), which is what makes lci the DFS leaf order. */
typedef struct {
  int nnodes, nleaves, nlevels;
  int* level;    /* per node */
  int* order;    /* nodes grouped by level, post-order within a level */
  int* lvlptr;   /* nlevels+1 */
  int* lci;      /* per node; -1 for internal */
} Sched;

static int sched_dfs(const int* kids, int node, Sched* s, int* post, int* npost) {
  int lvl = 0;
  if (kids[2*node] >= 0) {
    const int l0 = sched_dfs(kids, kids[2*node], s, post, npost);
    const int l1 = sched_dfs(kids, kids[2*node+1], s, post, npost);
    lvl = 1 + (l0 > l1 ? l0 : l1);
  } else {
    s->lci[node] = s->nleaves++;
  }
  s->level[node] = lvl;
  post[(*npost)++] = node;
  return lvl;
}

static int sched_init(Sched* s, int nnodes, int root, const int* kids) {
  memset(s, 0, sizeof(*s));
  s->nnodes = nnodes;
  s->level = (int*) malloc(sizeof(int)*nnodes);
  s->order = (int*) malloc(sizeof(int)*nnodes);
  s->lci = (int*) malloc(sizeof(int)*nnodes);
  int* post = (int*) malloc(sizeof(int)*nnodes);
  if (!s->level || !s->order || !s->lci || !post) return 1;
  for (int i = 0; i < nnodes; ++i) s->lci[i] = -1;
  int npost = 0;
  const int h = sched_dfs(kids, root, s, post, &npost);
  s->nlevels = h + 1;
  s->lvlptr = (int*) calloc(s->nlevels + 1, sizeof(int));
  if (!s->lvlptr) return 1;
  for (int i = 0; i < npost; ++i) ++s->lvlptr[s->level[post[i]] + 1];
  for (int l = 0; l < s->nlevels; ++l) s->lvlptr[l+1] += s->lvlptr[l];
  int* fill = (int*) malloc(sizeof(int)*s->nlevels);
  for (int l = 0; l < s->nlevels; ++l) fill[l] = s->lvlptr[l];
  for (int i = 0; i < npost; ++i) s->order[fill[s->level[post[i]]]++] = post[i];
  free(fill);
  free(post);
  return 0;
}

static void sched_free(Sched* s) {
  free(s->level); free(s->order); free(s->lvlptr); free(s->lci);
}

int oracle_leaf_order(int nnodes, int root, const int* kids, const int64_t* cellidx,
                      int64_t* lci2gci, int* nlevels) {
  Sched s;
  if (sched_init(&s, nnodes, root, kids)) return 1;
  for (int i = 0; i < nnodes; ++i)
    if (s.lci[i] >= 0) lci2gci[s.lci[i]] = cellidx[i];
  if (nlevels) *nlevels = s.nlevels;
  sched_free(&s);
  return 0;
}

/* -------------------------------------------------------------------- QLT */

/* cedr_qlt.cpp:85-96 + cedr_qlt_inl.hpp:101-108 */
int oracle_qlt_canonical_problem_type(int mask) {
  switch (mask) {
  case PT_S: case PT_S | PT_T: return PT_S | PT_T;
  case PT_C | PT_S: case PT_C | PT_S | PT_T: return PT_C | PT_S | PT_T;
  case PT_T: return PT_T;
  case PT_C | PT_T: return PT_C | PT_T;
  case PT_N: return PT_N;
  case PT_C | PT_N: return PT_C | PT_N;
  default: return -1;
  }
}

/* One tracer through QLT::run. l2r: 4 words per node, r2l: 3 words per node
 * (the reference packs them per problem type, cedr_qlt.cpp:104-157; the
 * arithmetic does not depend on the packing). */
static void qlt_one_tracer(const Sched* s, int root, const int* kids,
                           const int64_t* cellidx, int pt, int prefer_mass_con,
                           const double* node_rhom, const double* qm_min,
                           const double* qm, const double* qm_max,
                           const double* qm_prev, double* qm_out, double* l2r,
                           double* r2l) {
  const int nonneg = pt & PT_N, shape = pt & PT_S, conserve = pt & PT_C;
  const int consistent_only = (pt & PT_T) && !shape;
  /* DeviceOp::set_Qm, cedr_qlt_inl.hpp:21-58 */
  for (int i = 0; i < s->nnodes; ++i) {
    if (s->lci[i] < 0) continue;
    const int64_t g = cellidx[i];
    double* bd = l2r + 4*i;
    int next;
    if (shape) {
      bd[0] = qm_min[g]; bd[1] = qm[g]; bd[2] = qm_max[g]; next = 3;
    } else if (pt & PT_T) {
      const double rhom = node_rhom[i];
      bd[0] = qm_min[g]/rhom; bd[1] = qm[g]; bd[2] = qm_max[g]/rhom; next = 3;
    } else {
      bd[0] = qm[g]; next = 1;
    }
    if (conserve) bd[next] = qm_prev[g];
  }
  /* l2r_combine_kid_data, cedr_qlt.cpp:339-430 */
  for (int l = 1; l < s->nlevels; ++l)
    for (int j = s->lvlptr[l]; j < s->lvlptr[l+1]; ++j) {
      const int n = s->order[j];
      double* me = l2r + 4*n;
      const double* k0 = l2r + 4*kids[2*n];
      const double* k1 = l2r + 4*kids[2*n+1];
      if (nonneg) {
        me[0] = k0[0] + k1[0];
        if (conserve) me[1] = k0[1] + k1[1];
      } else {
        me[0] = shape ? k0[0] + k1[0] : dmin(k0[0], k1[0]);
        me[1] = k0[1] + k1[1];
        me[2] = shape ? k0[2] + k1[2] : dmax(k0[2], k1[2]);
        if (conserve) me[3] = k0[3] + k1[3];
      }
    }
  /* root_compute, cedr_qlt.cpp:441-476 */
  {
    const int l2rsz = nonneg ? (conserve ? 2 : 1) : (conserve ? 4 : 3);
    const int os = conserve ? l2rsz - 1 : (nonneg ? 0 : 1);
    r2l[3*root] = l2r[4*root + os];
    if (consistent_only) {
      r2l[3*root + 1] = l2r[4*root + 0];
      r2l[3*root + 2] = l2r[4*root + 2];
    }
  }
  /* r2l_solve_qp, cedr_qlt.cpp:525-604 */
  for (int l = s->nlevels - 1; l >= 1; --l)
    for (int j = s->lvlptr[l]; j < s->lvlptr[l+1]; ++j) {
      const int n = s->order[j];
      const int k0 = kids[2*n], k1 = kids[2*n+1];
      if (consistent_only) {
        const double q_min = r2l[3*n + 1], q_max = r2l[3*n + 2];
        l2r[4*n + 0] = q_min; l2r[4*n + 2] = q_max;
        l2r[4*k0 + 0] = q_min; l2r[4*k0 + 2] = q_max;
        r2l[3*k0 + 1] = q_min; r2l[3*k0 + 2] = q_max;
        l2r[4*k1 + 0] = q_min; l2r[4*k1 + 2] = q_max;
        r2l[3*k1 + 1] = q_min; r2l[3*k1 + 2] = q_max;
      }
      oracle_solve_node_problem(pt, node_rhom[n], l2r + 4*n, r2l[3*n],
                                node_rhom[k0], l2r + 4*k0, &r2l[3*k0],
                                node_rhom[k1], l2r + 4*k1, &r2l[3*k1],
                                prefer_mass_con);
    }
  /* DeviceOp::get_Qm, cedr_qlt_inl.hpp:60-66 */
  for (int i = 0; i < s->nnodes; ++i)
    if (s->lci[i] >= 0) qm_out[cellidx[i]] = r2l[3*i];
}

int oracle_qlt_run(int ncells, int nnodes, int root, const int* kids,
                   const int64_t* cellidx, int nt, const int* ptypes,
                   int prefer_mass_con, const double* rhom, const double* qm_min,
                   const double* qm, const double* qm_max, const double* qm_prev,
                   double* qm_out) {
  Sched s;
  if (sched_init(&s, nnodes, root, kids)) return 1;
  if (s.nleaves != ncells) { sched_free(&s); return 2; }
  for (int t = 0; t < nt; ++t)
    if (oracle_qlt_canonical_problem_type(ptypes[t]) < 0) { sched_free(&s); return 3; }
  /* rhom: word 0 of every slot, summed kid0 + kid1 (cedr_qlt.cpp:356-360). */
  double* node_rhom = (double*) malloc(sizeof(double)*nnodes);
  for (int i = 0; i < nnodes; ++i)
    if (s.lci[i] >= 0) node_rhom[i] = rhom[cellidx[i]];
  for (int l = 1; l < s.nlevels; ++l)
    for (int j = s.lvlptr[l]; j < s.lvlptr[l+1]; ++j) {
      const int n = s.order[j];
      node_rhom[n] = node_rhom[kids[2*n]] + node_rhom[kids[2*n+1]];
    }
  int err = 0;
#ifdef _OPENMP
# pragma omp parallel
#endif
  {
    double* l2r = (double*) malloc(sizeof(double)*4*nnodes);
    double* r2l = (double*) malloc(sizeof(double)*3*nnodes);
    if (!l2r || !r2l) {
#ifdef _OPENMP
#     pragma omp atomic write
#endif
      err = 1;
    } else {
#ifdef _OPENMP
#     pragma omp for schedule(static)
#endif
      for (int t = 0; t < nt; ++t) {
        const size_t os = (size_t) t*ncells;
        const int pt = oracle_qlt_canonical_problem_type(ptypes[t]);
        qlt_one_tracer(&s, root, kids, cellidx, pt, prefer_mass_con, node_rhom,
                       qm_min + os, qm + os, qm_max + os, qm_prev + os,
                       qm_out + os, l2r, r2l);
      }
    }
    free(l2r);
    free(r2l);
  }
  free(node_rhom);
  sched_free(&s);
  return err;
}

/* -------------------------------------------------------------------- BFB */

/* Tree-ordered sum of one field over the leaves, exactly as
 * BfbTreeAllReducer::allreduce accumulates it (cedr_bfb_tree_allreduce.cpp:
 * 86-124): leaf value copied; internal node d = 0; d += kid0; d += kid1. `val`
 * is indexed by lci. `wrk` has nnodes entries. */
static double bfb_sum(const Sched* s, int root, const int* kids, const double* val,
                      double* wrk) {
  for (int i = 0; i < s->nnodes; ++i)
    if (s->lci[i] >= 0) wrk[i] = val[s->lci[i]];
  for (int l = 1; l < s->nlevels; ++l)
    for (int j = s->lvlptr[l]; j < s->lvlptr[l+1]; ++j) {
      const int n = s->order[j];
      double d = 0;
      d += wrk[kids[2*n]];
      d += wrk[kids[2*n+1]];
      wrk[n] = d;
    }
  return wrk[root];
}

int oracle_bfb_allreduce(int nleaf, int nnodes, int root, const int* kids,
                         const int64_t* cellidx, int nfield, int transpose,
                         const double* send, double* recv) {
  (void) cellidx;
  Sched s;
  if (sched_init(&s, nnodes, root, kids)) return 1;
  if (s.nleaves != nleaf) { sched_free(&s); return 2; }
  double* val = (double*) malloc(sizeof(double)*nleaf);
  double* wrk = (double*) malloc(sizeof(double)*nnodes);
  for (int j = 0; j < nfield; ++j) {
    for (int i = 0; i < nleaf; ++i)
      val[i] = transpose ? send[(size_t) nleaf*j + i] : send[(size_t) nfield*i + j];
    recv[j] = bfb_sum(&s, root, kids, val, wrk);
  }
  free(val); free(wrk);
  sched_free(&s);
  return 0;
}

/* ------------------------------------------------------------------- CAAS */

int oracle_caas_run(int ncells, int reducer, int nnodes, int root, const int* kids,
                    const int64_t* cellidx, int nt, const int* ptypes,
                    const double* qm_min, const double* qm, const double* qm_max,
                    const double* qm_prev, double* qm_out) {
  Sched s;
  (void) cellidx;
  for (int t = 0; t < nt; ++t)
    if (!(ptypes[t] & PT_S)) return 3; /* cedr_caas.cpp:52-53 */
  if (reducer == 1) {
    if (sched_init(&s, nnodes, root, kids)) return 1;
    if (s.nleaves != ncells) { sched_free(&s); return 2; }
    /* CAAS cell i feeds the reducer's i-th local leaf, i.e. lci == i
     * (cedr_bfb_tree_allreduce.cpp:87-97 indexes send by position in
     * levels[0].nodes), whatever that leaf's cellidx is. */
  }
  int err = 0;
#ifdef _OPENMP
# pragma omp parallel
#endif
  {
    double* val = 0; double* wrk = 0;
    if (reducer == 1) {
      val = (double*) malloc(sizeof(double)*4*ncells);
      wrk = (double*) malloc(sizeof(double)*nnodes);
    }
#ifdef _OPENMP
#   pragma omp for schedule(static)
#endif
    for (int t = 0; t < nt; ++t) {
      const size_t os = (size_t) t*ncells;
      const double* lo = qm_min + os, * hi = qm_max + os, * q = qm + os;
      const double* pv = qm_prev + os;
      double* x = qm_out + os;
      const int conserve = ptypes[t] & PT_C;
      double sum[4];
      if (reducer == 0) {
        /* reduce_locally, cedr_caas.cpp:171-199: team size 1 => sequential. */
        double accum_clip = 0, accum_term = 0, accum_min = 0, accum_max = 0;
        for (int i = 0; i < ncells; ++i) {
          /* calc_Qm_scalars, cedr_caas_inl.hpp:44-57 */
          const double Qm = q[i];
          const double Qm_term = conserve ? pv[i] : Qm;
          const double Qm_clip = dmin(hi[i], dmax(lo[i], Qm));
          x[i] = Qm_clip;
          accum_clip += Qm_clip;
          accum_term += Qm_term;
        }
        for (int i = 0; i < ncells; ++i) accum_min += lo[i];
        for (int i = 0; i < ncells; ++i) accum_max += hi[i];
        sum[0] = accum_clip; sum[1] = accum_term; sum[2] = accum_min; sum[3] = accum_max;
      } else {
        /* user-reducer branch, cedr_caas.cpp:140-168 with n_accum_in_place 1:
         * each send entry is 0 + value; then the tree-ordered reduction. */
        for (int i = 0; i < ncells; ++i) {
          const double Qm = q[i];
          const double Qm_term = conserve ? pv[i] : Qm;
          const double Qm_clip = dmin(hi[i], dmax(lo[i], Qm));
          x[i] = Qm_clip;
          const int l = i;
          double a;
          a = 0; a += Qm_clip; val[l] = a;
          a = 0; a += Qm_term; val[ncells + l] = a;
          a = 0; a += lo[i];   val[2*ncells + l] = a;
          a = 0; a += hi[i];   val[3*ncells + l] = a;
        }
        for (int f = 0; f < 4; ++f)
          sum[f] = bfb_sum(&s, root, kids, val + (size_t) f*ncells, wrk);
      }
      /* finish_locally, cedr_caas.cpp:211-253 */
      const double Qm_clip_sum = sum[0], Qm_sum = sum[1];
      const double m = Qm_sum - Qm_clip_sum;
      if (m < 0) {
        const double Qm_min_sum = sum[2];
        double fac = Qm_clip_sum - Qm_min_sum;
        if (fac > 0) {
          fac = m/fac;
          for (int i = 0; i < ncells; ++i) {
            double Qm = x[i];
            Qm += fac*(Qm - lo[i]);
            x[i] = dmax(lo[i], Qm);
          }
        }
      } else if (m > 0) {
        const double Qm_max_sum = sum[3];
        double fac = Qm_max_sum - Qm_clip_sum;
        if (fac > 0) {
          fac = m/fac;
          for (int i = 0; i < ncells; ++i) {
            double Qm = x[i];
            Qm += fac*(hi[i] - Qm);
            x[i] = dmin(hi[i], Qm);
          }
        }
      }
    }
    free(val); free(wrk);
  }
  if (reducer == 1) sched_free(&s);
  return err;
}

/* --------------------------------------------------------------- workload */

/* The synthetic inputs of SURVEY.md 8(d) (same stream as
 * compose_b200/workloads.py::headline), for tracers [t0, t1). */
static double splitmix_u(uint64_t seed, uint64_t k) {
  uint64_t z = seed + (k + 1u)*0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30))*0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27))*0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double) (z >> 11)*0x1.0p-53;
}

void oracle_fill_headline(int ncells, int config_id, int t0, int t1, double* rhom,
                          double* qm_min, double* qm, double* qm_max, double* qm_prev) {
  const uint64_t seed = 0xCED20000ull + (uint64_t) config_id;
  for (int i = 0; i < ncells; ++i) rhom[i] = 0.5*(1 + splitmix_u(seed, (uint64_t) i));
#ifdef _OPENMP
# pragma omp parallel for schedule(static)
#endif
  for (int t = t0; t < t1; ++t)
    for (int i = 0; i < ncells; ++i) {
      const uint64_t p = (uint64_t) ncells + 4ull*((uint64_t) t*ncells + i);
      const double q_min = 0.1*splitmix_u(seed, p);
      const double q_max = q_min + splitmix_u(seed, p + 1);
      const double q = q_min + (q_max - q_min)*(1.4*splitmix_u(seed, p + 2) - 0.2);
      const double q_prev = q_min + (q_max - q_min)*splitmix_u(seed, p + 3);
      const size_t s = (size_t) (t - t0)*ncells + i;
      qm_min[s] = q_min*rhom[i];
      qm_max[s] = q_max*rhom[i];
      qm[s] = q*rhom[i];
      qm_prev[s] = q_prev*rhom[i];
    }
}
