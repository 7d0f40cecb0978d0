// cedr_b200_local.hpp -- the element-local solvers of COMPOSE's cedr::local
// (cedr/cedr_local.hpp:23-58, cedr_local_inl.hpp) as __host__ __device__ inline functions:
// what a HOMME-style caller runs per element right after CDR::run.
//
//   min_x sum_i w_i (x_i - y_i)^2   s.t.  a'x = b,  xlo <= x <= xhi,   a, w > 0
//
// Same names, arguments and return codes as the reference. The operation order follows
// the reference's exactly, so a caller compiled WITHOUT floating-point contraction
// (nvcc -fmad=false; host: -ffp-contract=off) gets the reference's bits;
// tests/cxx/test_cdr_mirror.cu checks that against the pinned oracle on the device.
#ifndef CEDR_B200_LOCAL_HPP
#define CEDR_B200_LOCAL_HPP

#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
# define CEDR_B200_LOCAL_HD __host__ __device__ inline
#else
# define CEDR_B200_LOCAL_HD inline
#endif

namespace cedr {
namespace local {
typedef int Int;
typedef double Real;

namespace impl {
// cedr_kokkos.hpp:136-139: ternaries, not fmin/fmax.
CEDR_B200_LOCAL_HD Real min (const Real a, const Real b) { return a < b ? a : b; }
CEDR_B200_LOCAL_HD Real max (const Real a, const Real b) { return a > b ? a : b; }

// cedr_local_inl.hpp:13-18
CEDR_B200_LOCAL_HD Real calc_r_tol (const Real b, const Real* a, const Real* y, const Int n) {
  Real ab = std::fabs(b);
  for (Int i = 0; i < n; ++i) ab = max(ab, std::fabs(a[i]*y[i]));
  return 1e1*DBL_EPSILON*std::fabs(ab);
}

// cedr_local_inl.hpp:23-41. 1: a corner solves the problem; -1: infeasible, x left at
// the violated corner; 0: neither.
CEDR_B200_LOCAL_HD Int check_lu (const Int n, const Real* a, const Real b, const Real* xlo,
                                 const Real* xhi, const Real r_tol, Real* x) {
  Real r = -b;
  for (Int i = 0; i < n; ++i) { x[i] = xlo[i]; r += a[i]*x[i]; }
  if (std::fabs(r) <= r_tol) return 1;
  if (r > 0) return -1;
  r = -b;
  for (Int i = 0; i < n; ++i) { x[i] = xhi[i]; r += a[i]*x[i]; }
  if (std::fabs(r) <= r_tol) return 1;
  if (r < 0) return -1;
  return 0;
}

// cedr_local_inl.hpp:43-64: x(lambda) clipped to the bounds, the constraint residual and
// its derivative.
CEDR_B200_LOCAL_HD void calc_r (const Int n, const Real* w, const Real* a, const Real b,
                                const Real* xlo, const Real* xhi, const Real* y,
                                const Real lambda, Real* x, Real& r, Real& r_lambda) {
  r = 0;
  r_lambda = 0;
  for (Int i = 0; i < n; ++i) {
    const Real q = a[i]/w[i];
    const Real x_trial = y[i] + lambda*q;
    if (x_trial < xlo[i]) x[i] = xlo[i];
    else if (x_trial > xhi[i]) x[i] = xhi[i];
    else {
      x[i] = x_trial;
      r_lambda += a[i]*q;
    }
    r += a[i]*x[i];
  }
  r -= b;
}
} // namespace impl

// cedr_local_inl.hpp:167-270: safeguarded Newton on the dual variable. Returns 0 if
// x == y solves it, 1 if solved with x != y, -1 if infeasible, -2 if max_its was hit.
CEDR_B200_LOCAL_HD Int
solve_1eq_bc_qp (const Int n, const Real* w, const Real* a, const Real b, const Real* xlo,
                 const Real* xhi, const Real* y, Real* x, const Int max_its = 100) {
  const Real r_tol = impl::calc_r_tol(b, a, y, n);
  Int info = impl::check_lu(n, a, b, xlo, xhi, r_tol, x);
  if (info != 0) return info;
  for (Int i = 0; i < n; ++i)
    if (x[i] != y[i]) { info = 1; x[i] = y[i]; }
  const Real wall_dist = 1e-3;
  // Bracket of the dual variable within which some x_i is strictly inside its bounds.
  Real lamlo = 0, lamhi = 0;
  for (Int i = 0; i < n; ++i) {
    const Real rq = w[i]/a[i];
    const Real lamlo_i = rq*(xlo[i] - y[i]), lamhi_i = rq*(xhi[i] - y[i]);
    if (i == 0) { lamlo = lamlo_i; lamhi = lamhi_i; }
    else { lamlo = impl::min(lamlo, lamlo_i); lamhi = impl::max(lamhi, lamhi_i); }
  }
  const Real lamlo_feas = lamlo, lamhi_feas = lamhi;
  Real lambda = lamlo <= 0 && lamhi >= 0 ? 0 : lamlo;
  bool prev_step_bisect = false;
  Int nbisect = 0;
  info = -2;
  for (Int iteration = 0; iteration < max_its; ++iteration) {
    Real r, r_lambda;
    impl::calc_r(n, w, a, b, xlo, xhi, y, lambda, x, r, r_lambda);
    if (std::fabs(r) <= r_tol) { info = 1; break; }
    if (nbisect > 64) {
      // Bisection has run out of precision: infeasible only if a bracket end never moved.
      info = (lamhi == lamhi_feas || lamlo == lamlo_feas) ? -1 : 1;
      break;
    }
    if (r > 0) lamhi = lambda; else lamlo = lambda;
    if (r_lambda != 0) lambda -= r/r_lambda; else lambda = lamlo;
    const Real D = prev_step_bisect ? 0 : wall_dist*(lamhi - lamlo);
    if (lambda - lamlo < D || lamhi - lambda < D) {
      lambda = 0.5*(lamlo + lamhi);
      ++nbisect;
      prev_step_bisect = true;
    } else {
      prev_step_bisect = false;
    }
  }
  return info;
}

// cedr_local_inl.hpp:68-165: the closed-form 2-variable case. With early_exit_on_tol only
// infeasibility returns early (the reference's inner `info` shadows the outer one).
CEDR_B200_LOCAL_HD Int
solve_1eq_bc_qp_2d (const Real* w, const Real* a, const Real b, const Real* xlo,
                    const Real* xhi, const Real* y, Real* x, const bool clip = true,
                    const bool early_exit_on_tol = true) {
  if (early_exit_on_tol) {
    const Real r_tol = impl::calc_r_tol(b, a, y, 2);
    if (impl::check_lu(2, a, b, xlo, xhi, r_tol, x) == -1) return -1;
  }
  { // Unconstrained optimum.
    Real qmass = 0, dm = b;
    for (Int i = 0; i < 2; ++i) {
      const Real qi = a[i]/w[i];
      qmass += a[i]*qi;
      dm -= a[i]*y[i];
    }
    const Real lambda = dm/qmass;
    bool ok = true;
    for (Int i = 0; i < 2; ++i) {
      x[i] = y[i] + lambda*(a[i]/w[i]);
      if (x[i] < xlo[i] || x[i] > xhi[i]) { ok = false; break; }
    }
    if (ok) return 1;
  }
  // The line a'x = b cuts the four bound lines at alphas[]; the feasible segment lies
  // between the two cuts that are neither the first minimum nor the first maximum.
  Real x_base[2];
  for (Int i = 0; i < 2; ++i) x_base[i] = 0.5*b/a[i];
  const Real x_dir[2] = {-a[1], a[0]};
  Real alphas[4];
  alphas[0] = (xlo[1] - x_base[1])/x_dir[1];   // bottom
  alphas[1] = (xhi[0] - x_base[0])/x_dir[0];   // right
  alphas[2] = (xhi[1] - x_base[1])/x_dir[1];   // top
  alphas[3] = (xlo[0] - x_base[0])/x_dir[0];   // left
  Real mn = alphas[0], mx = alphas[0];
  Int imin = 0, imax = 0;
  for (Int i = 1; i < 4; ++i) {
    if (alphas[i] < mn) { mn = alphas[i]; imin = i; }
    if (alphas[i] > mx) { mx = alphas[i]; imax = i; }
  }
  Int ais[2] = {0, 0}, cnt = 0;
  for (Int i = 0; i < 4 && cnt < 2; ++i)
    if (i != imin && i != imax) ais[cnt++] = i;
  Real objs[2];
  for (Int j = 0; j < 2; ++j) {
    const Real alpha = alphas[ais[j]];
    Real obj = 0;
    for (Int i = 0; i < 2; ++i) {
      x[i] = x_base[i] + alpha*x_dir[i];
      const Real d = y[i] - x[i];
      obj += w[i]*(d*d);
    }
    objs[j] = obj;
  }
  const Int ai = ais[objs[0] <= objs[1] ? 0 : 1];
  // Pin the coordinate whose bound line was chosen, solve the other from the constraint.
  Int i0;
  if (ai == 0 || ai == 2) { x[1] = ai == 0 ? xlo[1] : xhi[1]; i0 = 1; }
  else { x[0] = ai == 1 ? xhi[0] : xlo[0]; i0 = 0; }
  const Int i1 = (i0 + 1) % 2;
  x[i1] = (b - a[i0]*x[i0])/a[i1];
  if (clip) x[i1] = impl::min(xhi[i1], impl::max(xlo[i1], x[i1]));
  return 1;
}

// cedr_local_inl.hpp:272-305: clip, then spread the mass defect over the remaining
// capacities. Does not check feasibility.
CEDR_B200_LOCAL_HD void
caas (const Int n, const Real* a, const Real b, const Real* xlo, const Real* xhi,
      const Real* y, Real* x, const bool clip = true) {
  Real dm = b;
  for (Int i = 0; i < n; ++i) {
    x[i] = impl::max(xlo[i], impl::min(xhi[i], y[i]));
    dm -= a[i]*x[i];
  }
  if (dm == 0) return;
  if (dm > 0) {
    Real fac = 0;
    for (Int i = 0; i < n; ++i) fac += a[i]*(xhi[i] - x[i]);
    if (fac > 0) {
      fac = dm/fac;
      for (Int i = 0; i < n; ++i) x[i] += fac*(xhi[i] - x[i]);
    }
  } else if (dm < 0) {
    Real fac = 0;
    for (Int i = 0; i < n; ++i) fac += a[i]*(x[i] - xlo[i]);
    if (fac > 0) {
      fac = dm/fac;
      for (Int i = 0; i < n; ++i) x[i] += fac*(x[i] - xlo[i]);
    }
  }
  if (clip)
    for (Int i = 0; i < n; ++i) x[i] = impl::max(xlo[i], impl::min(xhi[i], x[i]));
}

struct Method { enum Enum { least_squares, caas }; };

// cedr_local_inl.hpp:307-330: x >= 0 instead of two-sided bounds; n <= 16.
CEDR_B200_LOCAL_HD Int
solve_1eq_nonneg (const Int n, const Real* a, const Real b, const Real* y, Real* x,
                  const Real* w, const Method::Enum lcl_method) {
  if (n > 16) return -3;
  if (b < 0) return -1;
  Real zero[16], xhi[16];
  for (Int i = 0; i < n; ++i) { zero[i] = 0; xhi[i] = b/a[i]; }
  if (lcl_method == Method::caas) {
    caas(n, a, b, zero, xhi, y, x);
    return 1;
  }
  if (n == 2) return solve_1eq_bc_qp_2d(w, a, b, zero, xhi, y, x);
  return solve_1eq_bc_qp(n, w, a, b, zero, xhi, y, x);
}

} // namespace local
} // namespace cedr

#endif
