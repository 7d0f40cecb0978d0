// cedr_b200_local.hpp -- element-local solvers for the B200 CEDR path:
// what a HOMME-style caller runs per element right after CDR::run,
//
//   min_x sum_i w_i (x_i - y_i)^2   s.t.  a'x = b,  xlo <= x <= xhi,   a, w > 0.
//
// The algorithms are those of COMPOSE's cedr::local (cedr/cedr_local.hpp:23-58,
// cedr_local_inl.hpp; COMPOSE version 1.0, Copyright 2018 NTESS, BSD license -- see the
// reference's LICENSE): a safeguarded Newton iteration on the dual variable, the closed form
// for two variables, and clip-and-redistribute. Results are bit-identical to the reference
// when the caller compiles without floating-point contraction (nvcc -fmad=false; host
// -ffp-contract=off): every sum and product is formed in the reference's order.
//
// Organisation (not the reference's): the algorithms are templates over a STORAGE VIEW of
// one element's problem, so the same code runs
//   * on a `Registers<N>` view -- the element's n <= N values held by value, which on the
//     device means in registers, gathered and scattered with arbitrary strides; this is what
//     the batched kernels of the library use (cedr_b200_local_solve: one thread per
//     element, SoA arrays with the element index fastest for coalesced access);
//   * on a `Pointers` view -- the reference's calling convention (contiguous n-vectors), any n.
//
// The reference's names and signatures are kept as thin wrappers at the end of the file.
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

struct Method { enum Enum { least_squares, caas }; };

// ---------------------------------------------------------------- storage views

// The reference's calling convention: caller-owned contiguous vectors.
struct Pointers {
  Int n;
  const Real* w_;
  const Real* a_;
  const Real* lo_;
  const Real* hi_;
  const Real* y_;
  Real* x_;
  Real b;
  CEDR_B200_LOCAL_HD Int size () const { return n; }
  CEDR_B200_LOCAL_HD Real w (Int i) const { return w_[i]; }
  CEDR_B200_LOCAL_HD Real a (Int i) const { return a_[i]; }
  CEDR_B200_LOCAL_HD Real lo (Int i) const { return lo_[i]; }
  CEDR_B200_LOCAL_HD Real hi (Int i) const { return hi_[i]; }
  CEDR_B200_LOCAL_HD Real y (Int i) const { return y_[i]; }
  CEDR_B200_LOCAL_HD Real& x (Int i) { return x_[i]; }
};

// One element held by value: with N a compile-time constant and the loops below fully
// unrolled, the arrays live in registers on the device. Entries i >= n are ignored.
template <int N> struct Registers {
  Int n;
  Real w_[N], a_[N], lo_[N], hi_[N], y_[N], x_[N];
  Real b;
  CEDR_B200_LOCAL_HD Int size () const { return n; }
  CEDR_B200_LOCAL_HD Real w (Int i) const { return w_[i]; }
  CEDR_B200_LOCAL_HD Real a (Int i) const { return a_[i]; }
  CEDR_B200_LOCAL_HD Real lo (Int i) const { return lo_[i]; }
  CEDR_B200_LOCAL_HD Real hi (Int i) const { return hi_[i]; }
  CEDR_B200_LOCAL_HD Real y (Int i) const { return y_[i]; }
  CEDR_B200_LOCAL_HD Real& x (Int i) { return x_[i]; }
  // Entry i of the element at base[i*stride]. Null w / a mean all ones; null lo / hi are
  // left for the caller to fill (the nonnegative problem derives them).
  CEDR_B200_LOCAL_HD void gather (const Int n_, const Real* w, const Real* a, const Real b_,
                                  const Real* lo, const Real* hi, const Real* y,
                                  const long long stride) {
    n = n_;
    b = b_;
#pragma unroll
    for (Int i = 0; i < N; ++i) {
      if (i >= n) break;
      const long long o = i*stride;
      w_[i] = w ? w[o] : 1.0;
      a_[i] = a ? a[o] : 1.0;
      if (lo) lo_[i] = lo[o];
      if (hi) hi_[i] = hi[o];
      y_[i] = y[o];
    }
  }
  CEDR_B200_LOCAL_HD void scatter (Real* x, const long long stride) const {
#pragma unroll
    for (Int i = 0; i < N; ++i) {
      if (i >= n) break;
      x[i*stride] = x_[i];
    }
  }
};

// ---------------------------------------------------------------- the algorithms

namespace impl {
// cedr_kokkos.hpp:136-139: ternaries, not fmin / fmax (same NaN and signed-zero behaviour).
CEDR_B200_LOCAL_HD Real lesser (const Real p, const Real q) { return p < q ? p : q; }
CEDR_B200_LOCAL_HD Real greater (const Real p, const Real q) { return p > q ? p : q; }

// cedr_local_inl.hpp:13-18: 10 eps max(|b|, max_i |a_i y_i|).
template <typename V> CEDR_B200_LOCAL_HD Real residual_tolerance (const V& v) {
  Real big = std::fabs(v.b);
  for (Int i = 0; i < v.size(); ++i) big = greater(big, std::fabs(v.a(i)*v.y(i)));
  return 1e1*DBL_EPSILON*std::fabs(big);
}

// cedr_local_inl.hpp:23-41: the constraint residual with every x at its lower, then its
// upper bound. +1: that corner solves the problem; -1: infeasible (x stays at the violated
// corner); 0: neither. x is left at the last corner tried.
template <typename V> CEDR_B200_LOCAL_HD Int try_corners (V& v, const Real tol) {
  for (Int side = 0; side < 2; ++side) {
    Real r = -v.b;
    for (Int i = 0; i < v.size(); ++i) {
      v.x(i) = side == 0 ? v.lo(i) : v.hi(i);
      r += v.a(i)*v.x(i);
    }
    if (std::fabs(r) <= tol) return 1;
    if (side == 0 ? r > 0 : r < 0) return -1;
  }
  return 0;
}
} // namespace impl

// cedr_local_inl.hpp:167-270. Returns 1 if the problem was solved (also when a corner of the
// box solves it), -1 if it is infeasible, -2 if max_its iterations did not suffice.
//
// The dual function r(lambda) = a'x(lambda) - b with x_i(lambda) = clamp(y_i + lambda
// a_i/w_i) is piecewise linear and nondecreasing: Newton from inside the bracket
// [lamlo, lamhi] outside of which every x_i sits on a bound, falling back to bisection when
// a step leaves the bracket or (every other time) lands within 1e-3 of its ends.
template <typename V>
CEDR_B200_LOCAL_HD Int solve_qp (V& v, const Int max_its = 100) {
  const Int n = v.size();
  const Real tol = impl::residual_tolerance(v);
  Int info = impl::try_corners(v, tol);
  if (info != 0) return info;
  for (Int i = 0; i < n; ++i)
    if (v.x(i) != v.y(i)) { info = 1; v.x(i) = v.y(i); }

  Real lamlo = 0, lamhi = 0;
  for (Int i = 0; i < n; ++i) {
    const Real s = v.w(i)/v.a(i);
    const Real l = s*(v.lo(i) - v.y(i)), h = s*(v.hi(i) - v.y(i));
    lamlo = i == 0 ? l : impl::lesser(lamlo, l);
    lamhi = i == 0 ? h : impl::greater(lamhi, h);
  }
  const Real lamlo0 = lamlo, lamhi0 = lamhi;
  Real lambda = (lamlo <= 0 && lamhi >= 0) ? 0 : lamlo;

  bool bisected_last = false;
  Int nbisect = 0;
  for (Int it = 0; it < max_its; ++it) {
    // x(lambda), r(lambda) and its slope over the coordinates strictly inside their bounds
    // (cedr_local_inl.hpp:43-64).
    Real r = 0, slope = 0;
    for (Int i = 0; i < n; ++i) {
      const Real q = v.a(i)/v.w(i);
      const Real xt = v.y(i) + lambda*q;
      if (xt < v.lo(i)) v.x(i) = v.lo(i);
      else if (xt > v.hi(i)) v.x(i) = v.hi(i);
      else {
        v.x(i) = xt;
        slope += v.a(i)*q;
      }
      r += v.a(i)*v.x(i);
    }
    r -= v.b;
    if (std::fabs(r) <= tol) return 1;
    // Bisection has run out of bits: infeasible only if a bracket end never moved.
    if (nbisect > 64) return (lamhi == lamhi0 || lamlo == lamlo0) ? -1 : 1;
    if (r > 0) lamhi = lambda; else lamlo = lambda;
    lambda = slope != 0 ? lambda - r/slope : lamlo;
    const Real margin = bisected_last ? 0 : 1e-3*(lamhi - lamlo);
    bisected_last = lambda - lamlo < margin || lamhi - lambda < margin;
    if (bisected_last) {
      lambda = 0.5*(lamlo + lamhi);
      ++nbisect;
    }
  }
  return -2;
}

// cedr_local_inl.hpp:68-165: two variables in closed form. With early_exit_on_tol only
// infeasibility returns early (the reference's inner `info` shadows the outer one, so a
// corner that solves the problem falls through to the general case).
template <typename V>
CEDR_B200_LOCAL_HD Int solve_qp2 (V& v, const bool clip = true,
                                  const bool early_exit_on_tol = true) {
  if (early_exit_on_tol && impl::try_corners(v, impl::residual_tolerance(v)) == -1) return -1;
  { // The optimum without the bounds, if it respects them.
    Real qmass = 0, dm = v.b;
    for (Int i = 0; i < 2; ++i) {
      qmass += v.a(i)*(v.a(i)/v.w(i));
      dm -= v.a(i)*v.y(i);
    }
    const Real lambda = dm/qmass;
    bool inside = true;
    for (Int i = 0; i < 2 && inside; ++i) {
      v.x(i) = v.y(i) + lambda*(v.a(i)/v.w(i));
      inside = ! (v.x(i) < v.lo(i) || v.x(i) > v.hi(i));
    }
    if (inside) return 1;
  }
  // Walk along the line a'x = b from its midpoint: the four bound lines cut it at cut[];
  // the feasible segment lies between the two cuts that are neither the first minimum nor
  // the first maximum. The better end of that segment is the solution.
  const Real mid[2] = {0.5*v.b/v.a(0), 0.5*v.b/v.a(1)};
  const Real dir[2] = {-v.a(1), v.a(0)};
  const Real cut[4] = {(v.lo(1) - mid[1])/dir[1],    // bottom
                       (v.hi(0) - mid[0])/dir[0],    // right
                       (v.hi(1) - mid[1])/dir[1],    // top
                       (v.lo(0) - mid[0])/dir[0]};   // left
  Int kmin = 0, kmax = 0;
  for (Int k = 1; k < 4; ++k) {
    if (cut[k] < cut[kmin]) kmin = k;
    if (cut[k] > cut[kmax]) kmax = k;
  }
  Int end[2] = {0, 0}, nend = 0;
  for (Int k = 0; k < 4 && nend < 2; ++k)
    if (k != kmin && k != kmax) end[nend++] = k;
  Real cost[2];
  for (Int e = 0; e < 2; ++e) {
    Real c = 0;
    for (Int i = 0; i < 2; ++i) {
      v.x(i) = mid[i] + cut[end[e]]*dir[i];
      const Real d = v.y(i) - v.x(i);
      c += v.w(i)*(d*d);
    }
    cost[e] = c;
  }
  const Int side = end[cost[0] <= cost[1] ? 0 : 1];
  // The chosen bound line pins one coordinate; the constraint gives the other.
  const Int pinned = (side == 0 || side == 2) ? 1 : 0, other = 1 - pinned;
  v.x(pinned) = (side == 0 || side == 3) ? v.lo(pinned) : v.hi(pinned);
  v.x(other) = (v.b - v.a(pinned)*v.x(pinned))/v.a(other);
  if (clip) v.x(other) = impl::lesser(v.hi(other), impl::greater(v.lo(other), v.x(other)));
  return 1;
}

// cedr_local_inl.hpp:272-305: clip y into the bounds, then spread the mass defect over the
// remaining capacities on the side it has to move to. Does not check feasibility.
template <typename V>
CEDR_B200_LOCAL_HD void clip_and_spread (V& v, const bool clip = true) {
  const Int n = v.size();
  Real dm = v.b;
  for (Int i = 0; i < n; ++i) {
    v.x(i) = impl::greater(v.lo(i), impl::lesser(v.hi(i), v.y(i)));
    dm -= v.a(i)*v.x(i);
  }
  if (dm == 0) return;
  const bool up = dm > 0;
  Real cap = 0;
  for (Int i = 0; i < n; ++i)
    cap += v.a(i)*(up ? v.hi(i) - v.x(i) : v.x(i) - v.lo(i));
  if (cap > 0) {
    const Real fac = dm/cap;
    for (Int i = 0; i < n; ++i)
      v.x(i) += fac*(up ? v.hi(i) - v.x(i) : v.x(i) - v.lo(i));
  }
  if (clip)
    for (Int i = 0; i < n; ++i)
      v.x(i) = impl::greater(v.lo(i), impl::lesser(v.hi(i), v.x(i)));
}

// cedr_local_inl.hpp:307-330: x >= 0 in place of two-sided bounds; the upper bound b/a_i is
// the value at which one entry holds all of the mass. The view's lo / hi are overwritten.
template <int N>
CEDR_B200_LOCAL_HD Int solve_nonneg (Registers<N>& v, const Method::Enum method) {
  if (v.b < 0) return -1;
#pragma unroll
  for (Int i = 0; i < N; ++i) {
    if (i >= v.n) break;
    v.lo_[i] = 0;
    v.hi_[i] = v.b/v.a_[i];
  }
  if (method == Method::caas) {
    clip_and_spread(v);
    return 1;
  }
  return v.n == 2 ? solve_qp2(v) : solve_qp(v);
}

// ---------------------------------------------------------------- the reference's names

CEDR_B200_LOCAL_HD Int
solve_1eq_bc_qp (const Int n, const Real* w, const Real* a, const Real b, const Real* xlo,
                 const Real* xhi, const Real* y, Real* x, const Int max_its = 100) {
  Pointers v = {n, w, a, xlo, xhi, y, x, b};
  return solve_qp(v, max_its);
}

CEDR_B200_LOCAL_HD Int
solve_1eq_bc_qp_2d (const Real* w, const Real* a, const Real b, const Real* xlo,
                    const Real* xhi, const Real* y, Real* x, const bool clip = true,
                    const bool early_exit_on_tol = true) {
  Pointers v = {2, w, a, xlo, xhi, y, x, b};
  return solve_qp2(v, clip, early_exit_on_tol);
}

CEDR_B200_LOCAL_HD void
caas (const Int n, const Real* a, const Real b, const Real* xlo, const Real* xhi,
      const Real* y, Real* x, const bool clip = true) {
  Pointers v = {n, nullptr, a, xlo, xhi, y, x, b};
  clip_and_spread(v, clip);
}

// n <= 16 (cedr_local_inl.hpp:310); -3 otherwise.
CEDR_B200_LOCAL_HD Int
solve_1eq_nonneg (const Int n, const Real* a, const Real b, const Real* y, Real* x,
                  const Real* w, const Method::Enum lcl_method) {
  if (n > 16) return -3;
  if (b < 0) return -1;      // (x untouched, as in the reference)
  Registers<16> v;
  v.gather(n, w, a, b, nullptr, nullptr, y, 1);
  const Int info = solve_nonneg(v, lcl_method);
  v.scatter(x, 1);
  return info;
}

} // namespace local
} // namespace cedr

#endif
