// Device node problem of the QLT down-sweep: the reference's
// impl::solve_node_problem (cedr_qlt_inl.hpp:119-203), r2l_nl_adjust_bounds
// (:69-99), local::solve_1eq_bc_qp_2d (cedr_local_inl.hpp:68-165) and
// local::solve_1eq_nonneg (:307-330), specialised to the only way QLT calls them
// (n = 2, a = {1, 1}).
//
// Every rewrite below is an IEEE-754 identity, so results are bit-identical to
// the reference compiled without FMA contraction (this file must be compiled
// with -fmad=false):
//   x*1 = x, x/1 = x, x/(-1) = -x, 0 + p = p for p >= +0, (-b) + x = x - b.
// The weights w_i = 1/rhom_i and their reciprocals q_i = 1/w_i depend only on
// the tree node, not on the tracer, so they are node constants computed once
// per run() by the rhom sweep (the reference recomputes them per node x tracer:
// 2 + 4 divisions); the only division left per node x tracer is lambda.
#ifndef CEDR_B200_NODE_SOLVE_CUH
#define CEDR_B200_NODE_SOLVE_CUH

namespace cedr_b200 {
namespace dev {

// cedr_kokkos.hpp:136-139 (ternaries, not fmin/fmax: same NaN/zero behaviour)
__device__ __forceinline__ double rmin (double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double rmax (double a, double b) { return a > b ? a : b; }

// Node constants produced by the rhom sweep for each internal node.
struct NodeConst {
  double w0, w1;    // 1/rhom_kid (cedr_qlt_inl.hpp:164)
  double q0, q1;    // 1/w_kid    (cedr_local_inl.hpp:83, :90 with a = 1)
  double rh0, rh1;  // rhom_kid   (bounds adjustment and consistent-only scaling)
};

// 10*eps as the reference evaluates it: 1e1*epsilon (exact in binary).
#define CEDR_B200_TEN_EPS 2.220446049250313080847263336181640625e-15

// solve_1eq_bc_qp_2d with a = {1,1}. Returns the reference's info value.
__device__ __forceinline__ int
qp2d (const double w0, const double w1, const double q0, const double q1,
      const double b, const double lo0, const double lo1, const double hi0,
      const double hi1, const double y0, const double y1, const bool clip,
      const bool early_exit_on_tol, double& x0, double& x1) {
  if (early_exit_on_tol) {
    // calc_r_tol, cedr_local_inl.hpp:13-18
    double ab = fabs(b);
    ab = rmax(ab, fabs(y0));
    ab = rmax(ab, fabs(y1));
    const double r_tol = CEDR_B200_TEN_EPS*fabs(ab);
    // check_lu, cedr_local_inl.hpp:23-41. Its "corner is a solution" result is
    // discarded by the caller's shadowed `info` (:73-77); only infeasibility
    // returns, leaving x at the violated corner.
    double r = (lo0 - b) + lo1;
    if ( ! (fabs(r) <= r_tol)) {
      if (r > 0) { x0 = lo0; x1 = lo1; return -1; }
      r = (hi0 - b) + hi1;
      if ( ! (fabs(r) <= r_tol)) {
        if (r < 0) { x0 = hi0; x1 = hi1; return -1; }
      }
    }
  }
  { // Unconstrained optimum, cedr_local_inl.hpp:80-97.
    const double qmass = q0 + q1;
    const double dm = (b - y0) - y1;
    const double lambda = dm/qmass;
    x0 = y0 + lambda*q0;
    if ( ! (x0 < lo0 || x0 > hi0)) {
      x1 = y1 + lambda*q1;
      if ( ! (x1 < lo1 || x1 > hi1)) return 1;
    }
  }
  // Intersections of the line x0 + x1 = b with the four bound lines,
  // cedr_local_inl.hpp:103-164. x_base = 0.5*b, x_dir = {-1, 1}.
  const double xb = 0.5*b;
  double al[4];
  al[0] = lo1 - xb;      // bottom
  al[1] = -(hi0 - xb);   // right
  al[2] = hi1 - xb;      // top
  al[3] = -(lo0 - xb);   // left
  double mn = al[0], mx = al[0];
  int imin = 0, imax = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i) {
    if (al[i] < mn) { mn = al[i]; imin = i; }
    if (al[i] > mx) { mx = al[i]; imax = i; }
  }
  // The two indices that are neither imin nor imax, in increasing order (the
  // first two if imin == imax).
  int ai0 = -1, ai1 = -1;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i != imin && i != imax) {
      if (ai0 < 0) ai0 = i;
      else if (ai1 < 0) ai1 = i;
    }
  const double alpha0 = ai0 == 0 ? al[0] : ai0 == 1 ? al[1] : ai0 == 2 ? al[2] : al[3];
  const double alpha1 = ai1 == 1 ? al[1] : ai1 == 2 ? al[2] : al[3];
  double obj0, obj1;
  {
    const double d0 = y0 - (xb - alpha0), d1 = y1 - (xb + alpha0);
    obj0 = w0*(d0*d0) + w1*(d1*d1);
  }
  {
    const double d0 = y0 - (xb - alpha1), d1 = y1 - (xb + alpha1);
    obj1 = w0*(d0*d0) + w1*(d1*d1);
    // The reference leaves x at this second trial point; both entries are
    // overwritten below.
  }
  const int ai = obj0 <= obj1 ? ai0 : ai1;
  if (ai == 0 || ai == 2) {
    x1 = ai == 0 ? lo1 : hi1;
    x0 = b - x1;
    if (clip) x0 = rmin(hi0, rmax(lo0, x0));
  } else {
    x0 = ai == 1 ? hi0 : lo0;
    x1 = b - x0;
    if (clip) x1 = rmin(hi1, rmax(lo1, x1));
  }
  return 1;
}

// r2l_nl_adjust_bounds, cedr_qlt_inl.hpp:69-99.
__device__ __forceinline__ void
adjust_bounds (double& bnd0, double& bnd1, const double rh0, const double rh1,
               const double extra) {
  const double qa = bnd0/rh0, qb = bnd1/rh1;
  if (extra < 0) {
    // i0 is the kid with the larger q (ties -> kid 0).
    const bool zero_first = qa >= qb;
    const double gap = zero_first ? (qb - qa)*rh0 : (qa - qb)*rh1;
    if (gap <= extra) {
      if (zero_first) bnd0 += extra; else bnd1 += extra;
      return;
    }
  } else {
    const bool zero_first = qa <= qb;
    const double gap = zero_first ? (qb - qa)*rh0 : (qa - qb)*rh1;
    if (gap >= extra) {
      if (zero_first) bnd0 += extra; else bnd1 += extra;
      return;
    }
  }
  const double tot = bnd0 + bnd1 + extra;
  const double rhtot = rh0 + rh1;
  const double qtot = tot/rhtot;
  bnd0 = qtot*rh0;
  bnd1 = qtot*rh1;
}

// The shape-preserving / consistent node problem, cedr_qlt_inl.hpp:119-173.
// (pmin, pqm, pmax) is the node's own l2r record, b its solved mass; (loK, yK,
// hiK) the kids' records.
__device__ __forceinline__ void
solve_node_bounded (const NodeConst& c, const bool prefer_mass_con,
                    const double pmin, const double pqm, const double pmax,
                    const double b, double lo0, const double y0, double hi0,
                    double lo1, const double y1, double hi1, double& x0, double& x1) {
  const bool lo = b < pmin, hi = b > pmax;
  if (lo || hi) {
    const double discrepancy = lo ? pmin - b : b - pmax;
    if (discrepancy > CEDR_B200_TEN_EPS*(pmax - pmin)) {
      if (lo) adjust_bounds(lo0, lo1, c.rh0, c.rh1, b - pmin);
      else adjust_bounds(hi0, hi1, c.rh0, c.rh1, b - pmax);
    }
  } else if (b == pqm && y0 >= lo0 && y0 <= hi0 && y1 >= lo1 && y1 <= hi1) {
    x0 = y0;
    x1 = y1;
    return;
  }
  qp2d(c.w0, c.w1, c.q0, c.q1, b, lo0, lo1, hi0, hi1, y0, y1, ! prefer_mass_con,
       ! prefer_mass_con, x0, x1);
}

// ---------------------------------------------------------------------------
// Lean form of the shape-preserving node problem for the fused kernels: the same
// operations on the same operands as solve_node_bounded / qp2d above (hence the same
// bits), arranged so that the common flow is short straight-line code:
//   - the node-level bound adjustment (b outside [pmin, pmax]) is a cold call;
//   - fabs is folded into compare / multiply operand modifiers;
//   - check_lu's two corner residuals are evaluated without nested early returns;
//   - in the boundary branch, the "drop the first minimum and the first maximum of the
//     four alphas, keep the other two in index order" scan of cedr_local_inl.hpp:112-131
//     is replaced by two compares whenever a0 < a2 and a1 < a3 (a box with nonzero
//     width on both sides): then the first minimum has index 0 or 1 (a2, a3 are
//     strictly above a0, a1) and the first maximum has index 2 or 3, so the scan keeps
//     {the larger of a0, a1; ties -> 1} and {the smaller of a2, a3; ties -> 3} whatever
//     their mutual order. Anything else takes the literal scan.
struct alignas(16) NodeWQ { double w0, w1, q0, q1; };
struct alignas(16) NodeRh { double rh0, rh1; };

__device__ __noinline__ void
solve_bounded_cold (const NodeWQ c, const NodeRh* rh, const bool prefer, const double pmin,
                    const double pqm, const double pmax, const double b, const double lo0,
                    const double y0, const double hi0, const double lo1, const double y1,
                    const double hi1, double* x) {
  NodeConst nc;
  nc.w0 = c.w0; nc.w1 = c.w1; nc.q0 = c.q0; nc.q1 = c.q1;
  const NodeRh r = *rh;
  nc.rh0 = r.rh0; nc.rh1 = r.rh1;
  double x0, x1;
  solve_node_bounded(nc, prefer, pmin, pqm, pmax, b, lo0, y0, hi0, lo1, y1, hi1, x0, x1);
  x[0] = x0; x[1] = x1;
}

// The literal boundary scan (any alpha order), out of line.
__device__ __noinline__ void
qp2d_boundary_cold (const double w0, const double w1, const double b, const double lo0,
                    const double lo1, const double hi0, const double hi1, const double y0,
                    const double y1, const bool clip, double* x) {
  const double xb = 0.5*b;
  double al[4];
  al[0] = lo1 - xb; al[1] = -(hi0 - xb); al[2] = hi1 - xb; al[3] = -(lo0 - xb);
  double mn = al[0], mx = al[0];
  int imin = 0, imax = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i) {
    if (al[i] < mn) { mn = al[i]; imin = i; }
    if (al[i] > mx) { mx = al[i]; imax = i; }
  }
  int ai0 = -1, ai1 = -1;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i != imin && i != imax) {
      if (ai0 < 0) ai0 = i;
      else if (ai1 < 0) ai1 = i;
    }
  const double alpha0 = ai0 == 0 ? al[0] : ai0 == 1 ? al[1] : ai0 == 2 ? al[2] : al[3];
  const double alpha1 = ai1 == 1 ? al[1] : ai1 == 2 ? al[2] : al[3];
  double obj0, obj1;
  {
    const double d0 = y0 - (xb - alpha0), d1 = y1 - (xb + alpha0);
    obj0 = w0*(d0*d0) + w1*(d1*d1);
  }
  {
    const double d0 = y0 - (xb - alpha1), d1 = y1 - (xb + alpha1);
    obj1 = w0*(d0*d0) + w1*(d1*d1);
  }
  const int ai = obj0 <= obj1 ? ai0 : ai1;
  double x0, x1;
  if (ai == 0 || ai == 2) {
    x1 = ai == 0 ? lo1 : hi1;
    x0 = b - x1;
    if (clip) x0 = rmin(hi0, rmax(lo0, x0));
  } else {
    x0 = ai == 1 ? hi0 : lo0;
    x1 = b - x0;
    if (clip) x1 = rmin(hi1, rmax(lo1, x1));
  }
  x[0] = x0; x[1] = x1;
}

// dm/qmass through the node constant rq = RN(1/qmass): q = dm rq; e = dm - qmass q
// (exact, one FMA); q + e rq rounds to the IEEE quotient (Markstein's correction step:
// with a correctly rounded reciprocal and no under/overflow in the intermediates the
// result is the correctly rounded quotient). The guards keep the exponents of dm and of
// the result away from the range ends; rq == 0 marks a qmass the rhom sweep declined
// (exponent outside 2^+-400, or fused kernels that do not carry rq). Otherwise the
// division instruction sequence runs. tools/microbench/div_check.cu compares the two
// bit for bit over 6e10 random and adversarial operand pairs on B200: 0 mismatches.
// A division costs 124 dependent cycles and 14 FP64-pipe slots on B200, this costs 25
// and 3.
__device__ __noinline__ double div_cold (const double a, const double b) { return a/b; }

__device__ __forceinline__ double
div_by_qmass (const double dm, const double qmass, const double rq) {
  const double q = dm*rq;
  const double e = fma(-qmass, q, dm);
  const double l = fma(e, rq, q);
  const unsigned ea = (static_cast<unsigned>(__double2hiint(dm)) >> 20) & 0x7ffu;
  const unsigned el = (static_cast<unsigned>(__double2hiint(l)) >> 20) & 0x7ffu;
#ifdef CEDR_B200_FASTDIV
  if (rq != 0 && (ea - 128u) < 1792u && (el - 128u) < 1792u) return l;
  return div_cold(dm, qmass);
#else
  // Off by default: on the current kernels the extra live registers (rq per node constant)
  // cost more in spills than the shorter dependency chain gains (measured: 9.1 vs 7.4 ms
  // down-sweep at ne120).
  return dm/qmass;
#endif
}
// The constant itself, computed once per node by the rhom sweep.
__device__ __forceinline__ double reciprocal_for_div (const double qmass) {
  const unsigned eq = (static_cast<unsigned>(__double2hiint(qmass)) >> 20) & 0x7ffu;
  return (eq - 623u) < 800u ? 1/qmass : 0.0;
}

template <bool PREFER>
__device__ __forceinline__ void
solve_bounded_lean (const NodeWQ& c, const double rq, const NodeRh* rh, const double pmin,
                    const double pqm,
                    const double pmax, const double b, const double lo0, const double y0,
                    const double hi0, const double lo1, const double y1, const double hi1,
                    double& x0, double& x1) {
  if (b < pmin || b > pmax) {
    double x[2];
    solve_bounded_cold(c, rh, PREFER, pmin, pqm, pmax, b, lo0, y0, hi0, lo1, y1, hi1, x);
    x0 = x[0]; x1 = x[1];
    return;
  }
  if (b == pqm) {
    if (y0 >= lo0 && y0 <= hi0 && y1 >= lo1 && y1 <= hi1) { x0 = y0; x1 = y1; return; }
  }
  if ( ! PREFER) {
    // calc_r_tol + check_lu (cedr_local_inl.hpp:13-41); only infeasibility returns.
    double ab = fabs(b) > fabs(y0) ? b : y0;
    ab = fabs(ab) > fabs(y1) ? ab : y1;
    const double r_tol = CEDR_B200_TEN_EPS*fabs(ab);
    const double r1 = (lo0 - b) + lo1;
    const double r2 = (hi0 - b) + hi1;
    const bool c1 = ! (fabs(r1) <= r_tol);
    const bool inf_lo = c1 && r1 > 0;
    const bool inf_hi = c1 && ! (fabs(r2) <= r_tol) && r2 < 0;
    if (inf_lo || inf_hi) {
      x0 = inf_lo ? lo0 : hi0;
      x1 = inf_lo ? lo1 : hi1;
      return;
    }
  }
  // The boundary branch's alphas do not depend on the division below; formed first, they
  // overlap its latency (in a warp some lane nearly always takes the boundary branch).
  // a0 bottom, a1 right, a2 top, a3 left.
  const double xb = 0.5*b;
  const double a0 = lo1 - xb, a1 = -(hi0 - xb), a2 = hi1 - xb, a3 = -(lo0 - xb);
  const bool L = a1 >= a0, H = a3 <= a2;       // kept pair = (max(a0,a1), min(a2,a3))
  const double aL = L ? a1 : a0, aH = H ? a3 : a2;
  const bool boxed = a0 < a2 && a1 < a3;
  { // Unconstrained optimum, cedr_local_inl.hpp:80-97.
    const double qmass = c.q0 + c.q1;
    const double dm = (b - y0) - y1;
    const double lambda = div_by_qmass(dm, qmass, rq);
    x0 = y0 + lambda*c.q0;
    x1 = y1 + lambda*c.q1;
    if ( ! (x0 < lo0 || x0 > hi0) && ! (x1 < lo1 || x1 > hi1)) return;
  }
  // Boundary branch.
  if ( ! boxed) {
    double x[2];
    qp2d_boundary_cold(c.w0, c.w1, b, lo0, lo1, hi0, hi1, y0, y1, ! PREFER, x);
    x0 = x[0]; x1 = x[1];
    return;
  }
  double obj0, obj1;
  {
    const double d0 = y0 - (xb - aL), d1 = y1 - (xb + aL);
    obj0 = c.w0*(d0*d0) + c.w1*(d1*d1);
  }
  {
    const double d0 = y0 - (xb - aH), d1 = y1 - (xb + aH);
    obj1 = c.w0*(d0*d0) + c.w1*(d1*d1);
  }
  const bool first = obj0 <= obj1;
  // ai = first ? (L ? 1 : 0) : (H ? 3 : 2); odd ai pins x0, even pins x1.
  const bool pin0 = first ? L : H;
  const double v = first ? (L ? hi0 : lo1) : (H ? lo0 : hi1);
  double o = b - v;
  if ( ! PREFER) {
    const double olo = pin0 ? lo1 : lo0, ohi = pin0 ? hi1 : hi0;
    o = rmin(ohi, rmax(olo, o));
  }
  x0 = pin0 ? v : o;
  x1 = pin0 ? o : v;
}

// The nonnegative node problem, cedr_qlt_inl.hpp:188-197 -> solve_1eq_nonneg
// (cedr_local_inl.hpp:307-330) with n = 2, least squares: bounds [0, b/a_i],
// default clip / early-exit flags (independent of the CDR option).
__device__ __forceinline__ void
solve_node_nonneg (const NodeConst& c, const double b, const double y0,
                   const double y1, double& x0, double& x1) {
  x0 = y0;
  x1 = y1;
  if (b < 0) return;
  qp2d(c.w0, c.w1, c.q0, c.q1, b, 0.0, 0.0, b, b, y0, y1, true, true, x0, x1);
}

} // namespace dev
} // namespace cedr_b200

#endif
