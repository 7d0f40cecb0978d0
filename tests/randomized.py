"""Randomized property-test data and checks, re-expressing the reference's test
contract (cedr_test_randomized.cpp) for one rank:

* 36 tracers = 6 problem types x 6 perturbations (cedr_test_randomized.cpp:26-51);
* data distributions of generate_rho / generate_Q (:53-93);
* perturbations of perturb_Q / add_const_to_Q / permute_Q (:95-197);
* the checks of TestRandomized::check (:293-418): local bounds, safety bounds,
  global mass, and bit-for-bit no-change for perturbation 0.

The random stream is numpy's, not glibc rand(): the reference holds no golden
numbers, only these properties. Like glibc's rand()/(RAND_MAX+1.0), every U here
carries only 31 random bits; the reference's no-change check for the nonnegative
types relies on that (sums of such values are exact, so the node QP's dm is 0).
"""
import numpy as np

C, S, T, N = 1, 2, 4, 8
EPS = np.finfo(np.float64).eps

PROBLEM_TYPES = [C | S | T, S, C | T, T, N, N | C]


class Tracer:
    def __init__(self, idx, problem_type, perturbation_type):
        self.idx = idx
        self.problem_type = problem_type
        self.perturbation_type = perturbation_type
        shapepreserve = bool(problem_type & S)
        nonnegative = bool(problem_type & N)
        self.no_change_should_hold = perturbation_type == 0
        self.safe_should_hold = True
        self.local_should_hold = perturbation_type < 4 and (shapepreserve or nonnegative)

    def __repr__(self):
        pt = self.problem_type
        s = "".join(c for c, m in (("c", C), ("s", S), ("t", T), ("n", N)) if pt & m)
        return "(ti %d %s pt %d)" % (self.idx, s, self.perturbation_type)


def tracers_vector():
    ts = []
    for perturb in range(6):
        for pt in PROBLEM_TYPES:
            ts.append(Tracer(len(ts), pt, perturb))
    return ts


class Values:
    def __init__(self, nt, ncells):
        self.rhom = np.zeros(ncells)
        self.Qm_min = np.zeros((nt, ncells))
        self.Qm = np.zeros((nt, ncells))
        self.Qm_max = np.zeros((nt, ncells))
        self.Qm_prev = np.zeros((nt, ncells))


def _permute_Q(rng, t, v):
    n = v.rhom.size
    p = np.arange(n)
    for _ in range(n):
        j, k = int(rng.random()*n), int(rng.random()*n)
        p[j], p[k] = p[k], p[j]
    v.Qm[t.idx] = v.Qm[t.idx][p].copy()


def _add_const_to_Q(rng, t, v, alpha, conserve_mass, safety_problem):
    ncells = v.rhom.size
    rhom = v.rhom.sum()
    Qm = v.Qm[t.idx].sum()
    Qm_max = v.Qm_max[t.idx].sum()
    Qm_max_safety = 0.0
    if safety_problem:
        Qm_max_safety = (v.Qm_max[t.idx]/v.rhom).max()*rhom
    if safety_problem:
        dQm = ((Qm_max - Qm) + alpha*(Qm_max_safety - Qm_max))/ncells
    else:
        dQm = alpha*(Qm_max - Qm)/ncells
    v.Qm[t.idx] += dQm
    _permute_Q(rng, t, v)
    relax = 0.9
    if conserve_mass:
        dQm_prev = dQm
    elif safety_problem:
        dQm_prev = ((Qm_max - Qm) + relax*alpha*(Qm_max_safety - Qm_max))/ncells
    else:
        dQm_prev = relax*alpha*(Qm_max - Qm)/ncells
    v.Qm_prev[t.idx] += dQm_prev


class _Rand31:
    """Uniform [0,1) with 31 random bits, like rand()/(RAND_MAX + 1.0)."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def random(self, n=None):
        if n is None:
            return float(self.rng.integers(0, 2**31))/2.0**31
        return self.rng.integers(0, 2**31, n).astype(np.float64)/2.0**31


def generate(ncells, seed, tracers=None):
    rng = _Rand31(seed)
    ts = tracers if tracers is not None else tracers_vector()
    v = Values(len(ts), ncells)
    v.rhom[:] = 0.5*(1 + rng.random(ncells))
    for t in ts:
        i = t.idx
        if t.problem_type & N:
            if t.no_change_should_hold:
                v.Qm[i] = rng.random(ncells)
            else:
                sgn = np.where(np.arange(ncells) % 2 == 0, 0.75, -0.75)
                v.Qm[i] = sgn + rng.random(ncells)
            v.Qm_min[i] = 0
            v.Qm_max[i] = 10
        else:
            q_min = -0.75 + rng.random(ncells)
            q_max = q_min + rng.random(ncells)
            q = q_min + (q_max - q_min)*rng.random(ncells)
            v.Qm_min[i] = q_min*v.rhom
            v.Qm_max[i] = q_max*v.rhom
            v.Qm[i] = np.maximum(v.Qm_min[i], np.minimum(v.Qm_max[i], q*v.rhom))
        v.Qm_prev[i] = v.Qm[i]
        # perturb_Q
        cm = not (t.problem_type & C)
        edg = 1 - ncells*EPS
        p = t.perturbation_type
        if p == 1:
            _permute_Q(rng, t, v)
        elif p == 2:
            _add_const_to_Q(rng, t, v, 0.5, cm, False)
        elif p == 3:
            _add_const_to_Q(rng, t, v, edg, cm, False)
        elif p == 4:
            _add_const_to_Q(rng, t, v, 0.5, cm, True)
        elif p == 5:
            _add_const_to_Q(rng, t, v, edg, cm, True)
    return ts, v


def check(ts, v, Qm_out, prefer_mass_con=False):
    """Return a list of failure strings (empty = pass)."""
    fails = []
    ulp3 = 3*EPS
    lv_tol = 100*EPS if prefer_mass_con else 0.0
    safety_tol = 100*EPS if prefer_mass_con else ulp3
    for k, t in enumerate(ts):
        Qm = Qm_out[k]
        Qm_min, Qm_max, Qm_prev = v.Qm_min[t.idx], v.Qm_max[t.idx], v.Qm_prev[t.idx]
        nonneg_only = bool(t.problem_type & N)
        safe_only = not t.local_should_hold
        if nonneg_only:
            lv = Qm < 0
        else:
            lv = (Qm < Qm_min - lv_tol) | (Qm > Qm_max + lv_tol)
        if not safe_only and lv.any():
            fails.append("local bounds violated %r" % t)
        if t.no_change_should_hold and not np.array_equal(Qm, Qm_prev):
            fails.append("changed but should not %r" % t)
        if safe_only:
            q_min = 0.0 if nonneg_only else (Qm_min/v.rhom).min()
            q_max = (Qm_max/v.rhom).max()
            delta = (q_max - q_min)*safety_tol
            if nonneg_only:
                sv = Qm < -ulp3
            else:
                sv = (Qm < q_min*v.rhom - delta) | (Qm > q_max*v.rhom + delta)
            if sv.any():
                fails.append("safety bounds violated %r" % t)
        desired, actual, den = Qm_prev.sum(), Qm.sum(), np.abs(Qm_prev).sum()
        rd = abs(actual - desired)/abs(den)
        if rd > 1e2*EPS:
            fails.append("mass re %.3e %r" % (rd, t))
    return fails
