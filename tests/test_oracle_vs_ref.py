"""Pin the plain-C oracle (oracle/cedr_oracle.c) bit-for-bit against the UNMODIFIED
reference sources compiled by oracle/Makefile into oracle/_ref/. Skipped where
oracle/_ref is absent (then tests/test_oracle_golden.py pins it through fixtures
generated from the same build).
"""
import numpy as np
import pytest

import randomized as R
from oracle.oracle_py import Oracle, Ref, Tree, ref_available

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def libs():
    return Oracle(), Ref()


def random_tree(rng, ncells):
    """Random binary tree over ncells leaves with a random leaf->cell permutation."""
    perm = rng.permutation(ncells)
    kids, cellidx = [], []

    def rec(lo, hi):
        me = len(cellidx)
        kids.extend([-1, -1])
        cellidx.append(-1)
        if hi - lo == 1:
            cellidx[me] = int(perm[lo])
            return me
        cut = int(rng.integers(lo + 1, hi))
        k0 = rec(lo, cut)
        k1 = rec(cut, hi)
        kids[2*me], kids[2*me + 1] = k0, k1
        return me

    rec(0, ncells)
    return Tree(np.array(kids, np.int32), np.array(cellidx, np.int64), 0)


def test_bisection_tree_matches_reference_builder(libs):
    o, r = libs
    for ncells in (1, 2, 3, 7, 21, 111, 675, 1000):
        for imb in (False, True):
            tree = o.bisection_tree(ncells, imb)
            lo, nlev_o = o.leaf_order(tree)
            lr, nlev_r, nslots = r.leaf_order(ncells, ("bisect", imb))
            assert np.array_equal(lo, lr)
            assert nlev_o == nlev_r
            assert nslots == 2*ncells - 1
            lr2, nlev_r2, _ = r.leaf_order(ncells, tree)
            assert np.array_equal(lo, lr2) and nlev_r2 == nlev_r


def test_leaf_order_random_trees(libs):
    o, r = libs
    rng = np.random.default_rng(5)
    for ncells in (1, 2, 5, 33, 200):
        tree = random_tree(rng, ncells)
        lo, nlev_o = o.leaf_order(tree)
        lr, nlev_r, _ = r.leaf_order(ncells, tree)
        assert np.array_equal(lo, lr) and nlev_o == nlev_r


@pytest.mark.parametrize("ncells", [1, 2, 7, 21, 111, 1000])
@pytest.mark.parametrize("imbalanced", [False, True])
@pytest.mark.parametrize("prefer", [False, True])
def test_qlt_randomized_bitwise(libs, ncells, imbalanced, prefer):
    o, r = libs
    ts, v = R.generate(ncells, seed=1000*ncells + 2*imbalanced + prefer)
    pts = [t.problem_type for t in ts]
    tree = o.bisection_tree(ncells, imbalanced)
    out_o = o.qlt(tree, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev, prefer)
    out_r, pto, _ = r.qlt(ncells, ("bisect", imbalanced), pts, v.rhom, v.Qm_min, v.Qm,
                          v.Qm_max, v.Qm_prev, prefer)
    assert np.array_equal(out_o, out_r)
    assert [o.canonical_problem_type(p) for p in pts] == list(pto)
    assert R.check(ts, v, out_r, prefer) == []


def test_qlt_random_trees_bitwise(libs):
    o, r = libs
    rng = np.random.default_rng(11)
    for ncells in (2, 3, 9, 64, 257):
        tree = random_tree(rng, ncells)
        ts, v = R.generate(ncells, seed=ncells)
        pts = [t.problem_type for t in ts]
        out_o = o.qlt(tree, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev)
        out_r, _, _ = r.qlt(ncells, tree, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev)
        assert np.array_equal(out_o, out_r)


def caas_tracers():
    return [t for t in R.tracers_vector()
            if (t.problem_type & R.S) and t.local_should_hold]


@pytest.mark.parametrize("ncells", [1, 2, 4, 11, 111, 1000])
def test_caas_bitwise_both_sum_orders(libs, ncells):
    o, r = libs
    ts, v = R.generate(ncells, seed=77 + ncells)
    sel = caas_tracers()
    idx = [t.idx for t in sel]
    pts = [t.problem_type for t in sel]
    a = [x[idx] for x in (v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev)]
    # default (sequential) sums
    out_o = o.caas(ncells, pts, *a)
    out_r, _ = r.caas(ncells, pts, v.rhom, *a)
    assert np.array_equal(out_o, out_r)
    # tree-ordered sums == reference CAAS + UserAllReducer -> BfbTreeAllReducer
    for imb in (False, True):
        tree = o.bisection_tree(ncells, imb)
        out_o = o.caas(ncells, pts, *a, tree=tree)
        out_r, _ = r.caas(ncells, pts, v.rhom, *a, tree=("bisect", imb))
        assert np.array_equal(out_o, out_r)
    # property check on the tree-ordered result (check() indexes v by t.idx and the
    # output by position in `sel`)
    assert R.check(sel, v, out_r) == []


def test_bfb_allreduce_bitwise(libs):
    o, r = libs
    rng = np.random.default_rng(3)
    for ncells in (1, 3, 24, 100):
        for imb in (False, True):
            for transpose in (False, True):
                nf = 3
                send = rng.random(ncells*nf)
                tree = o.bisection_tree(ncells, imb)
                a = o.bfb_allreduce(tree, send, nf, transpose)
                b = r.bfb_allreduce(ncells, ("bisect", imb), send, nf, transpose)
                assert np.array_equal(a, b)
                # vs a plain sum, the reference's own tolerance
                # (cedr_bfb_tree_allreduce.cpp:211-217)
                ref = (send.reshape(nf, ncells).sum(1) if transpose
                       else send.reshape(ncells, nf).sum(0))
                tol = 2*max(np.log(ncells), 1)*np.finfo(float).eps
                assert np.all(np.abs(a - ref) <= tol*np.abs(ref))


def test_local_solvers_bitwise(libs):
    o, r = libs
    rng = np.random.default_rng(9)
    for trial in range(2000):
        n = 2 if trial % 2 == 0 else int(rng.integers(2, 17))
        w = 0.1 + rng.random(n)
        a = 0.1 + rng.random(n) if trial % 4 else np.ones(n)
        xlo = rng.random(n) - 0.5
        xhi = xlo + rng.random(n)
        y = xlo + (xhi - xlo)*(1.6*rng.random(n) - 0.3)
        # mostly feasible b, sometimes not
        lo, hi = float(a @ xlo), float(a @ xhi)
        b = lo + (hi - lo)*(1.2*rng.random() - 0.1)
        if n == 2:
            for clip in (0, 1):
                for ee in (0, 1):
                    i1, x1 = o.solve_1eq_bc_qp_2d(w, a, b, xlo, xhi, y, clip, ee)
                    i2, x2 = r.solve_1eq_bc_qp_2d(w, a, b, xlo, xhi, y, clip, ee)
                    assert i1 == i2 and np.array_equal(x1, x2)
        i1, x1 = o.solve_1eq_bc_qp(w, a, b, xlo, xhi, y)
        i2, x2 = r.solve_1eq_bc_qp(w, a, b, xlo, xhi, y)
        assert i1 == i2 and np.array_equal(x1, x2)
        for clip in (0, 1):
            assert np.array_equal(o.local_caas(a, b, xlo, xhi, y, clip),
                                  r.local_caas(a, b, xlo, xhi, y, clip))
        bb = abs(b)
        for method in (0, 1):
            i1, x1 = o.solve_1eq_nonneg(a, bb, y, w, method)
            i2, x2 = r.solve_1eq_nonneg(a, bb, y, w, method)
            assert i1 == i2 and np.array_equal(x1, x2)


def test_solve_node_problem_bitwise(libs):
    o, r = libs
    rng = np.random.default_rng(21)
    pts = [R.C | R.S | R.T, R.S | R.T, R.C | R.T, R.T, R.N, R.C | R.N]
    for trial in range(6000):
        pt = pts[trial % 6]
        rhom0, rhom1 = 0.5*(1 + rng.random(2))
        rhom = rhom0 + rhom1
        kd = []
        for rh in (rhom0, rhom1):
            qmin = rng.random() - 0.75
            qmax = qmin + rng.random()
            q = qmin + (qmax - qmin)*(1.4*rng.random() - 0.2)
            if (pt & R.T) and not (pt & R.S):
                kd.append(np.array([qmin, q*rh, qmax, 0.0]))
            elif pt & R.N:
                kd.append(np.array([abs(q)*rh, 0.0, 0.0, 0.0]))
            else:
                kd.append(np.array([qmin*rh, q*rh, qmax*rh, 0.0]))
        k0d, k1d = kd
        if (pt & R.T) and not (pt & R.S):
            pd = np.array([min(k0d[0], k1d[0]), k0d[1] + k1d[1], max(k0d[2], k1d[2]), 0])
            lo, hi = pd[0]*rhom, pd[2]*rhom
        elif pt & R.N:
            pd = np.array([k0d[0] + k1d[0], 0, 0, 0.0])
            lo, hi = 0.0, 2*pd[0]
        else:
            pd = k0d + k1d
            lo, hi = pd[0], pd[2]
        mode = trial % 5
        if mode == 0:
            Qm = pd[0] if (pt & R.N) else pd[1]      # untouched mass -> quick exit path
        elif mode == 1:
            Qm = lo - 0.1*(hi - lo)*rng.random()     # below: safety problem
            if pt & R.N:
                Qm = abs(Qm)
        elif mode == 2:
            Qm = hi + 0.1*(hi - lo)*rng.random()     # above
        else:
            Qm = lo + (hi - lo)*rng.random()
        for prefer in (False, True):
            a = o.solve_node_problem(pt, rhom, pd, Qm, rhom0, k0d, rhom1, k1d, prefer)
            b = r.solve_node_problem(pt, rhom, pd, Qm, rhom0, k0d, rhom1, k1d, prefer)
            assert a == b
