"""TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.

ctypes front end to the two CPU checkers:

* ``Oracle``  -- oracle/liboracle_cedr.so, the plain-C restatement (cedr_oracle.c).
* ``Ref``     -- oracle/_ref/libcedr_ref[_omp].so, the UNMODIFIED reference sources
                 compiled by oracle/Makefile (present only if that build ran in a
                 container that has /root/reference; the .so ships to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module. compose_b200/ never does.

Array convention: SoA float64, tracer-major, global cell id fastest,
``a[t, gci]``; ``rhom[gci]``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

PT_CONSERVE, PT_SHAPEPRESERVE, PT_CONSISTENT, PT_NONNEGATIVE = 1, 2, 4, 8

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _l(a):
    return a.ctypes.data_as(_lp) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build(ref=True, quiet=True):
    """Run oracle/Makefile (C restatement always; _ref only where the reference is)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", HERE, "-j8"] + targets, check=True,
                   stdout=subprocess.DEVNULL if quiet else None,
                   stderr=subprocess.DEVNULL if quiet else None)


class Tree:
    """Flat binary tree: kids[2*i:2*i+2] (-1 for a leaf), cellidx[i], root."""

    def __init__(self, kids, cellidx, root=0):
        self.kids = np.ascontiguousarray(kids, dtype=np.int32).reshape(-1)
        self.cellidx = np.ascontiguousarray(cellidx, dtype=np.int64)
        self.root = int(root)
        self.nnodes = self.cellidx.size
        self.ncells = (self.nnodes + 1)//2


class Oracle:
    def __init__(self, path=None):
        path = path or os.path.join(HERE, "liboracle_cedr.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.oracle_make_bisection_tree.argtypes = [C.c_int, C.c_int, _ip, _lp]
        L.oracle_leaf_order.argtypes = [C.c_int, C.c_int, _ip, _lp, _lp, _ip]
        L.oracle_qlt_canonical_problem_type.argtypes = [C.c_int]
        L.oracle_qlt_run.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _lp, C.c_int, _ip,
                                     C.c_int] + [_dp]*6
        L.oracle_caas_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _ip, _lp,
                                      C.c_int, _ip] + [_dp]*5
        L.oracle_bfb_allreduce.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _lp, C.c_int,
                                           C.c_int, _dp, _dp]
        L.oracle_solve_1eq_bc_qp_2d.argtypes = [_dp, _dp, C.c_double, _dp, _dp, _dp, _dp,
                                                C.c_int, C.c_int]
        L.oracle_solve_1eq_bc_qp.argtypes = [C.c_int, _dp, _dp, C.c_double, _dp, _dp, _dp,
                                             _dp, C.c_int]
        L.oracle_local_caas.argtypes = [C.c_int, _dp, C.c_double, _dp, _dp, _dp, _dp,
                                        C.c_int]
        L.oracle_local_caas.restype = None
        L.oracle_solve_1eq_nonneg.argtypes = [C.c_int, _dp, C.c_double, _dp, _dp, _dp,
                                              C.c_int]
        L.oracle_solve_node_problem.argtypes = [C.c_int, C.c_double, _dp, C.c_double,
                                                C.c_double, _dp, _dp, C.c_double, _dp,
                                                _dp, C.c_int]
        L.oracle_solve_node_problem.restype = None
        L.oracle_fill_headline.argtypes = [C.c_int]*4 + [_dp]*5
        L.oracle_fill_headline.restype = None

    def fill_headline(self, ncells, config_id, t0, t1):
        """(rhom, qm_min, qm, qm_max, qm_prev) of SURVEY 8(d) for tracers [t0, t1)."""
        rhom = np.empty(ncells)
        a = [np.empty((t1 - t0, ncells)) for _ in range(4)]
        self.lib.oracle_fill_headline(ncells, config_id, t0, t1, _d(rhom), _d(a[0]),
                                      _d(a[1]), _d(a[2]), _d(a[3]))
        return (rhom,) + tuple(a)

    def bisection_tree(self, ncells, imbalanced=False):
        nn = 2*ncells - 1
        kids = np.empty(2*nn, np.int32)
        cellidx = np.empty(nn, np.int64)
        root = self.lib.oracle_make_bisection_tree(ncells, int(imbalanced), _i(kids),
                                                   _l(cellidx))
        assert root == 0
        return Tree(kids, cellidx, root)

    def leaf_order(self, tree):
        out = np.empty(tree.ncells, np.int64)
        nlev = C.c_int(0)
        rc = self.lib.oracle_leaf_order(tree.nnodes, tree.root, _i(tree.kids),
                                        _l(tree.cellidx), _l(out), C.byref(nlev))
        assert rc == 0
        return out, nlev.value

    def canonical_problem_type(self, mask):
        return self.lib.oracle_qlt_canonical_problem_type(int(mask))

    def qlt(self, tree, ptypes, rhom, qm_min, qm, qm_max, qm_prev, prefer_mass_con=False):
        pt = np.ascontiguousarray(ptypes, dtype=np.int32)
        nt = pt.size
        rhom, qm_min, qm, qm_max, qm_prev = map(_f64, (rhom, qm_min, qm, qm_max, qm_prev))
        out = np.empty((nt, tree.ncells))
        rc = self.lib.oracle_qlt_run(tree.ncells, tree.nnodes, tree.root, _i(tree.kids),
                                     _l(tree.cellidx), nt, _i(pt), int(prefer_mass_con),
                                     _d(rhom), _d(qm_min), _d(qm), _d(qm_max), _d(qm_prev),
                                     _d(out))
        if rc:
            raise RuntimeError("oracle_qlt_run failed: %d" % rc)
        return out

    def caas(self, ncells, ptypes, qm_min, qm, qm_max, qm_prev, tree=None):
        """tree=None: sequential sums (reducer 0); else tree-ordered sums (reducer 1)."""
        pt = np.ascontiguousarray(ptypes, dtype=np.int32)
        nt = pt.size
        qm_min, qm, qm_max, qm_prev = map(_f64, (qm_min, qm, qm_max, qm_prev))
        out = np.empty((nt, ncells))
        if tree is None:
            rc = self.lib.oracle_caas_run(ncells, 0, 0, 0, None, None, nt, _i(pt),
                                          _d(qm_min), _d(qm), _d(qm_max), _d(qm_prev),
                                          _d(out))
        else:
            rc = self.lib.oracle_caas_run(ncells, 1, tree.nnodes, tree.root,
                                          _i(tree.kids), _l(tree.cellidx), nt, _i(pt),
                                          _d(qm_min), _d(qm), _d(qm_max), _d(qm_prev),
                                          _d(out))
        if rc:
            raise RuntimeError("oracle_caas_run failed: %d" % rc)
        return out

    def bfb_allreduce(self, tree, send, nfield, transpose):
        send = _f64(send)
        recv = np.empty(nfield)
        rc = self.lib.oracle_bfb_allreduce(tree.ncells, tree.nnodes, tree.root,
                                           _i(tree.kids), _l(tree.cellidx), nfield,
                                           int(transpose), _d(send), _d(recv))
        assert rc == 0
        return recv

    def solve_1eq_bc_qp_2d(self, w, a, b, xlo, xhi, y, clip=True, early_exit=True):
        w, a, xlo, xhi, y = map(_f64, (w, a, xlo, xhi, y))
        x = np.zeros(2)
        info = self.lib.oracle_solve_1eq_bc_qp_2d(_d(w), _d(a), b, _d(xlo), _d(xhi), _d(y),
                                                  _d(x), int(clip), int(early_exit))
        return info, x

    def solve_1eq_bc_qp(self, w, a, b, xlo, xhi, y, max_its=100):
        w, a, xlo, xhi, y = map(_f64, (w, a, xlo, xhi, y))
        x = np.zeros(len(y))
        info = self.lib.oracle_solve_1eq_bc_qp(len(y), _d(w), _d(a), b, _d(xlo), _d(xhi),
                                               _d(y), _d(x), max_its)
        return info, x

    def local_caas(self, a, b, xlo, xhi, y, clip=True):
        a, xlo, xhi, y = map(_f64, (a, xlo, xhi, y))
        x = np.zeros(len(y))
        self.lib.oracle_local_caas(len(y), _d(a), b, _d(xlo), _d(xhi), _d(y), _d(x),
                                   int(clip))
        return x

    def solve_1eq_nonneg(self, a, b, y, w, method=0):
        a, y, w = map(_f64, (a, y, w))
        x = np.zeros(len(y))
        info = self.lib.oracle_solve_1eq_nonneg(len(y), _d(a), b, _d(y), _d(x), _d(w),
                                                method)
        return info, x

    def solve_node_problem(self, pt, rhom, pd, Qm, rhom0, k0d, rhom1, k1d, prefer=False):
        pd, k0d, k1d = map(_f64, (pd, k0d, k1d))
        q0, q1 = C.c_double(0), C.c_double(0)
        self.lib.oracle_solve_node_problem(pt, rhom, _d(pd), Qm, rhom0, _d(k0d),
                                           C.byref(q0), rhom1, _d(k1d), C.byref(q1),
                                           int(prefer))
        return q0.value, q1.value


def ref_available(omp=False):
    return os.path.exists(os.path.join(HERE, "_ref",
                                       "libcedr_ref_omp.so" if omp else "libcedr_ref.so"))


class Ref:
    """The reference's own QLT/CAAS (unmodified sources + stand-in Kokkos/MPI)."""

    def __init__(self, omp=False):
        path = os.path.join(HERE, "_ref", "libcedr_ref_omp.so" if omp else "libcedr_ref.so")
        self.lib = L = C.CDLL(path)
        L.cedr_ref_last_error.restype = C.c_char_p
        L.cedr_ref_qlt.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _lp, C.c_int, _ip,
                                   C.c_int] + [_dp]*6 + [_ip, C.c_int, _dp]
        L.cedr_ref_caas.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _ip, _lp, C.c_int,
                                    _ip] + [_dp]*6 + [C.c_int, _dp]
        L.cedr_ref_bfb_allreduce.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _lp, C.c_int,
                                             C.c_int, _dp, _dp]
        L.cedr_ref_qlt_leaf_order.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _lp, _lp, _ip,
                                              _ip]
        L.cedr_ref_solve_1eq_bc_qp_2d.argtypes = [_dp, _dp, C.c_double, _dp, _dp, _dp, _dp,
                                                  C.c_int, C.c_int]
        L.cedr_ref_solve_1eq_bc_qp.argtypes = [C.c_int, _dp, _dp, C.c_double, _dp, _dp,
                                               _dp, _dp, C.c_int]
        L.cedr_ref_local_caas.argtypes = [C.c_int, _dp, C.c_double, _dp, _dp, _dp, _dp,
                                          C.c_int]
        L.cedr_ref_local_caas.restype = None
        L.cedr_ref_solve_1eq_nonneg.argtypes = [C.c_int, _dp, C.c_double, _dp, _dp, _dp,
                                                C.c_int]
        L.cedr_ref_solve_node_problem.argtypes = [C.c_int, C.c_double, _dp, C.c_double,
                                                  C.c_double, _dp, _dp, C.c_double, _dp,
                                                  _dp, C.c_int]
        L.cedr_ref_solve_node_problem.restype = None

    def num_threads(self):
        return self.lib.cedr_ref_num_threads()

    def set_num_threads(self, n):
        """omp_set_num_threads (torchrun exports OMP_NUM_THREADS=1); returns the count."""
        self.lib.cedr_ref_set_num_threads.argtypes = [C.c_int]
        return self.lib.cedr_ref_set_num_threads(int(n))

    def _err(self, rc, what):
        if rc:
            raise RuntimeError("%s failed (%d): %s" %
                               (what, rc, self.lib.cedr_ref_last_error().decode()))

    @staticmethod
    def _tree_args(tree):
        """tree: ('bisect', imbalanced) or a Tree."""
        if isinstance(tree, Tree):
            return 2, tree.root, _i(tree.kids), _l(tree.cellidx)
        kind, imb = tree
        assert kind == "bisect"
        return (1 if imb else 0), 0, None, None

    def qlt(self, ncells, tree, ptypes, rhom, qm_min, qm, qm_max, qm_prev,
            prefer_mass_con=False, nrep=1):
        pt = np.ascontiguousarray(ptypes, dtype=np.int32)
        nt = pt.size
        rhom, qm_min, qm, qm_max, qm_prev = map(_f64, (rhom, qm_min, qm, qm_max, qm_prev))
        out = np.empty((nt, ncells))
        pto = np.empty(nt, np.int32)
        secs = np.zeros(max(nrep, 1))
        k, r, kd, ci = self._tree_args(tree)
        rc = self.lib.cedr_ref_qlt(ncells, k, r, kd, ci, nt, _i(pt), int(prefer_mass_con),
                                   _d(rhom), _d(qm_min), _d(qm), _d(qm_max), _d(qm_prev),
                                   _d(out), _i(pto), nrep, _d(secs))
        self._err(rc, "cedr_ref_qlt")
        return out, pto, secs

    def caas(self, ncells, ptypes, rhom, qm_min, qm, qm_max, qm_prev, tree=None, nrep=1):
        """tree=None: the reference's default sums; else its BfbTreeAllReducer."""
        pt = np.ascontiguousarray(ptypes, dtype=np.int32)
        nt = pt.size
        rhom, qm_min, qm, qm_max, qm_prev = map(_f64, (rhom, qm_min, qm, qm_max, qm_prev))
        out = np.empty((nt, ncells))
        secs = np.zeros(max(nrep, 1))
        if tree is None:
            red, (k, r, kd, ci) = 0, (0, 0, None, None)
        else:
            red, (k, r, kd, ci) = 1, self._tree_args(tree)
        rc = self.lib.cedr_ref_caas(ncells, red, k, r, kd, ci, nt, _i(pt), _d(rhom),
                                    _d(qm_min), _d(qm), _d(qm_max), _d(qm_prev), _d(out),
                                    nrep, _d(secs))
        self._err(rc, "cedr_ref_caas")
        return out, secs

    def bfb_allreduce(self, nleaf, tree, send, nfield, transpose):
        send = _f64(send)
        recv = np.empty(nfield)
        k, r, kd, ci = self._tree_args(tree)
        rc = self.lib.cedr_ref_bfb_allreduce(nleaf, k, r, kd, ci, nfield, int(transpose),
                                             _d(send), _d(recv))
        self._err(rc, "cedr_ref_bfb_allreduce")
        return recv

    def leaf_order(self, ncells, tree):
        out = np.empty(ncells, np.int64)
        nlev, nslots = C.c_int(0), C.c_int(0)
        k, r, kd, ci = self._tree_args(tree)
        rc = self.lib.cedr_ref_qlt_leaf_order(ncells, k, r, kd, ci, _l(out),
                                              C.byref(nlev), C.byref(nslots))
        self._err(rc, "cedr_ref_qlt_leaf_order")
        return out, nlev.value, nslots.value

    def solve_1eq_bc_qp_2d(self, w, a, b, xlo, xhi, y, clip=True, early_exit=True):
        w, a, xlo, xhi, y = map(_f64, (w, a, xlo, xhi, y))
        x = np.zeros(2)
        info = self.lib.cedr_ref_solve_1eq_bc_qp_2d(_d(w), _d(a), b, _d(xlo), _d(xhi),
                                                    _d(y), _d(x), int(clip),
                                                    int(early_exit))
        return info, x

    def solve_1eq_bc_qp(self, w, a, b, xlo, xhi, y, max_its=100):
        w, a, xlo, xhi, y = map(_f64, (w, a, xlo, xhi, y))
        x = np.zeros(len(y))
        info = self.lib.cedr_ref_solve_1eq_bc_qp(len(y), _d(w), _d(a), b, _d(xlo), _d(xhi),
                                                 _d(y), _d(x), max_its)
        return info, x

    def local_caas(self, a, b, xlo, xhi, y, clip=True):
        a, xlo, xhi, y = map(_f64, (a, xlo, xhi, y))
        x = np.zeros(len(y))
        self.lib.cedr_ref_local_caas(len(y), _d(a), b, _d(xlo), _d(xhi), _d(y), _d(x),
                                     int(clip))
        return x

    def solve_1eq_nonneg(self, a, b, y, w, method=0):
        a, y, w = map(_f64, (a, y, w))
        x = np.zeros(len(y))
        info = self.lib.cedr_ref_solve_1eq_nonneg(len(y), _d(a), b, _d(y), _d(x), _d(w),
                                                  method)
        return info, x

    def solve_node_problem(self, pt, rhom, pd, Qm, rhom0, k0d, rhom1, k1d, prefer=False):
        pd, k0d, k1d = map(_f64, (pd, k0d, k1d))
        q0, q1 = C.c_double(0), C.c_double(0)
        self.lib.cedr_ref_solve_node_problem(pt, rhom, _d(pd), Qm, rhom0, _d(k0d),
                                             C.byref(q0), rhom1, _d(k1d), C.byref(q1),
                                             int(prefer))
        return q0.value, q1.value
