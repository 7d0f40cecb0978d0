"""CPU-only: the C-ABI library loads and exports every declared symbol, the tree
plan (host C++) is self-consistent, and there is no CPU fallback."""
import re
import os

import numpy as np
import pytest

import compose_b200 as cb
from oracle.oracle_py import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    lib = cb.load_library()
    hdr = open(os.path.join(ROOT, "include", "cedr_b200.h")).read()
    declared = set(re.findall(r"\b(cedr_b200_[a-zA-Z0-9_]+)\s*\(", hdr))
    declared -= {"cedr_b200_allgather_fn"}
    bound = {s[0] for s in cb.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.cedr_b200_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.CedrError, match="no usable CUDA device"):
        cb.QLT(16)
    with pytest.raises(cb.CedrError, match="no usable CUDA device"):
        cb.CAAS(16)


def test_1d_tree_matches_oracle_builder():
    o = Oracle()
    for n in (1, 2, 3, 7, 21, 111, 675, 5400):
        for imb in (False, True):
            kids, cellidx, root = cb.make_1d_tree(n, imb)
            t = o.bisection_tree(n, imb)
            assert np.array_equal(kids, t.kids) and np.array_equal(cellidx, t.cellidx)


@pytest.mark.parametrize("ncells,imb,mbl", [
    (1, False, 1024), (2, False, 1024), (7, True, 2), (21, True, 4), (111, False, 16),
    (1000, True, 3), (5400, False, 1024), (86400, False, 1024), (86400, True, 1024),
    (393216, False, 1024)])
def test_plan_identity_sum(ncells, imb, mbl):
    """The reference's own check for partition code (test_comm_pattern,
    cedr_tree.cpp:279-348): sum of cell ids through the plan == n(n-1)/2."""
    r = cb.plan_probe(ncells, cb.make_1d_tree(ncells, imb), mbl)
    assert r["idsum"] == ncells*(ncells - 1)//2
    assert np.array_equal(r["lci2gci"], np.arange(ncells))
    assert r["nblocks"][-1] == 1


def test_plan_leaf_order_random_trees_matches_oracle():
    from test_oracle_vs_ref import random_tree
    o = Oracle()
    rng = np.random.default_rng(2)
    for n in (1, 2, 5, 33, 200, 3000):
        t = random_tree(rng, n)
        for mbl in (2, 7, 1024):
            r = cb.plan_probe(n, (t.kids, t.cellidx, t.root), mbl)
            lo, nlev = o.leaf_order(t)
            assert np.array_equal(r["lci2gci"], lo)
            assert r["nlevels_ref"] == nlev
            assert r["idsum"] == n*(n - 1)//2


def test_plan_rejects_malformed_trees():
    with pytest.raises(cb.CedrError):
        cb.plan_probe(3, (np.array([1, 2, -1, -1, -1, -1], np.int32),
                          np.array([-1, 0, 1], np.int64), 0))   # 3 nodes != 2*3-1
    with pytest.raises(cb.CedrError):
        cb.plan_probe(2, (np.array([1, -1, -1, -1, -1, -1], np.int32),
                          np.array([-1, 0, 1], np.int64), 0))   # one kid
