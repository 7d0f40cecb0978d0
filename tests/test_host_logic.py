"""CPU-only: the C-ABI library loads and exports every declared symbol, the tree
plan (host C++) is self-consistent, and there is no CPU fallback."""
import re
import os
import sys

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


# ---- partial trees (tree::Node::level, cedr_tree_caller.hpp:20-22) --------------------

def prune_for_rank(tree, leaf_rank_of_cell, me, rng=None):
    """Rank `me`'s partial tree: every subtree without a cell of `me` becomes a stub.

    Nodes are renumbered in a shuffled order (root not at 0) when `rng` is given, so
    that the merge cannot rely on the parts' numbering. Returns (kids, cellidx, root,
    node_rank)."""
    kids, cellidx, root = tree
    nn = cellidx.size
    sys.setrecursionlimit(max(10000, 4*nn))
    has_mine = np.zeros(nn, bool)
    first_rank = np.zeros(nn, np.int32)

    def scan(i):
        if kids[2*i] < 0:
            first_rank[i] = leaf_rank_of_cell[cellidx[i]]
            has_mine[i] = first_rank[i] == me
            return
        scan(kids[2*i]); scan(kids[2*i + 1])
        has_mine[i] = has_mine[kids[2*i]] or has_mine[kids[2*i + 1]]
        first_rank[i] = first_rank[kids[2*i]]

    scan(root)
    keep = []

    def collect(i, parent_kept):
        keep.append(i)
        if kids[2*i] >= 0 and has_mine[i]:
            collect(kids[2*i], True); collect(kids[2*i + 1], True)

    collect(root, True)
    order = np.array(keep)
    if rng is not None:
        order = order[rng.permutation(order.size)]
    new = {int(o): j for j, o in enumerate(order)}
    pk = np.full(2*order.size, -1, np.int32)
    pc = np.full(order.size, -1, np.int64)
    pr = np.zeros(order.size, np.int32)
    for o, j in new.items():
        if kids[2*o] >= 0 and has_mine[o]:
            pk[2*j], pk[2*j + 1] = new[int(kids[2*o])], new[int(kids[2*o + 1])]
            pr[j] = me if rng is None else int(first_rank[o])
        elif kids[2*o] >= 0:
            pr[j] = first_rank[o]           # a stub: some other rank, no cell
            pc[j] = 1000000 + o             # the reference uses a stub's cellidx as an id
        else:
            pc[j] = cellidx[o]
            pr[j] = leaf_rank_of_cell[cellidx[o]]
    return pk, pc, new[int(root)], pr


def preorder(tree, node_rank):
    kids, cellidx, root = tree
    ok, oc, orank, stack = [], [], [], [(int(root), -1)]
    while stack:
        i, patch = stack.pop()
        me = len(oc)
        if patch >= 0:
            ok[patch] = me
        ok.extend([-1, -1]); oc.append(int(cellidx[i])); orank.append(0)
        if kids[2*i] < 0:
            orank[me] = int(node_rank[i])
        else:
            oc[me] = -1
            stack.append((int(kids[2*i + 1]), 2*me + 1))
            stack.append((int(kids[2*i]), 2*me))
    return np.array(ok, np.int32), np.array(oc, np.int64), np.array(orank, np.int32)


@pytest.mark.parametrize("ncells,nranks,maps", [(1, 1, "contig"), (2, 2, "contig"),
                                                (37, 3, "contig"), (200, 4, "random"),
                                                (200, 7, "pseudo"), (5400, 8, "contig")])
def test_partial_trees_merge_to_the_whole_tree(ncells, nranks, maps):
    from test_oracle_vs_ref import random_tree
    rng = np.random.default_rng(ncells + nranks)
    rt = random_tree(rng, ncells)
    for tree in (cb.make_1d_tree(ncells), cb.make_1d_tree(ncells, True),
                 (rt.kids, rt.cellidx, rt.root)):
        kids, cellidx, root = tree
        if maps == "contig":       # cedr_tree.cpp:368-369
            rank_of_cell = np.minimum(np.arange(ncells)//max(1, ncells//nranks), nranks - 1)
        elif maps == "pseudo":     # cedr_tree.cpp:371-374
            ci = np.arange(ncells)
            rank_of_cell = (ci + ci//nranks) % nranks
        else:
            rank_of_cell = rng.integers(0, nranks, ncells)
            rank_of_cell[:nranks] = np.arange(nranks)
        rank_of_cell = rank_of_cell.astype(np.int32)
        node_rank = np.where(cellidx >= 0, rank_of_cell[np.maximum(cellidx, 0)], 0)
        parts = [prune_for_rank(tree, rank_of_cell, r, rng) for r in range(nranks)]
        if nranks > 1 and maps == "contig" and ncells >= 37:
            assert max(p[1].size for p in parts) < cellidx.size    # really partial
        (mk, mc, mroot), mrank = cb.merge_partial_trees(parts)
        ek, ec, er = preorder(tree, node_rank)
        assert mroot == 0
        assert np.array_equal(mk, ek) and np.array_equal(mc, ec) and np.array_equal(mrank, er)
        # and the plan takes it: same leaf order as from the caller's whole tree
        a = cb.plan_probe(ncells, (mk, mc, 0), 64)
        b = cb.plan_probe(ncells, tree, 64)
        assert np.array_equal(a["lci2gci"], b["lci2gci"]) and a["idsum"] == b["idsum"]


def test_partial_trees_merge_rejects_bad_parts():
    tree = cb.make_1d_tree(8)
    rank_of_cell = (np.arange(8)//4).astype(np.int32)
    p0 = prune_for_rank(tree, rank_of_cell, 0)
    p1 = prune_for_rank(tree, rank_of_cell, 1)
    cb.merge_partial_trees([p0, p1])
    # nobody expands rank 1's subtree: its stub passes for a leaf, and the plan rejects
    # the tree (8 cells need 15 nodes)
    (mk, mc, _), _ = cb.merge_partial_trees([p0, p0])
    with pytest.raises(cb.CedrError):
        cb.plan_probe(8, (mk, mc, 0), 64)
    k, c, r, nr = (np.array(a).copy() if not np.isscalar(a) else a for a in p1)
    nr[(k[0::2] < 0) & (nr == 1)] = 2              # rank 1's leaves claimed by nobody
    with pytest.raises(cb.CedrError, match="holds this leaf as its own"):
        cb.merge_partial_trees([p0, (k, c, r, nr)])
    k, c, r, nr = (np.array(a).copy() if not np.isscalar(a) else a for a in p1)
    k[2*r + 1] = -1                                # root keeps one kid slot only
    with pytest.raises(cb.CedrError, match="both kid slots"):
        cb.merge_partial_trees([p0, (k, c, r, nr)])
    k, c, r, nr = (np.array(a).copy() if not np.isscalar(a) else a for a in p1)
    k[2*r] = r                                     # a cycle
    with pytest.raises(cb.CedrError):
        cb.merge_partial_trees([p0, (k, c, r, nr)])


def test_allgather_host_one_rank_and_missing_hook():
    """cedr_b200_allgather_host: one rank is a copy (no device needed); several ranks
    without the all-gather hook is an error, not a silent local copy."""
    import ctypes as C
    lib = cb.load_library()
    send = np.arange(5, dtype=np.float64)
    recv = np.zeros(5)
    null_fn = cb.ALLGATHER_FN()
    assert lib.cedr_b200_allgather_host(null_fn, None, 1, send.ctypes.data_as(C.c_void_p),
                                        recv.ctypes.data_as(C.c_void_p), 5) == 0
    assert np.array_equal(send, recv)
    recv2 = np.zeros(10)
    assert lib.cedr_b200_allgather_host(null_fn, None, 2, send.ctypes.data_as(C.c_void_p),
                                        recv2.ctypes.data_as(C.c_void_p), 5) != 0
    assert b"all-gather hook" in lib.cedr_b200_last_error()
