"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle, bit for bit.

The contract (BASELINE.json north_star): results match the reference's CDR on
identical inputs within 1e-13 relative per cell, and bit-for-bit in the
deterministic-order mode. QLT's arithmetic order is fixed by the tree, and CAAS
runs in tree-ordered mode, so every comparison here demands EXACT equality
(np.array_equal); the 1e-13 tolerance is therefore met with margin 0.
"""
import numpy as np
import pytest

import randomized as R
from oracle.oracle_py import Oracle, Tree

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def random_tree(rng, ncells):
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


def test_library_is_loaded_and_device_present():
    import compose_b200 as cb
    lib = cb.load_library()
    assert lib.cedr_b200_device_available() == 1


def test_fill_headline_matches_numpy():
    import compose_b200 as cb
    from compose_b200 import workloads as W
    ncells, nt = 777, 5
    dev = cb.fill_headline(ncells, nt, 7)
    host = W.headline(ncells, nt, 7)
    for d, h in zip(dev, host):
        assert np.array_equal(d.cpu().numpy(), h)


@pytest.mark.parametrize("ncells", [1, 2, 7, 21, 111, 1000, 2731])
@pytest.mark.parametrize("imbalanced", [False, True])
@pytest.mark.parametrize("prefer", [False, True])
def test_qlt_randomized_bitwise(oracle, ncells, imbalanced, prefer):
    from gpu_util import run_qlt_gpu
    ts, v = R.generate(ncells, seed=31*ncells + 2*imbalanced + prefer)
    pts = [t.problem_type for t in ts]
    tree = oracle.bisection_tree(ncells, imbalanced)
    ref = oracle.qlt(tree, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev, prefer)
    got, q = run_qlt_gpu(ncells, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev,
                         imbalanced=imbalanced, prefer=prefer,
                         external_buffers=imbalanced)
    assert np.array_equal(got, ref)
    assert R.check(ts, v, got, prefer) == []
    # get_problem_type reports the canonical type (cedr_qlt.cpp:707-713)
    assert [q.get_problem_type(i) for i in range(len(pts))] == \
        [oracle.canonical_problem_type(p) for p in pts]


@pytest.mark.parametrize("ncells,mbl", [(7, 2), (21, 4), (111, 16), (1000, 32),
                                        (1000, 2), (5400, 128)])
def test_qlt_multi_tier_plans_bitwise(oracle, ncells, mbl):
    """Small max_block_leaves forces 3+ tier plans; results must not change."""
    from gpu_util import run_qlt_gpu
    ts, v = R.generate(ncells, seed=ncells + mbl)
    pts = [t.problem_type for t in ts]
    for imb in (False, True):
        tree = oracle.bisection_tree(ncells, imb)
        ref = oracle.qlt(tree, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev)
        got, q = run_qlt_gpu(ncells, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev,
                             imbalanced=imb, max_block_leaves=mbl)
        assert q.plan_info()["ntiers"] >= 2
        assert np.array_equal(got, ref)


def test_qlt_random_trees_bitwise(oracle):
    from gpu_util import run_qlt_gpu
    rng = np.random.default_rng(17)
    for ncells, mbl in ((2, None), (3, None), (9, 4), (64, 8), (257, None), (1500, 64)):
        tree = random_tree(rng, ncells)
        ts, v = R.generate(ncells, seed=ncells)
        pts = [t.problem_type for t in ts]
        ref = oracle.qlt(tree, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev)
        got, q = run_qlt_gpu(ncells, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev,
                             tree=(tree.kids, tree.cellidx, tree.root),
                             max_block_leaves=mbl)
        lo, _ = oracle.leaf_order(tree)
        assert np.array_equal(q.get_owned_glblcells(), lo)
        assert np.array_equal(got, ref)


def test_qlt_from_partial_trees_bitwise(oracle):
    """tree::Node::level partial trees (cedr_tree_caller.hpp:20-22): three ranks' pruned
    parts of a random tree under the pseudorandom map (cedr_tree.cpp:371-374), merged by
    cedr_b200_merge_partial_trees, give the oracle's results on the caller's whole tree."""
    import compose_b200 as cb
    from gpu_util import run_qlt_gpu
    from test_host_logic import prune_for_rank
    rng = np.random.default_rng(23)
    ncells, nranks = 700, 3
    tree = random_tree(rng, ncells)
    ci = np.arange(ncells)
    rank_of_cell = ((ci + ci//nranks) % nranks).astype(np.int32)
    rank_of_cell[100:400] = 2
    parts = [prune_for_rank((tree.kids, tree.cellidx, tree.root), rank_of_cell, r, rng)
             for r in range(nranks)]
    assert all(p[1].size < tree.cellidx.size for p in parts)
    whole, _ = cb.merge_partial_trees(parts)
    ts, v = R.generate(ncells, seed=5)
    pts = [t.problem_type for t in ts]
    ref = oracle.qlt(tree, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev)
    got, _ = run_qlt_gpu(ncells, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev,
                         tree=whole, max_block_leaves=64)
    assert np.array_equal(got, ref)


def test_qlt_headline_inputs_bitwise(oracle):
    """ne30-shaped tree (5,400 cells), a slice of the headline tracers, all `cst`."""
    from compose_b200 import workloads as W
    from gpu_util import run_qlt_gpu
    ncells, nt = 5400, 48
    rhom, lo, q, hi, prev = W.headline(ncells, nt, 1)
    pts = [7]*nt
    tree = oracle.bisection_tree(ncells)
    ref = oracle.qlt(tree, pts, rhom, lo, q, hi, prev)
    got, qlt = run_qlt_gpu(ncells, pts, rhom, lo, q, hi, prev, nrun=2)
    assert np.array_equal(got, ref)
    # properties: bounds exact, mass to 100 eps
    assert np.all(got >= lo) and np.all(got <= hi)
    rel = np.abs(got.sum(1) - prev.sum(1))/np.abs(prev).sum(1)
    assert rel.max() <= 100*np.finfo(float).eps


def caas_tracers():
    return [t for t in R.tracers_vector()
            if (t.problem_type & R.S) and t.local_should_hold]


@pytest.mark.parametrize("ncells", [1, 2, 4, 11, 111, 1000, 5400])
def test_caas_tree_sums_bitwise(oracle, ncells):
    from gpu_util import run_caas_gpu
    ts, v = R.generate(ncells, seed=5 + ncells)
    sel = caas_tracers()
    idx = [t.idx for t in sel]
    pts = [t.problem_type for t in sel]
    a = [x[idx] for x in (v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev)]
    tree = oracle.bisection_tree(ncells)
    ref = oracle.caas(ncells, pts, *a, tree=tree)
    for mbl, ext in ((None, False), (8, True)):
        if mbl and ncells < 3:
            continue
        got, c = run_caas_gpu(ncells, pts, v.rhom, a[0], a[1], a[2], a[3],
                              max_block_leaves=mbl, external_buffers=ext)
        assert np.array_equal(got, ref)
    assert R.check(sel, v, got) == []


def test_caas_headline_inputs_bitwise(oracle):
    from compose_b200 import workloads as W
    from gpu_util import run_caas_gpu
    ncells, nt = 5400, 48
    rhom, lo, q, hi, prev = W.headline(ncells, nt, 1)
    pts = [7]*nt
    ref = oracle.caas(ncells, pts, lo, q, hi, prev, tree=oracle.bisection_tree(ncells))
    got, _ = run_caas_gpu(ncells, pts, rhom, lo, q, hi, prev)
    assert np.array_equal(got, ref)
    assert np.all(got >= lo) and np.all(got <= hi)


@pytest.mark.parametrize("ncells,nt,conserve", [(5400, 37, True), (8*768, 20, True),
                                                (3*700 + 1, 9, False), (86400, 12, True)])
def test_caas_cluster_kernel_bitwise(oracle, ncells, nt, conserve):
    """Opt-in CAAS::run as ONE kernel of thread-block clusters (cluster_caas.cuh: a tracer
    held in the shared memory of up to 16 CTAs, block records exchanged through distributed
    shared memory, rows read once): one launch, the oracle's bits, for a cluster of one CTA
    (5,400 cells), of two (6,144 cells: 7 + 1 blocks), with and without a Qm_prev row, and
    the full 16-CTA cluster at 86,400 cells; run twice (the row ring's parity)."""
    import torch
    import compose_b200 as cb
    from compose_b200 import workloads as W
    rhom, lo, q, hi, prev = W.headline(ncells, nt, 1)
    pt = 7 if conserve else 6
    pts = [pt]*nt
    ref = oracle.caas(ncells, pts, lo, q, hi, prev, tree=oracle.bisection_tree(ncells))
    c = cb.CAAS(ncells)
    c.set_cluster_caas(1)
    for p in pts:
        c.declare_tracer(p)
    c.end_tracer_declarations()
    c.finish_setup()
    assert c.uses_cluster_caas()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    c.set_rhom(dev(rhom))
    for _ in range(2):
        c.set_Qm(dev(q), dev(lo), dev(hi), dev(prev))
        c.run()
        assert c.last_run_launches() == 1
        got = c.get_Qm().cpu().numpy()
        c.synchronize()
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("ncells", [1, 2, 21, 111, 256])
def test_single_block_problems_run_in_one_launch(oracle, ncells, monkeypatch):
    """A tree that is one small block (cedr_test_1d_transport's 111 cells, the randomized
    unit test's trees) runs as ONE launch per problem class (solo_kernel: rhom sums, node
    constants and the sweep in one CTA per tracer); bits equal to the oracle and to the
    two-launch path."""
    import os
    from compose_b200 import workloads as W
    from gpu_util import run_qlt_gpu, run_caas_gpu
    nt = 5
    rhom, lo, q, hi, prev = W.headline(ncells, nt, 4)
    pts = [7]*nt
    tree = oracle.bisection_tree(ncells)
    got, c = run_qlt_gpu(ncells, pts, rhom, lo, q, hi, prev)
    assert c.last_run_launches() == 1
    assert np.array_equal(got, oracle.qlt(tree, pts, rhom, lo, q, hi, prev))
    gotc, c = run_caas_gpu(ncells, pts, rhom, lo, q, hi, prev)
    assert c.last_run_launches() == 1
    assert np.array_equal(gotc, oracle.caas(ncells, pts, lo, q, hi, prev, tree=tree))
    monkeypatch.setenv("CEDR_B200_NO_SOLO", "1")
    got2, c = run_qlt_gpu(ncells, pts, rhom, lo, q, hi, prev)
    assert c.last_run_launches() >= 2 and np.array_equal(got, got2)
    gotc2, c = run_caas_gpu(ncells, pts, rhom, lo, q, hi, prev)
    assert c.last_run_launches() >= 2 and np.array_equal(gotc, gotc2)


def test_errors_mirror_reference():
    import compose_b200 as cb
    q = cb.QLT(8)
    with pytest.raises(cb.CedrError, match="rhomidx > 0 is not supported"):
        q.declare_tracer(7, 1)
    with pytest.raises(cb.CedrError, match="Invalid problem type"):
        q.declare_tracer(1, 0)   # conservation alone (cedr_qlt.hpp:82-83)
    q.declare_tracer(7)
    q.end_tracer_declarations()
    with pytest.raises(cb.CedrError, match="end_tracer_declarations was already called"):
        q.declare_tracer(7)
    c = cb.CAAS(4)
    with pytest.raises(cb.CedrError, match="does not support ! shapepreserve"):
        c.declare_tracer(cb.CONSERVE | cb.CONSISTENT)
    with pytest.raises(cb.CedrError, match="0 cells"):
        cb.CAAS(0)


@pytest.mark.parametrize("ncells", [5400, 8*513, 8*768, 8192, 2*1023, 3*700 + 1])
@pytest.mark.parametrize("prefer", [False, True])
@pytest.mark.parametrize("ring", [False, True, "down2"])
def test_fast_path_blocks_bitwise(oracle, ncells, prefer, ring, monkeypatch):
    """Tier-0 blocks of 513..1024 leaves run the fast kernels (TMA + register
    micro-subtrees) for the st/cst classes -- as separate launches (default: the down-sweep
    as midT_kernel + down3_kernel; "down2": the one-kernel down-sweep with a top warp) or as
    the persistent ring kernel -- with the same bits as the oracle, for every block size
    class (few pairs .. all pairs)."""
    from gpu_util import run_qlt_gpu
    # The randomized set has only a handful of tracers per class: below the default
    # threshold of the split down-sweep.
    monkeypatch.setenv("CEDR_B200_TRANSPOSED_MIN", "1")
    if ring == "down2":
        monkeypatch.setenv("CEDR_B200_TRANSPOSED", "0")
        ring = False
    ts, v = R.generate(ncells, seed=3*ncells + prefer)
    pts = [t.problem_type for t in ts]
    tree = oracle.bisection_tree(ncells)
    ref = oracle.qlt(tree, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev, prefer)
    got, q = run_qlt_gpu(ncells, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev,
                         prefer=prefer, nrun=2, ring=ring)
    assert q.uses_fast_path()
    assert q.uses_ring() == ring
    assert np.array_equal(got, ref)
    assert R.check(ts, v, got, prefer) == []


@pytest.mark.parametrize("ncells,nt", [(86400, 24), (86400, 9), (5400, 700), (5400, 333),
                                       (8*768, 150), (2*1023, 41), (49152, 17)])
def test_ring_kernel_many_tracers_bitwise(oracle, ncells, nt):
    """The persistent ring kernel (one cooperative launch per class: pieces per CTA, units
    of several tracers, the UP and DOWN passes, per-tracer arrival counters and flags, the
    in-kernel sweep above the sub-roots) for QLT `cst` and CAAS, against the oracle and
    against the multi-launch path."""
    from compose_b200 import workloads as W
    from gpu_util import run_qlt_gpu, run_caas_gpu
    rhom, lo, q, hi, prev = W.headline(ncells, nt, 3)
    pts = [7]*nt
    tree = oracle.bisection_tree(ncells)
    ref = oracle.qlt(tree, pts, rhom, lo, q, hi, prev)
    got, c = run_qlt_gpu(ncells, pts, rhom, lo, q, hi, prev, nrun=2, ring=True)
    assert c.uses_ring()
    assert c.last_run_launches() <= 5
    assert np.array_equal(got, ref)
    ref = oracle.caas(ncells, pts, lo, q, hi, prev, tree=tree)
    got, c = run_caas_gpu(ncells, pts, rhom, lo, q, hi, prev, ring=True)
    assert c.uses_ring() and c.last_run_launches() == 1
    assert np.array_equal(got, ref)
    got2, c2 = run_caas_gpu(ncells, pts, rhom, lo, q, hi, prev, ring=False)
    assert not c2.uses_ring()
    assert np.array_equal(got, got2)


def test_ring_mixed_classes_and_noop_inputs(oracle):
    """st and cst tracers go through the ring kernel, the other four classes through the
    multi-launch kernels, in one run(); and inputs already in bounds with Qm == Qm_prev must
    come back bit-for-bit (the quick exit, cedr_qlt_inl.hpp:145-160)."""
    from compose_b200 import workloads as W
    from gpu_util import run_qlt_gpu
    ncells = 5400
    ts, v = R.generate(ncells, seed=99)
    pts = [t.problem_type for t in ts]
    tree = oracle.bisection_tree(ncells)
    ref = oracle.qlt(tree, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev)
    got, c = run_qlt_gpu(ncells, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev, ring=True)
    assert c.uses_ring()
    assert np.array_equal(got, ref)
    nt = 40
    rhom, lo, q, hi, prev = W.headline(ncells, nt, 5)
    got, c = run_qlt_gpu(ncells, [7]*nt, rhom, lo, prev, hi, prev, ring=True)
    assert np.array_equal(got, prev)


@pytest.mark.parametrize("ring", [False, True])
@pytest.mark.parametrize("ncells", [5400, 8*768])
def test_caas_without_conserving_tracers_exact_buffers(oracle, ncells, ring):
    """A CAAS whose tracers all lack `conserve` allots three rows per tracer
    (cedr_caas.cpp:86-100): the fast kernels must not stage a fourth. The caller's buffers
    are sized exactly to get_buffers_sizes() at the very end of an allocation."""
    import torch
    import compose_b200 as cb
    from compose_b200 import workloads as W
    nt = 7
    rhom, lo, q, hi, prev = W.headline(ncells, nt, 4)
    pts = [cb.SHAPEPRESERVE]*nt
    ref = oracle.caas(ncells, pts, lo, q, hi, prev, tree=oracle.bisection_tree(ncells))
    c = cb.CAAS(ncells)
    c.set_ring(ring)
    for p in pts:
        c.declare_tracer(p)
    c.end_tracer_declarations()
    b1, b2 = c.get_buffers_sizes()
    assert b1 == (1 + 3*nt)*((ncells + 15)//16*16)
    # The buffer is the tail of a larger allocation: any read past it leaves the allocation.
    pool = torch.zeros(b1 + 4096, dtype=torch.float64, device="cuda")
    buf = pool[4096:]
    c.set_buffers(buf, torch.zeros(1, dtype=torch.float64, device="cuda"))
    c.finish_setup()
    assert c.uses_fast_path() and c.uses_ring() == ring
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    c.set_rhom(dev(rhom))
    c.set_Qm(dev(q), dev(lo), dev(hi), None)
    c.run()
    c.synchronize()
    assert np.array_equal(c.get_Qm().cpu().numpy(), ref)


@pytest.mark.parametrize("ncells", [5400, 2*1023, 8*513])
@pytest.mark.parametrize("prefer", [False, True])
def test_one_field_classes_fast_and_generic_kernels_agree(oracle, ncells, prefer, monkeypatch):
    """The consistent-only (t, ct) and nonnegative (nn, cnn) classes,
    cedr_qlt_inl.hpp:175-197, run fast::up_kernel + fast::down1_kernel on fast-shaped
    blocks; CEDR_B200_FAST_ST_ONLY sends them through the generic sweep instead. Both must
    equal the oracle bit for bit, for many tracers per class (several per CTA group)."""
    import compose_b200 as cb
    from gpu_util import run_qlt_gpu
    ts, v = R.generate(ncells, seed=11*ncells + prefer)
    S, C, T, N = cb.SHAPEPRESERVE, cb.CONSERVE, cb.CONSISTENT, cb.NONNEGATIVE
    keep = [i for i, t in enumerate(ts) if t.problem_type in (T, C | T, N, C | N)]
    assert len(keep) == 24
    keep = keep*3       # 72 tracers: 18 per class
    pts = [ts[i].problem_type for i in keep]
    lo, q, hi, prev = (np.ascontiguousarray(a[keep]) for a in
                       (v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev))
    tree = oracle.bisection_tree(ncells)
    ref = oracle.qlt(tree, pts, v.rhom, lo, q, hi, prev, prefer)
    got, c = run_qlt_gpu(ncells, pts, v.rhom, lo, q, hi, prev, prefer=prefer, nrun=2)
    assert c.uses_fast_path()
    assert np.array_equal(got, ref)
    monkeypatch.setenv("CEDR_B200_FAST_ST_ONLY", "1")
    got2, c2 = run_qlt_gpu(ncells, pts, v.rhom, lo, q, hi, prev, prefer=prefer)
    assert np.array_equal(got2, ref)


def test_fast_and_generic_paths_agree_on_headline_inputs(oracle):
    import torch
    import compose_b200 as cb
    from compose_b200 import workloads as W
    ncells, nt = 5400, 40
    rhom, lo, q, hi, prev = (torch.from_numpy(x).cuda() for x in W.headline(ncells, nt, 1))
    outs = []
    for fast in (True, False):
        for kind in ("qlt", "caas"):
            c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
            c.set_fast_path(fast)
            for _ in range(nt):
                c.declare_tracer(7)
            c.end_tracer_declarations()
            c.finish_setup()
            assert c.uses_fast_path() == fast
            c.set_rhom(rhom)
            c.set_Qm(q, lo, hi, prev)
            c.run()
            outs.append(c.get_Qm().cpu().numpy())
    assert np.array_equal(outs[0], outs[2])
    assert np.array_equal(outs[1], outs[3])


# ---------------------------------------------------------------- multi-rank
#
# P ranks emulated on one device: P CDR objects (rank r of P), each fed only its own
# cells; run() is split at the exchange (run_phase) and the test plays the all-gather.
# The single-rank oracle on the whole problem is the reference (SURVEY 8c "multi-rank
# reference"): the partition must not change a single bit.

def _emulate_exchange(cdrs):
    import torch
    for c in cdrs:
        c.run_phase(0)
    torch.cuda.synchronize()
    msg = torch.cat([c._xsend for c in cdrs])
    for c in cdrs:
        assert c._xrecv.numel() == msg.numel()
        c._xrecv.copy_(msg)
    for c in cdrs:
        c.run_phase(1)
    for c in cdrs:
        c.synchronize()


def run_qlt_multirank(ncells, P, pts, rhom, qm_min, qm, qm_max, qm_prev, mbl=None,
                      prefer=False):
    import torch
    import compose_b200 as cb
    cdrs, gcis = [], []
    for r in range(P):
        q = cb.QLT(ncells, rank=r, nranks=P,
                   prefer_numerical_mass_conservation_to_numerical_bounds=prefer)
        if mbl:
            q.set_max_block_leaves(mbl)
        for p in pts:
            q.declare_tracer(int(p))
        q.end_tracer_declarations()
        q.use_tensor_exchange_buffers(P)
        q.finish_setup()
        g = q.get_owned_glblcells()
        assert len(g) == q.nlclcells()
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a)[..., g])).cuda()
        q.set_rhom(dev(rhom))
        q.set_Qm(dev(qm), dev(qm_min), dev(qm_max), dev(qm_prev))
        cdrs.append(q)
        gcis.append(g)
    assert sorted(np.concatenate(gcis).tolist()) == list(range(ncells))
    _emulate_exchange(cdrs)
    res = np.empty((len(pts), ncells))
    for q, g in zip(cdrs, gcis):
        res[:, g] = q.get_Qm().cpu().numpy()
    return res


@pytest.mark.parametrize("ncells,P,mbl", [(5400, 2, None), (5400, 8, None), (8*768, 4, None),
                                          (64, 2, 4), (64, 8, 8), (1024, 4, 32),
                                          (86400, 8, None)])
def test_qlt_subtree_partition_bitwise(oracle, ncells, P, mbl):
    ts, v = R.generate(ncells, seed=7*ncells + P)
    if ncells > 10000:
        ts = ts[:12]
    pts = [t.problem_type for t in ts]
    n = len(pts)
    tree = oracle.bisection_tree(ncells)
    ref = oracle.qlt(tree, pts, v.rhom, v.Qm_min[:n], v.Qm[:n], v.Qm_max[:n], v.Qm_prev[:n])
    got = run_qlt_multirank(ncells, P, pts, v.rhom, v.Qm_min[:n], v.Qm[:n], v.Qm_max[:n],
                            v.Qm_prev[:n], mbl=mbl)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("ncells,P,mbl", [(5400, 2, None), (5400, 8, None), (64, 4, 4),
                                          (86400, 8, None)])
def test_caas_subtree_partition_bitwise(oracle, ncells, P, mbl):
    import torch
    import compose_b200 as cb
    from compose_b200 import workloads as W
    nt = 10
    rhom, lo, q, hi, prev = W.headline(ncells, nt, 11)
    pts = [7]*nt
    ref = oracle.caas(ncells, pts, lo, q, hi, prev, tree=oracle.bisection_tree(ncells))
    nl = ncells//P
    cdrs = []
    for r in range(P):
        c = cb.CAAS(nl, cell0=r*nl, ncells_global=ncells, rank=r, nranks=P)
        if mbl:
            c.set_max_block_leaves(mbl)
        for p in pts:
            c.declare_tracer(p)
        c.end_tracer_declarations()
        c.use_tensor_exchange_buffers(P)
        c.finish_setup()
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a[..., r*nl:(r + 1)*nl])).cuda()
        c.set_rhom(dev(rhom))
        c.set_Qm(dev(q), dev(lo), dev(hi), dev(prev))
        cdrs.append(c)
    _emulate_exchange(cdrs)
    got = np.concatenate([c.get_Qm().cpu().numpy() for c in cdrs], axis=1)
    assert np.array_equal(got, ref)


def test_partition_must_be_whole_blocks():
    import compose_b200 as cb
    # 3 ranks over 5400 cells: 1800-cell ranges cut the 675-leaf blocks. CAAS's built-in
    # tree-ordered sums need whole blocks (any cell sets work with a UserAllReducer); QLT
    # switches to its replicated mode (next test).
    c = cb.CAAS(1800, cell0=1800, ncells_global=5400, rank=1, nranks=3)
    c.declare_tracer(7)
    with pytest.raises(cb.CedrError, match="whole blocks"):
        c.end_tracer_declarations()
    q = cb.QLT(5400, rank=1, nranks=3)
    q.declare_tracer(7)
    q.end_tracer_declarations()
    assert "replicated mode" in q.print()


def _rank_of_cell(kind, ci, ncells, P):
    """oned::Mesh::rank, cedr_tree.cpp:366-375."""
    if kind == "pseudorandom":
        return (ci + ci//P) % P
    return min(P - 1, ci//(ncells//P))


@pytest.mark.parametrize("P", [2, 3, 8])
@pytest.mark.parametrize("mult", [1, 2, 7, 21, 167])
@pytest.mark.parametrize("decomp", ["contiguous", "pseudorandom"])
@pytest.mark.parametrize("imbalanced", [False, True])
def test_qlt_general_rank_maps_bitwise(oracle, P, mult, decomp, imbalanced):
    """The reference's own multi-rank sweep (cedr_qlt.cpp:745-772): ncells = {1, 2, 7,
    21} nranks (and a larger one) x {contiguous, pseudorandom cell -> rank map,
    cedr_tree.cpp:366-375} x {balanced, imbalanced tree} x prefer_mass_con, P ranks emulated
    on one device. Maps that cut blocks of the tree plan run in replicated mode; the
    single-rank oracle on the whole problem is the reference, bit for bit."""
    import torch
    import compose_b200 as cb
    ncells = mult*P
    prefer = (mult + P) % 2 == 1
    ts, v = R.generate(ncells, seed=31*ncells + P)
    pts = [t.problem_type for t in ts]
    tree = oracle.bisection_tree(ncells, imbalanced)
    ref = oracle.qlt(tree, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev, prefer)
    kids, cellidx, root = cb.make_1d_tree(ncells, imbalanced)
    node_rank = np.array([_rank_of_cell(decomp, int(ci), ncells, P) if ci >= 0 else 0
                          for ci in cellidx], np.int32)
    cdrs, gcis = [], []
    for r in range(P):
        q = cb.QLT(ncells, tree=(kids, cellidx, root), node_rank=node_rank, rank=r, nranks=P,
                   prefer_numerical_mass_conservation_to_numerical_bounds=prefer)
        for p in pts:
            q.declare_tracer(int(p))
        q.end_tracer_declarations()
        q.use_tensor_exchange_buffers(P)
        q.finish_setup()
        g = q.get_owned_glblcells()
        assert sorted(g.tolist()) == [ci for ci in range(ncells)
                                      if _rank_of_cell(decomp, ci, ncells, P) == r]
        for i in range(0, len(g), max(1, len(g)//5)):
            assert q.gci2lci(int(g[i])) == i
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a)[..., g])).cuda()
        q.set_rhom(dev(v.rhom))
        q.set_Qm(dev(v.Qm), dev(v.Qm_min), dev(v.Qm_max), dev(v.Qm_prev))
        cdrs.append(q)
        gcis.append(g)
    _emulate_exchange(cdrs)
    res = np.empty((len(pts), ncells))
    for q, g in zip(cdrs, gcis):
        res[:, g] = q.get_Qm().cpu().numpy()
    assert np.array_equal(res, ref)
    assert R.check(ts, v, res, prefer) == []


# ------------------------------------------------------------------ config 5 (1-D transport)

def test_transport1d_every_step_bitwise(oracle):
    """BASELINE.json config 5: the advection loop of cedr_test_1d_transport.cpp calling
    CDR::run every step (many tiny runs), for qltnn / qlt / caas. Each step's CUDA result
    must equal the oracle's bit for bit, so the two 351-step trajectories never part."""
    import torch
    import compose_b200 as cb
    import transport1d as T
    from test_oracle_golden import t1d_runners
    ncells = 111
    p, oruns = t1d_runners(oracle, ncells, caas_tree=True)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a[None])).cuda()

    def gpu_runner(kind, pt):
        c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
        c.declare_tracer(pt)
        c.end_tracer_declarations()
        c.finish_setup()
        g = c.get_owned_glblcells() if kind == "qlt" else np.arange(ncells)
        c.set_rhom(torch.from_numpy(np.ascontiguousarray(p.area[g])).cuda())

        def run(q, lo, hi, prev):
            c.set_Qm(dev(q[g]), dev(lo[g]), dev(hi[g]), dev(prev[g]))
            c.run()
            out = np.empty(ncells)
            out[g] = c.get_Qm().cpu().numpy()[0]
            return out
        return run

    gruns = {"yqltnn": gpu_runner("qlt", 1 | 8), "yqlt": gpu_runner("qlt", 1 | 2),
             "ycaas": gpu_runner("caas", 3)}
    nsteps = int(3.17*ncells)
    y0 = p.y0()
    for name in ("yqltnn", "yqlt", "ycaas"):
        def both(q, lo, hi, prev, name=name):
            a = gruns[name](q, lo, hi, prev)
            b = oruns[name](q, lo, hi, prev)
            assert np.array_equal(a, b), name
            return a
        yf = p.cycle(nsteps, y0, both)
        # mass is conserved over the whole run to rounding
        m0, m1 = (y0[:ncells]*p.area).sum(), (yf[:ncells]*p.area).sum()
        assert abs(m1 - m0) <= 1e-12*abs(m0)


@pytest.mark.parametrize("ncells", [1, 2, 11, 111, 1000, 5400])
def test_caas_sequential_sums_bitwise(oracle, ncells):
    """CEDR_B200_CAAS_SUM_SEQUENTIAL reproduces the reference's DEFAULT CAAS (host-order
    sums, cedr_caas.cpp:171-199) bit for bit."""
    import torch
    import compose_b200 as cb
    ts, v = R.generate(ncells, seed=900 + ncells)
    sel = caas_tracers()
    idx = [t.idx for t in sel]
    pts = [t.problem_type for t in sel]
    a = [x[idx] for x in (v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev)]
    ref = oracle.caas(ncells, pts, *a)          # tree=None: sequential sums
    c = cb.CAAS(ncells, sum_mode=cb.CAAS_SUM_SEQUENTIAL)
    for p in pts:
        c.declare_tracer(int(p))
    c.end_tracer_declarations()
    c.finish_setup()
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    c.set_rhom(dev(v.rhom))
    c.set_Qm(dev(a[1]), dev(a[0]), dev(a[2]), dev(a[3]))
    c.run()
    c.synchronize()
    assert np.array_equal(c.get_Qm().cpu().numpy(), ref)


# ------------------------------------------------------------- BfbTreeAllReducer (8f-2)

@pytest.mark.parametrize("nleaf,nfield,mbl", [(1, 3, 0), (2, 1, 0), (17, 5, 0), (111, 4, 8),
                                              (1000, 7, 32), (5400, 16, 0)])
@pytest.mark.parametrize("transpose", [False, True])
def test_bfb_tree_allreduce_bitwise(oracle, nleaf, nfield, mbl, transpose):
    """The device-resident BfbTreeAllReducer against the reference's (through the pinned
    oracle, oracle_bfb_allreduce): same tree order, same bits; and within the reference's
    own unit-test tolerance of a plain sum (cedr_bfb_tree_allreduce.cpp:211-217)."""
    import torch
    import compose_b200 as cb
    rng = np.random.default_rng(nleaf + nfield)
    data = rng.standard_normal((nleaf, nfield))*10.0**rng.integers(-3, 4, (nleaf, nfield))
    send = np.ascontiguousarray(data.T if transpose else data).reshape(-1)
    tree = oracle.bisection_tree(nleaf)
    ref = oracle.bfb_allreduce(tree, send, nfield, transpose)
    r = cb.BfbTreeAllReducer(nleaf, nfield, max_block_leaves=mbl)
    got = r.allreduce(torch.from_numpy(send).cuda(), transpose=transpose)
    r.synchronize()
    got = got.cpu().numpy()
    assert np.array_equal(got, ref)
    tol = 2*np.log(max(nleaf, 2))*np.finfo(float).eps*np.abs(data).sum(0)
    assert np.all(np.abs(got - data.sum(0)) <= tol + 1e-300)


def test_bfb_tree_allreduce_partition_invariant(oracle):
    """4 emulated ranks, each with its own leaves: same bits as one rank."""
    import torch
    import compose_b200 as cb
    nleaf, nfield, P, mbl = 4096, 6, 4, 64
    rng = np.random.default_rng(3)
    data = rng.standard_normal((nleaf, nfield))
    tree = oracle.bisection_tree(nleaf)
    ref = oracle.bfb_allreduce(tree, data.reshape(-1), nfield, False)
    nl = nleaf//P
    rs, recvs = [], []
    for r in range(P):
        red = cb.BfbTreeAllReducer(nleaf, nfield, max_block_leaves=mbl, rank=r, nranks=P)
        red.use_tensor_exchange_buffers(P)
        rs.append(red)
    sends = [torch.from_numpy(np.ascontiguousarray(data[r*nl:(r + 1)*nl]).reshape(-1)).cuda()
             for r in range(P)]
    recvs = [torch.empty(nfield, dtype=torch.float64, device="cuda") for _ in range(P)]
    for red, s, rv in zip(rs, sends, recvs):
        red.allreduce(s, rv, phase=0)
    torch.cuda.synchronize()
    msg = torch.cat([red._xsend for red in rs])
    for red in rs:
        red._xrecv.copy_(msg)
    for red, s, rv in zip(rs, sends, recvs):
        red.allreduce(s, rv, phase=1)
    torch.cuda.synchronize()
    for rv in recvs:
        assert np.array_equal(rv.cpu().numpy(), ref)


# ------------------------------------------------ full BASELINE size, size-independent properties

def test_headline_full_size_properties():
    """ne120 x 128 levels x 40 tracers (86,400 cells x 5,120 CDR tracers, BASELINE.json's
    headline config), where the oracle would take minutes: check the properties the
    reference's own tests check (cedr_test_randomized.cpp:293-418) on the device --
    bounds exact, per-tracer mass conserved to 100 eps -- plus idempotence (a second
    run on the solution, with Qm_prev unchanged, returns it bit for bit through the quick
    exit) and path independence (split / unsplit tier plans give identical bits on a
    slice)."""
    import os
    import torch
    import compose_b200 as cb
    from compose_b200.workloads import CONFIGS
    ncells, nt, cid = CONFIGS["ne120x128x40"]
    rhom, lo, q, hi, prev = cb.fill_headline(ncells, nt, cid)
    eps = np.finfo(float).eps

    def build(kind, ntr):
        c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
        for _ in range(ntr):
            c.declare_tracer(7)
        c.end_tracer_declarations()
        c.finish_setup()
        c.set_rhom(rhom)
        return c

    for kind in ("qlt", "caas"):
        c = build(kind, nt)
        c.set_Qm(q, lo, hi, prev)
        c.run()
        out = c.get_Qm()
        c.synchronize()
        assert bool((out >= lo).all()) and bool((out <= hi).all()), kind
        mass = (out.sum(1) - prev.sum(1)).abs()/prev.abs().sum(1)
        assert float(mass.max()) <= 100*eps, (kind, float(mass.max()))
        # idempotence: the solution is a fixed point
        c.set_Qm(out, lo, hi, prev)
        c.run()
        out2 = c.get_Qm()
        c.synchronize()
        if kind == "qlt":
            # sum(out) == sum(prev) only to rounding, so the root mass may move by an ulp;
            # the fixed point is exact when Qm_prev is the solution itself.
            c.set_Qm(out, lo, hi, out)
            c.run()
            out2 = c.get_Qm()
            c.synchronize()
            assert torch.equal(out2, out)
        else:
            assert float((out2 - out).abs().max()) <= 1e-13*float(out.abs().max())
        del c, out, out2
        torch.cuda.empty_cache()

    # path independence on the first 256 tracers
    n2 = 256
    outs = []
    for env in ({}, {"CEDR_B200_NO_SPLIT": "1"}, {"CEDR_B200_NO_FAST": "1"}):
        os.environ.update(env)
        try:
            c = build("qlt", n2)
            c.set_Qm(q[:n2], lo[:n2], hi[:n2], prev[:n2])
            c.run()
            outs.append(c.get_Qm().clone())
            c.synchronize()
        finally:
            for k in env:
                del os.environ[k]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_qlt_block_granular_rank_map_bitwise(oracle):
    """Any assignment of WHOLE blocks to ranks works, not only contiguous ranges: 3 ranks,
    the 16 blocks of a 16 x 675-cell mesh dealt round-robin with an uneven remainder
    (6 / 5 / 5), through a caller-provided tree with per-leaf ranks."""
    import torch
    import compose_b200 as cb
    ncells, P = 16*675, 3
    tree = oracle.bisection_tree(ncells)
    kids, cellidx = tree.kids, tree.cellidx
    rank = np.zeros(cellidx.size, np.int32)
    leaf = cellidx >= 0
    rank[leaf] = (cellidx[leaf]//675) % P
    ts, v = R.generate(ncells, seed=4242)
    ts = ts[:8]
    pts = [t.problem_type for t in ts]
    n = len(pts)
    ref = oracle.qlt(tree, pts, v.rhom, v.Qm_min[:n], v.Qm[:n], v.Qm_max[:n], v.Qm_prev[:n])
    cdrs, gcis = [], []
    for r in range(P):
        q = cb.QLT(ncells, tree=(kids, cellidx, 0), rank=r, nranks=P, node_rank=rank)
        for p in pts:
            q.declare_tracer(int(p))
        q.end_tracer_declarations()
        q.use_tensor_exchange_buffers(P)
        q.finish_setup()
        g = q.get_owned_glblcells()
        assert q.nlclcells() == len(g) == 675*(6 if r == 0 else 5)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a)[..., g])).cuda()
        q.set_rhom(dev(v.rhom))
        q.set_Qm(dev(v.Qm[:n]), dev(v.Qm_min[:n]), dev(v.Qm_max[:n]), dev(v.Qm_prev[:n]))
        cdrs.append(q)
        gcis.append(g)
    _emulate_exchange(cdrs)
    res = np.empty((n, ncells))
    for q, g in zip(cdrs, gcis):
        res[:, g] = q.get_Qm().cpu().numpy()
    assert np.array_equal(res, ref)


# ---------------------------------------------------------------- full sizes vs the oracle

@pytest.mark.parametrize("workload,nt", [("ne120x128x40", 64), ("ne256x128x10", 8)])
def test_default_path_full_size_bitwise_vs_oracle(oracle, workload, nt):
    """The DEFAULT single-rank path at the cell counts of BASELINE.json's configs 3 and 4
    (86,400 and 393,216 cells), on a slice of the workload's own tracers, bit for bit
    against the oracle (QLT, and CAAS with tree-ordered sums). Tracers are independent
    problems, so a slice of them is checked exactly as the whole would be."""
    import torch
    import compose_b200 as cb
    from compose_b200 import workloads as W
    ncells, _, cid = W.CONFIGS[workload]
    rhom, lo, q, hi, prev = W.headline(ncells, nt, cid)
    pts = [7]*nt
    tree = oracle.bisection_tree(ncells)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d = [dev(x) for x in (rhom, lo, q, hi, prev)]
    for kind in ("qlt", "caas"):
        c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
        for p in pts:
            c.declare_tracer(p)
        c.end_tracer_declarations()
        c.finish_setup()
        assert c.uses_fast_path() and not c.uses_ring()
        c.set_rhom(d[0])
        c.set_Qm(d[2], d[1], d[3], d[4])
        c.run()
        c.synchronize()
        got = c.get_Qm().cpu().numpy()
        ref = (oracle.qlt(tree, pts, rhom, lo, q, hi, prev) if kind == "qlt"
               else oracle.caas(ncells, pts, lo, q, hi, prev, tree=tree))
        assert np.array_equal(got, ref), kind


def test_against_the_reference_itself_when_present(oracle):
    """Straight against the UNMODIFIED reference (oracle/_ref, built where /root/reference
    exists and shipped to the GPU box), not through the C restatement: QLT bit for bit;
    CAAS bit for bit against the reference CAAS driven through its own BfbTreeAllReducer
    (tree-ordered sums), and its default sequential sums against our
    CAAS_SUM_SEQUENTIAL mode."""
    from oracle.oracle_py import Ref, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref is not built on this box")
    import compose_b200 as cb
    from compose_b200 import workloads as W
    from gpu_util import run_qlt_gpu, run_caas_gpu
    r = Ref()
    for ncells, nt in ((5400, 24), (86400, 6)):
        rhom, lo, q, hi, prev = W.headline(ncells, nt, 2)
        pts = [7]*nt
        ref, _, _ = r.qlt(ncells, ("bisect", False), pts, rhom, lo, q, hi, prev)
        got, _ = run_qlt_gpu(ncells, pts, rhom, lo, q, hi, prev)
        assert np.array_equal(got, ref)
        ref, _ = r.caas(ncells, pts, rhom, lo, q, hi, prev, tree=("bisect", False))
        got, _ = run_caas_gpu(ncells, pts, rhom, lo, q, hi, prev)
        assert np.array_equal(got, ref)
    # the randomized six-class set, balanced and imbalanced trees, both options
    for imb in (False, True):
        for prefer in (False, True):
            ncells = 1000
            ts, v = R.generate(ncells, seed=17 + imb + 2*prefer)
            pts = [t.problem_type for t in ts]
            ref, _, _ = r.qlt(ncells, ("bisect", imb), pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max,
                              v.Qm_prev, prefer_mass_con=prefer)
            got, _ = run_qlt_gpu(ncells, pts, v.rhom, v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev,
                                 imbalanced=imb, prefer=prefer)
            assert np.array_equal(got, ref)


# ---------------------------------------------------------------- CAAS UserAllReducer

def _numpy_caas(pts, lo, q, hi, prev, sums):
    """CAAS::run (cedr_caas.cpp:129-253) in numpy given the four per-tracer global sums
    `sums(values[nt, n]) -> [nt]` in whatever order the reducer under test uses."""
    pts = np.asarray(pts)
    clip = np.minimum(hi, np.maximum(lo, q))
    term = np.where((pts & 1)[:, None] != 0, prev, q)
    s_clip, s_term, s_min, s_max = sums(clip), sums(term), sums(lo), sums(hi)
    out = clip.copy()
    m = s_term - s_clip
    for t in range(len(pts)):
        if m[t] < 0:
            fac = s_clip[t] - s_min[t]
            if fac > 0:
                fac = m[t]/fac
                out[t] = np.maximum(lo[t], clip[t] + fac*(clip[t] - lo[t]))
        elif m[t] > 0:
            fac = s_max[t] - s_clip[t]
            if fac > 0:
                fac = m[t]/fac
                out[t] = np.minimum(hi[t], clip[t] + fac*(hi[t] - clip[t]))
    return out


@pytest.mark.parametrize("ncells,n_accum", [(11, 1), (1350, 1), (1350, 3), (5400, 8)])
def test_caas_user_reducer_sequential(oracle, ncells, n_accum):
    """A UserAllReducer that sums its nlocal partials one after the other on the host, as
    the reference's own TestAllReducer does (cedr_caas.cpp:276-300): with n_accum = 1 that
    is the reference's default summation order (bitwise the oracle's sequential mode and
    our CAAS_SUM_SEQUENTIAL); with n_accum > 1 the partials are block sums, checked
    against a numpy CAAS that sums in the same order."""
    import torch
    import compose_b200 as cb
    from compose_b200 import workloads as W
    nt = 6
    rhom, lo, q, hi, prev = W.headline(ncells, nt, 7)
    pts = [7, 3, 2, 7, 6, 3]
    calls = []

    def reducer(send, recv, nlocal, nfld):
        calls.append((nlocal, nfld))
        h = send.cpu().numpy()
        recv.copy_(torch.from_numpy(np.add.accumulate(h, axis=1)[:, -1].copy()).cuda())

    c = cb.CAAS(ncells, user_reducer=reducer, n_accum=n_accum)
    for p in pts:
        c.declare_tracer(p)
    c.end_tracer_declarations()
    b1, b2 = c.get_buffers_sizes()
    assert b2 == 4*nt*(ncells//n_accum + 1)       # send + recv, cedr_caas.cpp:75-90
    c.finish_setup()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    c.set_rhom(dev(rhom))
    c.set_Qm(dev(q), dev(lo), dev(hi), dev(prev))
    c.run()
    c.synchronize()
    got = c.get_Qm().cpu().numpy()
    assert calls == [(ncells//n_accum, 4*nt)]

    def sums(v):
        blocks = np.add.accumulate(v.reshape(nt, ncells//n_accum, n_accum), axis=2)[:, :, -1]
        return np.add.accumulate(blocks, axis=1)[:, -1]
    assert np.array_equal(got, _numpy_caas(pts, lo, q, hi, prev, sums))
    if n_accum == 1:
        assert np.array_equal(got, oracle.caas(ncells, pts, lo, q, hi, prev, tree=None))


@pytest.mark.parametrize("split", ["ragged", "pseudorandom"])
def test_caas_any_cell_sets_per_rank_through_the_reducer(oracle, split):
    """cedr_caas.cpp:37-48 takes only nlclcells: a rank's cells may be any set, the
    cross-rank sum is the reducer's (default MPI_Allreduce of one sequential partial per
    rank, :203-209). Three ranks emulated on one device with unequal, non-block-aligned
    (or interleaved) cell sets and n_accum = nlclcells; the reducer plays the all-reduce
    (partials added in rank order). Checked against a numpy CAAS summing in that order."""
    import torch
    import compose_b200 as cb
    from compose_b200 import workloads as W
    ncells, nt, P = 2731, 5, 3
    rhom, lo, q, hi, prev = W.headline(ncells, nt, 13)
    pts = [7, 3, 2, 7, 3]
    if split == "ragged":
        cuts = [0, 700, 701, ncells]
        own = [np.arange(cuts[r], cuts[r + 1]) for r in range(P)]
    else:
        ci = np.arange(ncells)
        rank_of = (ci + ci//P) % P          # cedr_tree.cpp:366-375
        own = [ci[rank_of == r] for r in range(P)]
    partials, total = {}, {}

    def make(r, phase):
        def reducer(send, recv, nlocal, nfld):
            assert (nlocal, nfld) == (1, 4*nt)
            if phase == 0:
                partials[r] = send[:, 0].clone()
                recv.copy_(send[:, 0])
            else:
                recv.copy_(total["v"])
        return reducer

    def run(phase):
        outs = []
        for r in range(P):
            g = own[r]
            c = cb.CAAS(len(g), user_reducer=make(r, phase), n_accum=len(g))
            for p in pts:
                c.declare_tracer(p)
            c.end_tracer_declarations()
            c.finish_setup()
            dev = lambda a: torch.from_numpy(np.ascontiguousarray(a[..., g])).cuda()
            c.set_rhom(dev(rhom))
            c.set_Qm(dev(q), dev(lo), dev(hi), dev(prev))
            c.run()
            c.synchronize()
            outs.append(c.get_Qm().cpu().numpy())
        return outs
    run(0)
    t = partials[0].clone()
    for r in range(1, P):
        t = t + partials[r]
    total["v"] = t
    outs = run(1)
    got = np.empty((nt, ncells))
    for r in range(P):
        got[:, own[r]] = outs[r]

    def sums(v):
        per_rank = [np.add.accumulate(v[:, own[r]], axis=1)[:, -1] for r in range(P)]
        tot = per_rank[0]
        for r in range(1, P):
            tot = tot + per_rank[r]
        return tot
    assert np.array_equal(got, _numpy_caas(pts, lo, q, hi, prev, sums))


def test_caas_user_reducer_backed_by_bfb_tree_allreducer(oracle):
    """the public BfbTreeAllReducer as the CAAS UserAllReducer (transpose =
    True is CAAS's (nlocal, nfld) send layout, cedr_caas.cpp:153-154): same bits as the
    built-in tree-ordered sums and as the oracle."""
    import torch
    import compose_b200 as cb
    from compose_b200 import workloads as W
    from gpu_util import run_caas_gpu
    ncells, nt = 5400, 9
    rhom, lo, q, hi, prev = W.headline(ncells, nt, 8)
    pts = [7, 3]*4 + [7]
    bfb = cb.BfbTreeAllReducer(ncells, 4*nt)

    def reducer(send, recv, nlocal, nfld):
        assert (nlocal, nfld) == (ncells, 4*nt)
        bfb.allreduce(send.reshape(-1), recv, transpose=True)

    c = cb.CAAS(ncells, user_reducer=reducer)
    for p in pts:
        c.declare_tracer(p)
    c.end_tracer_declarations()
    c.finish_setup()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    c.set_rhom(dev(rhom))
    c.set_Qm(dev(q), dev(lo), dev(hi), dev(prev))
    c.run()
    c.synchronize()
    got = c.get_Qm().cpu().numpy()
    ref = oracle.caas(ncells, pts, lo, q, hi, prev, tree=oracle.bisection_tree(ncells))
    assert np.array_equal(got, ref)
    got2, _ = run_caas_gpu(ncells, pts, rhom, lo, q, hi, prev)
    assert np.array_equal(got, got2)


@pytest.mark.parametrize("kind", ["qlt", "caas"])
@pytest.mark.parametrize("bound", [False, True])
def test_run_replayed_as_cuda_graph_bitwise(oracle, kind, bound):
    """cedr_b200_set_graph(1): run() is captured once and replayed. Every replay must read
    the inputs of ITS step (they change between runs) and give the oracle's bits; the launch
    count reported stays that of the plain run."""
    import torch
    import compose_b200 as cb
    from compose_b200 import workloads as W
    ncells, nt = 5400, 12
    pts = [7, 3]*(nt//2)
    tree = oracle.bisection_tree(ncells)
    c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
    c.set_graph(1)
    for p in pts:
        c.declare_tracer(p)
    c.end_tracer_declarations()
    c.finish_setup()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if bound:
        arrs = [torch.empty((nt, ncells), dtype=torch.float64, device="cuda") for _ in range(5)]
        lo_d, q_d, hi_d, prev_d, out_d = arrs
        if kind == "qlt":
            c.bind_arrays(q_d, lo_d, hi_d, prev_d, out=out_d)
        else:
            c.bind_arrays(q_d, lo_d, hi_d, prev_d)
    launches = None
    for step in range(5):
        rhom, lo, q, hi, prev = W.headline(ncells, nt, 40 + step)
        ref = (oracle.qlt(tree, pts, rhom, lo, q, hi, prev) if kind == "qlt"
               else oracle.caas(ncells, pts, lo, q, hi, prev, tree=tree))
        c.set_rhom(dev(rhom))
        if bound:
            for d, h in zip((lo_d, q_d, hi_d, prev_d), (lo, q, hi, prev)):
                d.copy_(dev(h))
        else:
            c.set_Qm(dev(q), dev(lo), dev(hi), dev(prev))
        c.run()
        c.synchronize()
        if bound:
            got = (out_d if kind == "qlt" else q_d).cpu().numpy()
        else:
            got = c.get_Qm().cpu().numpy()
        assert np.array_equal(got, ref), step
        assert c.uses_graph() == (step >= 1)
        if launches is None:
            launches = c.last_run_launches()
        assert c.last_run_launches() == launches
    c.set_graph(0)
    assert not c.uses_graph()
    c.run()
    c.synchronize()


@pytest.mark.parametrize("kind", ["qlt", "caas"])
def test_run_graph_auto_mode_bind_and_unbind(oracle, kind, monkeypatch):
    """Graph mode -1 (default) replays short one-rank runs by itself; binding and UNBINDING
    the caller's arrays drop the captured graph (its kernels carry the output array as an
    argument); a CDR above the auto threshold stays on plain launches."""
    import torch
    import compose_b200 as cb
    from compose_b200 import workloads as W
    ncells, nt = 5400, 10
    pts = [7]*nt
    tree = oracle.bisection_tree(ncells)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

    def make():
        c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
        for p in pts:
            c.declare_tracer(p)
        c.end_tracer_declarations()
        c.finish_setup()
        return c

    def reference(step):
        rhom, lo, q, hi, prev = W.headline(ncells, nt, 60 + step)
        ref = (oracle.qlt(tree, pts, rhom, lo, q, hi, prev) if kind == "qlt"
               else oracle.caas(ncells, pts, lo, q, hi, prev, tree=tree))
        return rhom, lo, q, hi, prev, ref

    c = make()
    arrs = [torch.empty((nt, ncells), dtype=torch.float64, device="cuda") for _ in range(5)]
    lo_d, q_d, hi_d, prev_d, out_d = arrs
    for step in range(9):
        rhom, lo, q, hi, prev, ref = reference(step)
        c.set_rhom(dev(rhom))
        phase = step//3                 # own buffers, bound arrays, own buffers again
        if step == 3:
            if kind == "qlt":
                c.bind_arrays(q_d, lo_d, hi_d, prev_d, out=out_d)
            else:
                c.bind_arrays(q_d, lo_d, hi_d, prev_d)
        if step == 6:
            c.bind_arrays(None, None, None)
        if phase == 1:
            for d, h in zip((lo_d, q_d, hi_d, prev_d), (lo, q, hi, prev)):
                d.copy_(dev(h))
            out_d.fill_(-1.0)
        else:
            c.set_Qm(dev(q), dev(lo), dev(hi), dev(prev))
        c.run()
        c.synchronize()
        if phase == 1:
            got = (out_d if kind == "qlt" else q_d).cpu().numpy()
        else:
            got = c.get_Qm().cpu().numpy()
        assert np.array_equal(got, ref), step
        assert c.uses_graph() == (step % 3 >= 1), step
    monkeypatch.setenv("CEDR_B200_GRAPH_AUTO_MAX", str(ncells*nt - 1))
    c = make()
    rhom, lo, q, hi, prev, ref = reference(0)
    c.set_rhom(dev(rhom))
    for _ in range(3):
        c.set_Qm(dev(q), dev(lo), dev(hi), dev(prev))
        c.run()
    c.synchronize()
    assert not c.uses_graph()
    assert np.array_equal(c.get_Qm().cpu().numpy(), ref)


# ---------------------------------------------------------------- zero-copy binding (8f-1)

@pytest.mark.parametrize("ncells", [111, 1000, 5400])
def test_bound_arrays_equal_set_get(oracle, ncells):
    """cedr_b200_bind_arrays: run() reads the caller's SoA arrays in place of set_Qm's copy
    and writes QLT's results to the caller's output array (CAAS in place on Qm): same bits
    as the set_Qm / run / get_Qm route and as the oracle, on the generic, the fast and
    the single-launch paths."""
    import torch
    import compose_b200 as cb
    from compose_b200 import workloads as W
    nt = 10
    rhom, lo, q, hi, prev = W.headline(ncells, nt, 9)
    pts = [7, 6]*5
    tree = oracle.bisection_tree(ncells)
    lda = (ncells + 15)//16*16
    def dev(a):
        t = torch.zeros((nt, lda), dtype=torch.float64, device="cuda")
        t[:, :ncells] = torch.from_numpy(np.ascontiguousarray(a))
        return t
    for kind in ("qlt", "caas"):
        ref = (oracle.qlt(tree, pts, rhom, lo, q, hi, prev) if kind == "qlt"
               else oracle.caas(ncells, [7, 3]*5, lo, q, hi, prev, tree=tree))
        c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
        for p in (pts if kind == "qlt" else [7, 3]*5):
            c.declare_tracer(p)
        c.end_tracer_declarations()
        c.finish_setup()
        g = c.get_owned_glblcells() if kind == "qlt" else np.arange(ncells)
        c.set_rhom(torch.from_numpy(np.ascontiguousarray(rhom[g])).cuda())
        d = [dev(a[:, g]) for a in (q, lo, hi, prev)]
        out = torch.zeros((nt, lda), dtype=torch.float64, device="cuda")
        if kind == "qlt":
            c.bind_arrays(d[0], d[1], d[2], d[3], out=out)
        else:
            c.bind_arrays(d[0], d[1], d[2], d[3])
        with pytest.raises(cb.CedrError, match="arrays are bound"):
            c.set_Qm(d[0], d[1], d[2], d[3])
        c.run()
        c.synchronize()
        res = (out if kind == "qlt" else d[0])[:, :ncells].cpu().numpy()
        got = np.empty_like(res)
        got[:, g] = res
        assert np.array_equal(got, ref), kind
        # inputs other than CAAS's Qm are untouched
        assert torch.equal(d[1][:, :ncells].cpu(), torch.from_numpy(np.ascontiguousarray(lo[:, g])))
        # unbind: back to the CDR's own buffer
        c.bind_arrays(None, None, None)
        d2 = [dev(a[:, g]) for a in (q, lo, hi, prev)]
        c.set_Qm(*d2)
        c.run()
        res = c.get_Qm().cpu().numpy()
        got[:, g] = res
        assert np.array_equal(got, ref), kind


def test_bind_arrays_rejects_what_it_cannot_read_in_place():
    import torch
    import compose_b200 as cb
    n = 64
    z = lambda: torch.zeros((2, n), dtype=torch.float64, device="cuda")
    q = cb.QLT(n)
    q.declare_tracer(cb.CONSERVE | cb.CONSISTENT)     # consistent-only: scaled bounds
    q.declare_tracer(7)
    q.end_tracer_declarations()
    q.finish_setup()
    with pytest.raises(cb.CedrError, match="shape-preserving"):
        q.bind_arrays(z(), z(), z(), z(), out=z())
    q = cb.QLT(n)
    q.declare_tracer(7)
    q.declare_tracer(7)
    q.end_tracer_declarations()
    q.finish_setup()
    a = z()
    with pytest.raises(cb.CedrError, match="alias"):
        q.bind_arrays(a, z(), z(), z(), out=a)
    with pytest.raises(cb.CedrError, match="Qm_prev"):
        q.bind_arrays(z(), z(), z(), None, out=z())


# ---------------------------------------------------------------- cedr::local (8f-4)

@pytest.mark.parametrize("n", [2, 3, 4, 7, 16])
def test_local_batched_solvers_bitwise(oracle, n):
    """cedr_b200_local_solve: the element-local solvers of cedr_local_inl.hpp:68-330 for a
    batch of elements in the coalesced SoA layout, one thread per element with the element
    in registers -- x and the return codes bit for bit against the oracle's restatement of
    the reference, for feasible, tight and infeasible problems."""
    import torch
    import compose_b200 as cb
    rng = np.random.default_rng(100 + n)
    nprob = 257
    w = 0.1 + rng.random((n, nprob))
    a = 0.1 + rng.random((n, nprob))
    xlo = rng.random((n, nprob)) - 0.5
    xhi = xlo + rng.random((n, nprob))*rng.choice([1.0, 1e-3], size=(1, nprob))
    y = xlo + (xhi - xlo)*(1.6*rng.random((n, nprob)) - 0.3)
    t = rng.random(nprob)*1.2 - 0.1          # some b outside [a'xlo, a'xhi]: infeasible
    b = (a*xlo).sum(0)*(1 - t) + (a*xhi).sum(0)*t
    dev = lambda v: torch.from_numpy(np.ascontiguousarray(v)).cuda()
    d = {k: dev(v) for k, v in dict(w=w, a=a, xlo=xlo, xhi=xhi, y=y, b=b).items()}

    def check(method, ref_fn, **kw):
        x, info = cb.local_solve(method, d["b"], d["y"], **kw)
        x, info = x.cpu().numpy(), info.cpu().numpy()
        for p in range(nprob):
            ri, rx = ref_fn(p)
            assert info[p] == ri, (method, p, info[p], ri)
            assert np.array_equal(x[:, p], rx), (method, p)

    check(cb.LOCAL_QP, lambda p: oracle.solve_1eq_bc_qp(w[:, p], a[:, p], b[p], xlo[:, p],
                                                        xhi[:, p], y[:, p]),
          xlo=d["xlo"], xhi=d["xhi"], w=d["w"], a=d["a"])
    check(cb.LOCAL_CAAS, lambda p: (0, oracle.local_caas(a[:, p], b[p], xlo[:, p], xhi[:, p],
                                                         y[:, p])),
          xlo=d["xlo"], xhi=d["xhi"], a=d["a"])
    bn = np.abs(b) * np.where(rng.random(nprob) < 0.1, -1.0, 1.0)
    d["b"] = dev(bn)
    yn = rng.random((n, nprob)) - 0.2
    d["y"] = dev(yn)

    def nn(method):
        def f(p):
            info, x = oracle.solve_1eq_nonneg(a[:, p], bn[p], yn[:, p], w[:, p], method)
            return info, (x if info != -1 or bn[p] >= 0 else np.zeros(n))
        return f
    check(cb.LOCAL_NONNEG_LS, nn(0), w=d["w"], a=d["a"])
    check(cb.LOCAL_NONNEG_CAAS, nn(1), w=d["w"], a=d["a"])
    if n == 2:
        d["b"], d["y"] = dev(b), dev(y)
        check(cb.LOCAL_QP_2D, lambda p: oracle.solve_1eq_bc_qp_2d(w[:, p], a[:, p], b[p],
                                                                  xlo[:, p], xhi[:, p], y[:, p]),
              xlo=d["xlo"], xhi=d["xhi"], w=d["w"], a=d["a"])


@pytest.mark.parametrize("use_graph", [False, True, 2])
def test_transport1d_device_harness_bitwise(oracle, use_graph):
    """BASELINE.json config 5 entirely on the device (cedr_b200_transport1d_cycle): the 351
    steps of cedr_test_1d_transport.cpp on 111 cells -- interpolation + set_Qm, CDR::run,
    get_Qm per step, optionally replayed from a CUDA graph, or (2) the whole cycle as one
    launch of a persistent CTA -- must end on the same bits as the host loop driving the
    oracle, for qltnn, qlt and caas."""
    import compose_b200 as cb
    import transport1d as T
    from test_oracle_golden import t1d_runners
    ncells = 111
    p, oruns = t1d_runners(oracle, ncells, caas_tree=True)
    nsteps = int(3.17*ncells)
    y0 = p.y0()
    for name, kind, pt in (("yqltnn", "qlt", 1 | 8), ("yqlt", "qlt", 1 | 2), ("ycaas", "caas", 3)):
        ref = p.cycle(nsteps, y0, oruns[name])
        c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
        c.declare_tracer(pt)
        c.end_tracer_declarations()
        c.finish_setup()
        yf, us = c.transport1d_cycle(nsteps, y0, use_graph=use_graph)
        assert np.array_equal(yf, ref), name
        assert us > 0
