"""Host-side logic of the multi-rank (subtree-partition) path, on CPU.

- the partition every rank derives from the tree tiles the cells and the tier-0 blocks
  exactly once, for 1/2/4/8 ranks on the BASELINE.json cubed-sphere sizes;
- a world_size-2 `gloo` run plays the exchange of run(): each rank packs one entry per
  owned block [global block index, payload], the entries are all-gathered, and every rank
  rebuilds the same replicated tier-1 leaf array -- the analogue of the reference's
  test_comm_pattern (cedr_tree.cpp:279-348): the sum of cell ids through the pattern
  must be n(n-1)/2 on every rank.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.mark.parametrize("ncells", [5400, 86400, 393216])
@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
def test_partition_tiles_cells_and_blocks(ncells, nranks):
    import compose_b200 as cb
    cells, blocks = [], []
    for r in range(nranks):
        p = cb.partition_probe(ncells, r, nranks)
        assert p["nlclcells"] == ncells//nranks
        assert p["nown"] <= p["nown_max"] or nranks == 1
        assert int(p["nl"].sum()) == p["nlclcells"]
        for g, l0, nl in zip(p["gidx"], p["leaf0"], p["nl"]):
            assert 513 <= nl <= 1024          # every block runs the fast kernels
            cells.extend(range(l0, l0 + nl))
            blocks.append(int(g))
    assert cells == list(range(ncells))
    assert sorted(blocks) == list(range(p["nblocks"]))


def test_partition_rejects_split_blocks():
    import compose_b200 as cb
    with pytest.raises(cb.CedrError, match="whole blocks"):
        cb.partition_probe(5400, 1, 3)


def _worker(rank, world, port, ncells, mbl, q):
    import torch
    import torch.distributed as dist
    import compose_b200 as cb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = cb.partition_probe(ncells, rank, world, max_block_leaves=mbl)
        # The kernels' message layout (kernels.cuh pack_kernel): word-major,
        # message[w*nown_max + j]; word 0 = global block index (-1 pads), word 1 = payload.
        nwords, nmax = 2, p["nown_max"]
        send = torch.full((nwords*nmax,), -1.0, dtype=torch.float64)
        for j, (g, l0, nl) in enumerate(zip(p["gidx"], p["leaf0"], p["nl"])):
            send[j] = float(g)
            send[nmax + j] = float(sum(range(l0, l0 + nl)))   # "up-sweep" of cell ids
        recv = torch.empty(world*send.numel(), dtype=torch.float64)
        dist.all_gather_into_tensor(recv, send)
        tier1 = np.full(p["nblocks"], np.nan)
        for msg in recv.numpy().reshape(world, nwords, nmax):
            for g, val in zip(msg[0], msg[1]):
                if g >= 0:
                    assert np.isnan(tier1[int(g)])      # each block arrives exactly once
                    tier1[int(g)] = val
        q.put((rank, float(tier1.sum()), bool(np.isnan(tier1).any())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ncells,mbl", [(5400, 1024), (64, 4)])
def test_exchange_pattern_gloo_world2(ncells, mbl):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ncells, mbl, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, total, missing in out:
        assert not missing
        assert total == ncells*(ncells - 1)/2


def _partial_worker(rank, world, port, ncells, q):
    import torch.distributed as dist
    import compose_b200 as cb
    from test_host_logic import prune_for_rank, preorder
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tree = cb.make_1d_tree(ncells, True)
        ci = np.arange(ncells)
        rank_of_cell = ((ci + ci//world) % world).astype(np.int32)   # cedr_tree.cpp:371-374
        rank_of_cell[:ncells//2] = 0     # ... with a long run of rank 0, so parts get cut
        rank_of_cell[ncells//2] = 1
        kids, cellidx, root, node_rank = prune_for_rank(tree, rank_of_cell, rank,
                                                        np.random.default_rng(rank))
        whole, whole_rank = cb.assemble_partial_tree((kids, cellidx, root), node_rank)
        full_rank = np.where(tree[1] >= 0, rank_of_cell[np.maximum(tree[1], 0)], 0)
        ek, ec, er = preorder(tree, full_rank)
        same = (np.array_equal(whole[0], ek) and np.array_equal(whole[1], ec) and
                np.array_equal(whole_rank, er))
        q.put((rank, bool(same), int(cellidx.size), int(ec.size)))
    finally:
        dist.destroy_process_group()


def test_partial_trees_assembled_over_gloo_world2():
    """tree::Node::level partial trees (cedr_tree_caller.hpp:20-22): each rank hands over
    its pruned part; after the setup-time gather both hold the caller's whole tree."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_partial_worker, args=(r, 2, port, 300, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, npart, nfull in out:
        assert same
        if rank == 1:
            assert npart < nfull      # rank 1's part really was cut
