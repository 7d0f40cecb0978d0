"""Helpers for the -m gpu parity tests: run the CUDA CDRs (through the C ABI, via
compose_b200) on numpy inputs given in global-cell order and return numpy outputs
in global-cell order, so they can be compared with the oracle directly."""
import numpy as np


def run_qlt_gpu(ncells, ptypes, rhom, qm_min, qm, qm_max, qm_prev, tree=None,
                imbalanced=False, prefer=False, max_block_leaves=None, nrun=1,
                external_buffers=False, ring=None):
    import torch
    import compose_b200 as cb
    q = cb.QLT(ncells, tree=tree, imbalanced=imbalanced,
               prefer_numerical_mass_conservation_to_numerical_bounds=prefer)
    if max_block_leaves:
        q.set_max_block_leaves(max_block_leaves)
    if ring is not None:
        q.set_ring(ring)
    for p in ptypes:
        q.declare_tracer(int(p))
    q.end_tracer_declarations()
    if external_buffers:
        b1, b2 = q.get_buffers_sizes()
        q.set_buffers(torch.zeros(b1, dtype=torch.float64, device="cuda"),
                      torch.zeros(max(b2, 1), dtype=torch.float64, device="cuda"))
    q.finish_setup()
    gcis = q.get_owned_glblcells()
    assert q.nlclcells() == ncells
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a)[..., gcis])).cuda()
    q.set_rhom(dev(rhom))
    d = [dev(a) for a in (qm, qm_min, qm_max, qm_prev)]
    out = None
    for _ in range(nrun):
        q.set_Qm(*d)
        q.run()
        out = q.get_Qm()
    q.synchronize()     # also raises if the persistent kernel's watchdog fired
    torch.cuda.synchronize()
    res = np.empty((len(ptypes), ncells))
    res[:, gcis] = out.cpu().numpy()
    return res, q


def run_caas_gpu(ncells, ptypes, rhom, qm_min, qm, qm_max, qm_prev, max_block_leaves=None,
                 external_buffers=False, ring=None, exact_buffers=False):
    import torch
    import compose_b200 as cb
    c = cb.CAAS(ncells)
    if max_block_leaves:
        c.set_max_block_leaves(max_block_leaves)
    if ring is not None:
        c.set_ring(ring)
    for p in ptypes:
        c.declare_tracer(int(p))
    c.end_tracer_declarations()
    if external_buffers:
        b1, b2 = c.get_buffers_sizes()
        c.set_buffers(torch.zeros(b1, dtype=torch.float64, device="cuda"),
                      torch.zeros(max(b2, 1), dtype=torch.float64, device="cuda"))
    c.finish_setup()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    c.set_rhom(dev(rhom))
    c.set_Qm(dev(qm), dev(qm_min), dev(qm_max), dev(qm_prev))
    c.run()
    out = c.get_Qm()
    c.synchronize()
    torch.cuda.synchronize()
    return out.cpu().numpy(), c
