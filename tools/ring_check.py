"""Quick GPU check of the persistent ring kernel against the oracle and the multi-launch
path (bitwise), then a timing of run() at a given size.

    python tools/ring_check.py [--time ne120x128x40] [--nt 640]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def check(kind, ncells, nt, pts=None, prefer=False, seed=0):
    import torch
    import compose_b200 as cb
    from compose_b200 import workloads as W
    from oracle.oracle_py import Oracle
    o = Oracle()
    rhom, lo, q, hi, prev = W.headline(ncells, nt, seed)
    pts = pts or [7]*nt
    tree = o.bisection_tree(ncells)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d = [dev(x) for x in (rhom, lo, q, hi, prev)]
    res = {}
    for ring in (True, False):
        c = (cb.QLT(ncells, prefer_numerical_mass_conservation_to_numerical_bounds=prefer)
             if kind == "qlt" else cb.CAAS(ncells))
        c.set_ring(ring)
        for p in pts:
            c.declare_tracer(p)
        c.end_tracer_declarations()
        c.finish_setup()
        c.set_rhom(d[0])
        c.set_Qm(d[2], d[1], d[3], d[4])
        t0 = time.time()
        c.run()
        try:
            c.synchronize()
        except Exception as e:
            print("  %s ring=%s: %s" % (kind, ring, e))
            return False
        res[ring] = c.get_Qm().cpu().numpy()
        info = c.ring_info() if ring else None
        if ring:
            print("  %s ncells=%d nt=%d ring=%s uses_ring=%s info=%s (%.3fs)"
                  % (kind, ncells, nt, ring, c.uses_ring(), info, time.time() - t0))
    ref = (o.qlt(tree, pts, rhom, lo, q, hi, prev, prefer_mass_con=prefer) if kind == "qlt"
           else o.caas(ncells, pts, lo, q, hi, prev, tree=tree))
    ok = True
    for ring in (True, False):
        bad = int((res[ring] != ref).sum())
        if bad:
            ok = False
            w = np.argwhere(res[ring] != ref)[:5]
            print("  MISMATCH ring=%s: %d of %d differ; first at %s" % (ring, bad, ref.size, w.tolist()))
    print("  -> %s" % ("bitwise OK" if ok else "FAILED"))
    return ok


def timeit(workload, nt_override, reps=5):
    import torch
    import compose_b200 as cb
    from compose_b200.workloads import CONFIGS
    ncells, nt, cid = CONFIGS[workload]
    nt = nt_override or nt
    rhom, lo, q, hi, prev = cb.fill_headline(ncells, nt, cid)
    for kind in ("qlt", "caas"):
        for ring in (True, False):
            c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
            c.set_ring(ring)
            for _ in range(nt):
                c.declare_tracer(7)
            c.end_tracer_declarations()
            c.finish_setup()
            c.set_rhom(rhom)
            c.set_Qm(q, lo, hi, prev)
            ts = []
            for i in range(reps + 2):
                if kind == "caas":
                    c.set_Qm(q, lo, hi, prev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                c.run()
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1))
            c.synchronize()
            ms = sum(ts)/len(ts)
            print("TIME %s %s nt=%d ring=%s: %.3f ms  (%.1f%% of 6525 GB/s at 40 B/update) %s"
                  % (workload, kind, nt, ring, ms, 100*40.0*ncells*nt/(ms*1e-3)/6525.2e9,
                     c.ring_info() if ring else ""), flush=True)
            del c
            torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", default=None)
    ap.add_argument("--nt", type=int, default=0)
    ap.add_argument("--skip-check", action="store_true")
    ap.add_argument("--case", action="append", default=[],
                    help="kind,ncells,nt (repeatable): check only these")
    a = ap.parse_args()
    ok = True
    if a.case:
        for cs in a.case:
            kind, nc, nt = cs.split(",")
            ok &= check(kind, int(nc), int(nt))
        print("ALL OK" if ok else "SOME FAILED")
    elif not a.skip_check:
        for kind in ("caas", "qlt"):
            for ncells, nt in ((5400, 3), (5400, 40), (86400, 7), (86400, 40), (2*1023, 5)):
                ok &= check(kind, ncells, nt)
        ok &= check("qlt", 86400, 9, prefer=True)
        ok &= check("qlt", 5400, 12, pts=[7, 6]*6)
        ok &= check("caas", 5400, 12, pts=[3, 2]*6)
        ok &= check("caas", 5400, 6, pts=[2]*6)
        print("ALL OK" if ok else "SOME FAILED")
    if a.time and ok:
        timeit(a.time, a.nt)
    sys.exit(0 if ok else 1)
