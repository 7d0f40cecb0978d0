"""Secondary runs of SURVEY 8(d): QLT::run at ne120 on (i) no-op inputs (Qm = Qm_prev in
bounds: every node takes the quick exit) and (ii) the mixed six-class tracer set of
cedr_test_randomized.cpp:28-35, against the headline all-cst active inputs.
Usage: python tools/secondary_runs.py [nt]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import compose_b200 as cb
from compose_b200.workloads import CONFIGS

ncells, nt, cid = CONFIGS["ne120x128x40"]
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
rhom, lo, q, hi, prev = cb.fill_headline(ncells, nt, cid)


def timed(c, qq):
    ts = []
    for i in range(5):
        c.set_Qm(qq, lo, hi, prev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c.run(); e1.record(); torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return min(ts)


def make(ptypes):
    c = cb.QLT(ncells)
    for p in ptypes:
        c.declare_tracer(p)
    c.end_tracer_declarations(); c.finish_setup(); c.set_rhom(rhom)
    return c


c = make([7]*nt)
print("all cst, active inputs: %.3f ms for %d tracers" % (timed(c, q), nt))
print("all cst, no-op inputs (Qm = Qm_prev): %.3f ms" % timed(c, prev))
del c
# st, cst, t, ct, nn, cnn in turn: shapepreserve=2, conserve=1, consistent=4, nonnegative=8
six = [2 | 4, 1 | 2 | 4, 4, 1 | 4, 8, 1 | 8]
c = make([six[i % 6] for i in range(nt)])
print("mixed six classes, active inputs: %.3f ms" % timed(c, q))
print("   launches", c.last_run_launches())
