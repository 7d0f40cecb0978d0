"""Per-launch times of one CAAS::run() and one QLT::run() (profiling mode) at ne120 x 1280."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import compose_b200 as cb
from compose_b200.workloads import CONFIGS
ncells, nt, cid = CONFIGS["ne120x128x40"]
nt = 1280
rhom, lo, q, hi, prev = cb.fill_headline(ncells, nt, cid)
for kind in ("caas", "qlt"):
    c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
    for _ in range(nt):
        c.declare_tracer(7)
    c.end_tracer_declarations(); c.finish_setup(); c.set_rhom(rhom)
    c.set_profiling(True)
    best = {}
    for _ in range(4):
        c.set_Qm(q, lo, hi, prev); c.run(); torch.cuda.synchronize()
        for n, tr, ms in c.launch_times():
            best[(n, tr)] = min(best.get((n, tr), 1e9), ms)
    print(kind, {k: round(v, 4) for k, v in best.items()})
    del c
