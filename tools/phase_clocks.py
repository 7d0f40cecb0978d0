"""Debug: per-phase clock shares of the specialised down-sweep kernel. Needs the library
built with CEDR_B200_EXTRA_NVCC_FLAGS=-DCEDR_B200_PHASE_CLOCKS.
Usage: python tools/phase_clocks.py [workload] [nt]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import compose_b200 as cb
from compose_b200.workloads import CONFIGS

wl = sys.argv[1] if len(sys.argv) > 1 else "ne30x72x40"
ncells, nt, cid = CONFIGS[wl]
if len(sys.argv) > 2:
    nt = int(sys.argv[2])
rhom, lo, q, hi, prev = cb.fill_headline(ncells, nt, cid)
c = cb.QLT(ncells)
for _ in range(nt):
    c.declare_tracer(7)
c.end_tracer_declarations()
c.finish_setup()
c.set_rhom(rhom)
if "--noop" in sys.argv:      # Qm = Qm_prev in bounds: every node takes the quick exit
    q = prev
c.set_Qm(q, lo, hi, prev)
c.run(); c.synchronize(); c.debug_phase_clocks()
c.run(); c.synchronize()
v = c.debug_phase_clocks()
leaf = ["loop/issue", "wait TMA", "re-sum", "wait T", "d7+d8 solves", "pairs", "barrier+writeback", ""]
top = ["loop/ldcg", "wait n7 / C", "up levels", "solves d0..6", "", "", "", ""]
for name, vals, labels in (("leaf warp 0", v[:8], leaf), ("top warp", v[8:], top)):
    tot = sum(vals) or 1
    print(name, "total clocks", tot)
    for l, x in zip(labels, vals):
        if l:
            print("   %-22s %6.1f%%" % (l, 100.0*x/tot))
