"""Small driver for ncu captures: builds one reconstructor on a named workload and calls
run() a few times. Usage: python tools/prof_run.py qlt|caas [workload] [nruns] [nt]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import compose_b200 as cb
from compose_b200.workloads import CONFIGS

kind = sys.argv[1] if len(sys.argv) > 1 else "qlt"
wl = sys.argv[2] if len(sys.argv) > 2 else "ne30x72x40"
nruns = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ncells, nt, cid = CONFIGS[wl]
if len(sys.argv) > 4:
    nt = int(sys.argv[4])
rhom, lo, q, hi, prev = cb.fill_headline(ncells, nt, cid)
c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
for _ in range(nt):
    c.declare_tracer(7)
c.end_tracer_declarations()
c.finish_setup()
c.set_rhom(rhom)
for _ in range(nruns):
    c.set_Qm(q, lo, hi, prev)
    c.run()
torch.cuda.synchronize()
print("done", kind, wl, ncells, nt, c.last_run_launches(), "launches/run")
