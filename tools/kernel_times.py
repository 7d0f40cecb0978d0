"""Per-kernel CUDA-event times of one run() (profiling on): python tools/kernel_times.py qlt|caas [workload] [nt]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import compose_b200 as cb
from compose_b200.workloads import CONFIGS

kind = sys.argv[1] if len(sys.argv) > 1 else "qlt"
wl = sys.argv[2] if len(sys.argv) > 2 else "ne120x128x40"
ncells, nt, cid = CONFIGS[wl]
if len(sys.argv) > 3:
    nt = int(sys.argv[3])
rhom, lo, q, hi, prev = cb.fill_headline(ncells, nt, cid)
c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
for _ in range(nt):
    c.declare_tracer(7)
c.end_tracer_declarations()
c.finish_setup()
c.set_rhom(rhom)
c.set_profiling(True)
for i in range(3):
    c.set_Qm(q, lo, hi, prev)
    c.run()
    torch.cuda.synchronize()
tot = 0.0
for name, tier, ms in c.launch_times():
    print("%-12s tier %d  %8.4f ms" % (name, tier, ms))
    tot += ms
print("sum %.4f ms" % tot)
