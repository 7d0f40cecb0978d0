"""Time run() of one reconstructor with CUDA events (min and mean of 4 after 2 warm-ups).
Usage: python tools/time_run.py qlt|caas [workload] [nt]"""
import json
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
ts = []
for i in range(6):
    c.set_Qm(q, lo, hi, prev)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    c.run()
    e1.record()
    torch.cuda.synchronize()
    if i >= 2:
        ts.append(e0.elapsed_time(e1))
print(json.dumps({"kind": kind, "workload": wl, "nt": nt, "min_ms": min(ts),
                  "mean_ms": sum(ts)/len(ts), "launches": c.last_run_launches()}))
