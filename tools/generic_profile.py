"""One run() of QLT on the imbalanced tree (generic sweep kernels), for ncu."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import compose_b200 as cb
ncells = int(sys.argv[1]) if len(sys.argv) > 1 else 86400
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rhom, lo, q, hi, prev = cb.fill_headline(ncells, nt, 3)
c = cb.QLT(ncells, imbalanced=True)
for _ in range(nt):
    c.declare_tracer(7)
c.end_tracer_declarations()
c.finish_setup()
c.set_rhom(rhom)
c.set_Qm(q, lo, hi, prev)
for _ in range(2):
    c.run()
c.synchronize()
torch.cuda.synchronize()
print(c.describe() if hasattr(c, "describe") else "ok")
