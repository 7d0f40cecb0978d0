"""Device time per run() of a tiny problem (config 5 size: 111 cells x 1 tracer), 200
back-to-back calls; with and without the single-launch path (CEDR_B200_NO_SOLO)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import compose_b200 as cb

for ncells, nt in ((111, 1), (21, 36), (256, 40)):
    r1, l1, q1, h1, p1 = cb.fill_headline(ncells, nt, 5)
    for kind in ("qlt", "caas"):
        c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
        for _ in range(nt):
            c.declare_tracer(3 if kind == "caas" else 7)
        c.end_tracer_declarations()
        c.finish_setup()
        c.set_rhom(r1)
        c.set_Qm(q1, l1, h1, p1)
        for _ in range(20):
            c.run()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            c.run()
        e1.record()
        torch.cuda.synchronize()
        print(ncells, nt, kind, "%.2f us/run, %d launches" % (1e3*e0.elapsed_time(e1)/200,
                                                             c.last_run_launches()))
