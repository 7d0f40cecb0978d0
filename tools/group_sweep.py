"""run() time of one rank's share of ne120 x 8 GPUs (10800 cells = 16 blocks, 5120 tracers) on
one GPU as a function of the fast kernels' tracers-per-CTA (CEDR_B200_GROUP)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import compose_b200 as cb

ncells = int(sys.argv[1]) if len(sys.argv) > 1 else 10800
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 5120
groups = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 8, 12, 16, 20, 24, 28, 32]
rhom, lo, q, hi, prev = cb.fill_headline(ncells, nt, 3)
for g in groups:
    if g:
        os.environ["CEDR_B200_GROUP"] = str(g)
    else:
        os.environ.pop("CEDR_B200_GROUP", None)
    out = []
    for kind in ("qlt", "caas"):
        c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
        for _ in range(nt):
            c.declare_tracer(7)
        c.end_tracer_declarations()
        c.finish_setup()
        c.set_rhom(rhom)
        c.set_Qm(q, lo, hi, prev)
        ts = []
        for i in range(8):
            if kind == "caas":
                c.set_Qm(q, lo, hi, prev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            c.run()
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        out.append(sum(ts)/len(ts))
        del c
    print("group %-5s qlt %.4f ms  caas %.4f ms" % (g or "auto", out[0], out[1]), flush=True)
