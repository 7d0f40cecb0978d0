"""run() time of QLT on trees the fast kernels do not take (the imbalanced n -> (n/3, n - n/3)
tree of cedr_tree.cpp:391-413, or the balanced one with the fast path off) as a function of
the block size of the plan (cedr_b200_set_max_block_leaves).

    python tools/generic_blocks.py [ncells] [ntracers] [mbl,mbl,...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import compose_b200 as cb
    ncells = int(sys.argv[1]) if len(sys.argv) > 1 else 86400
    nt = int(sys.argv[2]) if len(sys.argv) > 2 else 1280
    mbls = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 256, 64, 32]
    rhom, lo, q, hi, prev = cb.fill_headline(ncells, nt, 3)
    for kind in ("qlt", "caas"):
      for imb in ((False, True) if kind == "qlt" else (False,)):
        for mbl in mbls:
            if kind == "qlt":
                c = cb.QLT(ncells, imbalanced=imb)
            else:
                c = cb.CAAS(ncells)
            if mbl:
                c.set_max_block_leaves(mbl)
            for _ in range(nt):
                c.declare_tracer(7)
            c.end_tracer_declarations()
            c.finish_setup()
            c.set_rhom(rhom)
            c.set_Qm(q, lo, hi, prev)
            ts = []
            for i in range(6):
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
            print("%s %-10s max_block_leaves %-5s fast=%-5s launches %2d: %7.3f ms per run()"
                  % (kind, "imbalanced" if imb else "balanced", mbl or "dflt", c.uses_fast_path(),
                     c.last_run_launches(), sum(ts)/len(ts)), flush=True)
            del c
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
