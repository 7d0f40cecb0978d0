"""run() time of QLT for a tracer set that cycles through the six problem classes
(cedr_qlt_inl.hpp:175-203) against the same number of `cst` tracers, and against the generic
kernels for the four one-field classes (CEDR_B200_FAST_ST_ONLY=1).

    python tools/mixed_classes.py [ncells] [ntracers]
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
    rhom, lo, q, hi, prev = cb.fill_headline(ncells, nt, 3)
    S, C, T, N = cb.SHAPEPRESERVE, cb.CONSERVE, cb.CONSISTENT, cb.NONNEGATIVE
    six = [S | T, C | S | T, T, C | T, N, C | N]
    sets = {"all cst": [C | S | T]*nt, "six classes": [six[i % 6] for i in range(nt)]}
    for name, pts in sets.items():
        for st_only in ("", "1"):
            if st_only and name == "all cst":
                continue
            if st_only:
                os.environ["CEDR_B200_FAST_ST_ONLY"] = "1"
            else:
                os.environ.pop("CEDR_B200_FAST_ST_ONLY", None)
            c = cb.QLT(ncells)
            for p in pts:
                c.declare_tracer(p)
            c.end_tracer_declarations()
            c.finish_setup()
            c.set_rhom(rhom)
            c.set_Qm(q, lo, hi, prev)
            ts = []
            for i in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                c.run()
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1))
            c.synchronize()
            print("%-12s %-32s %7.3f ms per run() (%d cells x %d tracers)"
                  % (name, "generic kernels for t/ct/nn/cnn" if st_only else "fast kernels",
                     sum(ts)/len(ts), ncells, nt), flush=True)
            del c
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
