"""Generates the committed fixtures under tests/golden/ from the REFERENCE ITSELF
(oracle/_ref: COMPOSE's unmodified cedr/*.cpp built by oracle/Makefile). Run in the
container where /root/reference exists:   python tests/golden/make_golden.py

  qlt_caas_small.npz       inputs + reference outputs of QLT (6 problem classes x 6
                           perturbations, balanced/imbalanced trees, both option values) and
                           CAAS (default and BfbTreeAllReducer sums) on small meshes
  out_transport1d_nc111.py the file `cedr_test -t -t1d -nc 111` writes (config 5)
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import randomized as R                                    # noqa: E402
from oracle.oracle_py import Oracle, Ref, ref_available   # noqa: E402


def main():
    assert ref_available(), "build oracle/_ref first (python -c 'from oracle import oracle_py; oracle_py.build()')"
    o, r = Oracle(), Ref()
    out = {}
    cases = []
    for ncells in (1, 2, 7, 21, 111):
        for imb in (0, 1):
            for prefer in (0, 1):
                ts, v = R.generate(ncells, seed=1000*ncells + 2*imb + prefer)
                pts = np.array([t.problem_type for t in ts], np.int32)
                q, _, _ = r.qlt(ncells, ("bisect", bool(imb)), pts, v.rhom, v.Qm_min, v.Qm,
                                v.Qm_max, v.Qm_prev, bool(prefer))
                k = "qlt_%d_%d_%d" % (ncells, imb, prefer)
                cases.append(k)
                for name, a in (("pts", pts), ("rhom", v.rhom), ("lo", v.Qm_min), ("q", v.Qm),
                                ("hi", v.Qm_max), ("prev", v.Qm_prev), ("out", q)):
                    out[k + "/" + name] = a
    for ncells in (1, 2, 4, 11, 111):
        ts, v = R.generate(ncells, seed=77 + ncells)
        sel = [t for t in ts if (t.problem_type & R.S) and t.local_should_hold]
        idx = [t.idx for t in sel]
        pts = np.array([t.problem_type for t in sel], np.int32)
        a = [x[idx] for x in (v.Qm_min, v.Qm, v.Qm_max, v.Qm_prev)]
        seq, _ = r.caas(ncells, pts, v.rhom, *a)
        tree = o.bisection_tree(ncells)
        bfb, _ = r.caas(ncells, pts, v.rhom, *a, tree=("bisect", False))
        k = "caas_%d" % ncells
        cases.append(k)
        for name, x in (("pts", pts), ("rhom", v.rhom), ("lo", a[0]), ("q", a[1]), ("hi", a[2]),
                        ("prev", a[3]), ("out_seq", seq), ("out_tree", bfb)):
            out[k + "/" + name] = x
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "qlt_caas_small.npz"), **out)
    exe = os.path.join(ROOT, "oracle", "_ref", "cedr_test")
    with tempfile.TemporaryDirectory() as d:
        subprocess.run([exe, "-t", "-t1d", "-nc", "111"], cwd=d, check=True,
                       stdout=subprocess.DEVNULL)
        shutil.copy(os.path.join(d, "out_transport1d.py"),
                    os.path.join(HERE, "out_transport1d_nc111.py"))
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
