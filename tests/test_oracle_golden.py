"""Pin the plain-C oracle against fixtures generated from the reference itself
(tests/golden/make_golden.py ran the UNMODIFIED COMPOSE sources, oracle/_ref): these run
everywhere, also where neither /root/reference nor oracle/_ref exists.

- qlt_caas_small.npz: QLT and CAAS outputs, compared BITWISE;
- out_transport1d_nc111.py: config 5, the reference's 1-D transport test output after 351
  steps on 111 cells (printed with 16 significant digits, so compared to 1e-14).
"""
import os

import numpy as np
import pytest

import transport1d as T
from oracle.oracle_py import Oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "qlt_caas_small.npz"))


def test_qlt_matches_reference_fixtures(oracle, gold):
    n = 0
    for k in gold["cases"]:
        if not k.startswith("qlt_"):
            continue
        _, ncells, imb, prefer = k.split("_")
        g = lambda name: gold[k + "/" + name]
        tree = oracle.bisection_tree(int(ncells), bool(int(imb)))
        out = oracle.qlt(tree, g("pts"), g("rhom"), g("lo"), g("q"), g("hi"), g("prev"),
                         bool(int(prefer)))
        assert np.array_equal(out, g("out")), k
        n += 1
    assert n == 20


def test_caas_matches_reference_fixtures(oracle, gold):
    n = 0
    for k in gold["cases"]:
        if not k.startswith("caas_"):
            continue
        ncells = int(k.split("_")[1])
        g = lambda name: gold[k + "/" + name]
        seq = oracle.caas(ncells, g("pts"), g("lo"), g("q"), g("hi"), g("prev"))
        assert np.array_equal(seq, g("out_seq")), k
        tree = oracle.bisection_tree(ncells)
        tr = oracle.caas(ncells, g("pts"), g("lo"), g("q"), g("hi"), g("prev"), tree=tree)
        assert np.array_equal(tr, g("out_tree")), k
        n += 1
    assert n == 5


def load_t1d():
    s = {}
    src = open(os.path.join(GOLD, "out_transport1d_nc111.py")).read()
    exec(src.replace("s = {};", ""), {"s": s})
    return {k: np.array(v) for k, v in s.items()}


def t1d_runners(oracle, ncells, caas_tree):
    tree = oracle.bisection_tree(ncells)
    p = T.Problem1D(ncells)
    rhom = p.area

    def qlt(pt):
        return lambda q, lo, hi, prev: oracle.qlt(tree, [pt], rhom, lo[None], q[None], hi[None],
                                                  prev[None])[0]

    def caas(q, lo, hi, prev):
        return oracle.caas(ncells, [3], lo[None], q[None], hi[None], prev[None],
                           tree=tree if caas_tree else None)[0]
    return p, {"yqltnn": qlt(1 | 8), "yqlt": qlt(1 | 2), "ycaas": caas}


def test_transport1d_matches_reference_output(oracle):
    """The re-expressed harness + oracle reproduce the reference's own config-5 run."""
    ref = load_t1d()
    ncells = 111
    p, runners = t1d_runners(oracle, ncells, caas_tree=False)   # the reference's default sums
    assert np.allclose(p.xb, ref["xb"], rtol=0, atol=1e-15)
    y0 = p.y0()
    assert np.allclose(y0, ref["y0"], rtol=1e-15, atol=1e-15)
    nsteps = int(3.17*ncells)
    for name, run in runners.items():
        yf = p.cycle(nsteps, y0, run)
        err = np.abs(yf - ref[name]).max()
        assert err <= 1e-14, (name, err)
