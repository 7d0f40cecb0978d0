"""Config 5 of BASELINE.json: the reference's 1-D transport test
(cedr/cedr_test_1d_transport.cpp) re-expressed as a harness: periodic cubic-interpolation
semi-Lagrangian advection on 111 cells that calls CDR::run once per time step through
set_Qm / run / get_Qm, for a `conserve|nonnegative` QLT, a `conserve|shapepreserve` QLT
and a `conserve|shapepreserve` CAAS. The CDR is injected as a callable
    run(Qm, Qm_min, Qm_max, Qm_prev) -> Qm_new          (arrays over cells)
so the same loop drives the CPU oracle and the CUDA path.
"""
import bisect
import math

import numpy as np


def to_periodic_core(xl, xr, x):
    # cedr_test_1d_transport.cpp:14-18
    if xl <= x <= xr:
        return x
    w = xr - xl
    return x - w*math.floor((x - xl)/w)


def cubic_interp_periodic(x, y, xi):
    """cedr_test_1d_transport.cpp:38-85. Returns (yi, dod[nxi][4])."""
    nx = len(x)
    nc = nx - 1
    yi = np.empty(len(xi))
    dod = np.empty((len(xi), 4), np.int64)
    xl = x.tolist()

    def slope(i):
        return (y[i+1] - y[i])/(x[i+1] - x[i])
    for j in range(len(xi)):
        xp = to_periodic_core(x[0], x[nc], xi[j])
        ip1 = bisect.bisect_right(xl, xp)
        if ip1 == 0:
            ip1 += 1
        elif ip1 == nx:
            ip1 -= 1
        i = ip1 - 1
        for k in range(4):
            dod[j, k] = (i - 1 + k + nc) % nc
        smid = slope(i)
        if i == 0:
            a = (x[nc] - x[nc-1])/((x[1] - x[0]) + (x[nc] - x[nc-1]))
            s1 = (1 - a)*slope(nc - 1) + a*smid
        else:
            a = (x[i] - x[i-1])/(x[ip1] - x[i-1])
            s1 = (1 - a)*slope(i - 1) + a*smid
        if i == nc - 1:
            a = (x[ip1] - x[i])/((x[ip1] - x[i]) + (x[1] - x[0]))
            s2 = (1 - a)*smid + a*slope(0)
        else:
            a = (x[ip1] - x[i])/(x[i+2] - x[i])
            s2 = (1 - a)*smid + a*slope(ip1)
        # get_cubic, :24-36
        dx = x[ip1] - x[i]
        dx2 = dx*dx
        dx3 = dx2*dx
        den = -dx3
        c2, c3 = s1, y[i]
        b1 = y[ip1] - dx*c2 - c3
        b2 = s2 - c2
        c0 = (2.0*b1 - dx*b2)/den
        c1 = (-3.0*dx*b1 + dx2*b2)/den
        xij = xp - x[i]
        yi[j] = (((c0*xij + c1)*xij) + c2)*xij + c3
    return yi, dod


class Problem1D:
    def __init__(self, ncells):
        # init_mesh, uniform (:143-169)
        self.xb = np.array([i/ncells for i in range(ncells)] + [1.0])
        self.xcp = np.empty(ncells + 1)
        self.xcp[:ncells] = 0.5*(self.xb[:-1] + self.xb[1:])
        self.xcp[ncells] = 1 + self.xcp[0]
        self.area = self.xb[1:] - self.xb[:-1]
        self.n = ncells

    def y0(self):
        # transport1d::run, :291-296
        y = np.empty(self.n + 1)
        for i in range(self.n):
            x = self.xcp[i]
            if x < 0.4 or x > 0.9:
                y[i] = 0.1 + 0.8*0.5*(1 + math.sin(6*math.pi*x))
            else:
                y[i] = 0.0 if (x > 0.66 or x < 0.33) else 1.0
        y[self.n] = y[0]
        return y

    def cycle(self, nsteps, y0, run_cdr, on_step=None):
        """Problem1D::cycle (:231-254) with run_cdr (:170-189)."""
        n1 = self.n + 1
        xcpi = self.xcp + (-1.0/nsteps)
        ya = y0.copy()
        for ti in range(nsteps):
            yb, dod = cubic_interp_periodic(self.xcp, ya, xcpi)
            n = self.n
            nb = ya[dod[:n]]                    # the four values of each domain of dependence
            mn, mx = nb.min(axis=1), nb.max(axis=1)
            a = self.area
            q = run_cdr(yb[:n]*a, mn*a, mx*a, ya[:n]*a)
            yb[:n] = q/a
            yb[n] = yb[0]
            if on_step:
                on_step(ti, yb)
            ya = yb
        assert len(ya) == n1
        return ya
