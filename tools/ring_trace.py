"""Stage-by-stage latency breakdown of the ring kernel from its debug trace.

    CEDR_B200_RING_TRACE=1 python tools/ring_trace.py [qlt|caas] [ncells] [nt]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CEDR_B200_RING_TRACE", "1")

import torch
import compose_b200 as cb

kind = sys.argv[1] if len(sys.argv) > 1 else "caas"
ncells = int(sys.argv[2]) if len(sys.argv) > 2 else 86400
nt = int(sys.argv[3]) if len(sys.argv) > 3 else 640
rhom, lo, q, hi, prev = cb.fill_headline(ncells, nt, 2)
c = cb.QLT(ncells) if kind == "qlt" else cb.CAAS(ncells)
c.set_ring(True)
for _ in range(nt):
    c.declare_tracer(7)
c.end_tracer_declarations()
c.finish_setup()
c.set_rhom(rhom)
for i in range(3):
    c.set_Qm(q, lo, hi, prev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    c.run()
    e1.record()
    torch.cuda.synchronize()
print("run: %.3f ms" % e0.elapsed_time(e1), c.ring_info())
info = c.ring_info()
tr = c.ring_trace().astype(np.int64)
G, TB = info["grid"], info["TB"]
if not TB: raise SystemExit("ring kernel not used")
U = (nt + TB - 1)//TB
units = tr[:G*U*8].reshape(G, U, 8)
strace = tr[G*U*8:G*U*8 + 2*nt].reshape(nt, 2)
clk = tr[G*U*8 + 2*nt:].reshape(32, 2)
tr = tr[:G*U*8 + 2*nt]
t0 = units[units > 0].min()
names = ["load", "UPstart", "UPend", "T-UPend", "DNstart", "T-DNend", "L-DNend", "store"]
mid = slice(U//4, 3*U//4)
print("kernel span: %.1f us" % ((tr.max() - t0)/1e3))
u = units[:, mid, :].astype(np.float64)
def stat(x):
    x = x[np.isfinite(x)]
    return "mean %7.2f  p50 %7.2f  p90 %7.2f  max %7.2f us" % (x.mean()/1e3, np.median(x)/1e3,
                                                            np.percentile(x, 90)/1e3, x.max()/1e3)
pairs = [(0, 1, "load -> landed/UP start"), (1, 2, "UP"), (2, 3, "UPend -> T-UP end (arrival)"),
         (3, 4, "arrival -> DOWN start (hand-off)"), (4, 5, "T-DOWN"), (5, 6, "T-DNend -> L-DOWN end"),
         (6, 7, "L-DOWN end -> store issued"), (0, 7, "load -> store (whole unit)")]
if kind == "caas":
    pairs = [(0, 1, "load -> landed/UP start"), (1, 2, "UP"), (2, 3, "UPend -> T-UP end (arrival)"),
             (3, 4, "arrival -> DOWN start (hand-off)"), (4, 6, "DOWN"),
             (6, 7, "DOWN end -> store issued"), (0, 7, "load -> store (whole unit)")]
for a_, b_, nm in pairs:
    d = u[:, :, b_] - u[:, :, a_]
    d = d[(u[:, :, a_] > 0) & (u[:, :, b_] > 0)]
    print("%-36s %s" % (nm, stat(d)))
# per-CTA unit period
per = np.diff(units[:, mid, 7].astype(np.float64), axis=1)
print("%-36s %s" % ("store-to-store period per CTA", stat(per)))
# hand-off pieces: last arrival of tracer k -> S sees it -> serve end -> first DOWN start
if TB == 1:
    last_arr = units[:, :, 3].max(axis=0).astype(np.float64)
    first_dn = np.where(units[:, :, 4] > 0, units[:, :, 4], np.iinfo(np.int64).max).min(axis=0).astype(np.float64)
    k = np.arange(U)[mid]
    print("%-36s %s" % ("last arrival -> S sees count", stat(strace[k, 0] - last_arr[k])))
    print("%-36s %s" % ("S serve", stat((strace[k, 1] - strace[k, 0]).astype(np.float64))))
    print("%-36s %s" % ("S done -> first DOWN start", stat(first_dn[k] - strace[k, 1])))
    arr = units[:, mid, 3].astype(np.float64)
    print("%-36s %s" % ("arrival skew over CTAs (max-min)", stat(arr.max(axis=0) - arr.min(axis=0))))

names = {0: "L decision poll", 1: "L decision barrier", 2: "L UP smem loads+sums", 3: "L UP n7/shuffles/records",
         4: "L UP arrive", 5: "L DOWN compute (QLT: d7..d9)", 6: "L DOWN fence+arrive", 7: "L loop top",
         16: "L DOWN barrier", 17: "L DOWN pairs", 8: "T arrival (fence+red)", 9: "T-DOWN loads",
         10: "T-DOWN sums", 11: "T-DOWN solves", 18: "T idle/poll", 12: "S flag release", 13: "S loads + micro sums",
         14: "S tree over micro-roots", 15: "S micro solves", 20: "P store", 21: "P reload", 22: "P load", 23: "P idle"}
print("phase clocks, CTA 0 (cycles per occurrence, occurrences):")
for i in sorted(names):
    if clk[i, 1]:
        print("  %-32s %9.0f  x %d" % (names[i], clk[i, 0]/clk[i, 1], clk[i, 1]))
