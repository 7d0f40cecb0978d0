"""Rank the SASS instructions of one kernel in an .ncu-rep by shared-memory wavefronts
(source page): wavefronts, ideal wavefronts, executions, wavefronts per execution, and
the CUDA source line. Usage: python tools/ncu_smem.py report.ncu-rep kernel_regex [top]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name",
                      "regex:" + rx, "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if "L1 Wavefronts Shared" in r)
ix = {h: i for i, h in enumerate(hdr)}
src_cols = [i for i, h in enumerate(hdr) if h == "Source"]
items, tw, ti = [], 0.0, 0.0
line = ""
for r in rows[rows.index(hdr) + 1:]:
    if len(r) < len(hdr):
        continue
    if r[ix["Line No"]]:
        line = "%s: %s" % (r[ix["Line No"]], r[src_cols[0]].strip()[:70])
    try:
        w = float(r[ix["L1 Wavefronts Shared"]] or 0)
        wi = float(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
        ex = float(r[ix["Instructions Executed"]] or 0)
    except ValueError:
        continue
    if w > 0 and r[ix["Address"]]:
        items.append((w, wi, ex, r[src_cols[-1]].strip()[:40], line))
        tw += w
        ti += wi
print("shared wavefronts %.3g, ideal %.3g" % (tw, ti))
items.sort(reverse=True)
for w, wi, ex, sass, line in items[:top]:
    print("%10.0f %10.0f %9.0f %5.2f  %-40s %s" % (w, wi, ex, w/max(ex, 1), sass, line))
