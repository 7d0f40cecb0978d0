"""Per-source-line hot spots of one kernel from an .ncu-rep (needs -lineinfo and
--import-source on): samples, warp instructions and the main stall reasons per CUDA line.
Usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [topN]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + rx], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    num = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
    smp = num(r[ix["# Samples"]])
    ins = num(r[ix["Instructions Executed"]])
    stalls = {h[6:]: num(r[i]) for i, h in enumerate(hdr)
              if h.startswith("stall_") and "Not Issued" not in h}
    out.append((smp, ins, cur_file, int(r[0]), r[1].strip()[:70], stalls))
tot_s = sum(o[0] for o in out) or 1
tot_i = sum(o[1] for o in out) or 1
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
out.sort(reverse=True)
for smp, ins, f, ln, src, st in out[:top]:
    top3 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %-70s  %s" % (
        100.0*smp/tot_s, 100.0*ins/tot_i, f, ln, src,
        " ".join("%s=%d" % kv for kv in top3 if kv[1])))
