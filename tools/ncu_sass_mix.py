"""Instruction mix of one kernel from an .ncu-rep source page: warp-instructions executed
per SASS mnemonic, and the hottest stall samples. Usage:
python tools/ncu_sass_mix.py report.ncu-rep kernel_regex [topN]"""
import collections
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name",
                      "regex:" + rx], stdout=subprocess.PIPE, text=True).stdout
lines = raw.splitlines()
# first kernel only
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')),
           len(lines))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
mix = collections.Counter()
thr = collections.Counter()
samples = collections.Counter()
tot = 0
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    sass = r[ix["Source"]].strip()
    toks = sass.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0]
    n = int(r[ix["Instructions Executed"]] or 0)
    mix[op] += n
    thr[op] += int(r[ix["Thread Instructions Executed"]] or 0)
    samples[op] += int(r[ix["# Samples"]] or 0)
    tot += n
print("total warp instructions: %d" % tot)
print("%-10s %14s %7s %9s %9s" % ("op", "warp-instr", "share", "lanes", "samples"))
for op, n in mix.most_common(top):
    print("%-10s %14d %6.1f%% %9.1f %9d" % (op, n, 100.0*n/max(tot, 1), thr[op]/max(n, 1),
                                            samples[op]))
