"""Per-kernel SASS summary of the built library: instruction count and the mnemonics that
identify the Blackwell-native paths (UBLKCP = cp.async.bulk / TMA bulk copy, SYNCS = mbarrier
ops, LDGSTS = cp.async, UCGABAR / barrier.cluster, FP64 pipe ops, local-memory spills).
Usage: python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "compose_b200", "libcedr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
KEYS = ["UBLKCP", "SYNCS", "LDGSTS", "UCGABAR", "MAPA", "DFMA", "DADD", "DMUL", "DSETP", "MUFU",
        "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "SHFL"]
print("# cuobjdump -sass %s (sm_100a cubin): per kernel, instruction count and selected\n"
      "# mnemonics. UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops, LDGSTS =\n"
      "# cp.async, UCGABAR = barrier.cluster, MAPA = distributed-shared-memory address mapping,\n"
      "# DFMA/DADD/DMUL/DSETP = FP64 pipe, LDL/STL = local-memory spills. No tensor-core ops\n"
      "# (HMMA/UTC*MMA) are expected: the path has no dense contraction.\n"
      % os.path.relpath(lib, ROOT))
name, cnt = None, collections.Counter()


def flush():
    if name is None:
        return
    n = sum(cnt.values())
    short = re.sub(r"^_ZN\d+cedr_b200", "", name)
    print("%-88s n=%6d  %s" % (short[:88], n, " ".join("%s=%d" % (k, cnt[k]) for k in KEYS if cnt[k])))


for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, cnt = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        op = m.group(1)
        for k in ("UCGABAR", "UBLKCP", "SYNCS"):
            if op.startswith(k):
                op = k
        cnt[op] += 1
flush()
