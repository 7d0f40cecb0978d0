#!/bin/bash
# Build a named variant of the library for A/B timing: tools/ab_build.sh NAME [nvcc flags...]
# -> build/var/NAME.so (git-ignored, ships with gpurun). Prints registers / spills of down2.
name=$1; shift
cd "$(dirname "$0")/.."
python - "$name" "$@" <<'PY'
import sys
sys.path.insert(0, ".")
from compose_b200 import build as b
b.build(force=True, verbose=True, out="build/var/%s.so" % sys.argv[1], extra_flags=sys.argv[2:])
PY
