#!/bin/bash
# Dev: run the split path on small and cfg2-sized problems, each case in its own process with a short timeout
run() { echo "== $*"; timeout 40 python tools/dev/t_split.py "$@" 2>&1 | tail -3; }
export HMCMT_SPLIT=1
run 40 30 1 fwd
run 40 30 2 grad
python tools/dev/cmp_split.py 40 30 2
python tools/dev/cmp_split.py 200 100 3
python tools/dev/cmp_split.py 57 131 2
