#!/bin/bash
# compute-sanitizer passes (synccheck, racecheck, memcheck), each under its own timeout:
#   band kernels   : a small split problem, forward + gradient
#   multifrontal   : a mesh with every front-size class of the single-CTA kernel (1 / 4 / 8 / 16 warps), the global-memory path of
#                    the large fronts, the staged / direct solve kernels, groups of systems on streams (eager launches: the
#                    sanitizer serialises a graph replay anyway)
# writes a summary to gpurun_out/${TAG}_sanitizer.txt          bash tools/dev/sanitize.sh [TAG]
TAG=${1:-r02}
OUT=gpurun_out/${TAG}_sanitizer.txt
: > $OUT
run() {   # label, env, args
  for tool in synccheck racecheck memcheck; do
    echo "== $tool: $1" | tee -a $OUT
    env $2 timeout 600 compute-sanitizer --tool $tool python tools/dev/t_split.py $3 grad 2>&1 | grep -E "ERROR SUMMARY|grad ok|rror|hazard" | sort | uniq -c | head -8 | tee -a $OUT
  done
}
run "band kernels, two CTAs per system, 40x30 cells" "HMCMT_SOLVER=band HMCMT_SPLIT=1" "40 30 1"
run "multifrontal, 150x120 cells, 3 groups of systems" "HMCMT_SOLVER=mf HMCMT_GRAPH=0 HMCMT_GROUPS=3" "150 120 2"
