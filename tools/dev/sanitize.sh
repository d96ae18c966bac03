#!/bin/bash
# compute-sanitizer passes over a small split problem (forward + gradient) and a large-bandwidth one; each under its own timeout
export HMCMT_SPLIT=1
for tool in synccheck racecheck memcheck; do
  echo "== $tool split 40x30x1"
  timeout 300 compute-sanitizer --tool $tool python tools/dev/t_split.py 40 30 1 grad 2>&1 | grep -E "ERROR SUMMARY|grad ok|error|Error|hazard" | sort | uniq -c | head -8
done
echo "== synccheck big 60x124 (b > 104)"
timeout 300 compute-sanitizer --tool synccheck python tools/dev/t_split.py 124 118 1 grad 2>&1 | grep -E "ERROR SUMMARY|grad ok|error|Error" | sort | uniq -c | head -8
echo "== memcheck big"
timeout 300 compute-sanitizer --tool memcheck python tools/dev/t_split.py 124 118 1 grad 2>&1 | grep -E "ERROR SUMMARY|grad ok|error|Error" | sort | uniq -c | head -8
