#!/bin/bash
# ncu evidence for profiles/: launch list of 2 leapfrog steps at cfg2 + one --set full capture of the dominant launches
# (every capture under its own timeout: a replayed kernel that stalls must not eat the GPU budget)
TAG=${1:-r01s}
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py 200 100 30 2 > gpurun_out/${TAG}_prof.log 2>&1
timeout 180 ncu --set full --clock-control none --import-source on -k regex:band_factor -s 2 -c 1 -f -o gpurun_out/${TAG}_factor_own python tools/profile_step.py 200 100 30 2 >> gpurun_out/${TAG}_prof.log 2>&1
timeout 180 ncu --set full --clock-control none --import-source on -k regex:band_solve -s 5 -c 1 -f -o gpurun_out/${TAG}_solve_own python tools/profile_step.py 200 100 30 2 >> gpurun_out/${TAG}_prof.log 2>&1
tail -5 gpurun_out/${TAG}_prof.log
