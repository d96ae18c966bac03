#!/bin/bash
TAG=${1:-r02b}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_mf_cfg4_launches.csv python tools/profile_step.py 800 300 8 1 > gpurun_out/${TAG}_prof.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mf_small_kernel -c 1 -f -o gpurun_out/${TAG}_small python tools/profile_step.py 800 300 8 1 >> gpurun_out/${TAG}_prof.log 2>&1
for tw in 2 4; do HMCMT_MF_TINYWARPS=$tw python tools/dev/t_mf.py cfg4 2>&1 | grep "nfreq 8"; done
tail -3 gpurun_out/${TAG}_prof.log
