#!/bin/bash
# launch lists (ncu, serialised) of one leapfrog step on the multifrontal path: cfg4 mesh with 8 frequencies (16 systems, the
# share of one of 8 GPUs) and the cfg2 workload forced onto the multifrontal solver
TAG=${1:-r02}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_mf_cfg4_launches.csv python tools/profile_step.py 800 300 8 1 > gpurun_out/${TAG}_prof.log 2>&1
HMCMT_SOLVER=mf timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_mf_cfg2_launches.csv python tools/profile_step.py 200 100 30 1 >> gpurun_out/${TAG}_prof.log 2>&1
tail -3 gpurun_out/${TAG}_prof.log
