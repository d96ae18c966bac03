#!/bin/bash
TAG=${1:-r02c}
HMCMT_SOLVER=mf timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_mf_cfg2_launches.csv python tools/profile_step.py 200 100 30 1 > gpurun_out/${TAG}_prof.log 2>&1
HMCMT_SOLVER=mf timeout 600 python bench.py --steps 20 --warmup 3 --no-strong-scaling --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2_mf.json 2>> gpurun_out/${TAG}_prof.log
tail -2 gpurun_out/${TAG}_prof.log
