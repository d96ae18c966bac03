#!/bin/bash
# launch list of the large-bandwidth path (cfg4 mesh, 2 frequencies): a few hundred launches from the middle of the factorisation + the sweeps
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 2000 -c 300 --csv --log-file gpurun_out/big_launches.csv python tools/dev/t_big.py 800 300 2 > gpurun_out/big_prof.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"band_solve|k_" -c 40 --csv --log-file gpurun_out/big_launches2.csv python tools/dev/t_big.py 800 300 2 >> gpurun_out/big_prof.log 2>&1
tail -3 gpurun_out/big_prof.log
