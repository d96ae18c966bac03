#!/bin/bash
# ncu --set full of single launches of the multifrontal kernels inside one cfg2 step (un-grouped); raw + source pages exported as
# CSV on the box (the reports themselves stay in /tmp there: several launches exceed the 64 MiB return limit)
TAG=${1:-r02k}
export HMCMT_GROUPS=1
cap() {   # name, kernel regex, launch-skip
  timeout 300 ncu --set full --import-source on --clock-control none --kernel-name regex:$2 --launch-skip $3 --launch-count 1 -f -o /tmp/${TAG}_$1 python tools/profile_step.py 200 100 30 1 > /dev/null 2>&1
  ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_$1.ncu-rep --page source --csv > gpurun_out/${TAG}_$1_source.csv 2>/dev/null
  ncu -i /tmp/${TAG}_$1.ncu-rep --page details > gpurun_out/${TAG}_$1_details.txt 2>/dev/null
}
for spec in "$@"; do :; done
cap small2 mf_small_kernel 1
cap small16 mf_small_kernel 6
#cap gemm mf_gemm_kernel 3
ls -la gpurun_out/${TAG}_*
