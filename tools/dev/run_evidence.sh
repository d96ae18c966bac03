#!/bin/bash
# Evidence for profiles/: launch lists (ncu, serialised), DRAM bytes per launch, full captures of the dominant kernels, bench lines.
# Everything under its own timeout; reports of single launches are small enough to travel back.   bash tools/dev/run_evidence.sh TAG
TAG=${1:-r02}
O=gpurun_out
SER="HMCMT_GROUPS=1 HMCMT_GRAPH=0"
# 1. launch lists with DRAM bytes (one un-grouped, eagerly launched step)
env $SER timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_cfg2_launches.csv python tools/profile_step.py 200 100 30 1 > $O/${TAG}_cfg2_launches.log 2>&1
env $SER timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_cfg4x16_launches.csv python tools/profile_step.py 800 300 8 1 > $O/${TAG}_cfg4x16_launches.log 2>&1
# 2. full captures of single launches inside a cfg2 step: leaf level of the small-front kernel, its 16-warp instantiation, the
#    largest trailing-update GEMM, the leaf level of the backward substitution
cap() {   # name, kernel regex, launch-skip, workload args
  env $SER timeout 300 ncu --set full --import-source on --clock-control none --kernel-name regex:$2 --launch-skip $3 --launch-count 1 -f -o $O/${TAG}_$1 python tools/profile_step.py $4 > /dev/null 2>&1
}
cap small_leaf mf_small_kernel 1 "200 100 30 1"
cap small_w16 mf_small_kernel 6 "200 100 30 1"
cap gemm_cfg2 mf_gemm_kernel 3 "200 100 30 1"
cap bwd_leaf mf_bwd_warp_kernel 4 "200 100 30 1"
cap gemm_cfg4 mf_gemm_kernel 25 "800 300 8 1"
# 3. bench lines
timeout 600 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_cfg2_1gpu.json 2> $O/${TAG}_bench_cfg2_1gpu.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_cfg2_reference_arm.json 2> $O/${TAG}_bench_reference.err
ls -la $O/${TAG}_*
