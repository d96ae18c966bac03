#!/bin/bash
# build an instrumented library copy, run t_prof.py on the GPU, restore the sources
set -e
cd /root/repo
cp hmcmt2d_b200/csrc/band_factor.cuh /tmp/bf_orig.cuh
cp hmcmt2d_b200/csrc/hmcmt_b200.cu /tmp/hb_orig.cu
python tools/dev/prof_patch.py /tmp/bf_orig.cuh hmcmt2d_b200/csrc/band_factor.cuh
echo 'extern "C" int hmcmt_debug_prof(unsigned long long* out) { return cudaMemcpyFromSymbol(out, hmcmt::g_prof, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : -99; }' >> hmcmt2d_b200/csrc/hmcmt_b200.cu
python hmcmt2d_b200/build.py > /dev/null
/usr/local/graft/bin/gpurun --timeout 600 -- 'timeout 300 python tools/dev/t_prof.py' 2>&1 | grep -E "per-step|status"
cp /tmp/bf_orig.cuh hmcmt2d_b200/csrc/band_factor.cuh
cp /tmp/hb_orig.cu hmcmt2d_b200/csrc/hmcmt_b200.cu
python hmcmt2d_b200/build.py > /dev/null
