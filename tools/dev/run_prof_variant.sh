#!/bin/bash
# timing experiment: instrumented kernel with the pass-2 trailing updates removed (results are wrong on purpose)
set -e
cd /root/repo
cp hmcmt2d_b200/csrc/band_factor.cuh /tmp/bf_orig.cuh
cp hmcmt2d_b200/csrc/hmcmt_b200.cu /tmp/hb_orig.cu
restore() { cp /tmp/bf_orig.cuh hmcmt2d_b200/csrc/band_factor.cuh; cp /tmp/hb_orig.cu hmcmt2d_b200/csrc/hmcmt_b200.cu; python hmcmt2d_b200/build.py --force > /dev/null; }
trap restore EXIT
python tools/dev/prof_patch.py /tmp/bf_orig.cuh hmcmt2d_b200/csrc/band_factor.cuh
python - <<'PY'
p='/root/repo/hmcmt2d_b200/csrc/band_factor.cuh'
s=open(p).read()
old='''                    if (i == ip || i == ip1 || i == idg) continue;
                    update(i, buf);'''
assert old in s
s=s.replace(old,'''                    if (i == ip || i == ip1 || i == idg) continue;
                    if (s < 0) update(i, buf);''')
open(p,'w').write(s)
PY
echo 'extern "C" int hmcmt_debug_prof(unsigned long long* out) { return cudaMemcpyFromSymbol(out, hmcmt::g_prof, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : -99; }' >> hmcmt2d_b200/csrc/hmcmt_b200.cu
python hmcmt2d_b200/build.py --force > /dev/null
/usr/local/graft/bin/gpurun --timeout 600 -- 'timeout 300 python tools/dev/t_prof.py' 2>&1 | grep -E "per-step|status"
