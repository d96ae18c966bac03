"""Writes an instrumented copy of band_factor.cuh (clock64 phase timers, FM_OWN launch, CTA 0) to the path given; dev tool."""
import sys
src, dst = sys.argv[1], sys.argv[2]
s = open(src).read()
s = s.replace("enum { BAR_RAW = 1, BAR_INV = 2, BAR_M = 3, BAR_GJ = 4 /* and 5: alternates with the step parity */ };", "enum { BAR_RAW = 1, BAR_INV = 2, BAR_M = 3, BAR_GJ = 4 /* and 5: alternates with the step parity */ };\n__device__ unsigned long long g_prof[32];\n#define PT(k) do { if (blockIdx.x == 0 && lane == 0) { unsigned long long _n = clock64(); prof[k] += _n - tlast; tlast = _n; } } while (0)")
def rep(a, b):
    global s
    assert a in s, a[:60]
    s = s.replace(a, b, 1)
rep("            for (int s = s0; s < s1; ++s) {\n                const int p = s % T, p1 = (s + 1) % T, p2 = (s + 2) % T, buf = s & 1;",
    "            unsigned long long prof[8] = {0,0,0,0,0,0,0,0}, tlast = clock64();\n            for (int s = s0; s < s1; ++s) {\n                const int p = s % T, p1 = (s + 1) % T, p2 = (s + 2) % T, buf = s & 1;\n                PT(5);")
rep("                if (__any_sync(0xffffffffu, sm.fail != 0)) break;", "                if (__any_sync(0xffffffffu, sm.fail != 0)) break;\n                PT(0);")
rep("                bar_sync(BAR_M, NW * 32);\n", "                PT(1);\n                bar_sync(BAR_M, NW * 32);\n                PT(2);\n")
rep("                // pass 1: the tile of the next pivot column", "                PT(3);\n                // pass 1: the tile of the next pivot column")
rep("                if (sys.rhs && myX >= 0 && warp != 0) y_update(myX, buf);", "                PT(4);\n                if (sys.rhs && myX >= 0 && warp != 0) y_update(myX, buf);")
rep("            if (mode == FM_OWN) {\n                // end of the own lines: export the window",
    "            if (mode == FM_OWN && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 7)) for (int k = 0; k < 6; ++k) g_prof[(warp ? 8 : 0) + k] = prof[k];\n            if (mode == FM_OWN) {\n                // end of the own lines: export the window")
rep("            for (int s = s0; s < s1; ++s) {\n                const int p = s % T, p1 = (s + 1) % T, buf = s & 1;\n                const int rp = p * TS;",
    "            unsigned long long prof[8] = {0,0,0,0,0,0,0,0}, tlast = clock64();\n            for (int s = s0; s < s1; ++s) {\n                const int p = s % T, p1 = (s + 1) % T, buf = s & 1;\n                const int rp = p * TS;\n                PT(0);")
rep("                if (bad) break;\n", "                if (bad) break;\n                PT(1);\n")
rep("                    invert();\n                    publish(buf ^ 1);\n                }", "                    PT(2);\n                    invert();\n                    PT(3);\n                    publish(buf ^ 1);\n                }")
rep("                // ring slot of step s+kPre", "                PT(4);\n                // ring slot of step s+kPre")
rep("                if (s + 1 < s1) bar_sync(BAR_RAW, NTHR);        // raw(s+1) complete (y_p(s+1) final as well)\n            }",
    "                PT(5);\n                if (s + 1 < s1) bar_sync(BAR_RAW, NTHR);\n                PT(6);\n            }\n            if (mode == FM_OWN && blockIdx.x == 0 && lane == 0) for (int k = 0; k < 7; ++k) g_prof[16 + k] = prof[k];")
open(dst, 'w').write(s)
