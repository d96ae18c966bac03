"""gpurun_out/TAG_* (tools/dev/run_evidence.sh) -> tracked summaries under profiles/.   python tools/dev/make_profiles.py r02"""
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def launch_summary(name):
    src = os.path.join(G, f"{TAG}_{name}_launches.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(P, f"{TAG}_{name}_launches.csv"))
    rows = list(csv.DictReader([l for l in open(src) if not l.startswith("==")]))
    per = {}
    for r in rows:
        k = r["Kernel Name"].split("(")[0]
        v, u, m = float(r["Metric Value"].replace(",", "")), r["Metric Unit"], r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v = v / 1e3 if u == "ns" else (v if u == "us" else v * 1e3)
        else:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        e = per.setdefault(k, {"n": 0})
        e[m] = e.get(m, 0.0) + v
        e["n"] += m == "gpu__time_duration.sum"
    tot = sum(e["gpu__time_duration.sum"] for e in per.values())
    totb = sum(e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0) for e in per.values())
    with open(os.path.join(P, f"{TAG}_{name}_launches_summary.txt"), "w") as f:
        f.write(f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one leapfrog step,\n"
                f"un-grouped and eagerly launched (HMCMT_GROUPS=1 HMCMT_GRAPH=0): serialised, cold-cache per-launch times\n"
                f"total {tot / 1e3:.3f} ms over {sum(e['n'] for e in per.values())} launches, {totb / 1e9:.3f} GB of DRAM traffic\n")
        for k, e in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
            b = e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0)
            t = e["gpu__time_duration.sum"]
            f.write(f"{k[:46]:46s} n={e['n']:3d} total={t / 1e3:8.3f} ms share={100 * t / tot:5.1f}%  dram={b / 1e6:9.1f} MB ({b / t / 1e3:6.0f} GB/s)\n")


def report(name):
    rep = os.path.join(G, f"{TAG}_{name}.ncu-rep")
    if not os.path.exists(rep):
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "report", rep], capture_output=True, text=True).stdout
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    keep = [l for l in det.splitlines() if any(k in l for k in ("Duration", "Throughput", "Hit Rate", "Busy", "Occupancy", "Eligible", "Issued Warp",
                                                                  "bank conflict", "Registers Per", "Shared Memory Per", "Block Limit", "Grid Size", "Block Size"))]
    with open(os.path.join(P, f"{TAG}_{name}_ncu.txt"), "w") as f:
        f.write(f"ncu --set full --clock-control none --import-source on, one launch inside a leapfrog step ({TAG}_{name}.ncu-rep)\n\n")
        f.write(out + "\n--- details page (selection) ---\n" + "\n".join(keep) + "\n")


for n in ("cfg2", "cfg4x16"):
    launch_summary(n)
for n in ("small_leaf", "small_w16", "gemm_cfg2", "gemm_cfg4", "bwd_leaf"):
    report(n)
for f in os.listdir(G):
    if f.startswith(TAG + "_bench") and f.endswith(".json") and os.path.getsize(os.path.join(G, f)) > 0:
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
print("\n".join(sorted(x for x in os.listdir(P) if x.startswith(TAG))))
