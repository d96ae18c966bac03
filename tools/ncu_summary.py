"""Summaries of ncu outputs used for profiles/: launch list shares and key counters of one report."""
import csv, subprocess, sys

def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    tot = {}
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = row['Kernel Name'].split('(')[0][:48]
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e6 if u == 'ns' else (v / 1e3 if u == 'us' else v)
        tot.setdefault(k, [0.0, 0])
        tot[k][0] += v; tot[k][1] += 1
    s = sum(v[0] for v in tot.values())
    print(f"total {s:.3f} ms over {sum(v[1] for v in tot.values())} launches")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:50s} n={v[1]:3d} total={v[0]:9.3f} ms avg={v[0]/v[1]:8.3f} ms share={100*v[0]/s:5.1f}%")

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread ', 'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.max ', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__inst_executed.sum ', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit', 'launch__shared_mem_per_block_dynamic']

def report(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    h, units, v = r[0], r[1], r[2]
    for i, name in enumerate(h):
        if any((name + ' ').startswith(k) for k in KEYS) or name in ('Kernel Name',):
            print(f"{name} = {v[i]} {units[i]}")

if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    else:
        report(sys.argv[2])
