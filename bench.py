#!/usr/bin/env python
"""bench.py — HMC leapfrog steps/s of the HMCMT2D forward + adjoint hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one leapfrog step of the reference sampler (HMCSampler.jl:235-265): position drift with step
clipping and bound reflection, one compDataGradient-equivalent evaluation (forward + adjoint over all
frequencies x TE/TM), prior gradient, momentum kick.  Workload = BASELINE.json configs[1]/[2]: synthetic
200x100-cell mesh, 30 frequencies, TE+TM, one independent chain per GPU (weak scaling, no data-path
collective: "replicas only", as parallelHMC.jl).

    python bench.py --config cfg4 --gpus N [--nfreq F]       (BASELINE.json configs[3]; not the driver's default line)
the 800x300-cell mesh with 60 frequencies, frequency-sharded over N >= 2 GPUs (strong scaling; the 120 factors need
154 GB, so N = 1 only runs with a reduced --nfreq), one NCCL all-reduce of [gradient | misfit] per leapfrog step.

`value` : steps/s with the chain state resident in HBM (hmcmt_leapfrog_steps_device), CUDA events on the
          library's stream, max over ranks.
`e2e`   : the same step through the reference-facing call (compDataGradient-equivalent through the C ABI)
          with HOST buffers: the model goes host->device and predicted data / misfit / gradient come back
          every step; drift and kick run on the host exactly as the reference's proposeLeapfrog does.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG4 = dict(workload="cfg4: synthetic 800x300-cell mesh, 60 frequencies, TE+TM, 1 HMC chain, frequencies sharded over the GPUs "
                     "(one NCCL sum-all-reduce of [gradient | misfit] per step)",
            ny=800, nz=300, nfreq=60, nrx=40, modes="TE+TM", chains_per_gpu=1)
METRIC = "leapfrog_steps_per_sec"
UNIT = "steps/s"
WORKLOAD = dict(workload="cfg2: synthetic 200x100-cell mesh, 30 frequencies, TE+TM, 1 HMC chain per GPU",
                ny=200, nz=100, nfreq=30, nrx=40, modes="TE+TM", chains_per_gpu=1)
# dram__bytes_read + dram__bytes_write of the factorisation group per step: FM_OWN capture (24.3 MB + 2.251 GB,
# profiles/r01_final_factor_own_ncu.txt) + its back-substitution sweep, which streams the factor once exactly like the
# captured solve sweep (2.279 GB + 23.6 MB, profiles/r01_final_solve_own_ncu.txt); FM_SEP (13 of 2476 macro-steps) neglected
NCU_FACTOR_DRAM_BYTES = (24.28e6 + 2.250943e9) + (2.278771e9 + 23.64e6)
FP64_DMMA_PEAK_TFLOPS = 37.1     # measured on this pool's B200 with DMMA.8x8x4 (profiles/r01_fp64_peak_ubench.txt)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  nvidia-smi needs a moment
    to start, so it is launched ahead of the warm-up (`start`), `mark` records when the timed region begins, and `stop` keeps
    the samples whose own timestamps fall inside the region (or, if the region was shorter than one sampling period, the
    samples taken under the identical load right before it — `window` says which)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.t_mark = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            import select
            if select.select([self.proc.stdout], [], [], 5.0)[0]:      # wait (bounded) until nvidia-smi is up:
                self.proc.stdout.readline()                             # its first sample
        except Exception:
            self.proc = None

    def mark(self):
        self.t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.time()
        time.sleep(0.06)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        import datetime
        rows = []
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[2]), float(f[3]), f[6:10]))
            except ValueError:
                continue
        t0 = self.t_mark if self.t_mark is not None else 0.0
        inside = [r for r in rows if t0 <= r[0] <= t_end + 0.06]
        window = "timed region"
        if not inside:
            inside = rows[-3:]                   # same load (warm-up steps of the same workload) right before the region
            window = "timed region shorter than the sampling period: last samples under the same load before it"
        sm, mx, reasons = [], [], set()
        for _, a, b, flags in inside:
            sm.append(a); mx.append(b)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), flags):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def build_problem(cfg=None):
    from hmcmt2d_b200 import synthetic
    cfg = cfg or WORKLOAD
    mesh, data, inv, prior = synthetic.make_problem(cfg["ny"], cfg["nz"], cfg["nfreq"], cfg["nrx"])
    return mesh, data, inv, prior


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restated reference algorithm, SciPy SuperLU in place of UMFPACK/MUMPS) on host cores.

def _cpu_freq_job(args):
    """forward + adjoint for ONE frequency (TE+TM) of the workload — the unit the host cores are farmed over."""
    fidx, seed, cfg = args
    sys.path.insert(0, ROOT)
    from hmcmt2d_b200 import synthetic
    from oracle import fileio as ofio
    from oracle import sampler as osamp
    mesh, data, inv, prior = build_problem(cfg)
    keep = data.freqID == fidx + 1
    omesh = ofio.TensorMesh2D(mesh.yLen, mesh.zLen, mesh.airLayer, mesh.gridSize, mesh.origin, mesh.sigma)
    od = ofio.MTData(data.rxLoc, data.freqs[fidx:fidx + 1], data.dataType, data.dataComp, data.rxID[keep],
                     np.ones(int(keep.sum()), np.int64), data.dtID[keep], np.ones(int(keep.sum()), bool), True, True)
    oinv = osamp.setupInverseDataModel(omesh, [1e-8], inv.obsData[keep], inv.dataErr[keep])
    oinv.strModel = synthetic.stress_model(inv, seed)
    t0 = time.perf_counter()
    osamp.compDataGradient(omesh, od, oinv, ofio.HMCPrior())
    return time.perf_counter() - t0


def cpu_steps_per_sec(nworkers: int, nfreq_sample: int, repeats: int = 1, cfg=None):
    """Times `nfreq_sample` of the 30 frequencies farmed over `nworkers` processes and extrapolates to the full
    step (frequencies are independent and cost the same: identical sparsity pattern)."""
    import multiprocessing as mp
    # one process per core, ONE thread per process: the workers inherit these (spawn) before they import numpy / scipy —
    # without them every worker starts a full BLAS / OpenMP team and the oversubscribed host takes minutes per sample
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[var] = "1"
    cfg = cfg or WORKLOAD
    freqs = list(np.linspace(0, cfg["nfreq"] - 1, nfreq_sample).astype(int))
    ctx = mp.get_context("spawn")
    times = []
    with ctx.Pool(nworkers) as pool:
        if cfg["ny"] * cfg["nz"] < 100000:
            pool.map(_cpu_freq_job, [(freqs[0], 1, cfg)] * nworkers)   # warm the workers (imports, operator setup)
        for r in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_freq_job, [(int(f), 1, cfg) for f in freqs])
            times.append(time.perf_counter() - t0)
    t = min(times)
    step_time = t * cfg["nfreq"] / nfreq_sample
    return 1.0 / step_time, t


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nsample = min(WORKLOAD["nfreq"], cores)
    vals = []
    # every "step" of this arm is one bounded sample (nsample frequencies farmed over all host cores, extrapolated to the
    # full step); at most three samples, whatever --steps says, so that the arm ends within a minute or two
    for _ in range(max(1, min(args.steps, 3))):
        v, wall = cpu_steps_per_sec(cores, nsample)
        vals.append(v)
    value = float(np.median(vals))
    sample = (f"{nsample} of {WORKLOAD['nfreq']} frequencies (TE+TM forward+adjoint each) farmed over {cores} host processes, "
              f"extrapolated x{WORKLOAD['nfreq'] / nsample:.2f}; restated CPU path (SciPy SuperLU), not MUMPS/UMFPACK")
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000.0 / value, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64 (complex128)",
                data="synthetic", config=WORKLOAD, impl="reference",
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------

def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from hmcmt2d_b200 import api, synthetic
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — hmcmt2d_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    mesh, data, inv, prior = build_problem()
    pl = api.Plan(mesh, data, inv, prior, nChains=1, device=local_rank)
    rng = np.random.default_rng(100 + rank)
    m0 = synthetic.stress_model(inv, seed=1 + rank)            # chains differ per GPU (seeds 1..N, cfg3)
    p0 = np.clip(rng.standard_normal(len(m0)), -2.5, 2.5)
    dt = prior.dt
    K, W = args.steps, max(3, args.warmup)

    # ---------------- device-resident arm (`value`) ----------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    pl.set_state(m0, p0, m0)
    pl.leapfrog_steps_device(dt, W)
    pl.sync()
    pl.kernel_time(reset=True)
    launches0 = pl.info(10)
    barrier()
    sampler.mark()
    pl.timer_start()
    pl.leapfrog_steps_device(dt, K)
    ms = pl.timer_stop()
    barrier()
    clocks = sampler.stop()
    launches = pl.info(10) - launches0
    factor_ms, factor_n = pl.kernel_time(reset=True)
    ms = max_over_ranks(ms)
    value = world * K / (ms * 1e-3)

    # ---------------- end-to-end arm through the host-buffer C ABI ----------------
    lo, hi = np.log(prior.sigBounds[0]), np.log(prior.sigBounds[1])
    Wm, beta = inv.Wm, prior.regParam

    def host_step(m, p):
        dm = dt * p
        mx = np.abs(dm).max()
        if mx > 3.0:
            dm = dm / mx * 3.0
        m = m + dm
        for _ in range(500):                                   # checkParameterBound! (HMCSampler.jl:515-559)
            low, high = m < lo, m > hi
            if not (low.any() or high.any()):
                break
            m = np.where(low, 2 * lo - m, m); p = np.where(low, -p, p)
            high = m > hi
            m = np.where(high, 2 * hi - m, m); p = np.where(high, -p, p)
        pred, phi, g = pl.forward_gradient(m)                  # H2D m ; D2H pred, phi, grad  (pinned staging inside)
        p = p - dt * (g[0] + beta * (Wm @ (m - m0)))
        return m, p, float(phi[0])

    m, p = m0.copy(), p0.copy()
    for _ in range(W):
        m, p, _ = host_step(m, p)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        m, p, phi = host_step(m, p)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * K / e2e_s
    h2d = pl.nAC * 8
    d2h = pl.nData * 16 + 8 + pl.nAC * 8

    # ---------------- roofline of the dominant kernel (band_factor_kernel) ----------------
    peaks, peak_src = load_peaks()
    N, b, nsys = pl.info(0), pl.info(4), pl.info(7)
    # algorithmic work per launch: factor 4 N b^2 real flops per system + the fused forward
    # elimination / back-substitution 16 N b ; bytes: factor written once + read once by the fused back-substitution
    flops_launch = (4.0 * N * b * b + 16.0 * N * b) * nsys
    bytes_launch = 2.0 * 16.0 * N * (b + 1) * nsys
    fac_ms = factor_ms / max(1, factor_n)
    achieved = flops_launch / (fac_ms * 1e-3) / 1e12
    roofline = dict(bound="tensor", kernel="factorisation of the 60 systems = band_factor_kernel<14> FM_OWN + FM_SEP (FP64 DMMA.8x8x4 block LDL^T, "
                    "fused assembly + forward elimination) + band_solve_kernel<14> SM_BACKZ_OWN (its back-substitution), timed as one unit",
                    achieved=achieved, peak=FP64_DMMA_PEAK_TFLOPS, unit="TFLOP/s", frac=achieved / FP64_DMMA_PEAK_TFLOPS,
                    peak_source="measured FP64 DMMA m8n8k4 rate on this pool's B200 (profiles/r01_fp64_peak_ubench.txt); "
                                "MEASURED_PEAKS.json holds only bf16/HBM peaks: " + peak_src,
                    traffic=NCU_FACTOR_DRAM_BYTES, traffic_source="dram__bytes_read.sum + dram__bytes_write.sum of ncu --set full captures at this workload: FM_OWN launch "
                                   "(profiles/r01_final_factor_own_ncu.txt) + one factor-streaming sweep (profiles/r01_final_solve_own_ncu.txt)",
                    algorithmic_flops_per_launch=flops_launch, algorithmic_bytes_per_launch=bytes_launch,
                    hbm_achieved_gbs=bytes_launch / (fac_ms * 1e-3) / 1e9, hbm_peak_gbs=peaks.get("hbm_gbs"),
                    avg_launch_ms=fac_ms, launches_timed=factor_n, share_of_step=fac_ms / (ms / K))

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms / K,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64 (complex128)", data="synthetic",
                config=dict(WORKLOAD, l2_policy="inputs larger than L2: each step streams the 2.3 GB block-LDL^T factor of the "
                                                "60 systems (written once, read 3x) through the 126 MB L2",
                            parallelism=f"chains x{world} (replicas only, no data-path collective)"),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=1000 * e2e_s / K),
                gpu_launches=int(launches), clocks=clocks, roofline=roofline)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v1, wall1 = cpu_steps_per_sec(1, 1)
        line["cpu_baseline"] = dict(value=v1, unit=UNIT, cores=1, kind="port",
                                    sample=f"1 of {WORKLOAD['nfreq']} frequencies (TE+TM forward+adjoint, {wall1:.2f} s) on one host core, "
                                           f"extrapolated x{WORKLOAD['nfreq']}; restated CPU path (oracle, SciPy SuperLU), not MUMPS. "
                                           f"All-core figure: run `bench.py --impl reference` ({cores} cores on this box)")
    if rank == 0:
        print(json.dumps(line), flush=True)
    pl.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# cfg4: one chain, frequencies sharded over the ranks, one NCCL all-reduce per step (strong scaling)

def run_gpu_cfg4(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from hmcmt2d_b200 import api, synthetic
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — hmcmt2d_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cfg = dict(CFG4)
    if args.nfreq:
        cfg["nfreq"] = args.nfreq
        cfg["workload"] += f" [reduced to {args.nfreq} frequencies]"
    if world == 1 and cfg["nfreq"] > 48:
        raise SystemExit("bench.py --config cfg4: the 120 factors (1.28 GB each) do not fit one GPU; use --gpus >= 2 or --nfreq <= 48")
    mesh, data, inv, prior = build_problem(cfg)
    sp = api.FreqShardedPlan(mesh, data, inv, prior, rank, world, device=local_rank)
    pl = sp.plan
    m0 = synthetic.stress_model(inv, seed=1)                                   # the same chain state on every rank
    p0 = np.clip(np.random.default_rng(100).standard_normal(len(m0)), -2.5, 2.5)
    dt = prior.dt
    K, W = args.steps, max(3, args.warmup)

    sampler = ClockSampler(local_rank)
    sampler.start()
    sp.set_state(m0, p0, m0)
    sp.leapfrog_steps_device(dt, W)
    sp.sync()
    pl.kernel_time(reset=True)
    launches0 = pl.info(10)
    barrier()
    sampler.mark()
    pl.timer_start()
    sp.leapfrog_steps_device(dt, K)
    ms = pl.timer_stop()
    barrier()
    clocks = sampler.stop()
    launches = pl.info(10) - launches0
    factor_ms, factor_n = pl.kernel_time(reset=True)
    ms = max_over_ranks(ms)
    value = K / (ms * 1e-3)
    m_end, _ = sp.get_state()
    drift_between_ranks = max_over_ranks(float(np.abs(m_end).sum())) - (-max_over_ranks(-float(np.abs(m_end).sum())))

    # end to end: compDataGradient through host buffers (H2D model, D2H data / misfit / gradient, all-reduce) + host leapfrog
    lo, hi = np.log(prior.sigBounds[0]), np.log(prior.sigBounds[1])
    Wm, beta = inv.Wm, prior.regParam

    def host_step(m, p):
        dm = dt * p
        mx = np.abs(dm).max()
        if mx > 3.0:
            dm = dm / mx * 3.0
        m = np.clip(m + dm, lo, hi)
        pred, phi, g = sp.forward_gradient(m)
        return m, p - dt * (g + beta * (Wm @ (m - m0))), phi

    Ke = max(1, min(K, 5))
    m, p = m0.copy(), p0.copy()
    m, p, _ = host_step(m, p)
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        m, p, phi = host_step(m, p)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()

    N, b, nsys = pl.info(0), pl.info(4), pl.info(7)
    flops_launch = (4.0 * N * b * b + 16.0 * N * b) * nsys
    bytes_launch = 2.0 * 16.0 * N * (b + 1) * nsys
    fac_ms = factor_ms / max(1, factor_n)
    achieved = flops_launch / (fac_ms * 1e-3) / 1e12
    peaks, peak_src = load_peaks()
    roofline = dict(bound="tensor", kernel="large-bandwidth factorisation of this rank's systems: bigband_panel_kernel + "
                    "bigband_update_kernel (DMMA.8x8x4) per 32 columns + backward sweep (band_big.cuh)",
                    achieved=achieved, peak=FP64_DMMA_PEAK_TFLOPS, unit="TFLOP/s", frac=achieved / FP64_DMMA_PEAK_TFLOPS,
                    peak_source="measured FP64 DMMA m8n8k4 rate on this pool's B200 (profiles/r01_fp64_peak_ubench.txt); " + peak_src,
                    traffic=None, algorithmic_flops_per_launch=flops_launch, algorithmic_bytes_per_launch=bytes_launch,
                    hbm_achieved_gbs=bytes_launch / (fac_ms * 1e-3) / 1e9, hbm_peak_gbs=peaks.get("hbm_gbs"),
                    avg_launch_ms=fac_ms, launches_timed=factor_n, share_of_step=fac_ms / (ms / K),
                    note="latency-bound: ~7.5k stream-ordered panel/update launch pairs per factorisation")
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms / K, higher_is_better=True,
                scaling="strong", vs_baseline=None, dtype="f64 (complex128)", data="synthetic",
                config=dict(cfg, systems_per_gpu=int(nsys), l2_policy="inputs larger than L2: each rank streams its 1.28 GB-per-system factors",
                            parallelism=f"frequencies x{world} (NCCL sum-all-reduce of [gdata | phi_d], {8 * (pl.nAC + 1)} B per step)",
                            state_spread_between_ranks=drift_between_ranks),
                e2e=dict(value=Ke / e2e_s, unit=UNIT, h2d_bytes_per_step=pl.nAC * 8, d2h_bytes_per_step=pl.nData * 16 + 8 + pl.nAC * 8,
                         ms_per_step=1000 * e2e_s / Ke, steps=Ke),
                gpu_launches=int(launches), clocks=clocks, roofline=roofline)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v1, wall1 = cpu_steps_per_sec(1, 1, cfg=cfg)
        line["cpu_baseline"] = dict(value=v1, unit=UNIT, cores=1, kind="port",
                                    sample=f"1 of {cfg['nfreq']} frequencies (TE+TM forward+adjoint, {wall1:.1f} s) on one host core, extrapolated; "
                                           "matrix-free restatement (the reference's dense dBC needs 8.4 GB per system at this size)")
    if rank == 0:
        print(json.dumps(line), flush=True)
    sp.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg4"])
    ap.add_argument("--nfreq", type=int, default=0, help="cfg4 only: reduced number of frequencies (testing)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--config", args.config, "--nfreq", str(args.nfreq)]
        raise SystemExit(subprocess.call(cmd))
    if args.config == "cfg4":
        run_gpu_cfg4(args, rank, world, local_rank)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
