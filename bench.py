#!/usr/bin/env python
"""bench.py -- headline benchmark of the accelerated PEUQSE path (driver contract in the task statement).

Workload (BASELINE.json configs[1], "C2"): Example00 doMultiStart grid sampling -- quadratic model
y=(x-a)^2+b, 2 parameters, 3 observations, radii [5000,5000] => 100 020 001 grid points, MAP arg-max +
weighted moments + marginal histograms.  One "step" = one complete multistart over the grid:
  pass 1 (logP of every point materialised in HBM + first strict maximum), pass 2 (weights + moments),
  edges + pass 3 (marginal histograms).
Metric: fp64 log-posterior evals/s = grid points / step time (pass-1 evaluations only are counted).
With N GPUs the grid's first axis is widened so that every rank owns ~1e8 points (weak scaling); the three
tiny collectives (all-gather of the best point, all-reduce of moments / histograms) are inside the step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nproc-per-node N bench.py --gpus N ...
"""
import argparse
import ctypes
import json
import os
import statistics as stats
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

X00 = [600, 1100, 1400]
Y00 = [360500, 580500, 1620500]
S00 = [200000, 300000, 200000]
MU, SD = [200, 500], [100, 200]
INTERVAL = [0.08, 0.16]
METRIC = "fp64_log_posterior_evals_per_s"
UNIT = "evals/s"


def grid_for(n_gpus):
    """radii so that the grid has ~n_gpus * 100 020 001 points (exactly that at n_gpus = 1)."""
    r0 = (10001 * n_gpus - 1) // 2          # more rows of the same width: every rank sees the tiling of the 1-GPU run
    radii = [r0, 5000]
    total = (2 * radii[0] + 1) * (2 * radii[1] + 1)
    return radii, total


def workload_name(n_gpus):
    radii, total = grid_for(n_gpus)
    return (f"Example00 doMultiStart grid getLogP: quadratic y=(x-a)^2+b, k=2, n=3, center {MU}, interval {INTERVAL}, "
            f"radii {radii} => {total} points; MAP argmax + weighted moments + 10-bin marginal histograms")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# CPU side (checker + reported baseline only).  The UNMODIFIED reference is used when its package can be imported -- from
# baseline/_ref (pip install --no-deps --target baseline/_ref of a copy of /root/reference; travels to the GPU box) or from
# /root/reference -- under the import shim of SURVEY 8c (matplotlib is not installed: MagicMock modules; nothing in the
# reference is modified).  Otherwise the oracle port of its per-point path runs (kind "port").
# ------------------------------------------------------------------------------------------------
_W = {}


def _import_reference():
    """-> the reference's PEUQSE package, or None."""
    if 'ref' in _W:
        return _W['ref']
    _W['ref'] = None
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if not os.path.isdir(os.path.join(cand, "PEUQSE")):
            continue
        try:
            import tempfile
            from unittest.mock import MagicMock
            for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "matplotlib.ticker",
                      "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.mplot3d"]:
                sys.modules.setdefault(m, MagicMock())
            sys.path.insert(0, cand)
            os.chdir(tempfile.mkdtemp(prefix="peuqse_ref_"))        # the constructor creates ./graphs ./logs_and_csvs ./pickles
            import PEUQSE as REF
            _W['ref'], _W['ref_path'] = REF, cand
            break
        except Exception as e:                                        # noqa: BLE001
            _W['ref_error'] = repr(e)
            if cand in sys.path:
                sys.path.remove(cand)
    return _W['ref']


def _cpu_init():
    from oracle import peuqse_oracle as orc
    _W['orc'] = orc
    REF = _import_reference()
    if REF is not None:
        import contextlib
        import importlib
        import io
        sys.modules.pop('PEUQSE.UserInput', None)
        UI = importlib.import_module('PEUQSE.UserInput')
        x = np.asarray(X00, dtype=float)
        UI.responses['responses_abscissa'] = X00
        UI.responses['responses_observed'] = Y00
        UI.responses['responses_observed_uncertainties'] = S00
        UI.model['InputParameterPriorValues'] = MU
        UI.model['InputParametersPriorValuesUncertainties'] = SD
        # the user plugin of Examples/Example00/simulation_model_00.py:6-10 (the example files are not part of the package)
        UI.model['simulateByInputParametersOnlyFunction'] = lambda th: (x - th[0]) ** 2 + th[1]
        UI.model['exportResponses'] = False
        UI.parameter_estimation_settings['convergence_diagnostics'] = False
        UI.parameter_estimation_settings['mcmc_length'] = 10
        with contextlib.redirect_stdout(io.StringIO()):
            _W['PE'] = REF.parameter_estimation(UI)
        _W['kind'] = "reference"
    else:
        _W['P'] = orc.build_problem(MU, SD, Y00, S00)
        _W['pt'] = orc.sim_quadratic(X00)[0]
        _W['kind'] = "port"


def _cpu_eval_range(args):
    """getLogP one point at a time, the reference's loop body (IP.py:860): the unmodified parameter_estimation.getLogP when
    the reference is importable, else its oracle port (scipy logpdf per point)."""
    first, count, radii = args
    orc = _W['orc']
    th = orc.grid_decode([float(MU[0]), float(MU[1])], INTERVAL, radii, np.arange(first, first + count))
    best = -np.inf
    if _W['kind'] == "reference":
        f = _W['PE'].getLogP
        for t in th:
            v = f(t)
            if v > best:
                best = v
        return best
    for t in th:
        v = orc.log_p_point(_W['P'], t, _W['pt'])
        if v > best:
            best = v
    return best


def _cpu_how():
    if _W.get('kind') == "reference":
        where = os.path.relpath(_W['ref_path'], ROOT) if _W['ref_path'].startswith(ROOT) else _W['ref_path']
        return f"the UNMODIFIED reference's parameter_estimation.getLogP (imported from {where} under the matplotlib MagicMock shim)"
    return "oracle port of IP.py:1647-1654 (scipy.stats.multivariate_normal.logpdf per point; the reference package was not importable)"


def cpu_baseline_single_core(radii, budget_s=12.0):
    cwd = os.getcwd()
    _cpu_init()
    best0 = _cpu_eval_range((0, 14, radii))                         # includes grid index 13, the MAP of the Example00 grid
    t0 = time.perf_counter()
    _cpu_eval_range((0, 1500, radii))
    rate = 1500 / (time.perf_counter() - t0)
    n = int(max(2000, min(2e5, rate * budget_s)))
    t0 = time.perf_counter()
    _cpu_eval_range((1500, n, radii))
    dt = time.perf_counter() - t0
    os.chdir(cwd)
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": _W['kind'],
            "sample": f"grid points 1500..{1500 + n - 1} of the same grid, one point at a time through {_cpu_how()}, {dt:.1f} s",
            "first_points_max_logP": best0}


def run_reference_arm(args, emit=print):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, as independent workers each
    owning a slice of the grid (what the reference's mpiexec mode amounts to, IP.py:845-857)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    radii, total = grid_for(args.gpus)
    _import_reference()                                              # before the fork: the workers inherit the import
    ctx = mp.get_context("fork")
    pool = ctx.Pool(cores, initializer=_cpu_init)
    _cpu_init()
    # calibrate
    t0 = time.perf_counter()
    pool.map(_cpu_eval_range, [(i * 300, 300, radii) for i in range(cores)])
    rate = cores * 300 / (time.perf_counter() - t0)
    budget = 150.0 / max(1, args.steps + args.warmup)
    per_step = int(max(cores * 40, min(cores * 4000, rate * budget)))
    per_worker = max(1, per_step // cores)
    per_step = per_worker * cores

    def step(s):
        base = (s * per_step) % max(1, total - per_step)
        pool.map(_cpu_eval_range, [(base + i * per_worker, per_worker, radii) for i in range(cores)])
    for s in range(args.warmup):
        step(s)
    t0 = time.perf_counter()
    for s in range(args.steps):
        step(args.warmup + s)
    dt = time.perf_counter() - t0
    pool.close()
    value = per_step * args.steps / dt
    sample = (f"{per_step} grid points per step ({per_worker} per worker, {cores} fork()ed workers), getLogP one point at a "
              f"time through {_cpu_how()}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.gpus), "host": "cpu"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": _W['kind'], "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev_index):
        self.p = None
        try:
            import torch
            ident = str(dev_index)
            try:
                ident = "GPU-" + str(torch.cuda.get_device_properties(dev_index).uuid)
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                if vis:
                    ident = vis.split(",")[dev_index]
            self.p = subprocess.Popen(["nvidia-smi", "-i", ident, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": stats.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}



# ------------------------------------------------------------------------------------------------
# second half of BASELINE.json's metric: MH chain-steps/s on configs[2] ("C3": 4096 chains, 6-parameter
# TPD-style kinetics, 1024 observations, diagonal uncertainties).  Reported inside the same JSON line.
# ------------------------------------------------------------------------------------------------
C3 = {"n": 1024, "chains": 4096, "truth": [41.5, 45.0, 13.0, 13.0, 0.1, 0.1], "mu": [41.5, 41.5, 13.0, 13.0, 0.1, 0.1],
      "sd": [20.0, 20.0, 13.0, 13.0, 0.1, 0.1], "step": 2.5e-6, "seed": 0}
# per chain-step: n x (21 flop + 2 exp) + prior 24 + proposal ~ 21.5 kflop + 2048 exp.  Executed work, MEASURED
# (ncu --set full of k_mh_tpd<6,4,8,1>, profiles/r2_ncu_full_mh_c3_tpd_raw.csv, 4096 chains x 199 steps): 1169 warp instructions per
# chain-step of which 767 fp64 (24.6 k thread-level fp64 instructions; round 1: 3547 / 1780).  An fp64 warp instruction holds the
# sub-partition's issue port for two cycles, any other for one (profiles/fp64_latency.py), so a chain-step costs 2 x 767 + 402 = 1937
# issue cycles and the issue ports of one B200 deliver 2 x (fp64 warp-instruction peak) cycles per second.
C3_FP64_INSTR_PER_STEP = 24.6e3
C3_ISSUE_CYCLES_PER_STEP = 1937.0
C3_FLOP_PER_STEP = 23.6e3       # SURVEY 8d algorithmic figure (one exp counted as one flop)
FP64_INSTR_PEAK = 1.80e13     # measured DFMA / DMUL+DADD issue rate of one B200 (profiles/microbench.py), instructions/s


def c3_problem(PEUQSE, steps, chains):
    """The C3 problem built through the public API; the synthetic observations come from the product's own simulator."""
    import torch
    model = PEUQSE.Tpd2SiteModel.linear_ramp(C3["n"])

    def make(y, sig):
        UI = PEUQSE.UserInput.fresh()
        UI.responses['responses_observed_max_covmat_size'] = 5000
        UI.responses['responses_observed'] = list(y)
        UI.responses['responses_observed_uncertainties'] = list(sig)
        UI.model['InputParameterPriorValues'] = C3["mu"]
        UI.model['InputParametersPriorValuesUncertainties'] = C3["sd"]
        UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.Tpd2SiteModel.linear_ramp(C3["n"])
        pes = UI.parameter_estimation_settings
        pes.update({'mcmc_length': steps, 'mcmc_burn_in': steps // 10, 'mcmc_nchains': chains, 'mcmc_random_seed': C3["seed"],
                    'mcmc_relative_step_length': C3["step"], 'mcmc_store_samples': False,
                    'mcmc_threshold_filter_samples': False})
        UI.model['InputParameterInitialGuess'] = C3["truth"]
        return PEUQSE.parameter_estimation(UI)
    n = C3["n"]
    probe = make(np.ones(n), np.ones(n))
    sim_t = probe.simulate_batch(np.asarray(C3["truth"]))[0].cpu().numpy()
    sig = 0.02 * float(np.max(sim_t))
    y = sim_t + sig * np.random.default_rng(0).standard_normal(n)
    return make(y, np.full(n, sig)), y, sig


def mh_secondary(PEUQSE, torch, dist, rank, world, steps):
    """chain-steps/s: device-timed peuqse_mh_run over this rank's chains (moments in-kernel, samples not stored) and
    end to end through PE_object.doMetropolisHastings() (host start points in, pooled host statistics out)."""
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    from peuqse_b200 import distributed as D
    chains = C3["chains"] * world                                   # weak scaling: 4096 chains per GPU
    PE, y, sig = c3_problem(PEUQSE, steps, chains)
    c0, cn = D.shard_range(chains, rank, world)
    th0 = np.tile(np.asarray(C3["truth"]), (cn, 1))
    chol = step_cholesky(PE.Q_covmat, C3["step"])
    run = lambda: mh_launch(PE, th0, steps, steps // 10, C3["seed"], chol, chain_offset=c0, store_samples=False)
    run(); run(); run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 3
    a.record()
    for _ in range(reps):
        out = run()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    acc = float(out['stats'][:, 0].mean().item()) / max(steps - 1, 1)
    PE.doMetropolisHastings()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    PE.doMetropolisHastings()
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    value = chains * (steps - 1) / (ms * 1e-3)
    per_gpu = value / world
    sec = {"metric": "mh_chain_steps_per_s", "value": value, "unit": "chain-steps/s", "ms_per_run": ms,
           "config": {"workload": f"{chains} Metropolis-Hastings chains x {steps} steps, 6-parameter two-site TPD kinetics, "
                                  f"{C3['n']} observations, diagonal uncertainties, relative step {C3['step']}, in-kernel moments "
                                  "(samples not stored)", "chains_per_gpu": cn, "acceptance_rate": acc},
           "roofline": {"bound": "fp64", "kernel": "k_mh_tpd<k=6, up to 4 chains per warp>", "achieved": per_gpu * C3_FLOP_PER_STEP / 1e12,
                        "peak": 2.0 * FP64_INSTR_PEAK / 1e12, "unit": "TFLOP/s", "frac": per_gpu * C3_FLOP_PER_STEP / (2.0 * FP64_INSTR_PEAK),
                        "peak_source": "2 flop x measured fp64 issue rate 1.80e13 instr/s (profiles/microbench.py; MEASURED_PEAKS.json "
                                       "has no fp64 figure; nominal 37 TFLOP/s)",
                        "fp64_pipe_frac": per_gpu * C3_FP64_INSTR_PER_STEP / FP64_INSTR_PEAK,
                        "issue_frac": per_gpu * C3_ISSUE_CYCLES_PER_STEP / (2.0 * FP64_INSTR_PEAK / 32.0),
                        "note": "SURVEY 8d counts one exp as one flop (23.6 kflop per chain-step); the kernel executes 24.6 k fp64 "
                                "instructions per chain-step (ncu count; 2048 table-driven exps of 8 fp64 instructions), i.e. "
                                "fp64_pipe_frac = %.2f of the measured fp64 issue rate, and 1937 issue cycles per chain-step (fp64 = 2, "
                                "others = 1): issue_frac = %.2f of the issue ports -- the bound of this kernel; %.3g exp/s"
                                % (per_gpu * C3_FP64_INSTR_PER_STEP / FP64_INSTR_PEAK,
                                   per_gpu * C3_ISSUE_CYCLES_PER_STEP / (2.0 * FP64_INSTR_PEAK / 32.0), per_gpu * 2.0 * C3["n"]),
                        "traffic": None},
           "e2e": {"value": chains * (steps - 1) / float(te.item()), "unit": "chain-steps/s", "api": "PE_object.doMetropolisHastings()",
                   "h2d_bytes_per_step": int(th0.nbytes), "d2h_bytes_per_step": int(out['stats'].numel() * 8)}}
    if world == 1 and rank == 0:
        sec["cpu_baseline"] = mh_cpu_baseline(y, sig)
    return sec


def mh_cpu_baseline(y, sig, budget_s=8.0):
    """One chain of the oracle port of doMetropolisHastings (IP.py:2626-2675: per step one getLogP through scipy logpdf,
    n > 100 => one scalar logpdf per observation) on one host core."""
    from oracle import peuqse_oracle as orc
    n = C3["n"]
    kBT, c1, c2 = orc.tpd2site_constants(n)
    sim = orc.sim_tpd2site(kBT, c1, c2)[0]
    P = orc.build_problem(C3["mu"], C3["sd"], list(y), list(np.full(n, sig)))
    rng = np.random.default_rng(1)
    sd = np.sqrt(C3["step"]) * np.asarray(C3["sd"])
    nst, done, t0 = 6, 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        orc.mh_chain_point(P, np.asarray(C3["truth"]), rng.standard_normal((nst, 6)) * sd, rng.uniform(size=nst), sim)
        done += nst
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "chain-steps/s", "cores": 1, "kind": "port",
            "sample": f"{done} steps of one chain, point at a time (oracle.mh_chain_point), {dt:.1f} s"}

# ------------------------------------------------------------------------------------------------
# configs[0] ("C1"): Example00, doMetropolisHastings, 10 000 samples of ONE chain (the reference's own CPU-runnable case): the
# latency of the persistent kernel (thread shape) end to end through the public call, and the same chain replicated 65 536
# times per GPU (independent Philox streams) for the throughput of that kernel shape.
# ------------------------------------------------------------------------------------------------
C1 = {"steps": 10000, "replicas": 65536, "step": 0.1, "seed": 0}


def c1_secondary(PEUQSE, torch, dist, rank, world):
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    L = C1["steps"]
    UI = PEUQSE.UserInput.fresh()
    UI.responses['responses_abscissa'] = X00
    UI.responses['responses_observed'] = Y00
    UI.responses['responses_observed_uncertainties'] = S00
    UI.model['InputParameterPriorValues'] = MU
    UI.model['InputParametersPriorValuesUncertainties'] = SD
    UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.QuadraticModel(X00)
    UI.parameter_estimation_settings.update({'mcmc_length': L, 'mcmc_relative_step_length': C1["step"], 'mcmc_random_seed': C1["seed"],
                                             'mcmc_exportLog': False, 'exportLog': False, 'mcmc_nchains': world})
    PE = PEUQSE.parameter_estimation(UI)
    # one chain (per GPU: chains are what shards), end to end: host settings in, [MAP, mu_AP, stdap, evidence, info_gain, samples, logP] on the host out
    times = []
    for i in range(7):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = PE.doMetropolisHastings()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    te = torch.tensor([float(np.median(times[2:]))], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    one_s = float(te.item())
    rows = int(np.shape(res[5])[0])
    # the same chain as 65 536 replicas per GPU, in-kernel moments
    cn = C1["replicas"]
    th0 = np.tile(np.asarray(MU, dtype=float), (cn, 1))
    chol = step_cholesky(PE.Q_covmat, C1["step"])
    run = lambda: mh_launch(PE, th0, L, L // 10, C1["seed"], chol, chain_offset=rank * cn, store_samples=False)
    run(); run(); run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 3
    a.record()
    for _ in range(reps):
        out = run()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = cn * world * (L - 1) / (ms * 1e-3)
    per_gpu = value / world
    sec = {"metric": "mh_chain_steps_per_s", "value": value, "unit": "chain-steps/s", "ms_per_run": ms,
           "config": {"workload": f"Example00 quadratic, 2 parameters, 3 observations, doMetropolisHastings {L} samples: "
                                  f"{cn} replicas of the chain per GPU (independent Philox streams), in-kernel moments",
                      "acceptance_rate": float(out['stats'][:, 0].mean().item()) / (L - 1)},
           "roofline": {"bound": "fp64", "kernel": "k_mh<QUADRATIC,k=2> (a chain per thread)",
                        "achieved": per_gpu * 32.0 / 1e12, "peak": 2.0 * FP64_INSTR_PEAK / 1e12, "unit": "TFLOP/s",
                        "frac": per_gpu * 32.0 / (2.0 * FP64_INSTR_PEAK),
                        "note": "algorithmic 32 fp64 flop per log-posterior evaluation (SURVEY 8d, the C2 figure; one evaluation per "
                                "chain-step).  With 3 observations the step is not the likelihood: two Philox-4x32-10 blocks, a Box-Muller "
                                "pair (log, sqrt, sincos), log u and an exp on accept surround it, so the fraction of the fp64 roofline "
                                "is small by construction; the kernel's SASS holds 1205 fp64 instructions statically (slow paths of the "
                                "transcendentals included), no dynamic count was captured",
                        "traffic": None},
           "e2e": {"value": world * (L - 1) / one_s, "unit": "chain-steps/s", "api": "PE_object.doMetropolisHastings()", "chains": world,
                   "seconds_per_call": one_s, "post_burn_in_rows_returned": rows,
                   "h2d_bytes_per_step": 16, "d2h_bytes_per_step": int(rows * 8 * (PE.k + 2))}}
    if world == 1 and rank == 0:
        sec["cpu_baseline"] = c1_cpu_baseline()
    return sec



def c1_cpu_baseline(budget_s=4.0):
    """The oracle port of doMetropolisHastings' loop (IP.py:2626-2675: per step getLogP of the proposal AND of the current point,
    each through scipy) on one host core."""
    from oracle import peuqse_oracle as orc
    P = orc.build_problem(MU, SD, Y00, S00)
    sim = orc.sim_quadratic(X00)[0]
    rng = np.random.default_rng(1)
    sd = np.sqrt(C1["step"]) * np.asarray(SD, dtype=float)
    nst, done, t0 = 50, 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        orc.mh_chain_point(P, np.asarray(MU, dtype=float), rng.standard_normal((nst, 2)) * sd, rng.uniform(size=nst), sim)
        done += nst
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "chain-steps/s", "cores": 1, "kind": "port",
            "sample": f"{done} steps of one chain, point at a time (oracle.mh_chain_point), {dt:.1f} s"}


# ------------------------------------------------------------------------------------------------
# configs[3] ("C4"): doEnsembleSliceSampling, 2048 walkers, 12-parameter linear model, two responses of 4096 observations with
# full 4096 x 4096 covariances (the reference's route to a full covariance is >= 2 response dimensions); every
# stepping-out / shrink round is one batched log-posterior launch whose likelihood is the DMMA contraction.
# ------------------------------------------------------------------------------------------------
C4 = {"k": 12, "n": 4096, "walkers": 2048, "iters": 10, "seed": 7}


def fp64_probe(torch, mode):
    """The library's register-only probes (peuqse_bench_fp64): mode 0 DFMA instructions/s, mode 3 DMMA m8n8k4 flop/s."""
    from peuqse_b200 import _lib
    L = _lib.load()
    scratch = torch.zeros(64, dtype=torch.float64, device="cuda")
    best = 0.0
    for _ in range(2):
        r = ctypes.c_double()
        _lib.check(L.peuqse_bench_fp64(mode, ctypes.byref(r), _lib.ptr(scratch), _lib.current_stream()))
        best = max(best, r.value)
    return best


def c4_problem(PEUQSE, whitening):
    from oracle import peuqse_oracle as orc                            # only for the cos basis of the synthetic inputs
    k, n = C4["k"], C4["n"]
    x = np.linspace(0, 1, n)
    Phi = orc.cos_basis(2 * n, k)
    Sig = 0.1 ** 2 * np.exp(-(x[:, None] - x[None, :]) ** 2 / (2 * 0.02 ** 2)) + 0.05 ** 2 * np.eye(n)
    truth = np.random.default_rng(1).standard_normal(k)
    y = Phi @ truth + 0.1 * np.random.default_rng(2).standard_normal(2 * n)
    UI = PEUQSE.UserInput.fresh()
    UI.responses['responses_observed'] = [list(y[:n]), list(y[n:])]
    UI.responses['responses_observed_uncertainties'] = [Sig, Sig]
    UI.responses['responses_observed_max_covmat_size'] = n
    UI.responses['responses_observed_covmat_whitening'] = whitening
    UI.model['InputParameterPriorValues'] = [0.0] * k
    UI.model['InputParametersPriorValuesUncertainties'] = [2.0] * k
    UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.LinearModel(Phi)
    pes = UI.parameter_estimation_settings
    pes.update({'mcmc_length': C4["walkers"] * C4["iters"], 'mcmc_nwalkers': C4["walkers"], 'mcmc_burn_in': 0,
                'mcmc_random_seed': C4["seed"], 'mcmc_walkerInitialDistribution': 'sobol', 'mcmc_walkerInitialDistributionSpread': 0.25,
                'mcmc_threshold_filter_samples': False, 'mcmc_info_gain_returned': 'log_ratio', 'mcmc_exportLog': False})
    return PEUQSE.parameter_estimation(UI), Phi, Sig, y, truth


def c4_secondary(PEUQSE, torch, dist, rank, world):
    """Strong scaling: the ensemble is one; with N ranks every round's rows are shared out and the log-posteriors all-gathered
    (NCCL), the walkers stay those of the single-GPU run."""
    from peuqse_b200.mcmc import ess_launch
    from peuqse_b200 import _lib
    k, n, W, iters = C4["k"], C4["n"], C4["walkers"], C4["iters"]
    dmma_peak = fp64_probe(torch, 3)
    out = {}
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def tmax(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    for whitening in ("eigh", "cholesky"):
        PE, Phi, Sig, y, truth = c4_problem(PEUQSE, whitening)
        flops_eval = 2.0 * (2.0 * n * n if whitening == "eigh" else 1.0 * n * (n + 128))     # executed by the contraction, both responses
        start = 0.866 * 2.0 * 0.25 * np.random.default_rng(3).standard_normal((W, k))
        ess_launch(PE, start, 2, C4["seed"], store_chain=False)          # warm-up
        runs = []
        for _ in range(3):                                        # the rounds wait on the host: report the median of three runs
            barrier()
            a.record()
            run = ess_launch(PE, start, iters, C4["seed"], store_chain=False)
            b.record()
            torch.cuda.synchronize()
            runs.append(tmax(a.elapsed_time(b)) * 1e-3)
        s = sorted(runs)[1]
        info = run['info'].cpu().numpy()
        evals = float(info[0])
        # the dominant kernel alone: one batched log-posterior of 1024 rows (k_qf_tile<false>), CUDA events around 5 calls
        L = _lib.load()
        th = torch.from_numpy(truth + 0.05 * np.random.default_rng(4).standard_normal((1024, k))).cuda()
        lp = torch.empty(1024, dtype=torch.float64, device="cuda")
        call = lambda: _lib.check(L.peuqse_logp_batch(PE.device_problem.handle, _lib.ptr(th), 1024, _lib.ptr(lp), None, None, _lib.current_stream()))
        call(); call(); call()
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            call()
        b.record()
        torch.cuda.synchronize()
        ms1024 = a.elapsed_time(b) / 5
        # end to end: PE_object.doEnsembleSliceSampling() from host settings to host samples and statistics
        PE.doEnsembleSliceSampling()
        del PE.UserInput.last_MCMC_run_type, PE.mcmc_last_point_sampled
        barrier()
        t0 = time.perf_counter()
        PE.doEnsembleSliceSampling()
        torch.cuda.synchronize()
        te = tmax(time.perf_counter() - t0)
        blk = {"seconds": s, "seconds_of_each_run": runs, "logp_evals": evals, "logp_evals_per_s": evals / s, "walker_updates_per_s": W * iters / s,
               "tflops_executed": evals * flops_eval / s / 1e12, "mu": float(info[1]),
               "checksum_X": float(run['X'].sum().item()),
               "kernel": {"name": "k_qf_tile<false> inside peuqse_logp_batch, 1024 rows (simulate + residual + contraction + reduce)",
                          "ms": ms1024, "tflops": 1024 * flops_eval / (ms1024 * 1e-3) / 1e12,
                          "frac_of_dmma_peak": 1024 * flops_eval / (ms1024 * 1e-3) / dmma_peak},
               "e2e": {"seconds": te, "logp_evals_per_s": float(PE.ess_info['n_logp_evaluations']) / te,
                       "walker_updates_per_s": W * iters / te, "api": "PE_object.doEnsembleSliceSampling()",
                       "h2d_bytes": int(W * k * 8), "d2h_bytes": int(iters * W * (k + 1) * 8)}}
        out[whitening] = blk
        del PE
        torch.cuda.empty_cache()
    sec = {"metric": "ess_logp_evals_per_s", "value": out["eigh"]["logp_evals_per_s"], "unit": "evals/s", "scaling": "strong",
           "config": {"workload": f"doEnsembleSliceSampling: {W} walkers x {iters} ensemble iterations, {k}-parameter linear model, two "
                                  f"responses of {n} observations with full {n} x {n} covariances, differential move, round-synchronous "
                                  "device driver (csrc/ess_rounds.cu) over the DMMA contraction (csrc/fullcov.cu)",
                      "whitening": "eigh = scipy's _PSD arithmetic (the reference's, 2 n^2 flop per response and point); cholesky = "
                                   "triangular whitening (n^2 flop), same values to 1e-12 (tests/test_gpu_fullcov.py)"},
           "eigh": out["eigh"], "cholesky": out["cholesky"],
           "roofline": {"bound": "tensor", "kernel": "k_qf_tile<false,4> (fp64 DMMA m8n8k4)", "achieved": out["eigh"]["kernel"]["tflops"],
                        "peak": dmma_peak / 1e12, "unit": "TFLOP/s", "frac": out["eigh"]["kernel"]["frac_of_dmma_peak"],
                        "peak_source": "register-only mma.sync.m8n8k4.f64 probe of this run (peuqse_bench_fp64 mode 3); MEASURED_PEAKS.json "
                                       "has no fp64 figure", "traffic": None}}
    return sec


def c4_cpu_baseline(budget_evals=2):
    """The reference's per-point arithmetic at C4: scipy.stats.multivariate_normal.logpdf with the full covariance, which
    eigendecomposes the 4096 x 4096 matrix on EVERY call (IP.py:2960), two responses per evaluation; all BLAS threads."""
    from scipy.stats import multivariate_normal as mvn
    from oracle import peuqse_oracle as orc
    k, n = C4["k"], C4["n"]
    x = np.linspace(0, 1, n)
    Phi = orc.cos_basis(2 * n, k)
    Sig = 0.1 ** 2 * np.exp(-(x[:, None] - x[None, :]) ** 2 / (2 * 0.02 ** 2)) + 0.05 ** 2 * np.eye(n)
    truth = np.random.default_rng(1).standard_normal(k)
    y = Phi @ truth + 0.1 * np.random.default_rng(2).standard_normal(2 * n)
    t0 = time.perf_counter()
    for i in range(budget_evals):
        sim = Phi @ (truth + 0.01 * i)
        mvn.logpdf(x=y[:n], mean=sim[:n], cov=Sig) + mvn.logpdf(x=y[n:], mean=sim[n:], cov=Sig)
    dt = time.perf_counter() - t0
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return {"value": budget_evals / dt, "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": f"{budget_evals} log-likelihood evaluations (2 scipy logpdf calls each, LAPACK threads on all cores), {dt:.1f} s"}


# ------------------------------------------------------------------------------------------------
# configs[4] ("C5"): doMultiStart uniform sampling, 1e10 points of a 10-parameter linear model (64 observations), sharded over
# the ranks as contiguous index ranges (strong scaling: the total is fixed), MAP arg-max + weighted moments + histograms.
# ------------------------------------------------------------------------------------------------
C5 = {"k": 10, "n": 64, "total": 10_000_000_000, "seed": 0}


def c5_secondary(PEUQSE, torch, dist, rank, world, total):
    from oracle import peuqse_oracle as orc                            # cos basis of the synthetic inputs only
    from peuqse_b200.sweep import Sweep, make_spec
    k, n = C5["k"], C5["n"]
    Phi = orc.cos_basis(n, k)
    truth = np.random.default_rng(5).standard_normal(k) * 0.5
    y = Phi @ truth + 0.1 * np.random.default_rng(6).standard_normal(n)
    UI = PEUQSE.UserInput.fresh()
    UI.responses['responses_observed'] = list(y)
    UI.responses['responses_observed_uncertainties'] = list(np.full(n, 0.1))
    UI.model['InputParameterPriorValues'] = [0.0] * k
    UI.model['InputParametersPriorValuesUncertainties'] = [1.0] * k
    UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.LinearModel(Phi)
    pes = UI.parameter_estimation_settings
    pes.update({'multistart_initialPointsDistributionType': 'uniform', 'multistart_numStartPoints': int(total),
                'multistart_relativeInitialDistributionSpread': 0.866, 'multistart_searchType': 'getLogP', 'exportLog': False,
                'multistart_random_seed': C5["seed"], 'mcmc_length': 100})
    PE = PEUQSE.parameter_estimation(UI)
    spec = make_spec("uniform", k, np.zeros(k), spread=0.866, std_prior=np.ones(k), seed=C5["seed"])
    sw = Sweep(PE.device_problem, spec, int(total), store_logp=False)
    lr = float(np.log(10 ** -2.0))
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    warm = Sweep(PE.device_problem, spec, min(int(total), 1 << 24), store_logp=False, shard=(0, 1), local=True)
    warm.step(lr, True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    a.record()
    sw.pass1()
    b.record()
    sw.pass2(lr)
    sw.pass3(lr)
    c.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(c), a.elapsed_time(b)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms1 = float(t[0].item()), float(t[1].item())
    R = sw.unpack(sw.res.cpu().numpy(), 2.0)
    # end to end through the public call (one small untimed call first: module loading, allocator warm-up)
    PE.doMultiStart(numStartPoints=1 << 24)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    PE.doMultiStart()
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    # same MAP through the public call; below 2^31 points doMultiStart keeps logP and runs the bit-exact scalar kernel, whose
    # values agree with the no-store tensor-pipe kernel to 1e-13 (at the default 1e10 points both are the no-store kernel)
    assert PE.map_index == R.map_index and abs(PE.map_logP - R.map_logP) <= 1e-12 * abs(R.map_logP)
    fp64_instr = fp64_probe(torch, 0)
    flop_eval = 1.6e3                                                     # simulate 2kn + likelihood 4n + prior 4k + Philox
    per_gpu = (total / world) / (ms1 * 1e-3)
    sec = {"metric": "fp64_log_posterior_evals_per_s", "value": total / (ms * 1e-3), "unit": "evals/s", "scaling": "strong",
           "ms_per_sweep": ms, "pass1_ms": ms1, "pass1_evals_per_s": total / (ms1 * 1e-3),
           "config": {"workload": f"doMultiStart uniform getLogP: {int(total)} points, {k}-parameter linear model, {n} observations, "
                                  f"diagonal uncertainties, Philox points keyed by the global index, logP not stored (shortlist of "
                                  "near-maximal points instead of a second evaluation pass); MAP argmax + weighted moments + histograms",
                      "points_per_gpu": sw.count, "map_index": R.map_index, "map_logP": R.map_logP, "n_retained": R.n_retained},
           "roofline": {"bound": "fp64", "kernel": "k_sweep_pass1_lin_dmma<k=10,n=64> (simulation as mma.sync.m8n8k4.f64, which shares the fp64 pipe)",
                        "achieved": per_gpu * flop_eval / 1e12,
                        "peak": 2.0 * fp64_instr / 1e12, "unit": "TFLOP/s", "frac": per_gpu * flop_eval / (2.0 * fp64_instr),
                        "peak_source": "2 flop x DFMA issue rate measured in this run (peuqse_bench_fp64 mode 0)", "traffic": None},
           "e2e": {"value": total / float(te.item()), "unit": "evals/s", "api": "PE_object.doMultiStart()", "h2d_bytes_per_step": 3 * ctypes.sizeof(__import__('peuqse_b200')._lib.SweepSpec),
                   "d2h_bytes_per_step": sw.result_bytes()}}
    return sec, (Phi, y)


def c5_cpu_baseline(Phi, y, budget_s=8.0):
    """getLogP one point at a time through scipy logpdf (oracle port of the reference's loop) on one core."""
    from oracle import peuqse_oracle as orc
    from oracle import philox_np
    k, n = C5["k"], C5["n"]
    P = orc.build_problem([0.0] * k, [1.0] * k, list(y), list(np.full(n, 0.1)))
    pt = orc.sim_linear(Phi)[0]
    u = philox_np.sweep_uniforms(C5["seed"], 0, 200000, k)
    th = orc.uniform_points(u, 0.866, np.ones(k), np.zeros(k))
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s and done < len(th):
        for t in th[done:done + 500]:
            orc.log_p_point(P, t, pt)
        done += 500
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "evals/s", "cores": 1, "kind": "port",
            "sample": f"the first {done} points of the same Philox sequence, one point at a time (oracle.log_p_point), {dt:.1f} s"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, emit=print):
    import torch
    import peuqse_b200 as PEUQSE
    from peuqse_b200 import _lib
    from peuqse_b200.sweep import Sweep, make_spec

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world
    radii, total = grid_for(n_gpus)

    UI = PEUQSE.UserInput.fresh()
    UI.responses['responses_abscissa'] = X00
    UI.responses['responses_observed'] = Y00
    UI.responses['responses_observed_uncertainties'] = S00
    UI.model['InputParameterPriorValues'] = MU
    UI.model['InputParametersPriorValuesUncertainties'] = SD
    UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.QuadraticModel(X00)
    pes = UI.parameter_estimation_settings
    pes['multistart_initialPointsDistributionType'] = 'grid'
    pes['multistart_gridsearchSamplingInterval'] = INTERVAL
    pes['multistart_gridsearchSamplingRadii'] = radii
    pes['multistart_searchType'] = 'getLogP'
    pes['exportLog'] = False            # no log / CSV files inside the timed doMultiStart() calls (stated in the e2e block)
    pes['mcmc_length'] = 100
    PE = PEUQSE.parameter_estimation(UI)

    spec = make_spec("grid", 2, [float(MU[0]), float(MU[1])], interval=INTERVAL, radii=radii)
    pieces = int(os.environ.get('PEUQSE_BENCH_PIECES', '1'))
    sw = Sweep(PE.device_problem, spec, total, store_logp=True, nbins=10, pieces=pieces)
    log_rel = float(np.log(10 ** -2.0))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        sw.pass1_kernels()
        sw.rest_of_step(log_rel)
    torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    K = args.steps
    ev_a = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev_b = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = sw.launches
    barrier()
    start.record()
    for i in range(K):
        ev_a[i].record()
        sw.pass1_kernels()                          # dominant kernel, bracketed on its own stream
        ev_b[i].record()
        sw.rest_of_step(log_rel)                    # best-record exchange, pass 2, pass 3 and their exchanges: one graph launch
    end.record()
    barrier()
    launches = sw.launches - l0
    ms_total = start.elapsed_time(end)
    pass1_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(ev_a, ev_b)]))
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = total * K / (ms_total * 1e-3)

    # keep the GPU under the same load a little longer so that the clock sampler sees it
    if rank == 0:
        t_end = time.perf_counter() + 1.5
    probe_iters = 0
    while True:
        go = torch.tensor([1 if (rank != 0 or time.perf_counter() < t_end) else 0], device="cuda")
        if dist is not None:
            dist.broadcast(go, 0)
        if int(go.item()) == 0 or probe_iters > 100000:
            break
        for _ in range(20):
            sw.step(log_rel)
        torch.cuda.synchronize()
        probe_iters += 20
    clocks = sampler.stop() if sampler is not None else None

    # end-to-end through the public API: PE.doMultiStart() from host inputs to host results
    Ke = max(3, min(K, 50))
    PE.doMultiStart()
    PE.doMultiStart()                                   # the second identical call captures the launch sequence as a graph
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        PE.doMultiStart()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.barrier()
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    dt = float(te.item())
    e2e_value = total * Ke / dt
    h2d = 3 * ctypes.sizeof(_lib.SweepSpec) + 4 * 8      # sweep spec re-sent as kernel parameters by each pass + scalars
    d2h = sw.result_bytes()                              # packed (best, moments, histogram, edges, MAP point)
    assert PE.map_index >= 0 and np.isfinite(PE.map_logP)
    # multi-GPU correctness, outside the timed regions: rank 0 runs the WHOLE widened grid alone (no exchanges, logP not stored)
    # and the merged record of the sharded run must equal it -- MAP index and value, integer weight sum, retained points,
    # histogram counts
    multi_gpu_check = None
    if world > 1:
        ok = 1
        if rank == 0:
            alone = Sweep(PE.device_problem, spec, total, store_logp=False, nbins=10, shard=(0, 1), local=True)
            alone.step(log_rel, True)
            Ra = alone.unpack(alone.res.cpu().numpy(), 2.0)
            Rs = sw.unpack(sw.res.cpu().numpy(), 2.0)
            same = (Ra.map_index == Rs.map_index == PE.map_index and Ra.map_logP == Rs.map_logP and Ra.sum_weights == Rs.sum_weights
                    and Ra.n_retained == Rs.n_retained and np.array_equal(Ra.histogram_counts, Rs.histogram_counts)
                    and np.allclose(Ra.mu_AP_parameter_set, Rs.mu_AP_parameter_set, rtol=1e-12))
            ok = 1 if same else 0
            multi_gpu_check = "ok" if same else "MISMATCH: alone %r sharded %r" % ((Ra.map_index, Ra.map_logP, Ra.sum_weights),
                                                                                   (Rs.map_index, Rs.map_logP, Rs.sum_weights))
            del alone
        flag = torch.tensor([ok], device="cuda")
        dist.broadcast(flag, 0)
        if int(flag.item()) != 1:
            raise SystemExit("multi-GPU check failed: " + str(multi_gpu_check))

    secondary = sec_c4 = sec_c5 = None
    sec_c1 = None if args.no_c1 else c1_secondary(PEUQSE, torch, dist, rank, world)
    c5_inputs = None
    if not args.no_mh:
        secondary = mh_secondary(PEUQSE, torch, dist, rank, world, args.mh_steps)
    if not args.no_c5:
        sec_c5, c5_inputs = c5_secondary(PEUQSE, torch, dist, rank, world, args.c5_points)
    if not args.no_c4:
        sec_c4 = c4_secondary(PEUQSE, torch, dist, rank, world)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    pk, pk_kind = peaks()
    algo_bytes = 8.0 * sw.count                         # one fp64 logP written per evaluated point
    achieved = algo_bytes / (pass1_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_grid_pass1<QUADRATIC,k=2,n=3>", "achieved": achieved, "peak": pk["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "peak_source": pk_kind, "traffic": TRAFFIC_PASS1,
                "launch_ms": pass1_ms, "share_of_step": pass1_ms * K / ms_total,
                "fp64_note": "32 fp64 flop/eval algorithmic (SURVEY 8d): %.2f TFLOP/s" % (32.0 * sw.count / (pass1_ms * 1e-3) / 1e12)}
    cpu = cpu_baseline_single_core(radii) if n_gpus == 1 else None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(n_gpus), "points_per_gpu": sw.count,
                       "l2": "inputs larger than L2: the 800 MB logP buffer is rewritten and re-read every step",
                       "parallelism": (f"index-range sharding x{n_gpus}, 3 tiny exchanges per step ("
                                       + ("push-wait-merge kernels over NVLink peer memory" if getattr(sw, "peer", None) is not None
                                          else "NCCL") + ")") if n_gpus > 1 else "single GPU"},
            "roofline": roofline, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "PE_object.doMultiStart()", "file_export": "off (parameter_estimation_settings['exportLog'] = False)", "steps": Ke, "ms_per_step": 1e3 * dt / Ke},
            "gpu_launches": launches}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if multi_gpu_check is not None:
        line["multi_gpu_check"] = multi_gpu_check
    if sec_c1 is not None:
        line["secondary_c1"] = sec_c1
    if secondary is not None:
        line["secondary"] = secondary
    if sec_c4 is not None:
        if n_gpus == 1:
            sec_c4["cpu_baseline"] = c4_cpu_baseline()
        line["secondary_c4"] = sec_c4
    if sec_c5 is not None:
        if n_gpus == 1:
            sec_c5["cpu_baseline"] = c5_cpu_baseline(*c5_inputs)
        line["secondary_c5"] = sec_c5
    emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


# dram__bytes_read.sum + dram__bytes_write.sum of one pass-1 launch from the committed ncu --set full capture
# (profiles/); None until a capture exists.
TRAFFIC_PASS1 = 743.6e6   # 0.65 MB read + 742.95 MB written (profiles/r1_ncu_full_pass1_pass2_final_raw.csv); algorithmic 800.2 MB


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mh-steps", type=int, default=20000, help="steps per chain of the secondary (C3 Metropolis-Hastings) measurement")
    ap.add_argument("--no-mh", action="store_true", help="skip the secondary (C3) measurement")
    ap.add_argument("--no-c1", action="store_true", help="skip the C1 (Example00 doMetropolisHastings, 10 000 samples) measurement")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 (ensemble slice sampling, full covariance) measurement")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 (uniform multistart) measurement")
    ap.add_argument("--c5-points", type=float, default=C5["total"], help="points of the C5 sweep (default 1e10)")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout: libraries that print banners there (NCCL prints its version on the first
    # communicator) are sent to stderr by pointing fd 1 at fd 2 for the duration of the run; the line goes to the real stdout.
    sys.stdout.flush()
    real_out = os.dup(1)
    os.dup2(2, 1)
    lines = []
    try:
        rc = run_reference_arm(args, lines.append) if args.impl == "reference" else run_ours(args, lines.append)
    finally:
        sys.stdout.flush()
        os.dup2(real_out, 1)
        os.close(real_out)
    for ln in lines:
        print(ln, flush=True)
    return rc


if __name__ == "__main__":
    sys.exit(main())
