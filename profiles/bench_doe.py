"""Design-of-experiments throughput: the Example13 Langmuir problem, every (parameter modulation,
condition) pair one scenario of a device batch.  Reports
  * unit_test: the reference's own unit-test size (9 modulations x 25 conditions, 200-step chains) end to end through
    PE.doeParameterModulationPermutationsScanner(), wall clock, host results;
  * large: 9 modulations x (2r+1)^2 conditions, LEN-step chains: device-timed MH launch and statistics launch, chain-steps/s;
  * cpu: the oracle's point-at-a-time Metropolis-Hastings loop (scipy logpdf per evaluation, as the reference does) on one
    scenario, chain-steps/s on one host core.
usage: python profiles/bench_doe.py [radius=20] [length=10000]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE
from peuqse_b200 import _lib
from peuqse_b200.doe import ScenarioBatch, conditions_grid, _grid_points
from peuqse_b200.mcmc import step_cholesky

RAD = int(sys.argv[1]) if len(sys.argv) > 1 else 20
LEN = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
DH, DS = -0.687 - (-0.858), -1.50e-3 - (-1.45E-3)
DHU, DSU = (5.00e-2 ** 2 + 1.25e-1 ** 2) ** 0.5, (2.07e-4 ** 2 + 2.07e-4 ** 2) ** 0.5
SIG = float(np.log10(1.5))


def make(length, burn, radius, interval):
    UI = PEUQSE.UserInput.fresh()
    UI.responses['responses_observed'] = [0.0, 0.0, 0.0]
    UI.responses['responses_observed_uncertainties'] = [SIG] * 3
    UI.model['InputParameterPriorValues'] = [DH, DS]
    UI.model['InputParametersPriorValuesUncertainties'] = [DHU, DSU]
    UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.LangmuirLogModel(400, 0.1)
    UI.parameter_estimation_settings.update({'mcmc_random_seed': 0, 'mcmc_burn_in': burn, 'mcmc_length': length,
                                             'mcmc_relative_step_length': 0.05, 'mcmc_info_gain_returned': 'log_ratio',
                                             'mcmc_threshold_filter_samples': True})
    UI.doe_settings.update({'info_gains_matrices_array_format': 'xyz', 'independent_variable_grid_center': [500, 0.5],
                            'independent_variable_grid_interval_size': interval, 'independent_variable_grid_num_intervals': [radius, radius],
                            'parameter_modulation_grid_interval_size': [1, 1], 'parameter_modulation_grid_num_intervals': [1, 1]})
    return PEUQSE.parameter_estimation(UI)


out = {}
PE = make(200, 100, 2, [100, 0.1])
PE.doeParameterModulationPermutationsScanner()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    A = PE.doeParameterModulationPermutationsScanner()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
out['unit_test'] = dict(scenarios=225, steps=200, wall_ms=dt * 1e3, chain_steps_per_s=225 * 200 / dt,
                        best_condition=[float(v) for v in A[4][np.argmax(A[4][:, 2])]])

PE = make(LEN, LEN // 10, RAD, [200.0 / RAD, 0.4 / RAD])
conds = conditions_grid(PE, PE.UserInput.doe_settings)
mods = _grid_points(PE, [DH, DS], [DHU, DSU], [1, 1])
S = len(conds) * len(mods)
batch = ScenarioBatch(PE, np.tile(conds, (len(mods), 1)), np.repeat(mods, len(conds), axis=0))
chol = step_cholesky(PE.Q_covmat, 0.05)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
run = batch.run_mh([DH, DS], LEN, LEN // 10, 0, chol)           # warm-up
st = batch.statistics(run, LEN // 10, _lib.FILTER_AUTO, 2.0)
del run
torch.cuda.synchronize()
ev[0].record()
run = batch.run_mh([DH, DS], LEN, LEN // 10, 0, chol)
ev[1].record()
lay = _lib.scen_stats_layout(2)
d_st = torch.empty((S, lay['DOUBLES']), dtype=torch.float64, device='cuda')
_lib.check(_lib.load().peuqse_scenarios_stats(2, _lib.FILTER_AUTO, 2.0, S, 1, LEN, LEN // 10, _lib.ptr(run['samples']), _lib.ptr(run['logp']),
                                              _lib.ptr(run['logprior']), _lib.ptr(d_st), _lib.current_stream()))
ev[2].record()
st = d_st.cpu().numpy()                                         # the batch's whole read-back: S x (8 + 3k) doubles
ev[3].record()
torch.cuda.synchronize()
mh_ms, st_ms, d2h_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
batch._d_stats = d_st
kl = batch.kl_information_gain(run, LEN // 10, PE.samples_of_prior)                   # warm-up
ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ka.record()
kl = batch.kl_information_gain(run, LEN // 10, PE.samples_of_prior)
kb.record(); torch.cuda.synchronize()
rows = S * (LEN - LEN // 10)
out['large'] = dict(scenarios=S, steps=LEN, mh_ms=mh_ms, stats_ms=st_ms, readback_ms=d2h_ms, readback_bytes=S * lay['DOUBLES'] * 8, chain_steps_per_s=S * LEN / (mh_ms * 1e-3),
                    sample_bytes_written=S * LEN * 5 * 8, mh_write_GBps=S * LEN * 5 * 8 / (mh_ms * 1e-3) / 1e9,
                    stats_rows_per_s=rows / (st_ms * 1e-3),
                    stats_algorithmic_GBps=rows * (8 + 24 + 32) / (st_ms * 1e-3) / 1e9,   # logP | logP + theta | logP + theta + logPrior
                    accept_rate=float(run['stats'][:, 0].mean().item()) / (LEN - 1),
                    info_gain_range=[float(st[:, 6].min()), float(st[:, 6].max())],
                    e2e_chain_steps_per_s=S * LEN / ((mh_ms + st_ms + d2h_ms) * 1e-3),
                    kl_ms=ka.elapsed_time(kb), kl_range=[float(kl[:, 0].min()), float(kl[:, 0].max())])

from oracle import peuqse_oracle as orc
from oracle import philox_np
n = 400
z, us = philox_np.mh_draws(0, 0, np.arange(1, n), 2)
pt, _ = orc.sim_langmuir_log(500, 0.5)
P = orc.build_problem([DH, DS], [DHU, DSU], pt(np.array([DH + DHU, DS + DSU])), [SIG] * 3)
t0 = time.perf_counter()
orc.mh_chain_point(P, np.array([DH, DS]), z * np.sqrt(0.05) * np.array([DHU, DSU]), us, pt)
dt = time.perf_counter() - t0
out['cpu'] = dict(kind='port', cores=1, sample=f'one scenario, {n} steps of the point-at-a-time loop', chain_steps_per_s=(n - 1) / dt)
print(json.dumps(out))
