"""Multistart of Nelder-Mead optimisations on Example00: all start points in lock step (peuqse_b200/optimize.py, one batched
log-posterior launch per simplex stage) against one scipy run per start point (what the reference does, with its objective
also evaluated by the kernel), and against the oracle port of the reference's getNegLogP under scipy on one host core."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE

X00, Y00, S00 = [600, 1100, 1400], [360500, 580500, 1620500], [200000, 300000, 200000]


def make(n, batched):
    UI = PEUQSE.UserInput.fresh()
    UI.responses['responses_observed'] = Y00
    UI.responses['responses_observed_uncertainties'] = S00
    UI.model['InputParameterPriorValues'] = [200, 500]
    UI.model['InputParametersPriorValuesUncertainties'] = [100, 200]
    UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.QuadraticModel(X00)
    UI.parameter_estimation_settings.update({'multistart_initialPointsDistributionType': 'uniform', 'multistart_numStartPoints': n,
                                             'multistart_searchType': 'doOptimizeNegLogP', 'multistart_batched_optimizer': batched,
                                             'exportLog': False,
                                             'multistart_passThroughArgs': {'method': 'Nelder-Mead', 'printOptimum': False, 'verbose': False}})
    return PEUQSE.parameter_estimation(UI)


out = {}
for n in (64, 4096, 65536):
    PE = make(n, True)
    PE.doMultiStart()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    PE.doMultiStart()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    info = PE.multistart_optimizer_info
    out[f'lockstep_{n}'] = dict(start_points=n, wall_ms=dt * 1e3, launches=info['launches'], evaluations=int(info['nfev'].sum()),
                                starts_per_s=n / dt, evals_per_s=int(info['nfev'].sum()) / dt, map_logP=float(PE.map_logP))
PE = make(64, False)
t0 = time.perf_counter(); PE.doMultiStart(); dt = time.perf_counter() - t0
out['scipy_per_start_64'] = dict(start_points=64, wall_ms=dt * 1e3, starts_per_s=64 / dt, map_logP=float(PE.map_logP))
from oracle import peuqse_oracle as orc
import scipy.optimize
P = orc.build_problem([200, 500], [100, 200], Y00, S00)
pt, _ = orc.sim_quadratic(X00)
pts = PE.listOfPermutations[:16]
t0 = time.perf_counter()
for p in pts:
    scipy.optimize.minimize(lambda th: -1 * orc.log_p_point(P, th, pt), p, method='Nelder-Mead')
dt = time.perf_counter() - t0
out['cpu'] = dict(kind='port', cores=1, sample='16 start points, scipy Nelder-Mead over the oracle port of getNegLogP', starts_per_s=16 / dt)
print(json.dumps(out))
