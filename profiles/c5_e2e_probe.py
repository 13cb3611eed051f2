import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE
from peuqse_b200 import sweep as SW
from oracle import peuqse_oracle as orc
N = 4_000_000_000
k, n = 10, 64
Phi = orc.cos_basis(n, k)
truth = np.random.default_rng(5).standard_normal(k) * 0.5
y = Phi @ truth + 0.1 * np.random.default_rng(6).standard_normal(n)
UI = PEUQSE.UserInput.fresh()
UI.responses['responses_observed'] = list(y)
UI.responses['responses_observed_uncertainties'] = list(np.full(n, 0.1))
UI.model['InputParameterPriorValues'] = [0.0] * k
UI.model['InputParametersPriorValuesUncertainties'] = [1.0] * k
UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.LinearModel(Phi)
UI.parameter_estimation_settings.update({'multistart_initialPointsDistributionType': 'uniform', 'multistart_numStartPoints': N,
    'multistart_relativeInitialDistributionSpread': 0.866, 'multistart_searchType': 'getLogP', 'exportLog': False, 'multistart_random_seed': 0, 'mcmc_length': 100})
PE = PEUQSE.parameter_estimation(UI)
orig_step, orig_cap = SW.Sweep.step, SW.Sweep.capture
def step(self, *a, **k):
    torch.cuda.synchronize(); t=time.perf_counter(); r=orig_step(self,*a,**k); torch.cuda.synchronize(); print('step', time.perf_counter()-t); return r
def cap(self, fn):
    torch.cuda.synchronize(); t=time.perf_counter(); r=orig_cap(self,fn); torch.cuda.synchronize(); print('capture', time.perf_counter()-t, r is not None, getattr(self, '_graph_error', None)); return r
SW.Sweep.step, SW.Sweep.capture = step, cap
for i in range(3):
    torch.cuda.synchronize(); t=time.perf_counter(); PE.doMultiStart(); torch.cuda.synchronize(); print('doMultiStart', time.perf_counter()-t, N/(time.perf_counter()-t))
