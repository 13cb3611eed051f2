"""One C1-shaped launch: a single Example00 Metropolis-Hastings chain of 10 000 steps (for ncu / latency work)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE
from peuqse_b200.mcmc import mh_launch, step_cholesky
UI = PEUQSE.UserInput.fresh()
UI.responses['responses_observed'] = [360500, 580500, 1620500]
UI.responses['responses_observed_uncertainties'] = [200000, 300000, 200000]
UI.model['InputParameterPriorValues'] = [200, 500]
UI.model['InputParametersPriorValuesUncertainties'] = [100, 200]
UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.QuadraticModel([600, 1100, 1400])
PE = PEUQSE.parameter_estimation(UI)
th0 = np.array([[200.0, 500.0]])
for _ in range(2):
    st = mh_launch(PE, th0, 10000, 1000, 0, step_cholesky(PE.Q_covmat, 0.1), store_samples=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
st = mh_launch(PE, th0, 10000, 1000, 0, step_cholesky(PE.Q_covmat, 0.1), store_samples=True)
b.record(); torch.cuda.synchronize()
print("ms", a.elapsed_time(b), "accept", float(st['stats'][0, 0]) / 9999)
