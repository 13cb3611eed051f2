"""One no-store uniform sweep of the C5 problem (k = 10 linear model, 64 observations) over 2^28 points, for ncu.
usage: python profiles/c5_once.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE
from peuqse_b200.sweep import Sweep, make_spec
from oracle import peuqse_oracle as orc
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
PE = PEUQSE.parameter_estimation(UI)
spec = make_spec("uniform", k, np.zeros(k), spread=0.866, std_prior=np.ones(k), seed=0)
sw = Sweep(PE.device_problem, spec, 1 << 28, store_logp=False)
for _ in range(2):
    sw.pass1()
torch.cuda.synchronize()
print("best", sw.best.cpu())
