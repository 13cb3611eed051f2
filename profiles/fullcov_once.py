"""One batched log-posterior of the C4 problem at a large and at a small batch (for ncu: k_quadform_dmma<false>, then
k_quadform_dmma<true> + k_quadform_split_finish).  usage: python profiles/fullcov_once.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE
from oracle import peuqse_oracle as orc
k, n = 12, 4096
x = np.linspace(0, 1, n)
Phi = orc.cos_basis(2 * n, k)
Sig = 0.1 ** 2 * np.exp(-(x[:, None] - x[None, :]) ** 2 / (2 * 0.02 ** 2)) + 0.05 ** 2 * np.eye(n)
truth = np.random.default_rng(1).standard_normal(k)
y = Phi @ truth + 0.1 * np.random.default_rng(2).standard_normal(2 * n)
UI = PEUQSE.UserInput.fresh()
UI.responses['responses_observed'] = [list(y[:n]), list(y[n:])]
UI.responses['responses_observed_uncertainties'] = [Sig, Sig]
UI.model['InputParameterPriorValues'] = [0.0] * k
UI.model['InputParametersPriorValuesUncertainties'] = [2.0] * k
UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.LinearModel(Phi)
UI.responses['responses_observed_max_covmat_size'] = 4096
PE = PEUQSE.parameter_estimation(UI)
for B in (1024, 64):
    th = torch.from_numpy(truth + 0.05 * np.random.default_rng(3).standard_normal((B, k))).cuda()
    PE.getLogP_batch(th)
torch.cuda.synchronize()
print("done")
