"""One C3-shaped launch of the warp-shape Metropolis-Hastings kernel (4096 chains x 200 steps, n = 1024), for ncu."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE
from peuqse_b200.mcmc import mh_launch, step_cholesky
from oracle import peuqse_oracle as orc

n = 1024
kBT, c1, c2 = orc.tpd2site_constants(n)
truth = np.array([41.5, 45, 13, 13, 0.1, 0.1])
sim_t = orc.sim_tpd2site(kBT, c1, c2)[0](truth)
sig = 0.02 * np.max(sim_t)
y = sim_t + sig * np.random.default_rng(0).standard_normal(n)
UI = PEUQSE.UserInput.fresh()
UI.responses['responses_observed_max_covmat_size'] = 5000
UI.responses['responses_observed'] = list(y); UI.responses['responses_observed_uncertainties'] = list(np.full(n, sig))
UI.model['InputParameterPriorValues'] = [41.5, 41.5, 13.0, 13.0, 0.1, 0.1]
UI.model['InputParametersPriorValuesUncertainties'] = [20, 20, 13, 13, 0.1, 0.1]
UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.Tpd2SiteModel.linear_ramp(n)
PE = PEUQSE.parameter_estimation(UI)
C, Lm = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 200
th0 = np.tile(truth, (C, 1))
# PEUQSE_MH_SHAPES="1 7": one launch per listed CTA shape (see launch_mh) after a warm-up launch
for shape in os.environ.get('PEUQSE_MH_SHAPES', '0 0').split():
    os.environ['PEUQSE_MH_SHAPE'] = shape
    st = mh_launch(PE, th0, Lm, 20, 0, step_cholesky(PE.Q_covmat, 1e-5), store_samples=False)
torch.cuda.synchronize()
print("accept", float(st['stats'][:, 0].mean()) / (Lm - 1))
