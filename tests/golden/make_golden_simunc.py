"""tests/golden/simunc.npz: getLogP of the UNMODIFIED reference (/root/reference) when model['responses_simulation_uncertainties'] is a
FUNCTION of the parameters (IP.py:563-572, 2883-2896, 2925-2927: the simulation's standard deviations are squared and added to the
observed variances point by point, so the likelihood's covariance changes with every evaluation).  Two cases: Example00 (one
response of 3 values, one multivariate logpdf) and the same response with responses_observed_max_covmat_size = 2 (one scalar logpdf
per value, IP.py:2961-2975).  Run in the build container only:  python tests/golden/make_golden_simunc.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG                      # noqa: E402  (import shim, fresh_userinput, make_pe; chdirs to a scratch directory)

X = np.asarray(MG.X00, dtype=float)


def sim_sd(theta):
    """standard deviations that depend on the parameters: 3 % of the simulated value plus a floor that grows with theta[1]"""
    sim = (X - theta[0]) ** 2 + theta[1]
    return 0.03 * np.abs(sim) + 50.0 * (1.0 + np.abs(theta[1]) / 500.0)


rng = np.random.default_rng(12)
thetas = MG.theta_cloud(rng, np.array([200.0, 500.0]), np.array([100.0, 200.0]), 64)
out = {'simunc_thetas': thetas}
for tag, maxcov in (('mvn', 100), ('per_value', 2)):
    UI = MG.ui_example00(responses_simulation_uncertainties=sim_sd)
    UI.responses['responses_observed_max_covmat_size'] = maxcov
    PE = MG.make_pe(UI)
    out['simunc_' + tag] = MG.eval_points(PE, thetas)          # columns: logP, logPrior, logLikelihood
np.savez(os.path.join(HERE, 'simunc.npz'), **out)
print({k: v.shape for k, v in out.items()}, out['simunc_mvn'][:3], out['simunc_per_value'][:3])
