"""tests/golden/reduced_responses.npz: getLogP / getLogPrior / getLogLikelihood of the UNMODIFIED reference (/root/reference) with
responses['reducedResponseSpace'] set (IP.py:410-413, reduceResponseSpace 470-512): only the listed response dimensions enter the
likelihood.  Three responses of a linear model, keeping [0, 2] and keeping [1].
Run in the build container only:  python tests/golden/make_golden_reduced_responses.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG                      # noqa: E402  (import shim, fresh_userinput, make_pe; chdirs to a scratch directory)
from oracle import peuqse_oracle as orc       # noqa: E402

rng = np.random.default_rng(33)
k, n = 3, 6
Phi = rng.standard_normal((3 * n, k))
point, _ = orc.sim_linear(Phi)
truth = np.array([0.5, -1.0, 2.0])
sig = 0.1 + 0.4 * rng.random(3 * n)
y = point(truth) + sig * rng.standard_normal(3 * n)
th = truth + 0.3 * rng.standard_normal((30, k))
out = {'rr_Phi': Phi, 'rr_y': y, 'rr_sig': sig, 'rr_thetas': th}
for tag, keep in (('02', [0, 2]), ('1', [1])):
    UI = MG.fresh_userinput()
    UI.responses['responses_abscissa'] = [list(range(n))] * 3
    UI.responses['responses_observed'] = [list(y[r * n:(r + 1) * n]) for r in range(3)]
    UI.responses['responses_observed_uncertainties'] = [list(sig[r * n:(r + 1) * n]) for r in range(3)]
    UI.responses['reducedResponseSpace'] = keep
    UI.model['InputParameterPriorValues'] = [0.0, 0.0, 0.0]
    UI.model['InputParametersPriorValuesUncertainties'] = [2.0, 2.0, 3.0]
    UI.model['simulateByInputParametersOnlyFunction'] = lambda t: [point(t)[r * n:(r + 1) * n] for r in range(3)]
    PE = MG.make_pe(UI)
    out['rr_%s_vals' % tag] = MG.eval_points(PE, th)
np.savez(os.path.join(HERE, 'reduced_responses.npz'), **out)
print({a: b.shape for a, b in out.items()}, out['rr_02_vals'][:2], out['rr_1_vals'][:2])
