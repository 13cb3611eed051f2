"""tests/golden/reduced.npz: getLogP / getLogPrior / getLogLikelihood of the UNMODIFIED reference (/root/reference) with
model['reducedParameterSpace'] set (IP.py:406-408, 519-561, 574-606): only the listed parameters vary, the others stay at
model['InputParameterInitialGuess'].  Example00 with the second parameter alone, the first alone, and both in swapped order.
Run in the build container only:  python tests/golden/make_golden_reduced.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG                      # noqa: E402  (import shim, fresh_userinput, make_pe; chdirs to a scratch directory)

GUESS = [210.0, 480.0]
rng = np.random.default_rng(21)
out = {'reduced_guess': np.array(GUESS)}
for tag, idx in (('b', [1]), ('a', [0]), ('ba', [1, 0])):
    UI = MG.ui_example00()
    UI.model['InputParameterInitialGuess'] = list(GUESS)
    UI.model['reducedParameterSpace'] = idx
    PE = MG.make_pe(UI)
    mu = np.array([200.0, 500.0])[idx]
    sd = np.array([100.0, 200.0])[idx]
    th = MG.theta_cloud(rng, mu, sd, 40)
    out['reduced_%s_thetas' % tag] = th
    out['reduced_%s_vals' % tag] = MG.eval_points(PE, th)      # columns: logP, logPrior, logLikelihood
np.savez(os.path.join(HERE, 'reduced.npz'), **out)
print({k: v.shape for k, v in out.items()}, out['reduced_b_vals'][:2], out['reduced_ba_vals'][:2])
