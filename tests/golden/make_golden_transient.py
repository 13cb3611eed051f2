"""tests/golden/transient.npz: getLogP / getLogPrior / getLogLikelihood of the UNMODIFIED reference (/root/reference) with
responses['data_overcategory'] = 'transient_kinetics' (transform_responses, IP.py:609-658; littleEulerGivenArray 3501-3513;
littleEulerUncertaintyPropagation 3516-3525): rate-type responses are integrated over the abscissa (explicit Euler, starting at 0)
before they are compared, the observed uncertainties are propagated alongside.  Case 'P_r': one product response given as a rate.
Case 'mixed': an intermediate given as a rate (integrated twice), a reactant concentration (untouched) and a product rate, each on
its own unevenly spaced abscissa.  Run in the build container only:  python tests/golden/make_golden_transient.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG                      # noqa: E402  (import shim, fresh_userinput, make_pe; chdirs to a scratch directory)

rng = np.random.default_rng(44)
n = 9
t = [np.cumsum(0.05 + 0.1 * rng.random(n)) for _ in range(3)]


def rates(th):
    """three responses on their own abscissas, all smooth in the two parameters"""
    return [th[0] * np.exp(-th[1] * t[0]), 1.0 - 0.5 * th[0] * t[1] / (1.0 + th[1] * t[1]), th[0] * th[1] * np.exp(-th[1] * t[2])]


truth = np.array([2.0, 1.5])
out = {'tr_t': np.array(t), 'tr_truth': truth}
th = truth + 0.2 * rng.standard_normal((30, 2))
out['tr_thetas'] = th
for tag, resp, types, dtypes in (('P_r', [0], ['P'], ['r']), ('mixed', [0, 1, 2], ['I', 'R', 'P'], ['r', 'c', 'r'])):
    y = [rates(truth)[r] * (1.0 + 0.05 * rng.standard_normal(n)) for r in resp]
    sig = [0.05 * np.abs(v) + 0.01 for v in y]
    UI = MG.fresh_userinput()
    UI.responses['responses_abscissa'] = [list(t[r]) for r in resp] if len(resp) > 1 else list(t[resp[0]])
    UI.responses['responses_observed'] = [list(v) for v in y] if len(resp) > 1 else list(y[0])
    UI.responses['responses_observed_uncertainties'] = [list(v) for v in sig] if len(resp) > 1 else list(sig[0])
    UI.responses['data_overcategory'] = 'transient_kinetics'
    UI.responses['response_types'] = types
    UI.responses['response_data_types'] = dtypes
    UI.model['InputParameterPriorValues'] = [2.0, 1.5]
    UI.model['InputParametersPriorValuesUncertainties'] = [0.5, 0.4]
    UI.model['simulateByInputParametersOnlyFunction'] = (lambda p, resp=resp: [rates(p)[r] for r in resp]) if len(resp) > 1 else (lambda p, r=resp[0]: rates(p)[r])
    PE = MG.make_pe(UI)
    out['tr_%s_y' % tag] = np.array(y)
    out['tr_%s_sig' % tag] = np.array(sig)
    out['tr_%s_vals' % tag] = MG.eval_points(PE, th)
    out['tr_%s_obs_transformed' % tag] = np.array([np.asarray(v, dtype=float) for v in PE.UserInput.responses_observed_transformed])
    out['tr_%s_cov_transformed' % tag] = np.array([np.asarray(v, dtype=float) for v in PE.observed_responses_covmat_transformed])
np.savez(os.path.join(HERE, 'transient.npz'), **out)
print({a: b.shape for a, b in out.items()})
print(out['tr_P_r_vals'][:2], out['tr_mixed_vals'][:2])
