"""Generates tests/golden/geweke.npz: geweke_diagnostic and the Geweke half of calculateAndPlotConvergenceDiagnostics
(PEUQSE/InverseProblem.py:3902-4062) of the UNMODIFIED reference on seeded AR(1)-like chains.

statsmodels (yule_walker, imported inside geweke_diagnostic at IP.py:4044) and zeus (AutoCorrTime, IP.py:3921) are third-party
packages that are not installed here.  The reference's own code -- window and slice indices, z-score formula, the averaging
over parameters / chains and the outlier percentages -- is what this fixture pins; for that, a module named
statsmodels.regression.linear_model that carries the oracle's restatement of yule_walker is placed in sys.modules, and
zeus.autocorr.AutoCorrTime is stubbed (its outputs are not stored).

Run in the build container only:  python tests/golden/make_golden_geweke.py
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as G  # noqa: E402
from oracle import peuqse_oracle as orc  # noqa: E402

sm = types.ModuleType('statsmodels'); smr = types.ModuleType('statsmodels.regression'); sml = types.ModuleType('statsmodels.regression.linear_model')
sml.yule_walker = lambda x, order=1, **kw: orc.yule_walker(x, order)
sys.modules.update({'statsmodels': sm, 'statsmodels.regression': smr, 'statsmodels.regression.linear_model': sml})
zs = types.ModuleType('zeus'); zsa = types.ModuleType('zeus.autocorr')
zsa.AutoCorrTime = lambda x: np.ones(x.shape[2])
sys.modules.update({'zeus': zs, 'zeus.autocorr': zsa})

import PEUQSE.InverseProblem as IP  # noqa: E402


def chains(R, C, k, seed):
    rng = np.random.default_rng(seed)
    x = np.empty((R, C, k))
    phi = np.array([0.5, 0.9, 0.97, 0.0][:k])
    x[0] = rng.standard_normal((C, k))
    e = rng.standard_normal((R, C, k))
    for t in range(1, R):
        x[t] = phi * x[t - 1] + e[t]
    return x * np.array([100., 3., 0.01, 1.][:k]) + np.array([500., -20., 0.1, 0.][:k])


def main():
    out = {}
    x = chains(1500, 3, 3, 7)
    out['chains_a'] = x
    out['single_chain_z'] = IP.geweke_diagnostic(x[:, 0, 0])
    out['single_chain_z_short'] = IP.geweke_diagnostic(x[:333, 1, 2])
    names = {'a': 'a', 'b': 'b', 'c': 'c'}
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        res = IP.calculateAndPlotConvergenceDiagnostics(x, names, plot_settings={}, graphs_directory='./', createPlots=False)
    assert res[3] is not None, buf.getvalue()
    out['a_window_indices'] = np.array(res[3])
    out['a_final_combined'] = np.array(res[4])
    out['a_percent_outlier'] = np.array(res[5])
    path = os.path.join(G.HERE, "geweke.npz")
    np.savez_compressed(path, **out)
    print("geweke ->", path, {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
    os.chdir(G.REPO)
