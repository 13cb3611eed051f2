"""Generates tests/golden/doe.npz by running the UNMODIFIED reference's design-of-experiments functions
(doeGetInfoGainMatrix, doeParameterModulationPermutationsScanner; PEUQSE/InverseProblem.py:1363-1572) on the
configuration of Examples/Example13doeFunctionExample/runfile_for_unit_test_xyz.py / _meshgrid.py, with the
Metropolis-Hastings draws injected (the same Philox-derived increments and uniforms for every scenario: the reference
reseeds numpy in each __init__, IP.py:191-193, so its scenarios share one draw sequence as well).

Run in the build container only:  python tests/golden/make_golden_doe.py
"""
import contextlib
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as G  # noqa: E402  (sets up the matplotlib shim, the scratch cwd and the reference import)

from oracle import peuqse_oracle as orc  # noqa: E402
from oracle import philox_np  # noqa: E402

L_MCMC, BURN, STEP = 200, 100, 0.05
DH = -0.687 - (-0.858)
DS = -1.50e-3 - (-1.45E-3)
DH_U = (5.00e-2 ** 2 + 1.25e-1 ** 2) ** 0.5
DS_U = (2.07e-4 ** 2 + 2.07e-4 ** 2) ** 0.5
SIGMA = [np.log10(1.5)] * 3


class LangmuirModule:
    """Stand-in for the example's simulation module: conditions are pushed in by populate() (the reference's
    populateIndependentVariablesFunction plugin) and read by simulate()."""

    def __init__(self):
        self.cond = [400, 0.1]

    def populate(self, values):
        self.cond = list(values)

    def simulate(self, sample):
        point, _ = orc.sim_langmuir_log(self.cond[0], self.cond[1])
        return point(np.array(sample))


def ui_doe(fmt, filter_samples=False):
    UI = G.fresh_userinput()
    mod = LangmuirModule()
    UI.responses['responses_abscissa'] = [398.15 - 25, 398.15, 398.15 + 25]
    UI.responses['responses_observed'] = [0.00, 0.00, 0.00]
    UI.responses['responses_observed_uncertainties'] = SIGMA
    UI.responses['independent_variables_values'] = [400, 0.1]
    UI.responses['independent_variables_names'] = ['T(K)', 'P_A(bar)']
    UI.model['InputParameterPriorValues'] = [DH, DS]
    UI.model['InputParametersPriorValuesUncertainties'] = [DH_U, DS_U]
    UI.model['populateIndependentVariablesFunction'] = mod.populate
    UI.model['simulateByInputParametersOnlyFunction'] = mod.simulate
    pes = UI.parameter_estimation_settings
    pes['scaling_uncertainties_type'] = "std"
    pes['exportLog'] = False
    pes['verbose'] = False
    pes['mcmc_exportLog'] = False
    pes['mcmc_mode'] = 'unbiased'
    pes['mcmc_random_seed'] = 0
    pes['mcmc_burn_in'] = BURN
    pes['mcmc_length'] = L_MCMC
    pes['mcmc_relative_step_length'] = STEP
    pes['mcmc_modulate_accept_probability'] = 0
    pes['mcmc_info_gain_returned'] = 'log_ratio'
    pes['mcmc_threshold_filter_samples'] = filter_samples
    pes['convergence_diagnostics'] = False
    UI.doe_settings['info_gains_matrices_array_format'] = fmt
    UI.doe_settings['independent_variable_grid_center'] = [500, 0.5]
    UI.doe_settings['independent_variable_grid_interval_size'] = [100, 0.1]
    UI.doe_settings['independent_variable_grid_num_intervals'] = [2, 2]
    UI.doe_settings['parameter_modulation_grid_interval_size'] = [1, 1]
    UI.doe_settings['parameter_modulation_grid_num_intervals'] = [1, 1]
    return UI


def injected(fn):
    z, us = philox_np.mh_draws(0, 0, np.arange(1, L_MCMC), 2)
    deltas = z * (np.sqrt(STEP) * np.array([DH_U, DS_U]))
    it = {'i': 0}
    orig_mvn, orig_uni = np.random.multivariate_normal, np.random.uniform

    def fake_mvn(mean, cov, *a, **kw):
        if a or kw:
            return orig_mvn(mean, cov, *a, **kw)          # samples_of_prior in __init__ (IP.py:194)
        return deltas[it['i'] % (L_MCMC - 1)]

    def fake_uniform(*a, **kw):
        u = us[it['i'] % (L_MCMC - 1)]
        it['i'] += 1
        return u
    np.random.multivariate_normal, np.random.uniform = fake_mvn, fake_uniform
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            res = fn()
    finally:
        np.random.multivariate_normal, np.random.uniform = orig_mvn, orig_uni
    return res, deltas, us


def main():
    out = {}
    for fmt in ('xyz', 'meshgrid'):
        for filt in (False, True):
            UI = ui_doe(fmt, filt)
            PE = G.make_pe(UI)
            true = list(np.array([DH, DS]) + np.array([DH_U, DS_U]))       # the example's +1 SD perturbation
            mat, deltas, us = injected(lambda: PE.doeGetInfoGainMatrix(true))
            tag = f"{fmt}_{'filtered' if filt else 'unfiltered'}"
            out[f'{tag}_info_gain_matrix'] = np.array(mat)
            out[f'{tag}_last_observed'] = np.array(UI.responses_observed)
            out['deltas'], out['us'] = deltas, us
            out['theta_true'] = np.array(true)
    # scanner: every parameter modulation (3 x 3 grid of "true" parameters) x every condition
    UI = ui_doe('xyz', False)
    PE = G.make_pe(UI)
    arr, _, _ = injected(lambda: PE.doeParameterModulationPermutationsScanner())
    out['scanner_xyz_info_gains_matrices_array'] = np.array(arr)
    UI = ui_doe('meshgrid', False)
    PE = G.make_pe(UI)
    arr, _, _ = injected(lambda: PE.doeParameterModulationPermutationsScanner())
    out['scanner_meshgrid_info_gains_matrices_array'] = np.array(arr)
    out['settings'] = np.array([L_MCMC, BURN, STEP, DH, DS, DH_U, DS_U, SIGMA[0]])
    path = os.path.join(G.HERE, "doe.npz")
    np.savez_compressed(path, **out)
    print("doe ->", path, {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
    os.chdir(G.REPO)
