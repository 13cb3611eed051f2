"""Generates tests/golden/*.npz by importing and running the UNMODIFIED reference (/root/reference).

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py
The .npz files are committed; tests read them, never /root/reference.

Import shim: the reference imports matplotlib at package import
(PEUQSE/__init__.py:4 -> plotting_functions.py:3 -> mumpce/Project.py:9); matplotlib is not installed,
so MagicMock modules are placed in sys.modules first.  Nothing in the reference is modified.
"""
import importlib
import os
import shutil
import sys
import tempfile
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
REF = "/root/reference"

for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "matplotlib.ticker",
          "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.mplot3d"]:
    sys.modules.setdefault(m, MagicMock())
sys.path.insert(0, REF)
_scratch = tempfile.mkdtemp(prefix="peuqse_golden_")
os.chdir(_scratch)

import PEUQSE  # noqa: E402
import PEUQSE.UserInput as UserInput  # noqa: E402

from oracle import peuqse_oracle as orc  # noqa: E402
from oracle import philox_np  # noqa: E402


def fresh_userinput():
    # A genuinely fresh module object: importlib.reload keeps attributes that an earlier
    # parameter_estimation.__init__ wrote into the module (e.g. the uniform-prior index list, IP.py:150),
    # which would leak one case's prior structure into the next.
    sys.modules.pop('PEUQSE.UserInput', None)
    UserInput = importlib.import_module('PEUQSE.UserInput')
    UserInput.model['exportResponses'] = False
    UserInput.parameter_estimation_settings['convergence_diagnostics'] = False
    UserInput.parameter_estimation_settings['mcmc_length'] = 10
    return UserInput


def make_pe(UI):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return PEUQSE.parameter_estimation(UI)


def eval_points(PE, thetas):
    import contextlib
    import io
    out = np.empty((len(thetas), 3))
    with contextlib.redirect_stdout(io.StringIO()):
        for i, th in enumerate(thetas):
            out[i, 0] = PE.getLogP(th)
            out[i, 1] = PE.getLogPrior(th)
            out[i, 2] = PE.getLogLikelihood(th)[0]
    return out


X00 = [600, 1100, 1400]
Y00 = [360500, 580500, 1620500]
S00 = [200000, 300000, 200000]


def ui_example00(**kw):
    UI = fresh_userinput()
    UI.responses['responses_abscissa'] = X00
    UI.responses['responses_observed'] = Y00
    UI.responses['responses_observed_uncertainties'] = S00
    UI.model['InputParameterPriorValues'] = kw.get('mu', [200, 500])
    UI.model['InputParametersPriorValuesUncertainties'] = kw.get('unc', [100, 200])
    point, _ = orc.sim_quadratic(X00)
    UI.model['simulateByInputParametersOnlyFunction'] = point
    for key in ('InputParameterPriorValues_upperBounds', 'InputParameterPriorValues_lowerBounds',
                'responses_simulation_uncertainties'):
        if key in kw:
            UI.model[key] = kw[key]
    if 'scaling' in kw:
        UI.parameter_estimation_settings['scaling_uncertainties_type'] = kw['scaling']
    return UI


def theta_cloud(rng, mu, sd, n, wide=8.0):
    a = mu + sd * rng.standard_normal((n // 2, len(mu)))
    b = mu + wide * sd * (rng.random((n - n // 2, len(mu))) - 0.5)
    return np.vstack([a, b])


def case_example00(out):
    rng = np.random.default_rng(100)
    # --- informed prior, point evaluations
    PE = make_pe(ui_example00())
    csv = np.loadtxt(os.path.join(REF, "Examples/Example00/logs_and_csvs/multistart_MAP_logP_and_parameters_values.csv"), delimiter=",")
    known = np.array([[200, 500], [165.1707312, 500.06152225], [150, 400], [255, 139153]], dtype=float)
    thetas = np.vstack([csv[:, 1:], known, theta_cloud(rng, np.array([200., 500.]), np.array([100., 200.]), 240)])
    vals = eval_points(PE, thetas)
    out['ex00_thetas'] = thetas
    out['ex00_logP_prior_lik'] = vals
    out['ex00_csv_rows'] = csv                      # the reference's own stored (logP, a, b) rows
    # --- grid multistart, radii 50x50, interval 5x5 through the reference's doMultiStart
    UI = ui_example00()
    UI.parameter_estimation_settings['multistart_initialPointsDistributionType'] = 'grid'
    UI.parameter_estimation_settings['multistart_gridsearchSamplingInterval'] = [5, 5]
    UI.parameter_estimation_settings['multistart_gridsearchSamplingRadii'] = [50, 50]
    UI.parameter_estimation_settings['multistart_searchType'] = 'getLogP'
    PE = make_pe(UI)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        best = PE.doMultiStart()
    arr = np.array(PE.permutations_MAP_logP_and_parameters_values)
    out['ex00_grid_logP_theta'] = arr                # (10201, 3)
    out['ex00_grid_map_logP'] = np.array(float(PE.map_logP))
    out['ex00_grid_map_theta'] = np.array(PE.map_parameter_set, dtype=float)
    out['ex00_grid_mu_AP'] = np.array(PE.mu_AP_parameter_set)
    out['ex00_grid_stdap'] = np.array(PE.stdap_parameter_set)
    out['ex00_grid_evidence'] = np.array(float(PE.evidence))
    out['ex00_grid_nrows_expanded'] = np.array(len(PE.post_burn_in_samples))
    # per-point copy counts of the reference's expansion (row-run lengths)
    exp_rows = PE.post_burn_in_samples
    logps = PE.post_burn_in_log_posteriors_un_normed_vec.reshape(-1)
    out['ex00_grid_highest_logP'] = np.array(float(PE.highest_logP))
    # multiplicities by matching rows back onto the grid
    keys = {tuple(r[1:]): i for i, r in enumerate(arr)}
    mult = np.zeros(len(arr), dtype=np.int64)
    for r in exp_rows:
        mult[keys[tuple(r)]] += 1
    out['ex00_grid_multiplicity'] = mult
    # --- a small default grid (radii 1, interval = sigma) and an uneven one, list order only
    for name, c, h, r in [("g3", [200., 500.], [100., 200.], [1, 1]),
                          ("g3d", [1., 2., 3.], [0.5, 0.25, 2.], [2, 0, 3])]:
        import PEUQSE.CombinationGeneratorModule as CGM
        lst = CGM.combinationGenerator(np.array(c), np.array(h), np.array(r, dtype='int'), SpreadType="Addition")
        out[f'grid_{name}_center'] = np.array(c)
        out[f'grid_{name}_interval'] = np.array(h)
        out[f'grid_{name}_radii'] = np.array(r)
        out[f'grid_{name}_list'] = np.array([np.array(x, dtype=float) for x in lst])


def case_mh_example00(out):
    """MH with injected draws: np.random.multivariate_normal / np.random.uniform are looked
    up per call at IP.py:2630, 2650, so assigning them feeds our Philox-derived draws."""
    L = 1500
    k = 2
    UI = ui_example00()
    UI.parameter_estimation_settings['mcmc_length'] = L
    UI.parameter_estimation_settings['mcmc_burn_in'] = 100
    UI.parameter_estimation_settings['mcmc_relative_step_length'] = 0.1
    UI.parameter_estimation_settings['mcmc_threshold_filter_samples'] = False
    UI.parameter_estimation_settings['mcmc_info_gain_returned'] = 'log_ratio'
    PE = make_pe(UI)
    z, us = philox_np.mh_draws(0, 0, np.arange(1, L), k)
    sd = np.sqrt(0.1) * np.array([100., 200.])
    deltas = z * sd
    it = {'i': 0}
    orig_mvn, orig_uni = np.random.multivariate_normal, np.random.uniform

    def fake_mvn(mean, cov, *a, **kw):
        d = deltas[it['i']]
        return d

    def fake_uniform(*a, **kw):
        u = us[it['i']]
        it['i'] += 1
        return u
    np.random.multivariate_normal, np.random.uniform = fake_mvn, fake_uniform
    try:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            res = PE.doMetropolisHastings()
    finally:
        np.random.multivariate_normal, np.random.uniform = orig_mvn, orig_uni
    out['mh00_deltas'] = deltas
    out['mh00_us'] = us
    out['mh00_post_samples'] = PE.post_burn_in_samples
    out['mh00_post_logP'] = PE.post_burn_in_log_posteriors_un_normed_vec
    out['mh00_post_log_priors'] = PE.post_burn_in_log_priors_vec
    out['mh00_post_log_likelihoods'] = PE.post_burn_in_log_likelihoods_vec
    out['mh00_burn_samples'] = PE.during_burn_in_samples
    out['mh00_mu_AP'] = np.array(PE.mu_AP_parameter_set)
    out['mh00_stdap'] = np.array(PE.stdap_parameter_set)
    out['mh00_map_logP'] = np.array(float(PE.map_logP))
    out['mh00_map_theta'] = np.array(PE.map_parameter_set)
    out['mh00_evidence'] = np.array(float(PE.evidence))


def case_post_burn_in_statistics(out):
    """calculatePostBurnInStatistics (IP.py:1861-1937) of the UNMODIFIED reference on given sample arrays, for every
    filter benchmark; the info gain is the log-ratio form (the KL form needs a real matplotlib)."""
    import contextlib
    import io
    import warnings
    rng = np.random.default_rng(21)
    N = 3000
    mu, sd = np.array([200., 500.]), np.array([100., 200.])
    th = np.vstack([np.array([165., 500.]) + np.array([55., 130.]) * rng.standard_normal((N - 300, 2)),
                    mu + 6.0 * sd * (rng.random((300, 2)) - 0.5)])          # a posterior-like cloud plus a far tail
    th = th[rng.permutation(N)]
    UI = ui_example00()
    UI.parameter_estimation_settings['mcmc_info_gain_returned'] = 'log_ratio'
    PE = make_pe(UI)
    vals = eval_points(PE, th)
    out['pbs_samples'] = th
    out['pbs_logP'] = vals[:, 0:1]
    out['pbs_log_priors'] = vals[:, 1:2]
    odd = vals[:, 0:1].copy()                                                # logP values the log(-logP) rule cannot handle
    odd[5], odd[17], odd[123] = 0.5, 0.0, 3.25
    out['pbs_logP_odd'] = odd
    cases = [('auto', True, 2.0, 'pbs_logP'), ('auto1', True, 1.0, 'pbs_logP'), ('MAP', True, 2.0, 'pbs_logP'),
             ('mu_AP', True, 2.0, 'pbs_logP'), ('off', False, 2.0, 'pbs_logP'), ('odd', True, 2.0, 'pbs_logP_odd')]
    for name, on, coef, key in cases:
        pes = PE.UserInput.parameter_estimation_settings
        pes['mcmc_threshold_filter_samples'] = on
        pes['mcmc_threshold_filter_coefficient'] = coef if name != 'auto' else 'auto'
        pes['mcmc_threshold_filter_benchmark'] = {'auto1': 'auto', 'off': 'auto', 'odd': 'auto'}.get(name, name)
        PE.post_burn_in_samples = th.copy()
        PE.post_burn_in_log_posteriors_un_normed_vec = out[key].copy()
        PE.post_burn_in_log_priors_vec = vals[:, 1:2].copy()
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = PE.calculatePostBurnInStatistics()
        p = f'pbs_{name}_'
        out[p + 'map_theta'], out[p + 'mu_AP'], out[p + 'stdap'] = np.array(res[0]), np.array(res[1]), np.array(res[2])
        out[p + 'evidence'], out[p + 'info_gain'] = np.array(float(res[3])), np.array(float(res[4]))
        # the filtered arrays are row subsets of the inputs: store the kept row numbers (checked here, bit for bit)
        kept_s, kept_lp, kept_pr = np.array(res[5]), np.array(res[6]), np.array(PE.post_burn_in_log_priors_vec)
        lookup = {tuple(r): i for i, r in enumerate(th)}
        idx = np.array([lookup[tuple(r)] for r in kept_s], dtype=np.int32)
        assert np.all(np.diff(idx) > 0) and np.array_equal(th[idx], kept_s)
        assert np.array_equal(out[key][idx], kept_lp, equal_nan=True) and np.array_equal(vals[idx, 1:2], kept_pr)
        out[p + 'kept_rows'] = idx
        out[p + 'map_logP'] = np.array(float(PE.map_logP))
        out[p + 'map_index'] = np.array(int(PE.map_index))


def case_priors(out):
    rng = np.random.default_rng(101)
    # uniform prior on parameter 1 (uncertainty -1) with bounds; Gaussian on parameter 0
    UI = ui_example00(mu=[200, 500], unc=[100, -1],
                      InputParameterPriorValues_upperBounds=[None, 900],
                      InputParameterPriorValues_lowerBounds=[None, 100])
    PE = make_pe(UI)
    th = np.vstack([theta_cloud(rng, np.array([200., 500.]), np.array([100., 200.]), 60),
                    [[200, 900], [200, 100], [200, np.nextafter(900, 1e9)], [200, np.nextafter(100, 0)], [200, 500]]])
    out['unif_thetas'] = th
    out['unif_vals'] = eval_points(PE, th)
    # both uniform
    UI = ui_example00(mu=[200, 500], unc=[-1, -1],
                      InputParameterPriorValues_upperBounds=[400, 900],
                      InputParameterPriorValues_lowerBounds=[0, 100])
    PE = make_pe(UI)
    th = theta_cloud(rng, np.array([200., 500.]), np.array([100., 200.]), 40)
    out['unif2_thetas'] = th
    out['unif2_vals'] = eval_points(PE, th)
    # bounds with Gaussian priors
    UI = ui_example00(InputParameterPriorValues_upperBounds=[300, None],
                      InputParameterPriorValues_lowerBounds=[100, 0])
    PE = make_pe(UI)
    th = theta_cloud(rng, np.array([200., 500.]), np.array([100., 200.]), 60)
    out['bnd_thetas'] = th
    out['bnd_vals'] = eval_points(PE, th)
    # covariance-matrix prior
    UI = ui_example00(mu=[1.0, 2.0], unc=[[1, .3], [.3, 4]])
    PE = make_pe(UI)
    th = np.vstack([[[1.5, 1.0]], theta_cloud(rng, np.array([1., 2.]), np.array([1., 2.]), 60)])
    out['covp_thetas'] = th
    out['covp_vals'] = eval_points(PE, th)
    # scaling off, and scaling 'mu'
    for tag, sc in [("off", "off"), ("mu", "mu"), ("fix", 3.0)]:
        UI = ui_example00(scaling=sc)
        PE = make_pe(UI)
        th = theta_cloud(rng, np.array([200., 500.]), np.array([100., 200.]), 40)
        out[f'scal_{tag}_thetas'] = th
        out[f'scal_{tag}_vals'] = eval_points(PE, th)
    # simulation uncertainties added in quadrature
    UI = ui_example00(responses_simulation_uncertainties=[1e5, 5e4, 2e5])
    PE = make_pe(UI)
    th = theta_cloud(rng, np.array([200., 500.]), np.array([100., 200.]), 40)
    out['simunc_thetas'] = th
    out['simunc_vals'] = eval_points(PE, th)


def case_tpd(out):
    """C3-style: k=6, n=1024, constant sigma -> per-value path (IP.py:2969-2978)."""
    for n in (1024, 96):
        kBT, c1, c2 = orc.tpd2site_constants(n)
        point, batch = orc.sim_tpd2site(kBT, c1, c2)
        truth = np.array([41.5, 45, 13, 13, 0.1, 0.1])
        sim_t = point(truth)
        sig = 0.02 * np.max(sim_t)
        y = sim_t + sig * np.random.default_rng(0).standard_normal(n)
        UI = fresh_userinput()
        UI.responses['responses_abscissa'] = list(np.linspace(0, 400, n))
        UI.responses['responses_observed'] = list(y)
        UI.responses['responses_observed_uncertainties'] = list(np.full(n, sig))
        UI.model['InputParameterPriorValues'] = [41.5, 41.5, 13.0, 13.0, 0.1, 0.1]
        UI.model['InputParametersPriorValuesUncertainties'] = [20, 20, 13, 13, 0.1, 0.1]
        UI.model['simulateByInputParametersOnlyFunction'] = point
        PE = make_pe(UI)
        rng = np.random.default_rng(102)
        mu = np.array([41.5, 41.5, 13.0, 13.0, 0.1, 0.1])
        sd = np.array([20, 20, 13, 13, 0.1, 0.1])
        th = np.vstack([[[40, 44, 12.5, 13.2, 0.12, 0.08]], [truth], truth + 0.01 * sd * rng.standard_normal((20, 6)),
                        mu + 0.3 * sd * rng.standard_normal((20, 6))])
        out[f'tpd{n}_y'] = y
        out[f'tpd{n}_sigma'] = np.array(sig)
        out[f'tpd{n}_thetas'] = th
        out[f'tpd{n}_vals'] = eval_points(PE, th)
        if n == 1024:
            # short MH run, injected draws, step 1e-5
            L = 120
            UI.parameter_estimation_settings['mcmc_length'] = L
            UI.parameter_estimation_settings['mcmc_burn_in'] = 0
            UI.parameter_estimation_settings['mcmc_relative_step_length'] = 1e-5
            UI.parameter_estimation_settings['mcmc_threshold_filter_samples'] = False
            UI.parameter_estimation_settings['mcmc_info_gain_returned'] = 'log_ratio'
            UI.model['InputParameterInitialGuess'] = list(truth)
            PE = make_pe(UI)
            z, us = philox_np.mh_draws(0, 0, np.arange(1, L), 6)
            deltas = z * (np.sqrt(1e-5) * sd)
            it = {'i': 0}
            orig = np.random.multivariate_normal, np.random.uniform
            np.random.multivariate_normal = lambda *a, **kw: deltas[it['i']]

            def fu(*a, **kw):
                u = us[it['i']]
                it['i'] += 1
                return u
            np.random.uniform = fu
            try:
                import contextlib
                import io
                with contextlib.redirect_stdout(io.StringIO()):
                    PE.doMetropolisHastings()
            finally:
                np.random.multivariate_normal, np.random.uniform = orig
            out['tpdmh_deltas'] = deltas
            out['tpdmh_us'] = us
            out['tpdmh_samples'] = PE.post_burn_in_samples
            out['tpdmh_logP'] = PE.post_burn_in_log_posteriors_un_normed_vec


def case_multiresponse(out):
    """>= 2 responses: diagonal sigmas (Example07 style) and full covariance per response."""
    rng = np.random.default_rng(103)
    for tag, n, full in [("mr_diag", 7, False), ("mr_full5", 5, True), ("mr_full50", 50, True), ("mr_diag40", 40, False)]:
        k = 3
        Phi = rng.standard_normal((2 * n, k))
        point, batch = orc.sim_linear(Phi)
        truth = np.array([0.5, -1.0, 2.0])
        sims = point(truth)
        if full:
            covs = []
            for r in range(2):
                A = rng.standard_normal((n, n))
                covs.append(0.05 * (A @ A.T / n + np.eye(n)))
            y = sims + 0.2 * rng.standard_normal(2 * n)
            unc = covs
        else:
            sig = 0.1 + 0.4 * rng.random(2 * n)
            y = sims + sig * rng.standard_normal(2 * n)
            unc = [list(sig[:n]), list(sig[n:])]
        UI = fresh_userinput()
        UI.responses['responses_abscissa'] = [list(range(n)), list(range(n))]
        UI.responses['responses_observed'] = [list(y[:n]), list(y[n:])]
        UI.responses['responses_observed_uncertainties'] = unc
        UI.model['InputParameterPriorValues'] = [0.0, 0.0, 0.0]
        UI.model['InputParametersPriorValuesUncertainties'] = [2.0, 2.0, 3.0]
        UI.model['simulateByInputParametersOnlyFunction'] = lambda th, n=n, point=point: [point(th)[:n], point(th)[n:]]
        PE = make_pe(UI)
        th = truth + 0.3 * rng.standard_normal((30, k))
        out[f'{tag}_Phi'] = Phi
        out[f'{tag}_y'] = y
        out[f'{tag}_unc'] = np.array(unc, dtype=float)
        out[f'{tag}_thetas'] = th
        out[f'{tag}_vals'] = eval_points(PE, th)


def case_uniform_multistart(out):
    """doMultiStart 'uniform' on a C5-shaped problem (k=10, n=64 cos basis), injected Philox uniforms."""
    k, n, N = 10, 64, 400
    Phi = orc.cos_basis(n, k)
    point, _ = orc.sim_linear(Phi)
    truth = np.random.default_rng(5).standard_normal(k) * 0.5
    y = point(truth) + 0.1 * np.random.default_rng(6).standard_normal(n)
    UI = fresh_userinput()
    UI.responses['responses_abscissa'] = list(np.linspace(0, 1, n))
    UI.responses['responses_observed'] = list(y)
    UI.responses['responses_observed_uncertainties'] = list(np.full(n, 0.1))
    UI.model['InputParameterPriorValues'] = [0.0] * k
    UI.model['InputParametersPriorValuesUncertainties'] = [1.0] * k
    UI.model['simulateByInputParametersOnlyFunction'] = point
    UI.parameter_estimation_settings['multistart_initialPointsDistributionType'] = 'uniform'
    UI.parameter_estimation_settings['multistart_numStartPoints'] = N
    UI.parameter_estimation_settings['multistart_searchType'] = 'getLogP'
    PE = make_pe(UI)
    u01 = philox_np.sweep_uniforms(0, 0, N, k)
    orig = np.random.uniform
    np.random.uniform = lambda lo, hi, shape: lo + (hi - lo) * u01
    try:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            PE.doMultiStart()
    finally:
        np.random.uniform = orig
    arr = np.array(PE.permutations_MAP_logP_and_parameters_values)
    out['c5_y'] = y
    out['c5_u01'] = u01
    out['c5_logP_theta'] = arr
    out['c5_map_logP'] = np.array(float(PE.map_logP))
    out['c5_map_theta'] = np.array(PE.map_parameter_set, dtype=float)
    if hasattr(PE, 'mu_AP_parameter_set'):
        out['c5_mu_AP'] = np.array(PE.mu_AP_parameter_set)
        out['c5_stdap'] = np.array(PE.stdap_parameter_set)


def case_example26(out):
    for f, key in [("Examples/Example26/permutations_MAP_logP_and_parameters_values.csv", "ex26_perm_csv"),
                   ("Examples/Example26/logs_and_csvs/multistart_MAP_logP_and_parameters_values.csv", "ex26_ms_csv")]:
        out[key] = np.loadtxt(os.path.join(REF, f), delimiter=",")


if __name__ == "__main__":
    groups = {
        "example00": [case_example00, case_example26],
        "mh_example00": [case_mh_example00],
        "priors": [case_priors],
        "tpd": [case_tpd],
        "multiresponse": [case_multiresponse],
        "uniform_multistart": [case_uniform_multistart],
        "post_burn_in_statistics": [case_post_burn_in_statistics],
    }
    only = sys.argv[1:]
    for name, fns in groups.items():
        if only and name not in only:
            continue
        out = {}
        for fn in fns:
            fn(out)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "->", path, {k: np.shape(v) for k, v in out.items()})
    os.chdir(REPO)
    shutil.rmtree(_scratch, ignore_errors=True)
