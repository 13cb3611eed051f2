"""GPU parity tests of the design-of-experiments scenario batches (csrc/doe.cu, peuqse_b200/doe.py) against
(a) tests/golden/doe.npz, produced by the UNMODIFIED reference's doeGetInfoGainMatrix /
doeParameterModulationPermutationsScanner on the configuration of Examples/Example13doeFunctionExample with injected draws
(tests/golden/make_golden_doe.py), (b) the numpy oracle replaying the device's own exported draws, and (c) the
single-problem Metropolis-Hastings path (itself pinned against the reference) scenario by scenario.
Tolerances: per-point values 1e-12 relative (exp / log10 of the CUDA math library are within 1-2 ulp of numpy's);
information gains 1e-9 (a mean over post-burn-in rows of differences of log-posteriors)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import peuqse_b200 as PEUQSE
    from oracle import peuqse_oracle as orc
    from oracle import philox_np
    return PEUQSE, orc, philox_np, torch


X00 = [600, 1100, 1400]
Y00 = [360500, 580500, 1620500]
S00 = [200000, 300000, 200000]


def make_pe(PEUQSE, model, observed, unc, mu, punc, settings=None):
    UI = PEUQSE.UserInput.fresh()
    UI.responses['responses_observed'] = observed
    UI.responses['responses_observed_uncertainties'] = unc
    UI.model['InputParameterPriorValues'] = mu
    UI.model['InputParametersPriorValuesUncertainties'] = punc
    UI.model['simulateByInputParametersOnlyFunction'] = model
    for key, v in (settings or {}).items():
        (UI.responses if key in UI.responses else UI.parameter_estimation_settings)[key] = v
    return PEUQSE.parameter_estimation(UI)


def doe_pe(PEUQSE, g, fmt='xyz', filtered=False, info='log_ratio', **extra):
    L, BURN, STEP, DH, DS, DHU, DSU, SIG = g['settings']
    UI = PEUQSE.UserInput.fresh()
    model = PEUQSE.LangmuirLogModel(400, 0.1)
    UI.responses['responses_observed'] = [0.0, 0.0, 0.0]
    UI.responses['responses_observed_uncertainties'] = [SIG] * 3
    UI.responses['independent_variables_values'] = [400, 0.1]
    UI.model['InputParameterPriorValues'] = [DH, DS]
    UI.model['InputParametersPriorValuesUncertainties'] = [DHU, DSU]
    UI.model['simulateByInputParametersOnlyFunction'] = model
    UI.model['populateIndependentVariablesFunction'] = model.populateIndependentVariables
    pes = UI.parameter_estimation_settings
    pes.update({'scaling_uncertainties_type': 'std', 'mcmc_mode': 'unbiased', 'mcmc_random_seed': 0, 'mcmc_burn_in': int(BURN),
                'mcmc_length': int(L), 'mcmc_relative_step_length': float(STEP), 'mcmc_modulate_accept_probability': 0,
                'mcmc_info_gain_returned': info, 'mcmc_threshold_filter_samples': filtered})
    pes.update(extra)
    UI.doe_settings.update({'info_gains_matrices_array_format': fmt, 'independent_variable_grid_center': [500, 0.5],
                            'independent_variable_grid_interval_size': [100, 0.1], 'independent_variable_grid_num_intervals': [2, 2],
                            'parameter_modulation_grid_interval_size': [1, 1], 'parameter_modulation_grid_num_intervals': [1, 1]})
    return PEUQSE.parameter_estimation(UI)


def test_langmuir_points_and_scenario_logp(env, golden):
    PEUQSE, orc, _, torch = env
    g = golden("doe")
    L, BURN, STEP, DH, DS, DHU, DSU, SIG = g['settings']
    PE = doe_pe(PEUQSE, g)
    rng = np.random.default_rng(5)
    th = np.array([DH, DS]) + np.array([DHU, DSU]) * rng.standard_normal((500, 2))
    # single problem under the model's own conditions
    P = orc.build_problem([DH, DS], [DHU, DSU], [0.0, 0.0, 0.0], [SIG] * 3)
    _, bt = orc.sim_langmuir_log(400, 0.1)
    ref = orc.log_p_batch(P, th, bt)
    got = [t.cpu().numpy() for t in PE.getLogP_batch(th, return_parts=True)]
    for a, b in zip(got, ref):
        np.testing.assert_allclose(a, b, rtol=1e-12)
    np.testing.assert_allclose(PE.simulate_batch(th).cpu().numpy(), bt(th), rtol=1e-13)
    # scenario batch: synthetic observations and per-scenario log-posteriors
    from peuqse_b200.doe import ScenarioBatch, conditions_grid
    conds = conditions_grid(PE, PE.UserInput.doe_settings)
    assert np.array_equal(conds, orc.doe_conditions([500, 0.5], [100, 0.1], [2, 2], 'xyz'))
    batch = ScenarioBatch(PE, conds, g['theta_true'])
    obs = batch.observed.cpu().numpy()
    lp, pr, ll = (t.cpu().numpy() for t in batch.logp(np.tile(th[:64], (len(conds), 1, 1)), return_parts=True))
    for s, c in enumerate(conds):
        pt, bt = orc.sim_langmuir_log(c[0], c[1])
        np.testing.assert_allclose(obs[s], pt(g['theta_true']), rtol=1e-13)
        Ps = orc.build_problem([DH, DS], [DHU, DSU], obs[s], [SIG] * 3)
        r = orc.log_p_batch(Ps, th[:64], bt)
        np.testing.assert_allclose(lp[s], r[0], rtol=1e-12)
        np.testing.assert_allclose(pr[s], r[1], rtol=1e-12)
        np.testing.assert_allclose(ll[s], r[2], rtol=1e-12)
    assert np.allclose(obs[-1], g['xyz_unfiltered_last_observed'][0], rtol=1e-13)


@pytest.mark.parametrize("fmt", ["xyz", "meshgrid"])
@pytest.mark.parametrize("filtered", [False, True])
def test_info_gain_matrix_vs_reference_fixture(env, golden, fmt, filtered):
    PEUQSE, orc, _, _ = env
    g = golden("doe")
    L, BURN, STEP, DH, DS, DHU, DSU, SIG = g['settings']
    PE = doe_pe(PEUQSE, g, fmt, filtered)
    M = PE.doeGetInfoGainMatrix(g['theta_true'])
    ref = g[f"{fmt}_{'filtered' if filtered else 'unfiltered'}_info_gain_matrix"]
    assert M.shape == ref.shape == (25, 3)
    assert np.array_equal(M[:, :2], ref[:, :2])                      # the conditions, in the reference's loop order
    np.testing.assert_allclose(M[:, 2], ref[:, 2], rtol=1e-9)        # the reference's information gains
    assert np.array_equal(PE.info_gain_matrix, M)
    # the oracle replaying the device's own draws (exported by a second launch of the same batch)
    from peuqse_b200.mcmc import step_cholesky
    batch = PE._doe_batch
    run = batch.run_mh([DH, DS], int(L), int(BURN), 0, step_cholesky(PE.Q_covmat, STEP), export_draws=True)
    deltas, us = run['deltas'].cpu().numpy(), run['us'].cpu().numpy()
    assert np.array_equal(us[:, 0], g['us']) and all(np.array_equal(us[:, 0], us[:, s]) for s in range(25))   # common draws
    np.testing.assert_allclose(deltas[:, 0], g['deltas'], rtol=1e-13, atol=1e-18)
    assert np.array_equal(run['samples'].cpu().numpy(), PE._doe_run['samples'].cpu().numpy())
    conds = orc.doe_conditions([500, 0.5], [100, 0.1], [2, 2], fmt)
    fk = dict(filter_samples=filtered, benchmark='auto', coef=2.0)
    Mo, st = orc.doe_info_gain_matrix(lambda y: orc.build_problem([DH, DS], [DHU, DSU], y, [SIG] * 3),
                                      lambda c: orc.sim_langmuir_log(c[0], c[1]), conds, g['theta_true'], [DH, DS],
                                      deltas[:, 0], us[:, 0], int(BURN), fk)
    np.testing.assert_allclose(M[:, 2], Mo[:, 2], rtol=1e-9)
    S = PE.doe_scenario_statistics
    from peuqse_b200 import _lib
    SS = _lib.scen_stats_layout(2)
    for s in range(25):
        assert int(S[s, SS['N_KEPT']]) == st[s]['samples'].shape[0]
        np.testing.assert_allclose(S[s, SS['MU_AP']:SS['MU_AP'] + 2], st[s]['mu_AP'], rtol=1e-11)
        np.testing.assert_allclose(S[s, SS['STDAP']:SS['STDAP'] + 2], st[s]['stdap'], rtol=1e-9)
        np.testing.assert_allclose(S[s, SS['MAP_LOGP']], st[s]['map_logP'], rtol=1e-12)
        assert int(S[s, SS['MAP_INDEX']]) == st[s]['map_index']
        np.testing.assert_allclose(S[s, SS['EVIDENCE']], st[s]['evidence'], rtol=1e-11)


@pytest.mark.parametrize("fmt", ["xyz", "meshgrid"])
def test_parameter_modulation_scanner_vs_reference_fixture(env, golden, fmt):
    PEUQSE, orc, _, _ = env
    g = golden("doe")
    PE = doe_pe(PEUQSE, g, fmt, False)
    A = PE.doeParameterModulationPermutationsScanner()
    ref = g[f'scanner_{fmt}_info_gains_matrices_array']
    assert A.shape == ref.shape == (9, 25, 3)
    assert np.array_equal(A[:, :, :2], ref[:, :, :2])
    np.testing.assert_allclose(A[:, :, 2], ref[:, :, 2], rtol=1e-9)
    assert np.array_equal(PE.info_gains_matrices_array, A) and np.array_equal(PE.info_gain_matrix, A[-1])
    assert PE.numParModulationPermutations == 9


def test_map_filter_and_kl_information_gain(env, golden):
    PEUQSE, orc, _, _ = env
    g = golden("doe")
    L, BURN, STEP, DH, DS, DHU, DSU, SIG = g['settings']
    conds = orc.doe_conditions([500, 0.5], [100, 0.1], [2, 2], 'xyz')
    build = lambda y: orc.build_problem([DH, DS], [DHU, DSU], y, [SIG] * 3)
    sim = lambda c: orc.sim_langmuir_log(c[0], c[1])
    # 'MAP' benchmark through the batched statistics kernel
    PE = doe_pe(PEUQSE, g, 'xyz', True, mcmc_threshold_filter_benchmark='MAP', mcmc_threshold_filter_coefficient=1.5)
    M = PE.doeGetInfoGainMatrix(g['theta_true'])
    Mo, _ = orc.doe_info_gain_matrix(build, sim, conds, g['theta_true'], [DH, DS], g['deltas'], g['us'], int(BURN),
                                     dict(filter_samples=True, benchmark='MAP', coef=1.5))
    np.testing.assert_allclose(M[:, 2], Mo[:, 2], rtol=1e-9)
    # KL-divergence form (per-scenario streaming statistics kernels), 'each' matrices
    PE = doe_pe(PEUQSE, g, 'xyz', False, info='KL_divergence')
    PE.UserInput.doe_settings['info_gains_matrices_multiple_parameters'] = 'each'
    M = PE.doeGetInfoGainMatrix(g['theta_true'])
    Mo, st = orc.doe_info_gain_matrix(build, sim, conds, g['theta_true'], [DH, DS], g['deltas'], g['us'], int(BURN),
                                      dict(filter_samples=False), prior_samples=PE.samples_of_prior, info_gain_returned='KL_divergence')
    np.testing.assert_allclose(M[:, 2], Mo[:, 2], rtol=1e-9)
    for d in range(2):
        np.testing.assert_allclose(PE.info_gain_matrices_each_parameter[d][:, 2],
                                   [s['info_gain_KL_each_parameter'][d] for s in st], rtol=1e-9, atol=1e-12)


def test_scenario_chains_equal_single_problem_chains(env):
    """Thread-shape scenario kernel (more than 8192 chains) and the warp shape (few chains): every scenario's chains are
    bit-identical to the chains of a stand-alone problem built from that scenario's conditions and observations."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200.doe import ScenarioBatch
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    rng = np.random.default_rng(11)
    L, burn = 120, 20
    for S, cps in ((2200, 4), (37, 3), (40, 1)):          # cps = 3: one CTA per scenario; powers of two: column tiles
        PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200], settings={'mcmc_length': L})
        conds = np.array(X00, dtype=float) + 50.0 * rng.standard_normal((S, 3))
        truth = np.array([200., 500.]) + np.array([100., 200.]) * rng.standard_normal((S, 2))
        batch = ScenarioBatch(PE, conds, truth)
        chol = step_cholesky(PE.Q_covmat, 0.1)
        run = batch.run_mh([200., 500.], L, burn, 7, chol, chains_per_scenario=cps)
        obs = batch.observed.cpu().numpy()
        stats = batch.statistics(run, burn)
        for s in rng.choice(S, 5, replace=False):
            assert np.array_equal(obs[s], (conds[s] - truth[s, 0]) ** 2 + truth[s, 1])
            PE1 = make_pe(PEUQSE, PEUQSE.QuadraticModel(conds[s]), list(obs[s]), S00, [200, 500], [100, 200], settings={'mcmc_length': L})
            one = mh_launch(PE1, np.tile([200., 500.], (cps, 1)), L, burn, 7, chol)
            sl = slice(s * cps, (s + 1) * cps)
            assert np.array_equal(run['samples'][:, sl].cpu().numpy(), one['samples'].cpu().numpy())
            assert np.array_equal(run['logp'][:, sl].cpu().numpy(), one['logp'].cpu().numpy())
            assert np.array_equal(run['logprior'][:, sl].cpu().numpy(), one['logprior'].cpu().numpy())
            # per-scenario statistics against the oracle on the stacked chains
            smp = one['samples'][burn:].permute(1, 0, 2).reshape(-1, 2).cpu().numpy()
            lp = one['logp'][burn:].permute(1, 0).reshape(-1).cpu().numpy()
            pr = one['logprior'][burn:].permute(1, 0).reshape(-1).cpu().numpy()
            st = orc.post_burn_in_statistics_filtered(smp, lp, pr, filter_samples=False)
            from peuqse_b200 import _lib
            SS = _lib.scen_stats_layout(2)
            np.testing.assert_allclose(stats[s, SS['MU_AP']:SS['MU_AP'] + 2], st['mu_AP'], rtol=1e-12)
            np.testing.assert_allclose(stats[s, SS['STDAP']:SS['STDAP'] + 2], st['stdap'], rtol=1e-10)
            assert int(stats[s, SS['MAP_INDEX']]) == st['map_index'] and stats[s, SS['MAP_LOGP']] == st['map_logP']
            np.testing.assert_allclose(stats[s, SS['INFO_GAIN_LOG_RATIO']], st['info_gain_log_ratio'], rtol=1e-10)
            assert np.array_equal(stats[s, SS['MAP_THETA']:SS['MAP_THETA'] + 2], st['map_parameter_set'])


def test_scenarios_with_many_observations_tpd(env, golden):
    """Scenario batches of a model with 1024 observations (the warp shape: lanes stride over the scenario's own rows of
    constants): chains bit-identical to stand-alone runs of the same problem, per-point values against the oracle."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200.doe import ScenarioBatch
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    n = 1024
    mu, sd = [41.5, 41.5, 13., 13., 0.1, 0.1], [20., 20., 13., 13., 0.1, 0.1]
    truth = np.array([41.5, 45., 13., 13., 0.1, 0.1])
    kBT, c1, c2 = orc.tpd2site_constants(n)

    class RampModel(PEUQSE.Tpd2SiteModel):
        """Independent variables: heating rate and initial coverage of the prescribed-coverage TPD model."""

        def with_conditions(self, values):
            return RampModel.linear_ramp(n, beta=values[0], cov0=values[1])

    base = RampModel.linear_ramp(n)
    pt, _ = orc.sim_tpd2site(kBT, c1, c2)
    sig = [0.02 * float(np.max(pt(truth)))] * n
    PE = make_pe(PEUQSE, base, list(pt(truth)), sig, mu, sd, settings={'mcmc_length': 60, 'responses_observed_max_covmat_size': 5000})
    conds = np.array([[0.5, 0.5], [0.4, 0.6], [0.7, 0.3], [0.55, 0.45], [0.3, 0.5]])
    batch = ScenarioBatch(PE, conds, truth)
    obs = batch.observed.cpu().numpy()
    rng = np.random.default_rng(2)
    th = truth + 0.01 * np.array(sd) * rng.standard_normal((5, 16, 6))
    lp, pr, ll = (t.cpu().numpy() for t in batch.logp(th, return_parts=True))
    chol = step_cholesky(PE.Q_covmat, 2.5e-6)
    run = batch.run_mh(truth, 60, 10, 5, chol, chains_per_scenario=2)
    for s, c in enumerate(conds):
        m = RampModel.linear_ramp(n, beta=c[0], cov0=c[1])
        pts, bts = orc.sim_tpd2site(m.kBT, m.cov1, m.cov2)
        np.testing.assert_allclose(obs[s], pts(truth), rtol=1e-12)
        Ps = orc.build_problem(mu, sd, obs[s], sig, max_covmat_size=5000)
        r = orc.log_p_batch(Ps, th[s], bts)
        np.testing.assert_allclose(lp[s], r[0], rtol=1e-12)
        np.testing.assert_allclose(ll[s], r[2], rtol=1e-12)
        PE1 = make_pe(PEUQSE, m, list(obs[s]), sig, mu, sd, settings={'mcmc_length': 60, 'responses_observed_max_covmat_size': 5000})
        one = mh_launch(PE1, np.tile(truth, (2, 1)), 60, 10, 5, chol)
        assert np.array_equal(run['samples'][:, 2 * s:2 * s + 2].cpu().numpy(), one['samples'].cpu().numpy())
        # the stand-alone kernel of this model takes the fast rate-law form (csrc/mh.cu: FAST; fused exponent, table-driven
        # exp), the scenario kernel the reference's operation order: same accept / reject decisions, log-posteriors to 1e-12
        np.testing.assert_allclose(run['logp'][:, 2 * s:2 * s + 2].cpu().numpy(), one['logp'].cpu().numpy(), rtol=1e-12)


def test_scenario_statistics_edge_cases(env):
    """peuqse_scenarios_stats on hand-made MH-layout arrays: non-negative and NaN logP (the 'auto' threshold becomes NaN and
    nothing is removed, IP.py:1894-1915 with arrayThresholdFilter's NaN comparisons), -inf log-priors (infinite ratios count
    as 0, IP.py:1801), identical logP values, two chains per scenario and three (the one-CTA-per-scenario shape)."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(8)
    for cps in (2, 3):
        S, R, k, burn = 6, 300, 3, 20
        C = S * cps
        smp = rng.normal([1.0, -50.0, 3e3], [0.5, 4.0, 100.0], size=(R, C, k))
        lp = -0.5 * rng.chisquare(3, size=(R, C)) - 40.0
        pr = -1.0 - 0.1 * rng.random((R, C))
        lp[100, 1 * cps] = 0.5                      # scenario 1: a positive logP
        lp[150, 2 * cps + 1] = np.nan               # scenario 2: a NaN
        pr[60:70, 3 * cps] = -np.inf                # scenario 3: -inf priors
        lp[:, 4 * cps:5 * cps] = -41.25             # scenario 4: identical values
        lp[30, 5 * cps] = -1e4                      # scenario 5: one far outlier (removed by 'auto' and by 'MAP')
        d_smp, d_lp, d_pr = (torch.as_tensor(a).cuda().contiguous() for a in (smp, lp, pr))     # kept alive across the launches
        lay = _lib.scen_stats_layout(k)
        for mode, bench in ((_lib.FILTER_AUTO, 'auto'), (_lib.FILTER_MAP, 'MAP'), (_lib.FILTER_NONE, None)):
            out = torch.empty((S, lay['DOUBLES']), dtype=torch.float64, device='cuda')
            _lib.check(L.peuqse_scenarios_stats(k, mode, 2.0, S, cps, R, burn, _lib.ptr(d_smp), _lib.ptr(d_lp), _lib.ptr(d_pr),
                                                _lib.ptr(out), _lib.current_stream()))
            got = out.cpu().numpy()
            for s in range(S):
                if s == 2 and mode != _lib.FILTER_AUTO:
                    continue                         # python's max() over a NaN is order dependent: only the 'auto' rule is defined
                cols = slice(s * cps, (s + 1) * cps)
                flat = lambda a: np.moveaxis(a[burn:, cols], 1, 0).reshape((R - burn) * cps, *a.shape[2:])
                with np.errstate(all='ignore'):
                    st = orc.post_burn_in_statistics_filtered(flat(smp), flat(lp), flat(pr), benchmark=bench or 'auto', coef=2.0,
                                                              filter_samples=bench is not None)
                assert int(got[s, lay['N_KEPT']]) == st['samples'].shape[0], (cps, bench, s)
                np.testing.assert_allclose(got[s, lay['MU_AP']:lay['MU_AP'] + k], st['mu_AP'], rtol=1e-12)
                if s == 2:
                    continue
                np.testing.assert_allclose(got[s, lay['STDAP']:lay['STDAP'] + k], st['stdap'], rtol=1e-10)
                assert int(got[s, lay['MAP_INDEX']]) == st['map_index'] and got[s, lay['MAP_LOGP']] == st['map_logP']
                np.testing.assert_allclose(got[s, lay['EVIDENCE']], st['evidence'], rtol=1e-12)
                np.testing.assert_allclose(got[s, lay['INFO_GAIN_LOG_RATIO']], st['info_gain_log_ratio'], rtol=1e-11)
            if mode == _lib.FILTER_AUTO:
                assert int(got[1, lay['N_KEPT']]) == (R - burn) * cps        # NaN threshold: nothing removed
                assert int(got[5, lay['N_KEPT']]) < (R - burn) * cps         # the far outlier is gone


def test_large_scenario_batch_properties(env, golden):
    """Size-independent properties at the size of a real design study (81 x 81 conditions x 3 modulations = 19 683 scenarios,
    thread-shape kernel): (i) a scenario's result does not depend on where it sits in the batch or on what surrounds it
    (the batch reversed gives the results reversed, bit for bit), (ii) duplicated scenarios give duplicated results,
    (iii) spot-checked scenarios equal the 25-scenario reference fixture's values where conditions coincide."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200 import _lib
    from peuqse_b200.doe import ScenarioBatch
    from peuqse_b200.mcmc import step_cholesky
    g = golden("doe")
    L, BURN, STEP, DH, DS, DHU, DSU, SIG = g['settings']
    PE = doe_pe(PEUQSE, g, 'xyz', True)
    T = np.linspace(300, 700, 81)
    pA = np.linspace(0.3, 0.7, 81)
    conds = np.array([[t, p] for t in T for p in pA])
    truth = np.array([DH, DS]) + np.array([DHU, DSU]) * np.array([[1, 1], [0, 0], [-1, 1]])
    all_c = np.tile(conds, (3, 1))
    all_t = np.repeat(truth, len(conds), axis=0)
    all_c[7], all_t[7] = all_c[12345], all_t[12345]                   # a duplicate far away in the batch
    S = len(all_c)
    chol = step_cholesky(PE.Q_covmat, STEP)
    lay = _lib.scen_stats_layout(2)

    def run(c, t):
        b = ScenarioBatch(PE, c, t)
        r = b.run_mh([DH, DS], int(L), int(BURN), 0, chol)
        st = b.statistics(r, int(BURN), _lib.FILTER_AUTO, 2.0)
        kl = b.kl_information_gain(r, int(BURN), PE.samples_of_prior)
        return st, kl
    st, kl = run(all_c, all_t)
    assert st.shape == (S, lay['DOUBLES']) and S == 19683
    st_r, kl_r = run(all_c[::-1].copy(), all_t[::-1].copy())
    assert np.array_equal(st, st_r[::-1], equal_nan=True) and np.array_equal(kl, kl_r[::-1], equal_nan=True)
    assert np.array_equal(st[7], st[12345], equal_nan=True) and np.array_equal(kl[7], kl[12345])
    assert np.isfinite(st[:, lay['INFO_GAIN_LOG_RATIO']]).all()
    # conditions shared with the reference's 5 x 5 grid, "true" parameters = the fixture's +1 SD perturbation (first block)
    ref = g['xyz_filtered_info_gain_matrix']
    hits = 0
    for row in ref:
        i = np.nonzero((np.abs(conds[:, 0] - row[0]) < 1e-9) & (np.abs(conds[:, 1] - row[1]) < 1e-12))[0]
        if len(i) and i[0] != 7:
            np.testing.assert_allclose(st[i[0], lay['INFO_GAIN_LOG_RATIO']], row[2], rtol=1e-9)
            hits += 1
    assert hits >= 9


def test_info_gain_cutoff_in_scenario_batches(env, golden):
    """mcmc_info_gain_cutoff != 0 (IP.py:1805-1843) inside design-of-experiments batches: the truncated log-ratio information
    gain of every scenario against the oracle (whose cutoff routine is pinned bit for bit by tests/golden/info_gain_cutoff.npz)
    replaying the same draws; a cutoff above every prior density leaves no rows: NaN, as numpy's mean of nothing."""
    PEUQSE, orc, _, _ = env
    g = golden("doe")
    L, BURN, STEP, DH, DS, DHU, DSU, SIG = g['settings']
    conds = orc.doe_conditions([500, 0.5], [100, 0.1], [2, 2], 'xyz')
    build = lambda y: orc.build_problem([DH, DS], [DHU, DSU], y, [SIG] * 3)
    sim = lambda c: orc.sim_langmuir_log(c[0], c[1])
    plain = doe_pe(PEUQSE, g, 'xyz', True).doeGetInfoGainMatrix(g['theta_true'])
    for cutoff in (0.13, 0.4):
        PE = doe_pe(PEUQSE, g, 'xyz', True, mcmc_info_gain_cutoff=cutoff)
        M = PE.doeGetInfoGainMatrix(g['theta_true'])
        Mo, _ = orc.doe_info_gain_matrix(build, sim, conds, g['theta_true'], [DH, DS], g['deltas'], g['us'], int(BURN),
                                         dict(filter_samples=True, benchmark='auto', coef=2.0, info_gain_cutoff=cutoff))
        np.testing.assert_allclose(M[:, 2], Mo[:, 2], rtol=1e-9, equal_nan=True)
        if cutoff == 0.13:
            fin = np.isfinite(M[:, 2])
            assert 10 <= fin.sum() < 25 and not np.allclose(M[fin, 2], plain[fin, 2])        # the cutoff removed rows, in some scenarios all
        else:
            assert np.isnan(M[:, 2]).all()
