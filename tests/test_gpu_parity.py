"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI by the Python
boundary class, against (a) fixtures produced by the unmodified reference and (b) the numpy oracle on the
same seeded inputs.  Tolerances: bit-exact wherever the reference's summation order is reproduced (diagonal
covariances, n <= 128, thread shape); 1e-12 relative (north-star budget) where a parallel reduction or a
dense BLAS order is involved."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

X00 = [600, 1100, 1400]
Y00 = [360500, 580500, 1620500]
S00 = [200000, 300000, 200000]


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import peuqse_b200 as PEUQSE
    from oracle import peuqse_oracle as orc
    from oracle import philox_np
    return PEUQSE, orc, philox_np, torch


def make_pe(PEUQSE, model, observed, unc, mu, punc, settings=None, **modelkw):
    UI = PEUQSE.UserInput.fresh()
    UI.responses['responses_observed'] = observed
    UI.responses['responses_observed_uncertainties'] = unc
    UI.model['InputParameterPriorValues'] = mu
    UI.model['InputParametersPriorValuesUncertainties'] = punc
    UI.model['simulateByInputParametersOnlyFunction'] = model
    for k, v in modelkw.items():
        UI.model[k] = v
    for k, v in (settings or {}).items():
        if k in UI.responses:
            UI.responses[k] = v
        else:
            UI.parameter_estimation_settings[k] = v
    return PEUQSE.parameter_estimation(UI)


def parts(PE, th):
    lp, pr, ll = PE.getLogP_batch(th, return_parts=True)
    return lp.cpu().numpy(), pr.cpu().numpy(), ll.cpu().numpy()


def test_philox_blocks(env):
    PEUQSE, orc, philox_np, torch = env
    from peuqse_b200 import _lib
    L = _lib.load()
    for seed, stream, first, block in [(0, 1, 0, 0), (12345678901234567, 2, 2 ** 33 + 5, 7), (7, 3, 10, 4000000000)]:
        out = torch.empty((1000, 4), dtype=torch.int32, device="cuda")
        _lib.check(L.peuqse_philox_blocks(seed, stream, first, 1000, block, _lib.ptr(out), _lib.current_stream()))
        got = out.cpu().numpy().view(np.uint32)
        want = philox_np.raw_block(seed, stream, np.uint64(first) + np.arange(1000, dtype=np.uint64), block)
        assert np.array_equal(got, want)


def test_example00_points_bit_exact(env, golden):
    PEUQSE, orc, _, _ = env
    g = golden("example00")
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200])
    lp, pr, ll = parts(PE, g['ex00_thetas'])
    vals = g['ex00_logP_prior_lik']            # produced by the reference's getLogP/getLogPrior/getLogLikelihood
    assert np.array_equal(lp, vals[:, 0])
    assert np.array_equal(pr, vals[:, 1])
    assert np.array_equal(ll, vals[:, 2])
    # the reference repo's own stored rows
    rows = g['ex00_csv_rows']
    assert np.array_equal(PE.getLogP_batch(rows[:, 1:]).cpu().numpy(), rows[:, 0])
    # single-point API
    assert PE.getLogP([200, 500]) == -42.81726459961094
    assert PE.getLogPrior([200, 500]) == -1.8378770664093453
    ll1, sim = PE.getLogLikelihood([200, 500])
    assert ll1 == -40.97938753320159 and np.array_equal(sim[0], (np.array(X00) - 200.) ** 2 + 500)


def test_priors_bounds_scaling(env, golden):
    PEUQSE, orc, _, _ = env
    g = golden("priors")
    Q = PEUQSE.QuadraticModel

    def chk(PE, tag, exact=True):
        lp, pr, ll = parts(PE, g[f'{tag}_thetas'])
        v = g[f'{tag}_vals']
        if exact:
            assert np.array_equal(lp, v[:, 0]) and np.array_equal(pr, v[:, 1]) and np.array_equal(ll, v[:, 2]), tag
        else:
            assert np.array_equal(np.isfinite(lp), np.isfinite(v[:, 0]))
            f = np.isfinite(v[:, 0])
            np.testing.assert_allclose(lp[f], v[f, 0], rtol=1e-13)
            np.testing.assert_allclose(pr, v[:, 1], rtol=1e-13)
            assert np.array_equal(ll, v[:, 2])
    chk(make_pe(PEUQSE, Q(X00), Y00, S00, [200, 500], [100, -1], InputParameterPriorValues_upperBounds=[None, 900],
                InputParameterPriorValues_lowerBounds=[None, 100]), 'unif')
    chk(make_pe(PEUQSE, Q(X00), Y00, S00, [200, 500], [-1, -1], InputParameterPriorValues_upperBounds=[400, 900],
                InputParameterPriorValues_lowerBounds=[0, 100]), 'unif2')
    chk(make_pe(PEUQSE, Q(X00), Y00, S00, [200, 500], [100, 200], InputParameterPriorValues_upperBounds=[300, None],
                InputParameterPriorValues_lowerBounds=[100, 0]), 'bnd')
    chk(make_pe(PEUQSE, Q(X00), Y00, S00, [1.0, 2.0], [[1, .3], [.3, 4]]), 'covp', exact=False)
    for tag, sc in [("off", "off"), ("mu", "mu"), ("fix", 3.0)]:
        chk(make_pe(PEUQSE, Q(X00), Y00, S00, [200, 500], [100, 200], settings={'scaling_uncertainties_type': sc}),
            f'scal_{tag}', exact=False)
    chk(make_pe(PEUQSE, Q(X00), Y00, S00, [200, 500], [100, 200], responses_simulation_uncertainties=[1e5, 5e4, 2e5]), 'simunc')


@pytest.mark.parametrize("n", [1024, 96])
def test_tpd_points(env, golden, n):
    PEUQSE, orc, _, _ = env
    g = golden("tpd")
    PE = make_pe(PEUQSE, PEUQSE.Tpd2SiteModel.linear_ramp(n), list(g[f'tpd{n}_y']), list(np.full(n, float(g[f'tpd{n}_sigma']))),
                 [41.5, 41.5, 13.0, 13.0, 0.1, 0.1], [20, 20, 13, 13, 0.1, 0.1])
    lp, pr, ll = parts(PE, g[f'tpd{n}_thetas'])
    v = g[f'tpd{n}_vals']
    assert np.array_equal(pr, v[:, 1])
    np.testing.assert_allclose(ll, v[:, 2], rtol=1e-12)      # north-star tolerance (exp ulp + reduction order)
    np.testing.assert_allclose(lp, v[:, 0], rtol=1e-12)
    assert np.max(np.abs(lp - v[:, 0]) / np.abs(v[:, 0])) < 2e-14   # what is actually achieved


def test_multiresponse_builtin_and_external(env, golden):
    PEUQSE, orc, _, torch = env
    g = golden("multiresponse")
    for tag, n, full in [("mr_diag", 7, False), ("mr_diag40", 40, False), ("mr_full5", 5, True), ("mr_full50", 50, True)]:
        Phi, y, unc = g[f'{tag}_Phi'], g[f'{tag}_y'], g[f'{tag}_unc']
        obs = [list(y[:n]), list(y[n:])]
        u = [unc[0], unc[1]] if full else [list(unc[0]), list(unc[1])]
        th, v = g[f'{tag}_thetas'], g[f'{tag}_vals']
        Phi_t = torch.from_numpy(Phi).cuda()
        for model in (PEUQSE.LinearModel(Phi), PEUQSE.TorchSimulator(lambda t, P=Phi_t: t @ P.T)):
            PE = make_pe(PEUQSE, model, obs, u, [0., 0., 0.], [2., 2., 3.])
            lp, pr, ll = parts(PE, th)
            np.testing.assert_allclose(pr, v[:, 1], rtol=1e-13)
            np.testing.assert_allclose(ll, v[:, 2], rtol=1e-12)
            np.testing.assert_allclose(lp, v[:, 0], rtol=1e-12)
        if not full:
            # built-in linear model + diagonal responses: same order as the oracle -> bit exact
            P = orc.build_problem([0., 0., 0.], [2., 2., 3.], [y[:n], y[n:]], [unc[0], unc[1]])
            _, bt = orc.sim_linear(Phi)
            want = orc.log_p_batch(P, th, lambda t: [bt(t)[:, :n], bt(t)[:, n:]])
            PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), obs, u, [0., 0., 0.], [2., 2., 3.])
            lp, pr, ll = parts(PE, th)
            assert np.array_equal(lp, want[0]) and np.array_equal(ll, want[2]), tag


def test_nan_and_failure_conventions(env):
    PEUQSE, orc, _, torch = env
    # simulator returning NaN for some rows -> logLik = -inf, never an exception (IP.py:2880-2882)
    def sim(t):
        out = (torch.tensor(X00, dtype=torch.float64, device=t.device)[None, :] - t[:, 0:1]) ** 2 + t[:, 1:2]
        out[t[:, 0] > 250] = float('nan')
        return out
    PE = make_pe(PEUQSE, sim, Y00, S00, [200, 500], [100, 200])
    lp, pr, ll = parts(PE, np.array([[200., 500.], [300., 500.]]))
    assert lp[0] == -42.81726459961094 and lp[1] == -np.inf and ll[1] == -np.inf and np.isfinite(pr[1])
    assert PE.getLogLikelihood([300., 500.]) == (float('-inf'), None)
    # empty batch
    assert PE.getLogP_batch(np.zeros((0, 2))).shape[0] == 0


def test_grid_multistart_example00(env, golden):
    PEUQSE, orc, _, _ = env
    g = golden("example00")
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                 settings={'multistart_initialPointsDistributionType': 'grid', 'multistart_gridsearchSamplingInterval': [5, 5],
                           'multistart_gridsearchSamplingRadii': [50, 50]})
    best = PE.doMultiStart()
    arr = g['ex00_grid_logP_theta']                     # the reference's permutations_MAP_logP_and_parameters_values
    assert np.array_equal(PE.permutations_MAP_logP_and_parameters_values, arr)
    assert np.array_equal(PE.listOfPermutations, arr[:, 1:])
    assert PE.map_index == 13
    assert PE.map_logP == float(g['ex00_grid_map_logP']) == PE.highest_logP
    assert np.array_equal(PE.map_parameter_set, g['ex00_grid_map_theta']) and np.array_equal(best[0], g['ex00_grid_map_theta'])
    assert PE.sweep_result.sum_weights == int(g['ex00_grid_nrows_expanded']) == 225166
    np.testing.assert_allclose(PE.mu_AP_parameter_set, g['ex00_grid_mu_AP'], rtol=1e-12)
    np.testing.assert_allclose(PE.stdap_parameter_set, g['ex00_grid_stdap'], rtol=1e-10)
    np.testing.assert_allclose(PE.evidence, float(g['ex00_grid_evidence']), rtol=1e-11)
    # expanded samples == the reference's post_burn_in_samples row multiplicities
    S = PE.expandedPostBurnInSamples()
    assert S.shape == (225166, 2)
    P = orc.build_problem([200, 500], [100, 200], Y00, S00)
    st = orc.weighted_statistics(arr[:, 0], arr[:, 1:], g['ex00_grid_multiplicity'])
    assert np.array_equal(PE.marginal_histograms[0], st['hist'].astype(np.int64))
    np.testing.assert_allclose(PE.marginal_histograms[1], st['edges'], rtol=0, atol=0)


def test_grid_order_and_shapes(env, golden):
    PEUQSE, orc, _, _ = env
    g = golden("example00")
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200])
    for name in ("g3", "g3d"):
        c, h, r = g[f'grid_{name}_center'], g[f'grid_{name}_interval'], g[f'grid_{name}_radii']
        if len(c) != 2:
            from peuqse_b200.sweep import make_spec
            spec = make_spec("grid", len(c), c, interval=h, radii=r)
            pts = PE._points_from_spec(spec, 0, len(g[f'grid_{name}_list']))
        else:
            pts = PE.generateInitialPoints(initialPointsDistributionType='grid', centerPoint=c, gridsearchSamplingInterval=list(h),
                                           gridsearchSamplingRadii=list(r))
        assert np.array_equal(pts, g[f'grid_{name}_list'])
    # default grid: radii 1, interval = prior sigma (IP.py:758-767)
    PE.UserInput.parameter_estimation_settings['multistart_initialPointsDistributionType'] = 'grid'
    PE.doMultiStart()
    assert PE.sweep_result.total == 9
    assert np.array_equal(PE.listOfPermutations, g['grid_g3_list'])
    # big random window of a 5-axis grid against the closed-form oracle decode
    from peuqse_b200.sweep import make_spec
    c, h, r = [1., -2., 3., 0.5, 10.], [0.1, 0.2, 0.3, 0.05, 1.5], [7, 0, 13, 1, 40]
    spec = make_spec("grid", 5, c, interval=h, radii=r)
    N = orc.grid_size(r)
    first = N - 60000
    pts = PE._points_from_spec(spec, first, 60000)
    assert np.array_equal(pts, orc.grid_decode(c, h, r, np.arange(first, N)))
    assert np.array_equal(PE._points_from_spec(spec, 0, 5000), orc.grid_decode(c, h, r, np.arange(5000)))


def test_uniform_multistart_c5(env, golden):
    PEUQSE, orc, philox_np, _ = env
    g = golden("uniform_multistart")
    k, n, N = 10, 64, 400
    Phi = orc.cos_basis(n, k)
    PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), list(g['c5_y']), list(np.full(n, 0.1)), [0.0] * k, [1.0] * k,
                 settings={'multistart_initialPointsDistributionType': 'uniform', 'multistart_numStartPoints': N,
                           'multistart_random_seed': 0})
    PE.doMultiStart()
    arr = g['c5_logP_theta']                       # reference fed the same Philox uniforms
    got = PE.permutations_MAP_logP_and_parameters_values
    assert np.array_equal(got[:, 1:], arr[:, 1:])  # points bit-exact
    np.testing.assert_allclose(got[:, 0], arr[:, 0], rtol=1e-13)
    assert np.array_equal(PE.map_parameter_set, g['c5_map_theta'])
    np.testing.assert_allclose(PE.map_logP, float(g['c5_map_logP']), rtol=1e-13)
    np.testing.assert_allclose(PE.mu_AP_parameter_set, g['c5_mu_AP'], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(PE.stdap_parameter_set, g['c5_stdap'], rtol=1e-9, atol=1e-13)
    # built-in linear model follows the oracle's accumulation order -> bit-exact against the oracle
    P = orc.build_problem([0.0] * k, [1.0] * k, g['c5_y'], np.full(n, 0.1))
    _, bt = orc.sim_linear(Phi)
    assert np.array_equal(got[:, 0], orc.log_p_batch(P, got[:, 1:], bt)[0])


def test_mh_example00_step_for_step(env, golden):
    PEUQSE, orc, philox_np, _ = env
    g = golden("mh_example00")
    L = 1500
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                 settings={'mcmc_length': L, 'mcmc_burn_in': 100, 'mcmc_relative_step_length': 0.1, 'mcmc_random_seed': 0,
                           'mcmc_threshold_filter_samples': False, 'mcmc_info_gain_returned': 'log_ratio'})
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    out = mh_launch(PE, [[200., 500.]], L, 100, 0, step_cholesky(PE.Q_covmat, 0.1), export_draws=True)
    deltas, us = out['deltas'].cpu().numpy()[:, 0], out['us'].cpu().numpy()[:, 0]
    # the device draws are the numpy-Philox draws the reference was fed when the fixture was made
    assert np.array_equal(us, g['mh00_us'])
    np.testing.assert_allclose(deltas, g['mh00_deltas'], rtol=1e-13, atol=1e-11)
    # replay the device's own draws through the oracle: step-for-step identical trajectory
    P = orc.build_problem([200, 500], [100, 200], Y00, S00)
    _, bt = orc.sim_quadratic(X00)
    ref = orc.mh_chains_batch(P, np.array([[200., 500.]]), deltas[:, None, :], us[:, None], bt)
    assert np.array_equal(out['samples'].cpu().numpy(), ref['samples'])
    assert np.array_equal(out['logp'].cpu().numpy(), ref['logP'])
    assert np.array_equal(out['logprior'].cpu().numpy()[1:], ref['log_priors'][1:])
    assert np.array_equal(out['loglik'].cpu().numpy()[1:], ref['log_likelihoods'][1:])
    same_draws = np.array_equal(deltas, g['mh00_deltas'])
    # through the reference-named API
    res = PE.doMetropolisHastings()
    if same_draws:       # identical draws -> identical to the reference run stored in the fixture
        assert np.array_equal(PE.post_burn_in_samples, g['mh00_post_samples'])
        assert np.array_equal(PE.post_burn_in_log_posteriors_un_normed_vec, g['mh00_post_logP'])
        assert np.array_equal(PE.post_burn_in_log_priors_vec, g['mh00_post_log_priors'])
        assert np.array_equal(PE.during_burn_in_samples, g['mh00_burn_samples'])
        # statistics come from the streaming device reductions (csrc/stats.cu): summation order differs from numpy's
        np.testing.assert_allclose(PE.mu_AP_parameter_set, g['mh00_mu_AP'], rtol=1e-12)
        np.testing.assert_allclose(PE.stdap_parameter_set, g['mh00_stdap'], rtol=1e-12)
        assert float(PE.map_logP) == float(g['mh00_map_logP'])
        np.testing.assert_allclose(float(PE.evidence), float(g['mh00_evidence']), rtol=1e-12)
    else:
        acc_ref = np.mean(np.any(np.diff(g['mh00_post_samples'], axis=0) != 0, axis=1))
        acc_got = np.mean(np.any(np.diff(PE.post_burn_in_samples, axis=0) != 0, axis=1))
        assert abs(acc_ref - acc_got) < 0.05
    assert PE.post_burn_in_samples.shape == (1400, 2) and len(res) == 7
    # in-kernel statistics == numpy statistics of the stored rows
    mom = PE.mcmc_chain_moments
    np.testing.assert_allclose(mom['mean'], np.mean(PE.post_burn_in_samples, axis=0), rtol=1e-12)
    np.testing.assert_allclose(mom['std'], np.std(PE.post_burn_in_samples, axis=0), rtol=1e-10)
    assert mom['map_logP'] == float(np.ravel(PE.map_logP)[0])


def test_mh_tpd_step_for_step(env, golden):
    PEUQSE, orc, _, _ = env
    g = golden("tpd")
    n, L = 1024, 120
    sd = np.array([20, 20, 13, 13, 0.1, 0.1])
    PE = make_pe(PEUQSE, PEUQSE.Tpd2SiteModel.linear_ramp(n), list(g['tpd1024_y']), list(np.full(n, float(g['tpd1024_sigma']))),
                 [41.5, 41.5, 13.0, 13.0, 0.1, 0.1], list(sd))
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    truth = np.array([[41.5, 45, 13, 13, 0.1, 0.1]])
    out = mh_launch(PE, truth, L, 0, 0, step_cholesky(PE.Q_covmat, 1e-5), export_draws=True)
    deltas, us = out['deltas'].cpu().numpy(), out['us'].cpu().numpy()
    assert np.array_equal(us[:, 0], g['tpdmh_us'])
    np.testing.assert_allclose(deltas[:, 0], g['tpdmh_deltas'], rtol=1e-12, atol=1e-15)
    kBT, c1, c2 = orc.tpd2site_constants(n)
    _, bt = orc.sim_tpd2site(kBT, c1, c2)
    P = orc.build_problem([41.5, 41.5, 13.0, 13.0, 0.1, 0.1], sd, g['tpd1024_y'], np.full(n, float(g['tpd1024_sigma'])))
    ref = orc.mh_chains_batch(P, truth, deltas, us, bt)
    # accept/reject sequence step for step, states identical (they are sums of the same doubles)
    assert np.array_equal(out['samples'].cpu().numpy(), ref['samples'])
    np.testing.assert_allclose(out['logp'].cpu().numpy(), ref['logP'], rtol=1e-12)
    assert 0.05 < ref['accepted'][1:].mean() < 0.95


def test_mh_many_chains_statistics(env):
    """4096 Example00 chains: pooled posterior mean/std against the grid-quadrature posterior (MC error)."""
    PEUQSE, orc, _, _ = env
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                 settings={'mcmc_length': 2000, 'mcmc_burn_in': 500, 'mcmc_relative_step_length': 0.5, 'mcmc_random_seed': 3,
                           'mcmc_nchains': 4096, 'mcmc_threshold_filter_samples': False, 'mcmc_info_gain_returned': 'log_ratio'})
    PE.doMetropolisHastings()
    assert PE.post_burn_in_samples.shape == (4096 * 1500, 2)
    # quadrature reference
    a = np.linspace(-300, 700, 801)
    b = np.linspace(-700, 1700, 801)
    A, B = np.meshgrid(a, b, indexing='ij')
    P = orc.build_problem([200, 500], [100, 200], Y00, S00)
    _, bt = orc.sim_quadratic(X00)
    lp = orc.log_p_batch(P, np.stack([A.ravel(), B.ravel()], axis=1), bt)[0]
    w = np.exp(lp - lp.max())
    mean = np.array([(w * A.ravel()).sum(), (w * B.ravel()).sum()]) / w.sum()
    std = np.sqrt(np.array([(w * (A.ravel() - mean[0]) ** 2).sum(), (w * (B.ravel() - mean[1]) ** 2).sum()]) / w.sum())
    np.testing.assert_allclose(PE.mcmc_chain_moments['mean'], mean, atol=0.02 * std.max())
    np.testing.assert_allclose(PE.mcmc_chain_moments['std'], std, rtol=0.02)
    np.testing.assert_allclose(np.mean(PE.post_burn_in_samples, axis=0), PE.mcmc_chain_moments['mean'], rtol=1e-10)


def _oracle_grid_stats(orc, P, bt, c, h, r, coef=2.0, nbins=10):
    N = orc.grid_size(r)
    th = orc.grid_decode(c, h, r, np.arange(N))
    lp = orc.log_p_batch(P, th, bt)[0]
    idx, best = orc.first_max_np(lp)
    w, thr = orc.permutation_weights(lp, best, coef)
    return th, lp, idx, best, w, orc.weighted_statistics(lp, th, w, nbins)


@pytest.mark.parametrize("radii", [[20, 300], [3, 1200], [150, 128]])
def test_grid_specialised_kernels_example00(env, radii):
    """Long last axis -> the specialised row/column-hoisted grid kernels (sweep_grid.cu); bit-exact logP,
    exact MAP index / weights / histograms, moments to rounding."""
    PEUQSE, orc, _, torch = env
    c, h = [200., 500.], [2.5, 0.4]
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                 settings={'multistart_initialPointsDistributionType': 'grid', 'multistart_gridsearchSamplingInterval': h,
                           'multistart_gridsearchSamplingRadii': radii})
    PE.doMultiStart()
    P = orc.build_problem([200, 500], [100, 200], Y00, S00)
    _, bt = orc.sim_quadratic(X00)
    th, lp, idx, best, w, st = _oracle_grid_stats(orc, P, bt, c, h, radii)
    got = PE.permutations_MAP_logP_and_parameters_values
    assert np.array_equal(got[:, 1:], th)
    assert np.array_equal(got[:, 0], lp)
    assert PE.map_index == idx and PE.map_logP == best
    assert PE.sweep_result.sum_weights == w.sum() and PE.sweep_result.n_retained == int((w > 0).sum())
    np.testing.assert_allclose(PE.mu_AP_parameter_set, st['mu_AP'], rtol=1e-12)
    np.testing.assert_allclose(PE.stdap_parameter_set, st['stdap'], rtol=1e-10)
    np.testing.assert_allclose(PE.evidence, st['evidence'], rtol=1e-11)
    assert np.array_equal(PE.marginal_histograms[0], st['hist'].astype(np.int64))
    np.testing.assert_allclose(PE.marginal_histograms[1], st['edges'], rtol=1e-15)
    # sharded exactly like ranks would: three index ranges, merged with the rule of distributed.merge_best
    from peuqse_b200.sweep import Sweep, make_spec
    from peuqse_b200 import distributed as D
    spec = make_spec("grid", 2, c, interval=h, radii=radii)
    N = orc.grid_size(radii)
    bests, accs, hists = [], [], []
    log_rel = float(np.log(10 ** -2.0))
    for rnk in range(3):
        sw = Sweep(PE.device_problem, spec, N, shard=(rnk, 3))
        sw.pass1()
        b = sw.best.cpu()
        bests.append((float(b[0]), int(b.view(torch.int64)[1])))
        assert np.array_equal(sw.logp.cpu().numpy(), lp[sw.first:sw.first + sw.count])
    gb = D.merge_best(bests)
    assert gb == (best, idx)
    # passes 2 and 3 per shard with the merged best record, reduced like the collectives would
    from peuqse_b200 import _lib
    lay = _lib.acc_layout(2)
    tot = None
    hist = np.zeros((2, 10), dtype=np.int64)
    sws = []
    for rnk in range(3):
        sw = Sweep(PE.device_problem, spec, N, shard=(rnk, 3))
        sw.pass1()
        sw.best[0] = gb[0]
        sw.best[1:2].view(torch.int64)[0] = gb[1]
        sw.pass2(log_rel)
        a = sw.acc.cpu().numpy()
        if tot is None:
            tot = a.copy()
        else:
            tot[:lay['MIN']] += a[:lay['MIN']]
            tot[lay['MIN']:lay['MIN'] + 2] = np.minimum(tot[lay['MIN']:lay['MIN'] + 2], a[lay['MIN']:lay['MIN'] + 2])
            tot[lay['MAX']:lay['MAX'] + 2] = np.maximum(tot[lay['MAX']:lay['MAX'] + 2], a[lay['MAX']:lay['MAX'] + 2])
        sws.append(sw)
    assert tot[lay['W']] == w.sum() and tot[lay['NKEPT']] == int((w > 0).sum())
    np.testing.assert_allclose(th[idx] + tot[lay['S1']:lay['S1'] + 2] / tot[lay['W']], st['mu_AP'], rtol=1e-12)
    for sw in sws:                         # every rank bins with the globally reduced min/max
        sw.acc.copy_(torch.from_numpy(tot).cuda())
        sw.pass3(log_rel, reuse_pass2_sums=True)
        hist += sw.hist.cpu().numpy().reshape(2, 10)
    assert np.array_equal(hist, st['hist'].astype(np.int64))
    # stand-alone histogram pass (no fused sums) gives the same counts
    sw = sws[0]
    sw.pass3(log_rel, reuse_pass2_sums=False)
    h0 = sw.hist.cpu().numpy().reshape(2, 10).copy()
    sw.pass2(log_rel)
    sw.acc.copy_(torch.from_numpy(tot).cuda())
    sw.pass3(log_rel, reuse_pass2_sums=True)
    assert np.array_equal(h0, sw.hist.cpu().numpy().reshape(2, 10))


def test_grid_specialised_linear_k3(env):
    PEUQSE, orc, _, _ = env
    rng = np.random.default_rng(11)
    k, n = 3, 5
    Phi = rng.standard_normal((n, k))
    truth = np.array([0.3, -0.2, 0.5])
    sig = 0.2 + 0.1 * rng.random(n)
    y = Phi @ truth + sig * rng.standard_normal(n)
    c, h, r = [0.0, 0.1, -0.1], [0.05, 0.07, 0.004], [6, 5, 400]
    PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), list(y), list(sig), [0.0] * k, [1.0, 2.0, 0.5],
                 settings={'multistart_initialPointsDistributionType': 'grid', 'multistart_gridsearchSamplingInterval': h,
                           'multistart_gridsearchSamplingRadii': r},
                 InputParameterInitialGuess=c, InputParameterPriorValues_upperBounds=[None, 0.4, None],
                 InputParameterPriorValues_lowerBounds=[-0.2, None, None])
    PE.doMultiStart()
    P = orc.build_problem([0.0] * k, [1.0, 2.0, 0.5], y, sig, upper_bounds=[None, 0.4, None], lower_bounds=[-0.2, None, None])
    _, bt = orc.sim_linear(Phi)
    th, lp, idx, best, w, st = _oracle_grid_stats(orc, P, bt, c, h, r)
    got = PE.permutations_MAP_logP_and_parameters_values
    assert np.array_equal(got[:, 1:], th)
    assert np.array_equal(got[:, 0], lp)
    assert np.isneginf(lp).any() and PE.map_index == idx and PE.map_logP == best
    assert PE.sweep_result.sum_weights == w.sum()
    np.testing.assert_allclose(PE.mu_AP_parameter_set, st['mu_AP'], rtol=1e-11)
    np.testing.assert_allclose(PE.stdap_parameter_set, st['stdap'], rtol=1e-9)
    assert np.array_equal(PE.marginal_histograms[0], st['hist'].astype(np.int64))


def _linear_gaussian_case(orc, k=4, n=30, sig=0.3, seed=3):
    rng = np.random.default_rng(seed)
    Phi = orc.cos_basis(n, k)
    truth = rng.standard_normal(k)
    y = Phi @ truth + sig * rng.standard_normal(n)
    mean, cov = orc.linear_gaussian_posterior(Phi, Phi / sig ** 2, y, 4.0)
    return Phi, y, mean, cov


def test_ess_kernel_linear_gaussian(env):
    """Ensemble slice sampling kernels (parity unpinned: zeus is not available) against the closed-form posterior of
    a linear-Gaussian problem, within Monte-Carlo error."""
    PEUQSE, orc, _, torch = env
    k, n, sig = 4, 30, 0.3
    Phi, y, mean, cov = _linear_gaussian_case(orc, k, n, sig)
    PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), list(y), list(np.full(n, sig)), [0.0] * k, [2.0] * k,
                 settings={'mcmc_length': 64 * 1500, 'mcmc_nwalkers': 64, 'mcmc_burn_in': 64 * 300, 'mcmc_random_seed': 11,
                           'mcmc_threshold_filter_samples': False, 'mcmc_info_gain_returned': 'log_ratio'})
    res = PE.doEnsembleSliceSampling()
    S = PE.post_burn_in_samples
    assert S.shape == (64 * 1200, k) and len(res) == 7
    assert PE.discrete_chains_post_burn_in_samples.shape == (1200, 64, k)
    ess_n = len(S) / 10.0                                   # generous autocorrelation allowance
    se = np.sqrt(np.diag(cov) / ess_n)
    assert np.all(np.abs(S.mean(axis=0) - mean) < 5 * se), (S.mean(axis=0), mean, se)
    d = np.sqrt(np.diag(cov))
    np.testing.assert_allclose(np.cov(S.T) / np.outer(d, d), cov / np.outer(d, d), atol=0.08)
    # stored logP of the chain is the logP of the stored positions
    lp = PE.getLogP_batch(S[:500]).cpu().numpy()
    np.testing.assert_allclose(lp, PE.post_burn_in_log_posteriors_un_normed_vec[:500, 0], rtol=1e-13)
    assert PE.ess_info['n_logp_evaluations'] > 64 * 1500 and 0.05 < PE.ess_info['mu'] < 50
    # same seed -> same chain (counter-based RNG), different seed -> different chain
    from peuqse_b200.mcmc import ess_launch
    start = PE.discrete_chains_post_burn_in_samples[-1]
    a = ess_launch(PE, start, 20, 5)['chain'].cpu().numpy()
    b = ess_launch(PE, start, 20, 5)['chain'].cpu().numpy()
    c = ess_launch(PE, start, 20, 6)['chain'].cpu().numpy()
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_ess_batched_full_covariance(env):
    """Round-synchronous ESS driver + full-covariance contraction (the C4 path at a small size): two responses with
    dense covariance, posterior mean/cov against the closed form."""
    PEUQSE, orc, _, torch = env
    rng = np.random.default_rng(21)
    k, n = 3, 24
    Phi = orc.cos_basis(2 * n, k)
    x = np.linspace(0, 1, n)
    Sig = 0.1 ** 2 * np.exp(-(x[:, None] - x[None, :]) ** 2 / (2 * 0.05 ** 2)) + 0.05 ** 2 * np.eye(n)
    truth = rng.standard_normal(k)
    Lc = np.linalg.cholesky(Sig)
    y = Phi @ truth + np.concatenate([Lc @ rng.standard_normal(n), Lc @ rng.standard_normal(n)])
    Sinv = np.linalg.inv(Sig)
    big = np.block([[Sinv, np.zeros((n, n))], [np.zeros((n, n)), Sinv]])
    mean, cov = orc.linear_gaussian_posterior(Phi, big @ Phi, y, 4.0)
    PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), [list(y[:n]), list(y[n:])], [Sig, Sig], [0.0] * k, [2.0] * k,
                 settings={'mcmc_length': 32 * 700, 'mcmc_nwalkers': 32, 'mcmc_burn_in': 32 * 200, 'mcmc_random_seed': 4,
                           'mcmc_threshold_filter_samples': False, 'mcmc_info_gain_returned': 'log_ratio'})
    # per-point parity of the full-covariance likelihood against scipy's logpdf (what IP.py:2960 computes)
    from scipy.stats import multivariate_normal as mvn
    th = truth + 0.1 * rng.standard_normal((20, k))
    want = np.array([mvn.logpdf(x=y[:n], mean=(Phi @ t)[:n], cov=Sig) + mvn.logpdf(x=y[n:], mean=(Phi @ t)[n:], cov=Sig) for t in th])
    _, _, ll = parts(PE, th)
    np.testing.assert_allclose(ll, want, rtol=1e-12)
    PE.doEnsembleSliceSampling()
    S = PE.post_burn_in_samples
    se = np.sqrt(np.diag(cov) / (len(S) / 10.0))
    assert np.all(np.abs(S.mean(axis=0) - mean) < 5 * se)
    d = np.sqrt(np.diag(cov))
    np.testing.assert_allclose(np.cov(S.T) / np.outer(d, d), cov / np.outer(d, d), atol=0.12)


def test_ess_warp_shape_tpd(env, golden):
    PEUQSE, orc, _, _ = env
    g = golden("tpd")
    n = 1024
    PE = make_pe(PEUQSE, PEUQSE.Tpd2SiteModel.linear_ramp(n), list(g['tpd1024_y']), list(np.full(n, float(g['tpd1024_sigma']))),
                 [41.5, 41.5, 13.0, 13.0, 0.1, 0.1], [20, 20, 13, 13, 0.1, 0.1])
    from peuqse_b200.mcmc import ess_launch
    truth = np.array([41.5, 45, 13, 13, 0.1, 0.1])
    start = truth + 1e-3 * np.array([20, 20, 13, 13, 0.1, 0.1]) * np.random.default_rng(2).standard_normal((24, 6))
    out = ess_launch(PE, start, 30, 9)
    Z = out['chain_logp'].cpu().numpy()
    assert np.all(np.isfinite(Z)) and Z[-1].mean() > Z[0].mean() - 5
    lp = PE.getLogP_batch(out['chain'][-1]).cpu().numpy()
    np.testing.assert_allclose(lp, Z[-1], rtol=1e-12)


def test_fullcov_dmma_contraction(env):
    """DMMA quadratic form at sizes that are not multiples of the tile (n_r = 520, B = 130), two responses with dense
    covariances (the reference's working full-covariance route), against scipy's logpdf per point and the
    oracle batch; also through a torch-callable simulator."""
    PEUQSE, orc, _, torch = env
    from scipy.stats import multivariate_normal as mvn
    rng = np.random.default_rng(31)
    k, n = 5, 520
    x = np.linspace(0, 1, n)
    Phi = orc.cos_basis(2 * n, k)
    covs = []
    for ell in (0.02, 0.05):
        covs.append(0.1 ** 2 * np.exp(-(x[:, None] - x[None, :]) ** 2 / (2 * ell ** 2)) + 0.05 ** 2 * np.eye(n))
    truth = rng.standard_normal(k)
    y = Phi @ truth + 0.1 * rng.standard_normal(2 * n)
    B = 130
    th = truth + 0.05 * rng.standard_normal((B, k))
    P = orc.build_problem([0.0] * k, [2.0] * k, [y[:n], y[n:]], covs, max_covmat_size=600)
    _, bt = orc.sim_linear(Phi)
    want = orc.log_p_batch(P, th, lambda t: [bt(t)[:, :n], bt(t)[:, n:]])
    Phi_t = torch.from_numpy(Phi).cuda()
    for model in (PEUQSE.LinearModel(Phi), PEUQSE.TorchSimulator(lambda t, Pm=Phi_t: t @ Pm.T)):
        PE = make_pe(PEUQSE, model, [list(y[:n]), list(y[n:])], covs, [0.0] * k, [2.0] * k,
                     settings={'responses_observed_max_covmat_size': 600})
        lp, pr, ll = parts(PE, th)
        np.testing.assert_allclose(ll, want[2], rtol=1e-12)
        np.testing.assert_allclose(lp, want[0], rtol=1e-12)
        assert np.array_equal(pr, want[1])
    for i in (0, 7, 129):
        s = Phi @ th[i]
        ref = mvn.logpdf(x=y[:n], mean=s[:n], cov=covs[0]) + mvn.logpdf(x=y[n:], mean=s[n:], cov=covs[1])
        np.testing.assert_allclose(ll[i], ref, rtol=1e-12)
    # NaN parameter: logLik = -inf (NaN simulated values, IP.py:2880-2882); logP = -inf + NaN prior = NaN as in the reference
    bad = th.copy()
    bad[3, 0] = np.nan
    lpb, _, llb = parts(PE, bad)
    assert llb[3] == float('-inf') and np.isnan(lpb[3]) and np.array_equal(np.delete(lpb, 3), np.delete(lp, 3))


def test_c2_full_size_properties(env):
    """BASELINE config C2 at full size (100 020 001 grid points): size-independent properties + spot checks."""
    PEUQSE, orc, _, torch = env
    c, h, r = [200., 500.], [0.08, 0.16], [5000, 5000]
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                 settings={'multistart_initialPointsDistributionType': 'grid', 'multistart_gridsearchSamplingInterval': h,
                           'multistart_gridsearchSamplingRadii': r})
    PE.doMultiStart()
    N = orc.grid_size(r)
    assert N == 100020001 and PE.sweep_result.total == N
    P = orc.build_problem([200, 500], [100, 200], Y00, S00)
    _, bt = orc.sim_quadratic(X00)
    sw = PE._sweep
    # (i) the MAP point: decoded parameters and logP are the oracle's, bit for bit
    th_map = orc.grid_decode(c, h, r, np.array([PE.map_index]))[0]
    assert np.array_equal(th_map, PE.map_parameter_set)
    assert orc.log_p_batch(P, th_map[None, :], bt)[0][0] == PE.map_logP
    # (ii) 200 000 random indices (plus both ends) of the stored logP are bit-exact
    rng = np.random.default_rng(0)
    idx = np.unique(np.concatenate([rng.integers(0, N, 200000), [0, 1, 10000, 10001, 10002, N - 1]]))
    got = sw.logp[torch.from_numpy(idx).cuda()].cpu().numpy()
    assert np.array_equal(got, orc.log_p_batch(P, orc.grid_decode(c, h, r, idx), bt)[0])
    # (iii) the maximum of the stored array is the reported MAP, at its first occurrence
    m = sw.logp.max()
    assert float(m.item()) == PE.map_logP
    assert int(torch.nonzero(sw.logp == m)[0].item()) == PE.map_index
    # (iv) weights: retained count and weight sum recomputed from the stored logP
    thr = PE.sweep_result.threshold
    kept = sw.logp > thr
    assert int(kept.sum().item()) == PE.sweep_result.n_retained
    wsum = float(torch.floor(torch.exp(sw.logp[kept] - thr)).sum().item())
    assert abs(wsum - PE.sweep_result.sum_weights) <= 4          # two exp implementations may differ on a few exact-integer edges
    # (v) every marginal histogram holds all the weight; edges span the retained range
    H, E = PE.marginal_histograms
    assert np.all(H.sum(axis=1) == int(PE.sweep_result.sum_weights))
    assert np.array_equal(E[:, 0], PE.sweep_result.retained_min) and np.array_equal(E[:, -1], PE.sweep_result.retained_max)
    # (vi) moments against a float64 torch reduction over the stored array (independent code path)
    ii = torch.nonzero(kept)[:, 0]
    th = sw.points(0, N)[ii]
    w = torch.floor(torch.exp(sw.logp[ii] - thr))
    mu = (w[:, None] * th).sum(0) / w.sum()
    sd = torch.sqrt((w[:, None] * (th - mu) ** 2).sum(0) / w.sum())
    np.testing.assert_allclose(PE.mu_AP_parameter_set, mu.cpu().numpy(), rtol=1e-9)
    np.testing.assert_allclose(PE.stdap_parameter_set, sd.cpu().numpy(), rtol=1e-8)
    # (vii) two-rank style sharding at full size gives the same best record
    from peuqse_b200.sweep import Sweep, make_spec
    from peuqse_b200 import distributed as D
    spec = make_spec("grid", 2, c, interval=h, radii=r)
    bests = []
    for rnk in range(2):
        s2 = Sweep(PE.device_problem, spec, N, store_logp=False, shard=(rnk, 2))
        s2.pass1()
        b = s2.best.cpu()
        bests.append((float(b[0]), int(b.view(torch.int64)[1])))
        del s2
    assert D.merge_best(bests) == (PE.map_logP, PE.map_index)


def test_c5_uniform_sweep_properties(env):
    """Uniform (Philox) sweep of the C5-shaped problem at 2^24 points: index-addressed draws make any sharding give the
    same points and the same best; spot-checked against the oracle."""
    PEUQSE, orc, philox_np, torch = env
    k, n, N = 10, 64, 1 << 24
    Phi = orc.cos_basis(n, k)
    truth = np.random.default_rng(5).standard_normal(k) * 0.5
    y = Phi @ truth + 0.1 * np.random.default_rng(6).standard_normal(n)
    PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), list(y), list(np.full(n, 0.1)), [0.0] * k, [1.0] * k,
                 settings={'multistart_initialPointsDistributionType': 'uniform', 'multistart_numStartPoints': N,
                           'multistart_random_seed': 7})
    PE.doMultiStart()
    sw = PE._sweep
    P = orc.build_problem([0.0] * k, [1.0] * k, y, np.full(n, 0.1))
    _, bt = orc.sim_linear(Phi)
    idx = np.array([0, 1, 2, 12345, 99999, N // 2, N - 2, N - 1, PE.map_index])
    u = np.vstack([philox_np.sweep_uniforms(7, int(i), 1, k) for i in idx])
    th = orc.uniform_points(u, 0.866, np.ones(k), np.zeros(k))
    got_th = np.vstack([sw.points(int(i), 1).cpu().numpy() for i in idx])
    assert np.array_equal(got_th, th)
    assert np.array_equal(sw.logp[torch.from_numpy(idx).cuda()].cpu().numpy(), orc.log_p_batch(P, th, bt)[0])
    assert float(sw.logp.max().item()) == PE.map_logP and np.array_equal(PE.map_parameter_set, th[-1])
    from peuqse_b200.sweep import Sweep
    from peuqse_b200 import distributed as D
    bests = []
    for rnk in range(3):
        s2 = Sweep(PE.device_problem, sw.spec, N, store_logp=False, shard=(rnk, 3))
        s2.pass1()
        b = s2.best.cpu()
        bests.append((float(b[0]), int(b.view(torch.int64)[1])))
    # (sweeps that do not keep logP simulate a linear model on the tensor pipe: another summation order, values to 1e-12)
    mb = D.merge_best(bests)
    assert mb[1] == PE.map_index and abs(mb[0] - PE.map_logP) <= 1e-12 * abs(PE.map_logP)
    assert PE.sweep_result.sum_weights >= 100 - 1


@pytest.mark.parametrize("case", ["sparse", "dense"])
def test_uniform_sweep_without_stored_logp(env, case):
    """multistart_store_logP=False: pass 1 keeps a shortlist of near-maximal points and passes 2-3 read it ("sparse":
    10-D, a handful of retained points); when more than the shortlist's sort capacity are retained ("dense": 2-D, broad
    posterior) the recomputing kernels take over.  Either way the results must be those of the stored-logP sweep."""
    PEUQSE, orc, philox_np, torch = env
    from peuqse_b200.sweep import Sweep
    if case == "sparse":
        k, n, N = 10, 64, 1 << 22
        Phi = orc.cos_basis(n, k)
        truth = np.random.default_rng(5).standard_normal(k) * 0.5
        y = Phi @ truth + 0.1 * np.random.default_rng(6).standard_normal(n)
        PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), list(y), list(np.full(n, 0.1)), [0.0] * k, [1.0] * k,
                     settings={'multistart_initialPointsDistributionType': 'uniform', 'multistart_numStartPoints': N,
                               'multistart_random_seed': 3})
    else:
        k, N = 2, 1 << 20
        PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                     settings={'multistart_initialPointsDistributionType': 'uniform', 'multistart_numStartPoints': N,
                               'multistart_random_seed': 3})
    PE.doMultiStart()
    A = PE.sweep_result
    spec = PE._sweep.spec
    for shard in (None, (1, 3)):
        ref = Sweep(PE.device_problem, spec, N, store_logp=True, shard=shard)
        new = Sweep(PE.device_problem, spec, N, store_logp=False, shard=shard)
        ref.use_graph = new.use_graph = False
        Rr, Rn = ref.run(2.0), new.run(2.0)
        # the no-store sweep of a linear model simulates on the tensor pipe (csrc/sweep.cu: k_sweep_pass1_lin_dmma): another
        # summation order than the reference's, values within the 1e-12 budget; every other model: the same bits
        assert Rn.map_index == Rr.map_index
        if case == "sparse":
            assert abs(Rn.map_logP - Rr.map_logP) <= 1e-12 * abs(Rr.map_logP)
        else:
            assert Rn.map_logP == Rr.map_logP
        assert Rn.n_retained == Rr.n_retained and Rn.sum_weights == Rr.sum_weights
        assert (Rn.n_retained > 2048) == (case == "dense")
        assert np.array_equal(Rn.histogram_counts, Rr.histogram_counts)
        assert np.array_equal(Rn.histogram_edges, Rr.histogram_edges)
        np.testing.assert_allclose(Rn.mu_AP_parameter_set, Rr.mu_AP_parameter_set, rtol=1e-13)
        np.testing.assert_allclose(Rn.stdap_parameter_set, Rr.stdap_parameter_set, rtol=1e-10)
        np.testing.assert_allclose(Rn.evidence, Rr.evidence, rtol=1e-12)
        if shard is None:
            assert Rn.map_index == A.map_index and abs(Rn.map_logP - A.map_logP) <= 1e-12 * abs(A.map_logP)
        # a second run on the same work buffer (header and list are rebuilt by pass 1)
        Rn2 = new.run(2.0)
        assert Rn2.n_retained == Rr.n_retained and np.array_equal(Rn2.histogram_counts, Rr.histogram_counts)
        np.testing.assert_array_equal(Rn2.mu_AP_parameter_set, Rn.mu_AP_parameter_set)
    # a threshold wider than the shortlist margin (coef 25 -> 57.6 nats) must fall back to recomputation
    ref = Sweep(PE.device_problem, spec, N, store_logp=True)
    new = Sweep(PE.device_problem, spec, N, store_logp=False)
    ref.use_graph = new.use_graph = False
    Rr, Rn = ref.run(25.0), new.run(25.0)
    assert Rn.n_retained == Rr.n_retained
    np.testing.assert_allclose(Rn.mu_AP_parameter_set, Rr.mu_AP_parameter_set, rtol=1e-12)


def test_mh_kernel_shapes_agree(env):
    """The two Metropolis-Hastings kernel shapes for small n -- one warp per chain with thread-order evaluation
    (<= 8192 chains) and one thread per chain (more) -- must produce the same chains bit for bit: draws are
    addressed by (seed, chain, step) and every value goes through the same operations in the same order."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200])
    rng = np.random.default_rng(11)
    th0 = np.array([200.0, 500.0]) + rng.standard_normal((8200, 2)) * [30.0, 60.0]
    L = step_cholesky(PE.Q_covmat, 0.3)
    few = mh_launch(PE, th0[:300], 400, 100, 5, L, thin=3, export_draws=True)
    many = mh_launch(PE, th0, 400, 100, 5, L, thin=3, export_draws=True)
    for key in ('samples', 'logp', 'logprior', 'loglik', 'deltas', 'us'):
        a, b = few[key].cpu().numpy(), many[key].cpu().numpy()[:, :300]
        assert np.array_equal(a, b), key
    assert np.array_equal(few['stats'].cpu().numpy(), many['stats'].cpu().numpy()[:300])
    # k = 10 linear model: more than 32 moment slots per chain (two per lane)
    k, n = 10, 64
    Phi = orc.cos_basis(n, k)
    y = Phi @ (0.5 * rng.standard_normal(k)) + 0.1 * rng.standard_normal(n)
    PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), list(y), list(np.full(n, 0.1)), [0.0] * k, [1.0] * k)
    th0 = 0.1 * rng.standard_normal((8200, k))
    L = step_cholesky(PE.Q_covmat, 1e-3)
    few = mh_launch(PE, th0[:200], 150, 30, 9, L, store_samples=False)
    many = mh_launch(PE, th0, 150, 30, 9, L, store_samples=False)
    assert np.array_equal(few['stats'].cpu().numpy(), many['stats'].cpu().numpy()[:200])


def test_grid_sweep_with_unaligned_logp_buffer(env):
    """d_logp only 8-byte aligned: pass 2 cannot use the bulk-copy (TMA) kernel and takes the plain fused one (48 KB of
    dynamic shared memory next to its static arrays); results must be those of the aligned run."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200.sweep import Sweep, make_spec
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200])
    spec = make_spec("grid", 2, [200., 500.], interval=[2.5, 0.4], radii=[37, 2900])
    N = 75 * 5801
    ref = Sweep(PE.device_problem, spec, N)
    ref.use_graph = False
    Rr = ref.run(2.0)
    sw = Sweep(PE.device_problem, spec, N)
    sw.use_graph = False
    big = torch.empty(N + 1, dtype=torch.float64, device='cuda')
    sw.logp = big[1:]
    sw.logps = [sw.logp]
    R = sw.run(2.0)
    assert (R.map_index, R.map_logP, R.sum_weights, R.n_retained) == (Rr.map_index, Rr.map_logP, Rr.sum_weights, Rr.n_retained)
    assert np.array_equal(R.histogram_counts, Rr.histogram_counts)
    assert torch.equal(sw.logp, ref.logp)
    np.testing.assert_allclose(R.mu_AP_parameter_set, Rr.mu_AP_parameter_set, rtol=1e-13)
    np.testing.assert_allclose(R.stdap_parameter_set, Rr.stdap_parameter_set, rtol=1e-11)


# ----------------------------------------------------------------------------------------------
# post-burn-in statistics on the device (csrc/stats.cu)
# ----------------------------------------------------------------------------------------------
PBS_CASES = [('auto', True, 'auto', 'auto', 'pbs_logP'), ('auto1', True, 1.0, 'auto', 'pbs_logP'), ('MAP', True, 2.0, 'MAP', 'pbs_logP'),
             ('mu_AP', True, 2.0, 'mu_AP', 'pbs_logP'), ('off', False, 2.0, 'auto', 'pbs_logP'), ('odd', True, 2.0, 'auto', 'pbs_logP_odd')]


@pytest.mark.parametrize("case", PBS_CASES, ids=[c[0] for c in PBS_CASES])
def test_device_statistics_vs_reference_fixture(env, golden, case):
    """calculatePostBurnInStatistics of the unmodified reference on stored arrays vs the device passes: the kept rows,
    the MAP row and its index are exact; means / std / evidence / info gain within 1e-12 relative (summation order)."""
    PEUQSE, orc, _, torch = env
    name, on, coef, bench, key = case
    g = golden("post_burn_in_statistics")
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                 settings={'mcmc_threshold_filter_samples': on, 'mcmc_threshold_filter_coefficient': coef,
                           'mcmc_threshold_filter_benchmark': bench, 'mcmc_info_gain_returned': 'log_ratio'})
    PE.post_burn_in_samples = g['pbs_samples'].copy()
    PE.post_burn_in_log_posteriors_un_normed_vec = g[key].copy()
    PE.post_burn_in_log_priors_vec = g['pbs_log_priors'].copy()
    res = PE.calculatePostBurnInStatistics()
    rows = g[f'pbs_{name}_kept_rows']
    assert np.array_equal(PE.post_burn_in_samples, g['pbs_samples'][rows])
    assert np.array_equal(PE.post_burn_in_log_posteriors_un_normed_vec, g[key][rows], equal_nan=True)
    assert np.array_equal(PE.post_burn_in_log_priors_vec, g['pbs_log_priors'][rows])
    assert float(PE.map_logP[0]) == float(g[f'pbs_{name}_map_logP']) and PE.map_index == int(g[f'pbs_{name}_map_index'])
    assert np.array_equal(PE.map_parameter_set, g[f'pbs_{name}_map_theta'])
    np.testing.assert_allclose(PE.mu_AP_parameter_set, g[f'pbs_{name}_mu_AP'], rtol=1e-12)
    np.testing.assert_allclose(PE.stdap_parameter_set, g[f'pbs_{name}_stdap'], rtol=1e-12)
    np.testing.assert_allclose(PE.evidence, float(g[f'pbs_{name}_evidence']), rtol=1e-12)
    np.testing.assert_allclose(PE.info_gain, float(g[f'pbs_{name}_info_gain']), rtol=1e-11)
    assert len(res) == 7 and res[5] is PE.post_burn_in_samples


@pytest.mark.parametrize("k,N", [(6, 1_000_003), (3, 70_001), (1, 513), (16, 40_000)])
def test_device_statistics_vs_oracle_large(env, k, N):
    """Seeded heavy-tailed samples: all three benchmarks, KL-divergence info gain against the numpy restatement
    (histogram counts are integers, so the KL value only differs by the final 100-term sums)."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200 import stats as ST, _lib
    rng = np.random.default_rng(100 + k)
    th = rng.standard_normal((N, k)) * (1.0 + np.arange(k)) + 3.0
    th[rng.random(N) < 0.02] *= 4.0
    logp = -0.5 * np.sum(((th - 3.0) / (1.0 + np.arange(k))) ** 2, axis=1) - 7.0
    logpr = -0.5 * np.sum((th / 10.0) ** 2, axis=1) - 2.0
    prior = rng.standard_normal((5000, k)) * 10.0
    d = lambda a: torch.as_tensor(a, dtype=torch.float64).cuda()
    for mode, bench in ((_lib.FILTER_AUTO, 'auto'), (_lib.FILTER_MAP, 'MAP'), (_lib.FILTER_NONE, None)):
        R = ST.device_statistics(d(th), d(logp), d(logpr), mode, 2.0, prior_samples=prior)
        with np.errstate(all='ignore'):
            want = orc.post_burn_in_statistics_filtered(th, logp, logpr, benchmark=bench or 'auto', coef=2.0,
                                                        filter_samples=bench is not None, prior_samples=prior)
        assert R.n_kept == len(want['samples']) and R.n_total == N
        ks, kl, kp = R.kept_rows()
        assert np.array_equal(ks.cpu().numpy(), want['samples']) and np.array_equal(kl.cpu().numpy(), want['logP'])
        assert np.array_equal(kp.cpu().numpy(), want['log_priors'])
        assert R.map_logP == float(want['map_logP']) and R.map_index == want['map_index']
        assert np.array_equal(R.map_parameter_set, want['map_parameter_set'])
        np.testing.assert_allclose(R.mu_AP_parameter_set, want['mu_AP'], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(R.stdap_parameter_set, want['stdap'], rtol=1e-12)
        np.testing.assert_allclose(R.evidence, want['evidence'], rtol=1e-12)
        np.testing.assert_allclose(R.info_gain_log_ratio, want['info_gain_log_ratio'], rtol=1e-11)
        np.testing.assert_allclose(R.info_gain_KL_each_parameter, want['info_gain_KL_each_parameter'], rtol=1e-11, atol=1e-14)
        np.testing.assert_allclose(R.info_gain_KL, want['info_gain_KL'], rtol=1e-11)


def test_mh_with_default_tail_filter(env):
    """doMetropolisHastings with the reference's default settings (filter on, benchmark 'auto', KL-divergence info gain):
    the returned arrays are the oracle-filtered rows of the chain the kernel produced."""
    PEUQSE, orc, _, torch = env
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                 settings={'mcmc_length': 6000, 'mcmc_burn_in': 500, 'mcmc_relative_step_length': 0.3, 'mcmc_random_seed': 3,
                           'mcmc_nchains': 4})
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    raw = mh_launch(PE, np.tile([200., 500.], (4, 1)), 6000, 500, 3, step_cholesky(PE.Q_covmat, 0.3))
    S = raw['samples'][500:].permute(1, 0, 2).reshape(-1, 2).cpu().numpy()
    lp = raw['logp'][500:].permute(1, 0).reshape(-1, 1).cpu().numpy()
    pr = raw['logprior'][500:].permute(1, 0).reshape(-1, 1).cpu().numpy()
    res = PE.doMetropolisHastings()
    with np.errstate(all='ignore'):
        want = orc.post_burn_in_statistics_filtered(S, lp, pr, benchmark='auto', coef=2.0, prior_samples=PE.samples_of_prior)
    assert 0 < len(want['samples']) < len(S)
    assert np.array_equal(PE.post_burn_in_samples, want['samples']) and np.array_equal(res[6], want['logP'])
    assert np.array_equal(PE.post_burn_in_log_priors_vec, want['log_priors'])
    assert PE.map_index == want['map_index'] and float(PE.map_logP[0]) == float(want['map_logP'])
    np.testing.assert_allclose(res[1], want['mu_AP'], rtol=1e-12)
    np.testing.assert_allclose(res[2], want['stdap'], rtol=1e-12)
    np.testing.assert_allclose(res[3], want['evidence'], rtol=1e-12)
    np.testing.assert_allclose(res[4], want['info_gain_KL'], rtol=1e-11)
    assert PE.post_burn_in_log_likelihoods_vec.shape == (len(S), 1)     # the reference leaves this array unfiltered (IP.py:2726)


def test_device_statistics_full_size_properties(env):
    """C3-sized slice (4e7 rows x 6 parameters = 2.6 GB): filtering then taking statistics of the kept rows without a
    filter gives the same numbers; the kept count matches the cut; traffic-bound passes finish in milliseconds."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200 import stats as ST, _lib
    N, k = 40_000_000, 6
    gen = torch.Generator(device="cuda").manual_seed(5)
    th = torch.randn((N, k), dtype=torch.float64, device="cuda", generator=gen)
    logp = -0.5 * (th * th).sum(dim=1) - 5.0
    logpr = -0.02 * (th * th).sum(dim=1)
    R = ST.device_statistics(th, logp, logpr, _lib.FILTER_AUTO, 2.0)
    assert 0 < R.n_kept < N
    assert int((~(logp < R.cut)).sum().item()) == R.n_kept
    ks, kl, kp = R.kept_rows()
    R2 = ST.device_statistics(ks, kl, kp, _lib.FILTER_NONE, 2.0)
    assert R2.n_kept == R.n_kept and R2.map_logP == R.map_logP and R2.map_index == R.map_index
    np.testing.assert_allclose(R2.mu_AP_parameter_set, R.mu_AP_parameter_set, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(R2.stdap_parameter_set, R.stdap_parameter_set, rtol=1e-12)
    np.testing.assert_allclose(R2.evidence, R.evidence, rtol=1e-12)
    np.testing.assert_allclose(R2.info_gain_log_ratio, R.info_gain_log_ratio, rtol=1e-11)
    np.testing.assert_allclose(R.mu_AP_parameter_set, ks.mean(dim=0).cpu().numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(R.stdap_parameter_set, ks.std(dim=0, unbiased=False).cpu().numpy(), rtol=1e-10)


def test_on_disk_formats_round_trip(env, golden):
    """the samplers / multistart write the reference's files (byte-level format pinned on the CPU in
    tests/test_host_cpu.py); here the files of real device runs are read back by reload_samples and by the reference's
    own readers (np.genfromtxt / pickle)."""
    import os
    import pickle
    PEUQSE, orc, _, _ = env
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                 settings={'mcmc_length': 800, 'mcmc_burn_in': 100, 'mcmc_relative_step_length': 0.1, 'mcmc_random_seed': 0,
                           'mcmc_info_gain_returned': 'log_ratio', 'responses_abscissa': X00})
    PE.UserInput.model['exportResponses'] = True
    res = PE.doMetropolisHastings()
    files = set(os.listdir('logs_and_csvs')) | set(os.listdir('pickles'))
    for f in ('mcmc_logP_and_parameter_samples.csv', 'mcmc_log_file.txt', 'mcmc_burn_in_logP_and_parameter_samples.csv',
              'mcmc_post_burn_in_statistics.pkl', 'mcmc_map_logP.pkl', 'mcmc_map_parameter_set.pkl', 'mcmc_initial_point_parameters.pkl',
              'mcmc_last_point_sampled.pkl', 'mcmc_logP_and_parameter_samples.pkl', 'ExportedResponsesData__map__ResponsesConsolidated.csv',
              'ExportedResponsesData__muap__ResponsesConsolidated.csv', 'ExportedResponsesData__prior__ResponsesConsolidated.csv'):
        assert f in files, f
    arr = np.genfromtxt('logs_and_csvs/mcmc_logP_and_parameter_samples.csv', delimiter=",")
    assert np.array_equal(arr[:, 1:], PE.post_burn_in_samples) and np.array_equal(arr[:, 0:1], PE.post_burn_in_log_posteriors_un_normed_vec)
    st = pickle.load(open('pickles/mcmc_post_burn_in_statistics.pkl', 'rb'))
    assert np.array_equal(st[0], res[0]) and np.array_equal(st[5], res[5]) and st[4] == res[4]
    sim = np.genfromtxt('logs_and_csvs/ExportedResponsesData__map__ResponsesConsolidated.csv', delimiter=",", skip_header=1)
    assert np.array_equal(sim[:, 0], X00) and np.array_equal(sim[:, 1], (np.array(X00) - res[0][0]) ** 2 + res[0][1])
    # the last point of this run, by file name, as the start of a new object (IP.py:355-380)
    PE4 = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200], InputParameterInitialGuess='mcmc_last_point_sampled.pkl')
    assert np.array_equal(PE4.UserInput.InputParameterInitialGuess, PE.mcmc_last_point_sampled)
    # a fresh object continues the exported run (mcmc_continueSampling = True reads the pickles, IP.py:2583-2596)
    PE5 = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                  settings={'mcmc_length': 300, 'mcmc_relative_step_length': 0.1, 'mcmc_random_seed': 0, 'mcmc_continueSampling': True,
                            'mcmc_threshold_filter_samples': False, 'mcmc_info_gain_returned': 'log_ratio', 'mcmc_exportLog': False})
    r5 = PE5.doMetropolisHastings()
    n_old = PE.post_burn_in_samples.shape[0]
    assert r5[5].shape == (n_old + 300, 2) and np.array_equal(r5[5][:n_old], PE.post_burn_in_samples)
    assert np.array_equal(r5[5][n_old], PE.post_burn_in_samples[-1])         # row 0 of the continuation is the last exported row
    # a fresh object rebuilds the statistics from the files (the filter already ran: the kept rows are what was written)
    PE2 = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                  settings={'mcmc_length': 800, 'mcmc_threshold_filter_samples': False, 'mcmc_info_gain_returned': 'log_ratio'})
    r2 = PE2.reload_samples('mcmc')
    assert np.array_equal(r2[0], res[0]) and np.array_equal(PE2.post_burn_in_samples, PE.post_burn_in_samples)
    np.testing.assert_allclose(r2[1], res[1], rtol=1e-13)
    np.testing.assert_allclose(r2[2], res[2], rtol=1e-12)
    np.testing.assert_allclose(r2[4], res[4], rtol=1e-12)
    # multistart: the reference's stored grid (10 201 points)
    g = golden("example00")
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                 settings={'multistart_initialPointsDistributionType': 'grid', 'multistart_gridsearchSamplingInterval': [5, 5],
                           'multistart_gridsearchSamplingRadii': [50, 50], 'responses_abscissa': X00})
    PE.doMultiStart()
    arr = np.genfromtxt('logs_and_csvs/multistart_MAP_logP_and_parameters_values.csv', delimiter=",")
    assert arr.shape == (10201, 3) and np.array_equal(arr, PE.permutations_MAP_logP_and_parameters_values)
    assert np.array_equal(pickle.load(open('pickles/multistart_MAP_logP_and_parameters_values.pkl', 'rb')), arr)
    pts = np.genfromtxt('logs_and_csvs/multistart_initial_points_parameters_values.csv', delimiter=",")
    assert np.array_equal(pts, arr[:, 1:])
    log = open('logs_and_csvs/multistart_log_file.txt').read()
    assert "highest_MAP_initial_point_index: 13" in log and "highest_MAP_logP: [-42.63739101]" in log
    PE3 = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200], settings={'mcmc_info_gain_returned': 'log_ratio'})
    r3 = PE3.reload_samples('multistart')
    assert PE3.post_burn_in_samples.shape[0] == int(g['ex00_grid_nrows_expanded'])
    np.testing.assert_allclose(r3[1], g['ex00_grid_mu_AP'], rtol=1e-12)
    np.testing.assert_allclose(r3[2], g['ex00_grid_stdap'], rtol=1e-10)


def test_geweke_on_device_vs_reference_fixture_and_oracle(env, golden):
    """Geweke z-scores from one streaming prefix-sum pass (csrc/diag.cu) against the reference's own
    outputs (tests/golden/geweke.npz) and the oracle on larger seeded chains.  Tolerance 1e-8 relative: slice statistics are
    differences of prefix sums instead of numpy's pairwise sums over each slice."""
    import warnings
    PEUQSE, orc, _, torch = env
    from peuqse_b200.diagnostics import geweke_z_scores, convergence_diagnostics
    g = golden("geweke")
    x = g['chains_a']
    z, windows, starts = geweke_z_scores(x)
    z = z.cpu().numpy()
    assert np.array_equal(windows, g['a_window_indices']) and z.shape == (3, 3, 20, 20)
    ref_full = g['single_chain_z']
    assert np.array_equal(starts, ref_full[:, 0].astype(int))
    np.testing.assert_allclose(z[0, 0, -1], ref_full[:, 1], rtol=1e-8)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        r = orc.geweke_convergence(x)
    fin = np.isfinite(r['z'])
    assert np.array_equal(np.isfinite(z), fin)
    np.testing.assert_allclose(z[fin], r['z'][fin], rtol=1e-8)
    import types
    pe = types.SimpleNamespace(gpu_launches=0)
    out = convergence_diagnostics(pe, x)
    np.testing.assert_allclose(out[4], g['a_final_combined'], rtol=1e-8, equal_nan=True)
    np.testing.assert_allclose(out[5], g['a_percent_outlier'], rtol=1e-15)
    assert np.array_equal(np.asarray(out[3]), g['a_window_indices'])
    assert pe.convergence['Geweke']['z_scores_each_parameter'].shape == (3, 20, 20)
    # many columns and several row segments; values with a large offset (the pivot keeps the prefix sums small)
    rng = np.random.default_rng(21)
    R, C, k = 6000, 40, 4
    e = rng.standard_normal((R, C, k))
    y = np.empty_like(e)
    y[0] = e[0]
    phi = np.array([0.2, 0.8, 0.95, 0.5])
    for t in range(1, R):
        y[t] = phi * y[t - 1] + e[t]
    y = y * np.array([1e-3, 1.0, 50.0, 1e4]) + np.array([7.0, -3e3, 1e5, 1e7])
    z2 = geweke_z_scores(torch.as_tensor(y).cuda())[0].cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for c, d in ((0, 0), (17, 2), (39, 3), (5, 1)):
            for w, n in ((19, R), (7, 2400), (1, 600)):
                ref = orc.geweke_diagnostic(y[:n, c, d])[:, 1]
                np.testing.assert_allclose(z2[d, c, w], ref, rtol=1e-7, equal_nan=True)
    # through the sampler: convergence_diagnostics=True fills PE_object.convergence after doMetropolisHastings
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                 settings={'mcmc_length': 3000, 'mcmc_burn_in': 500, 'mcmc_nchains': 8, 'mcmc_random_seed': 1, 'mcmc_relative_step_length': 0.3,
                           'mcmc_info_gain_returned': 'log_ratio', 'convergence_diagnostics': True, 'mcmc_exportLog': False})
    PE.doMetropolisHastings()
    G = PE.convergence['Geweke']
    assert np.array_equal(G['window_indices'], np.linspace(0, 2500, 21).astype(int)[1:])
    A = PE.convergence['AutoCorrTime']                          # the autocorrelation half (unpinned: oracle restatement of zeus)
    ra = orc.autocorr_time_convergence(PE.discrete_chains_post_burn_in_samples)
    assert np.array_equal(A['window_indices'], ra['window_indices']) and list(A['parameter_act_for_each_window']) == ['0', '1']
    np.testing.assert_allclose(A['final_parameter_values'], ra['final_parameter_values'], rtol=1e-9)
    np.testing.assert_allclose(A['parameter_act_for_each_window']['1'], ra['taus'][:, 1], rtol=1e-9, atol=1e-12, equal_nan=True)
    zz = PE.geweke_z_scores.cpu().numpy()
    chains = PE.discrete_chains_post_burn_in_samples
    assert chains.shape == (2500, 8, 2) and zz.shape == (2, 8, 20, 20)
    checked = 0
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for d in range(2):
            for c in range(8):
                for w in (6, 13, 19):
                    try:       # MH chains repeat rows: a short slice of identical values makes the reference's 2x2 solve singular
                        ref = orc.geweke_diagnostic(chains[:G['window_indices'][w], c, d])[:, 1]
                    except np.linalg.LinAlgError:
                        continue
                    np.testing.assert_allclose(zz[d, c, w], ref, rtol=1e-7, equal_nan=True)
                    checked += 1
    assert checked >= 40
    comb = np.mean(np.mean(np.abs(zz), axis=0), axis=0)
    np.testing.assert_allclose(G['final_combined_parameter_values'], comb[:, -1], rtol=1e-12, equal_nan=True)


def test_optimizers_walk_the_reference_iterates(env, golden):
    """doOptimizeNegLogP / doOptimizeLogP and a multistart of optimiser runs (IP.py:1661-1694, 912-921): scipy's optimiser on
    the host with every objective value from the log-posterior kernel.  getLogP is bit-exact for Example00, so the optimiser
    sees the reference's values and must end on the reference's optimum to the last bit (tests/golden/make_golden_optimize.py)."""
    import contextlib
    import io
    PEUQSE, _, _, _ = env
    g = golden("optimize")
    for tag, method in (('nm', 'Nelder-Mead'), ('bfgs', None), ('powell', 'Powell')):
        PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200])
        with contextlib.redirect_stdout(io.StringIO()):
            x, f = PE.doOptimizeNegLogP(method=method, printOptimum=False, verbose=False)
        assert np.array_equal(x, g[f'{tag}_x']) and float(f) == float(g[f'{tag}_negLogP']), tag
        assert float(PE.map_logP) == -float(g[f'{tag}_negLogP'])
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200])
    x, lp = PE.find_MAP(method='Nelder-Mead', printOptimum=False, verbose=False, maxiter=30)
    assert np.array_equal(x, g['nm30_x']) and float(lp) == float(g['nm30_logP'])
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                 settings={'multistart_initialPointsDistributionType': 'grid', 'multistart_gridsearchSamplingInterval': [50, 100],
                           'multistart_gridsearchSamplingRadii': [1, 1], 'multistart_searchType': 'doOptimizeNegLogP',
                           'multistart_passThroughArgs': {'method': 'Nelder-Mead', 'printOptimum': False, 'verbose': False},
                           'responses_abscissa': X00})
    best = PE.doMultiStart()
    assert np.array_equal(best[0], g['ms_best_x']) and float(PE.map_logP) == float(g['ms_map_logP'][0]) and PE.map_index == 6
    assert np.array_equal(PE.permutations_MAP_logP_and_parameters_values, g['ms_perm_array'])
    assert np.array_equal(PE.listOfPermutations, g['ms_list'])
    assert open('logs_and_csvs/multistart_log_file.txt', 'rb').read() == bytes(g['ms_log'])
    assert PE.multistart_optimizer_info['launches'] < 400          # the 9 start points advanced in lock step
    # 64 uniform start points: lock-step batch == one scipy run per start point (multistart_batched_optimizer False)
    res = {}
    for flag in (True, False):
        PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                     settings={'multistart_initialPointsDistributionType': 'uniform', 'multistart_numStartPoints': 64,
                               'multistart_searchType': 'doOptimizeLogP', 'multistart_batched_optimizer': flag, 'exportLog': False,
                               'multistart_passThroughArgs': {'method': 'Nelder-Mead', 'printOptimum': False, 'verbose': False}})
        best = PE.doMultiStart()
        res[flag] = (best[0], PE.map_logP, PE.map_index, PE.permutations_MAP_logP_and_parameters_values)
    assert np.array_equal(res[True][0], res[False][0]) and res[True][1] == res[False][1] and res[True][2] == res[False][2]
    assert np.array_equal(res[True][3], res[False][3])


def test_multistart_of_metropolis_hastings_runs(env):
    """doMultiStart(searchType='doMetropolisHastings') (IP.py:865-879, 975-990): every start point is one chain of a single
    launch.  The oracle replays each chain with the kernel's own draws (step for step), picks the best start with the
    reference's strict '>' rule and filters its rows twice, as the reference does with bestResultSoFar."""
    PEUQSE, orc, _, _ = env
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    settings = {'mcmc_length': 400, 'mcmc_burn_in': 50, 'mcmc_relative_step_length': 0.1, 'mcmc_random_seed': 3,
                'mcmc_info_gain_returned': 'log_ratio', 'multistart_searchType': 'doMetropolisHastings',
                'multistart_initialPointsDistributionType': 'grid', 'multistart_gridsearchSamplingInterval': [60, 120],
                'multistart_gridsearchSamplingRadii': [1, 1], 'multistart_keep_cumulative_post_burn_in_data': True,
                'responses_abscissa': X00}
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200], settings=settings)
    best = PE.doMultiStart()
    pts = PE.listOfPermutations
    assert pts.shape == (9, 2) and np.array_equal(pts[0], [200., 500.])
    raw = mh_launch(PE, pts, 400, 50, 3, step_cholesky(PE.Q_covmat, 0.1), export_draws=True)
    P = orc.build_problem([200, 500], [100, 200], Y00, S00)
    _, bt = orc.sim_quadratic(X00)
    ref = orc.mh_chains_batch(P, pts, raw['deltas'].cpu().numpy(), raw['us'].cpu().numpy(), bt)
    assert np.array_equal(raw['samples'].cpu().numpy(), ref['samples'])
    post_lp = ref['logP'][50:]                                                  # [rows, starts]
    maps = post_lp.max(axis=0)
    i = int(np.argmax(maps))
    arr = PE.permutations_MAP_logP_and_parameters_values
    assert np.array_equal(arr[:, 0], maps)
    for c in range(9):
        assert np.array_equal(arr[c, 1:], ref['samples'][50:, c][int(np.argmax(post_lp[:, c]))])
    with np.errstate(all='ignore'):
        st = orc.post_burn_in_statistics_filtered(ref['samples'][50:, i], ref['logP'][50:, i], ref['log_priors'][50:, i])
        st = orc.post_burn_in_statistics_filtered(st['samples'], st['logP'], st['log_priors'])
    assert np.array_equal(best[5], st['samples']) and np.array_equal(best[0], st['map_parameter_set'])
    np.testing.assert_allclose(best[1], st['mu_AP'], rtol=1e-12)
    np.testing.assert_allclose(best[2], st['stdap'], rtol=1e-12)
    np.testing.assert_allclose(best[4], st['info_gain_log_ratio'], rtol=1e-11)
    assert PE.cumulative_post_burn_in_samples.shape == (9 * 350, 2)
    assert np.array_equal(PE.cumulative_post_burn_in_samples[350 * 4:350 * 5], ref['samples'][50:, 4])
    log = open('logs_and_csvs/multistart_log_file.txt').read()
    assert f"highest_MAP_initial_point_index: {i}" in log and "(for the above initial point)" in log
    assert np.array_equal(PE.UserInput.InputParameterInitialGuess, [200., 500.])


def test_geweke_short_and_single_column_chains(env):
    """Chains too short for the early windows give NaN scores there (slices of fewer than three rows) and the later windows
    still agree with the oracle; a single chain with a single parameter; a window list given by the caller."""
    import warnings
    PEUQSE, orc, _, torch = env
    from peuqse_b200.diagnostics import geweke_z_scores
    rng = np.random.default_rng(4)
    x = np.cumsum(rng.standard_normal((260, 1, 1)), axis=0) * 0.1 + rng.standard_normal((260, 1, 1))
    z, windows, _ = geweke_z_scores(x)
    z = z.cpu().numpy()
    assert z.shape == (1, 1, 20, 20) and np.array_equal(windows, np.linspace(0, 260, 21).astype(int)[1:])
    assert np.isnan(z[0, 0, 0]).all()                # 13 rows: every early slice holds a single row
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for w in (9, 14, 19):
            np.testing.assert_allclose(z[0, 0, w], orc.geweke_diagnostic(x[:windows[w], 0, 0])[:, 1], rtol=1e-7, equal_nan=True)
        z2, w2, _ = geweke_z_scores(torch.as_tensor(x).cuda(), windows=[200, 260], initial_window=0.2, comparison_window=0.4, intervals=7)
        z2 = z2.cpu().numpy()
        assert z2.shape == (1, 1, 2, 7)
        np.testing.assert_allclose(z2[0, 0, 0], orc.geweke_diagnostic(x[:200, 0, 0], 0.2, 0.4, 7)[:, 1], rtol=1e-7)
        np.testing.assert_allclose(z2[0, 0, 1], orc.geweke_diagnostic(x[:260, 0, 0], 0.2, 0.4, 7)[:, 1], rtol=1e-7)
    with pytest.raises(ValueError):
        geweke_z_scores(x, initial_window=0.6, comparison_window=0.5)


@pytest.mark.gpu
def test_autocorr_time_on_device_vs_oracle(env):
    """SURVEY 8f rank 4, the integrated-autocorrelation-time half of the convergence diagnostics (IP.py:3916-3948):
    k_act_concat + FFT autocorrelation + automated window on the device against the oracle's numpy restatement of
    zeus.autocorr.AutoCorrTime (PARITY UNPINNED at the zeus boundary: zeus is not installed and not vendored).  Tolerance
    1e-9 relative on tau: cuFFT and numpy's pocketfft round differently; the window index itself must agree (a tau taken at
    a neighbouring lag differs by far more than the tolerance).  AR(1) chains also pin the estimate to (1+rho)/(1-rho)."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200.diagnostics import autocorr_times
    rng = np.random.default_rng(5)
    R, C, k = 4000, 37, 5                                                  # ragged: 37 * 5 columns, tiles of 64; rows % 32 != 0
    rho = np.array([0.0, 0.5, 0.9, 0.99, 0.3])
    e = rng.standard_normal((R, C, k))
    x = np.empty_like(e)
    x[0] = e[0]
    for t in range(1, R):
        x[t] = rho * x[t - 1] + e[t]
    x = x * np.array([1.0, 1e-3, 40.0, 1.0, 1e5]) + np.array([0.0, 7.0, -3e3, 1e2, 1e7])
    taus, windows, launches = autocorr_times(torch.as_tensor(x).cuda())
    taus = taus.cpu().numpy()
    ref = orc.autocorr_time_convergence(x)
    assert np.array_equal(windows, ref['window_indices']) and launches == 20 and taus.shape == (20, 5)
    np.testing.assert_allclose(taus, ref['taus'], rtol=1e-9, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(taus[-1, :3], (1 + rho[:3]) / (1 - rho[:3]), rtol=0.1)
    # the concatenated series itself, bit for bit (the layout kernel subtracts the mean it is given)
    import ctypes
    from peuqse_b200 import _lib
    L = _lib.load()
    n = 1234
    d_x = torch.as_tensor(x).cuda()
    mean = torch.as_tensor(x[:n].mean(axis=(0, 1))).cuda()
    series = torch.full((k, 65536), 9.0, dtype=torch.float64, device='cuda')
    _lib.check(L.peuqse_act_concat(_lib.ptr(d_x), n, C, k, _lib.ptr(mean), _lib.ptr(series), 65536, _lib.current_stream()))
    s = series.cpu().numpy()
    for d in range(k):
        assert np.array_equal(s[d, :n * C], x[:n, :, d].T.reshape(-1) - mean[d].item())
        assert not s[d, n * C:].any()
    # short and degenerate inputs: one chain, windows of one row, a constant parameter (0/0 -> NaN like numpy)
    y = rng.standard_normal((50, 1, 2))
    y[:, :, 1] = 3.0
    t2, w2, _ = autocorr_times(y)
    r2 = orc.autocorr_time_convergence(y)
    assert np.array_equal(w2, r2['window_indices'])
    np.testing.assert_allclose(t2.cpu().numpy(), r2['taus'], rtol=1e-9, atol=1e-12, equal_nan=True)   # tau = 0 for two-row windows
    t3, _, _ = autocorr_times(x[:, :4, :2], windows=[100, 1000], c=3.0)
    for i, n in enumerate((100, 1000)):
        np.testing.assert_allclose(t3[i].cpu().numpy(), orc.autocorr_time(x[:n, :4, :2], c=3.0), rtol=1e-9)
    with pytest.raises(ValueError):
        autocorr_times(x, windows=[R + 1])


@pytest.mark.gpu
def test_full_covariance_value_independent_of_batch_size(env):
    """The dense-covariance contraction (csrc/fullcov.cu) runs one CTA per tile for large batches and one CTA per (tile,
    K chunk) for small ones; both accumulate the same four K chunks and combine them in the same order, so a row's
    log-likelihood must not depend on the batch it is evaluated in (the sharded ensemble-slice rounds rely on it), and both
    must agree with scipy's logpdf (IP.py:2960) to 1e-12 relative."""
    PEUQSE, orc, _, torch = env
    from scipy.stats import multivariate_normal as mvn
    rng = np.random.default_rng(8)
    k, n = 3, 300
    Phi = orc.cos_basis(2 * n, k)
    x = np.linspace(0, 1, n)
    Sig = 0.1 ** 2 * np.exp(-(x[:, None] - x[None, :]) ** 2 / (2 * 0.05 ** 2)) + 0.05 ** 2 * np.eye(n)
    truth = rng.standard_normal(k)
    y = Phi @ truth + 0.1 * rng.standard_normal(2 * n)
    PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), [list(y[:n]), list(y[n:])], [Sig, Sig], [0.0] * k, [2.0] * k,
                 settings={'responses_observed_max_covmat_size': 400})
    B = 64 * 64                                                   # 5 column tiles x 64 row tiles: above two tiles per SM
    th = torch.from_numpy(truth + 0.1 * rng.standard_normal((B, k))).cuda()
    big = PE.getLogP_batch(th)
    for lo, hi in ((0, 100), (1000, 1001), (4000, 4096), (37, 700)):
        assert torch.equal(PE.getLogP_batch(th[lo:hi].contiguous()), big[lo:hi])
    _, _, ll = parts(PE, th[:6].cpu().numpy())
    want = np.array([mvn.logpdf(x=y[:n], mean=(Phi @ t)[:n], cov=Sig) + mvn.logpdf(x=y[n:], mean=(Phi @ t)[n:], cov=Sig)
                     for t in th[:6].cpu().numpy()])
    np.testing.assert_allclose(ll, want, rtol=1e-12)


@pytest.mark.gpu
def test_info_gain_cutoff_vs_reference_fixture(env, golden):
    """parameter_estimation_settings['mcmc_info_gain_cutoff'] != 0 (IP.py:1805-1843): the truncated log-ratio information gain
    of the unmodified reference (tests/golden/info_gain_cutoff.npz) from the device vectors of the kept rows; 1e-11 relative
    (summation order, device exp / log), NaN when the cutoff removes every row."""
    PEUQSE, orc, _, torch = env
    g, src = golden("info_gain_cutoff"), golden("post_burn_in_statistics")
    for key in g.files:
        filt, cutoff, info_gain, evidence, n_kept = g[key]
        PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                     settings={'mcmc_threshold_filter_samples': bool(filt), 'mcmc_threshold_filter_coefficient': 'auto',
                               'mcmc_threshold_filter_benchmark': 'auto', 'mcmc_info_gain_returned': 'log_ratio',
                               'mcmc_info_gain_cutoff': float(cutoff)})
        PE.post_burn_in_samples = src['pbs_samples'].copy()
        PE.post_burn_in_log_posteriors_un_normed_vec = src['pbs_logP'].copy()
        PE.post_burn_in_log_priors_vec = src['pbs_log_priors'].copy()
        res = PE.calculatePostBurnInStatistics()
        assert len(PE.post_burn_in_samples) == int(n_kept)
        np.testing.assert_allclose(PE.evidence, evidence, rtol=1e-12)
        np.testing.assert_allclose(res[4], info_gain, rtol=1e-11, equal_nan=True)
        np.testing.assert_allclose(PE.info_gain_log_ratio, info_gain, rtol=1e-11, equal_nan=True)


@pytest.mark.gpu
def test_c3_benched_shape_in_kernel_statistics_vs_oracle(env, golden):
    """The shape bench.py times for C3: many chains of the 1024-observation TPD model with the samples NOT stored, only the
    per-chain statistics block of k_mh_warp (accept count, MAP row / logP / theta, last state, first and second moments).  64 chains x
    300 steps from different start points: the oracle replays the kernel's exported draws (a second launch with the same seed
    that stores everything) and every statistic is rebuilt from the oracle's rows."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200 import _lib
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    g = golden("tpd")
    n, L, C, burn, k = 1024, 300, 64, 30, 6
    sd = np.array([20, 20, 13, 13, 0.1, 0.1])
    # as in bench.py the 1024 observations are ONE response with a diagonal covariance (responses_observed_max_covmat_size above n;
    # the default 100 makes it a sum of per-observation densities: test_mh_tpd_step_for_step and test_c3_many_chain_kernel_shapes_agree)
    PE = make_pe(PEUQSE, PEUQSE.Tpd2SiteModel.linear_ramp(n), list(g['tpd1024_y']), list(np.full(n, float(g['tpd1024_sigma']))),
                 [41.5, 41.5, 13.0, 13.0, 0.1, 0.1], list(sd), settings={'responses_observed_max_covmat_size': 5000})
    truth = np.array([41.5, 45, 13, 13, 0.1, 0.1])
    start = truth + 2e-4 * sd * np.random.default_rng(12).standard_normal((C, k))
    chol = step_cholesky(PE.Q_covmat, 2.5e-6)
    full = mh_launch(PE, start, L, burn, 5, chol, export_draws=True, chain_offset=1000)
    lean = mh_launch(PE, start, L, burn, 5, chol, store_samples=False, chain_offset=1000)          # the benched shape
    assert lean['samples'] is None and torch.equal(full['stats'], lean['stats'])
    kBT, c1, c2 = orc.tpd2site_constants(n)
    _, bt = orc.sim_tpd2site(kBT, c1, c2)
    P = orc.build_problem([41.5, 41.5, 13.0, 13.0, 0.1, 0.1], sd, g['tpd1024_y'], np.full(n, float(g['tpd1024_sigma'])), max_covmat_size=5000)
    ref = orc.mh_chains_batch(P, start, full['deltas'].cpu().numpy(), full['us'].cpu().numpy(), bt)
    S, LP = ref['samples'], ref['logP']                                  # [L, C, k], [L, C]
    assert np.array_equal(full['samples'].cpu().numpy(), S)             # accept / reject step for step, all 64 chains
    st = lean['stats'].cpu().numpy()
    lay = _lib.mh_stat_layout(k)
    nacc = (np.abs(np.diff(S, axis=0)).sum(axis=2) > 0).sum(axis=0)
    assert np.array_equal(st[:, lay['NACC']], ref['accepted'][1:].sum(axis=0)) and np.all(nacc <= st[:, lay['NACC']])
    assert 0.05 < st[:, lay['NACC']].mean() / (L - 1) < 0.95
    assert np.array_equal(st[:, lay['LAST']:lay['LAST'] + k], S[-1])
    np.testing.assert_allclose(st[:, lay['LASTLOGP']], LP[-1], rtol=1e-12)
    # first maximum among rows >= burn_in (IP.py:1925-1927).  The device's logP differs from the oracle's in the last digits
    # (n = 1024: warp reduction), so two distinct states whose logP tie to 1e-12 could swap: require that the device's MAP row
    # holds a state whose oracle logP is the maximum to 1e-12, that it is the FIRST row holding that state, and the exact row
    # wherever the oracle's best state is unique
    lp_post = LP[burn:]
    mrow = st[:, lay['MAPROW']].astype(int)
    cols = np.arange(C)
    assert np.all(mrow >= burn) and np.all(mrow < L)
    np.testing.assert_allclose(LP[mrow, cols], lp_post.max(axis=0), rtol=1e-12)
    np.testing.assert_allclose(st[:, lay['MAPLOGP']], lp_post.max(axis=0), rtol=1e-12)
    assert np.array_equal(st[:, lay['MAPTHETA']:lay['MAPTHETA'] + k], S[mrow, cols])
    first_of_state = (mrow == burn) | np.any(S[np.maximum(mrow - 1, 0), cols] != S[mrow, cols], axis=1)
    assert np.all(first_of_state)
    exact = burn + np.argmax(lp_post, axis=0)
    assert np.mean(mrow == exact) > 0.9
    d = S[burn:] - start[None]                                            # moments about each chain's start point
    s1 = np.zeros((C, k))
    for r_ in range(d.shape[0]):                                          # the kernel adds row after row
        s1 = s1 + d[r_]
    assert np.array_equal(st[:, lay['S1']:lay['S1'] + k], s1)
    s2 = np.einsum('rci,rcj->cij', d, d)
    scale = np.sqrt(np.einsum('cii->ci', s2)[:, :, None] * np.einsum('cii->ci', s2)[:, None, :])      # cross terms may cancel
    assert np.all(np.abs(st[:, lay['S2']:lay['S2'] + k * k].reshape(C, k, k) - s2) <= 1e-12 * scale)
    # pooled moments through the public call with store off == numpy over the oracle's rows
    PE.UserInput.parameter_estimation_settings.update(mcmc_length=L, mcmc_burn_in=burn, mcmc_nchains=C, mcmc_random_seed=5,
                                                      mcmc_relative_step_length=2.5e-6, mcmc_store_samples=False)
    PE.UserInput.InputParameterInitialGuess = truth
    out = PE.doMetropolisHastings()
    ref2 = orc.mh_chains_batch(P, np.tile(truth, (C, 1)), *(t.cpu().numpy() for t in
                               (lambda o: (o['deltas'], o['us']))(mh_launch(PE, np.tile(truth, (C, 1)), L, burn, 5, chol, export_draws=True))), bt)
    rows2 = ref2['samples'][burn:].reshape(-1, k)
    np.testing.assert_allclose(out[1], rows2.mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(out[2], rows2.std(axis=0), rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("max_covmat", [5000, 100], ids=["one_diagonal_response", "per_observation_densities"])
def test_c3_many_chain_kernel_shapes_agree(env, golden, monkeypatch, max_covmat):
    """k_mh_tpd (config C3: rate-law model, one diagonal response) deals up to four chains to a warp, the chains of an SM evenly to
    its sub-partitions.  A chain's trajectory must not depend on the dealing: 4100 chains (warps of 4 + 3 chains, a last CTA
    that is not full) x 40 steps on 160 observations (not a multiple of the 32 a warp strides over) through every compiled shape
    give bit-identical rows, draws and statistics, and agree with the general kernel k_mh_warp (library exp, other summation
    grouping): same accept / reject decisions, log-posteriors to 1e-12."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200 import _lib
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    n, L, C, burn, k = 160, 40, 4100, 5, 6
    kBT, c1, c2 = orc.tpd2site_constants(n)
    truth = np.array([41.5, 45, 13, 13, 0.1, 0.1])
    sim_t = orc.sim_tpd2site(kBT, c1, c2)[0](truth)
    sig = 0.02 * np.max(sim_t)
    y = sim_t + sig * np.random.default_rng(3).standard_normal(n)
    sd = np.array([20, 20, 13, 13, 0.1, 0.1])
    PE = make_pe(PEUQSE, PEUQSE.Tpd2SiteModel.linear_ramp(n), list(y), list(np.full(n, sig)), [41.5, 41.5, 13.0, 13.0, 0.1, 0.1], list(sd),
                 settings={'responses_observed_max_covmat_size': max_covmat})   # 5000: one response with a diagonal covariance, as in
    # bench.py; 100 (the reference's default): the 160 observations are a sum of per-observation densities (IP.py:2969-2978)
    start = truth + 5e-4 * sd * np.random.default_rng(4).standard_normal((C, k))
    chol = step_cholesky(PE.Q_covmat, 2e-5)
    runs = {}
    for shape in (1, 4, 5, 7, 8):
        monkeypatch.setenv("PEUQSE_MH_SHAPE", str(shape))
        runs[shape] = mh_launch(PE, start, L, burn, 9, chol, thin=2, export_draws=True, chain_offset=77)
    monkeypatch.delenv("PEUQSE_MH_SHAPE")
    ref = runs[4]
    for shape in (5, 7, 8):
        for key in ('samples', 'logp', 'logprior', 'loglik', 'deltas', 'us', 'stats'):
            assert torch.equal(ref[key], runs[shape][key]), (shape, key)
    gen = runs[1]
    for key in ('samples', 'deltas', 'us'):
        assert torch.equal(ref[key], gen[key]), key
    np.testing.assert_allclose(ref['logp'].cpu().numpy(), gen['logp'].cpu().numpy(), rtol=1e-12)
    assert PE.host_problem.resp_mode[0] == (_lib.RESP_MVN_DIAG if max_covmat > n else _lib.RESP_PER_VALUE)
    assert not torch.equal(ref['logp'], gen['logp'])                    # two kernels did run
    lay = _lib.mh_stat_layout(k)
    a, b = ref['stats'].cpu().numpy(), gen['stats'].cpu().numpy()
    assert np.array_equal(a[:, lay['NACC']], b[:, lay['NACC']]) and 0.02 < a[:, lay['NACC']].mean() / (L - 1) < 0.98
    assert np.array_equal(a[:, lay['LAST']:lay['LAST'] + k], b[:, lay['LAST']:lay['LAST'] + k])
    np.testing.assert_allclose(a[:, lay['MAPLOGP']], b[:, lay['MAPLOGP']], rtol=1e-12)
    assert np.array_equal(a[:, lay['S1']:lay['S1'] + k], b[:, lay['S1']:lay['S1'] + k])
    # mcmc_mode 'MAP_finding' (per-chain step coefficient adapted every hundred steps) in the dealt kernel: the general kernel's chains
    mf = {}
    for shape in (1, 4, 7):
        monkeypatch.setenv("PEUQSE_MH_SHAPE", str(shape))
        mf[shape] = mh_launch(PE, start, 230, 10, 9, chol, store_samples=False, map_finding=True)['stats'].cpu().numpy()
    monkeypatch.delenv("PEUQSE_MH_SHAPE")
    assert np.array_equal(mf[4], mf[7])
    for key, width in (('NACC', 1), ('DYN', 1), ('LAST', k), ('MAPROW', 1)):
        assert np.array_equal(mf[4][:, lay[key]:lay[key] + width], mf[1][:, lay[key]:lay[key] + width]), key
    assert len(np.unique(mf[4][:, lay['DYN']])) >= 3                 # two adaptations in 230 steps: 0.95^2, 0.95 x 1.05, 1.05^2
    # the launcher's own choice: one chain per warp up to 12 chains per SM, the dealt shape above
    few = mh_launch(PE, start[:200], L, burn, 9, chol, thin=2, export_draws=True, chain_offset=77)
    assert torch.equal(few['samples'], ref['samples'][:, :200]) and torch.equal(few['stats'], ref['stats'][:200])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2050, 2300])
def test_c3_many_observations_shared_memory_limit(env, monkeypatch, n):
    """k_mh_tpd stages 48 bytes per observation above its 128 KB of per-chain blocks and exp table: 2050 observations still fit
    the 227 KB of an SM (zero rows pad the last group of 32), 2300 do not and the launcher must take the general kernel.  Either
    way the chains are the general kernel's: same accept / reject decisions, log-posteriors to 1e-12."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    L, C, k = 25, 70, 6
    kBT, c1, c2 = orc.tpd2site_constants(n)
    truth = np.array([41.5, 45, 13, 13, 0.1, 0.1])
    sim_t = orc.sim_tpd2site(kBT, c1, c2)[0](truth)
    sig = 0.02 * np.max(sim_t)
    y = sim_t + sig * np.random.default_rng(8).standard_normal(n)
    sd = np.array([20, 20, 13, 13, 0.1, 0.1])
    PE = make_pe(PEUQSE, PEUQSE.Tpd2SiteModel.linear_ramp(n), list(y), list(np.full(n, sig)), [41.5, 41.5, 13.0, 13.0, 0.1, 0.1], list(sd),
                 settings={'responses_observed_max_covmat_size': 5000})       # one response with a diagonal covariance, as in bench.py
    start = truth + 2e-4 * sd * np.random.default_rng(9).standard_normal((C, k))
    chol = step_cholesky(PE.Q_covmat, 2e-6)
    auto = mh_launch(PE, start, L, 3, 4, chol, export_draws=True)
    monkeypatch.setenv("PEUQSE_MH_SHAPE", "1")
    gen = mh_launch(PE, start, L, 3, 4, chol, export_draws=True)
    monkeypatch.delenv("PEUQSE_MH_SHAPE")
    for key in ('samples', 'deltas', 'us'):
        assert torch.equal(auto[key], gen[key]), key
    np.testing.assert_allclose(auto['logp'].cpu().numpy(), gen['logp'].cpu().numpy(), rtol=1e-12)
    assert torch.equal(auto['logp'], gen['logp']) == (n > 2080)          # the general kernel ran both times only above the limit
    acc = (np.abs(np.diff(auto['samples'].cpu().numpy(), axis=0)).sum(axis=2) > 0).mean()
    assert 0.02 < acc < 0.98


@pytest.mark.gpu
def test_c5_window_above_two_to_the_33(env):
    """C5 is 1e10 points: global indices beyond 2^33.  A 4096-point window that starts above 2^33 (and one straddling 2^32) must
    be the oracle's points (Philox counter = the 64-bit global index) and log-posteriors, bit for bit, and the window's best
    record must carry the global index."""
    PEUQSE, orc, philox_np, torch = env
    from peuqse_b200.sweep import Sweep, make_spec
    k, n = 10, 64
    Phi = orc.cos_basis(n, k)
    truth = np.random.default_rng(5).standard_normal(k) * 0.5
    y = Phi @ truth + 0.1 * np.random.default_rng(6).standard_normal(n)
    PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), list(y), list(np.full(n, 0.1)), [0.0] * k, [1.0] * k)
    P = orc.build_problem([0.0] * k, [1.0] * k, y, np.full(n, 0.1))
    _, bt = orc.sim_linear(Phi)
    spec = make_spec("uniform", k, np.zeros(k), spread=0.866, std_prior=np.ones(k), seed=0)
    total, pieces = 10_000_000_000, 2441407                      # 1e10 points in pieces of ~4096
    for piece in (int((2 ** 33 + 12345) * pieces / total) + 1, int(2 ** 32 * pieces / total), pieces - 1):
        sw = Sweep(PE.device_problem, spec, total, store_logp=True, shard=(piece, pieces), local=True)
        first, count = sw.ranges[0]
        assert first == (total * piece) // pieces and 4090 <= count <= 4100
        if piece == pieces - 1:
            assert first + count == total
        sw.step(float(np.log(10 ** -2.0)), True)
        u = philox_np.sweep_uniforms(0, first, count, k)
        th = orc.uniform_points(u, 0.866, np.ones(k), np.zeros(k))
        assert np.array_equal(sw.points(first, count).cpu().numpy(), th)
        lp = orc.log_p_batch(P, th, bt)[0]
        assert np.array_equal(sw.logp.cpu().numpy(), lp)
        R = sw.unpack(sw.res.cpu().numpy(), 2.0)
        idx, best = orc.first_max_np(lp)
        assert R.map_index == first + idx and R.map_logP == best and np.array_equal(R.map_parameter_set, th[idx])
        w, _ = orc.permutation_weights(lp, best)
        assert R.sum_weights == w.sum()
        # the no-store path (the one the 1e10-point run uses) finds the same record
        s2 = Sweep(PE.device_problem, spec, total, store_logp=False, shard=(piece, pieces), local=True)
        s2.step(float(np.log(10 ** -2.0)), True)
        R2 = s2.unpack(s2.res.cpu().numpy(), 2.0)
        assert (R2.map_index, R2.sum_weights) == (R.map_index, R.sum_weights)
        assert abs(R2.map_logP - R.map_logP) <= 1e-12 * abs(R.map_logP)         # tensor-pipe simulation: another summation order
        # ... and with the tensor-pipe kernel switched off the no-store path is the scalar one: the same bits
    import subprocess, sys, os
    code = ("import numpy as np, sys; sys.path.insert(0, %r); import peuqse_b200 as PEUQSE; from oracle import peuqse_oracle as orc;"
            "from peuqse_b200.sweep import Sweep, make_spec;"
            "k, n = 10, 64; Phi = orc.cos_basis(n, k); truth = np.random.default_rng(5).standard_normal(k) * 0.5;"
            "y = Phi @ truth + 0.1 * np.random.default_rng(6).standard_normal(n); UI = PEUQSE.UserInput.fresh();"
            "UI.responses['responses_observed'] = list(y); UI.responses['responses_observed_uncertainties'] = list(np.full(n, 0.1));"
            "UI.model['InputParameterPriorValues'] = [0.0] * k; UI.model['InputParametersPriorValuesUncertainties'] = [1.0] * k;"
            "UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.LinearModel(Phi); PE = PEUQSE.parameter_estimation(UI);"
            "spec = make_spec('uniform', k, np.zeros(k), spread=0.866, std_prior=np.ones(k), seed=0);"
            "a = Sweep(PE.device_problem, spec, 1 << 20, store_logp=True).run(2.0); b = Sweep(PE.device_problem, spec, 1 << 20, store_logp=False).run(2.0);"
            "assert (a.map_index, a.map_logP, a.sum_weights) == (b.map_index, b.map_logP, b.sum_weights); print('SCALAR_SAME')"
            % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PEUQSE_SWEEP_NO_DMMA="1"), capture_output=True, text=True, timeout=600)
    assert "SCALAR_SAME" in r.stdout, r.stderr[-2000:]


@pytest.mark.gpu
def test_mh_map_finding_vs_reference_fixture_and_oracle(env, golden):
    """mcmc_mode 'MAP_finding' (IP.py:2631-2635, 2688-2713) inside the Metropolis-Hastings kernels.  (i) The kernel's draws are the
    fixture's (the unmodified reference ran on them): same modulations and accept uniforms bit for bit, increments to 1e-12.  (ii) The
    oracle replays the kernel's own exported draws: chain, log-posteriors and the final step coefficient identical, and every
    exported increment is sqrt(dyn * mod) times the increment of the unbiased mode with the oracle's dyn.  (iii) Both kernel shapes
    (thread per chain, warp per chain) agree; the reference-named API returns the chain."""
    PEUQSE, orc, philox_np, torch = env
    from peuqse_b200 import _lib
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    g = golden("mapfinding")
    L, seed, step = int(g['L']), int(g['seed']), float(g['step'])
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200],
                 settings={'mcmc_length': L, 'mcmc_burn_in': 100, 'mcmc_relative_step_length': step, 'mcmc_random_seed': seed,
                           'mcmc_mode': 'MAP_finding', 'mcmc_threshold_filter_samples': False, 'mcmc_info_gain_returned': 'log_ratio'})
    chol = step_cholesky(PE.Q_covmat, step)
    out = mh_launch(PE, [[200., 500.]], L, 100, seed, chol, export_draws=True, map_finding=True)
    plain = mh_launch(PE, [[200., 500.]], L, 100, seed, chol, export_draws=True)
    us, mods = out['us'].cpu().numpy()[:, 0, 0], out['us'].cpu().numpy()[:, 0, 1]
    deltas, base = out['deltas'].cpu().numpy()[:, 0], plain['deltas'].cpu().numpy()[:, 0]
    assert np.array_equal(us, g['us']) and np.array_equal(mods, g['mods']) and np.array_equal(us, plain['us'].cpu().numpy()[:, 0])
    np.testing.assert_allclose(base, g['base_deltas'], rtol=1e-13, atol=1e-11)
    P = orc.build_problem([200, 500], [100, 200], Y00, S00)
    _, bt = orc.sim_quadratic(X00)
    ref = orc.mh_chain_map_finding(P, [200.0, 500.0], deltas, mods, us, bt, scaled=True)
    assert np.array_equal(out['samples'].cpu().numpy()[:, 0], ref['samples'])
    assert np.array_equal(out['logp'].cpu().numpy()[:, 0], ref['logP'])
    lay = _lib.mh_stat_layout(2)
    assert float(out['stats'][0, lay['DYN']].item()) == ref['dyn_final'] and float(plain['stats'][0, lay['DYN']].item()) == 1.0
    assert len(np.unique(ref['dyn'])) > 5
    assert np.array_equal(deltas, np.sqrt(ref['dyn'][1:] * mods)[:, None] * base)
    if np.array_equal(base, g['base_deltas']):                        # identical draws: the reference's own chain
        assert np.array_equal(ref['samples'], g['samples']) and ref['dyn_final'] == float(g['dyn_final'])
    # warp-per-chain shape (<= 8192 chains of a small problem) against thread-per-chain (more chains): same chains
    th0 = np.array([200.0, 500.0]) + np.random.default_rng(4).standard_normal((8200, 2)) * [30.0, 60.0]
    few = mh_launch(PE, th0[:64], 450, 50, 9, chol, store_samples=False, map_finding=True)
    many = mh_launch(PE, th0, 450, 50, 9, chol, store_samples=False, map_finding=True)
    assert torch.equal(few['stats'], many['stats'][:64])
    assert len(torch.unique(many['stats'][:, lay['DYN']])) > 3
    res = PE.doMetropolisHastings()
    assert np.array_equal(PE.post_burn_in_samples, out['samples'].cpu().numpy()[100:, 0]) and len(res) == 7


@pytest.mark.gpu
def test_custom_log_prior_and_log_likelihood(env):
    """UserInput.model['custom_logPrior'] / ['custom_logLikelihood'] (IP.py:2757-2759, 2854-2857) as vectorised torch callables: a
    custom prior / likelihood that restates the built-in one must give the built-in log-posterior; the drivers that evaluate in
    batches (multistart, ensemble slice sampling, lock-step Metropolis-Hastings) run on them."""
    PEUQSE, orc, _, torch = env
    rng = np.random.default_rng(3)
    k, n, sig = 3, 20, 0.3
    Phi = orc.cos_basis(n, k)
    y = Phi @ rng.standard_normal(k) + sig * rng.standard_normal(n)
    PhiT, yt = torch.from_numpy(Phi.T.copy()).cuda(), torch.from_numpy(y).cuda()
    LOG2PI = float(np.log(2 * np.pi))

    def my_prior(th):                                                  # N(0, 2^2) in the reference's scaled form
        z = th / 2.0
        return -0.5 * (k * LOG2PI + (z * z).sum(dim=1))

    def my_lik(th):
        sim = th @ PhiT
        r = (yt - sim) / sig
        return -0.5 * (n * LOG2PI + n * float(np.log(sig ** 2)) + (r * r).sum(dim=1)), sim
    base = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), list(y), list(np.full(n, sig)), [0.0] * k, [2.0] * k)
    th = rng.standard_normal((500, k))
    want = parts(base, th)
    for kw in (dict(custom_logPrior=my_prior), dict(custom_logLikelihood=my_lik), dict(custom_logPrior=my_prior, custom_logLikelihood=my_lik)):
        model = None if 'custom_logLikelihood' in kw else PEUQSE.LinearModel(Phi)
        PE = make_pe(PEUQSE, model, list(y), list(np.full(n, sig)), [0.0] * k, [2.0] * k,
                     settings={'mcmc_length': 16 * 40, 'mcmc_nwalkers': 16, 'mcmc_burn_in': 16 * 5, 'mcmc_random_seed': 2,
                               'mcmc_threshold_filter_samples': False, 'multistart_initialPointsDistributionType': 'uniform',
                               'multistart_numStartPoints': 5000, 'multistart_random_seed': 1, 'exportLog': False}, **kw)
        got = parts(PE, th)
        for a, b in zip(got, want):
            np.testing.assert_allclose(a, b, rtol=1e-12)
        assert abs(PE.getLogP(th[0]) - want[0][0]) <= 1e-12 * abs(want[0][0])
        ll, sim = PE.getLogLikelihood(th[0])
        assert abs(ll - want[2][0]) <= 1e-12 * abs(want[2][0])
        if 'custom_logLikelihood' in kw:
            np.testing.assert_allclose(sim, Phi @ th[0], rtol=1e-12)
        PE.doMultiStart()
        base.UserInput.parameter_estimation_settings.update({'multistart_initialPointsDistributionType': 'uniform',
                                                             'multistart_numStartPoints': 5000, 'multistart_random_seed': 1, 'exportLog': False})
        base.doMultiStart()
        assert np.array_equal(PE.map_parameter_set, base.map_parameter_set) and abs(PE.map_logP - base.map_logP) < 1e-9
        PE.doEnsembleSliceSampling()
        assert PE.post_burn_in_samples.shape == (16 * 35, k) and np.all(np.isfinite(PE.post_burn_in_log_posteriors_un_normed_vec))
        PE.UserInput.parameter_estimation_settings.update({'mcmc_length': 300, 'mcmc_burn_in': 30, 'mcmc_relative_step_length': 0.01})
        del PE.UserInput.last_MCMC_run_type                            # another sampler on the same UserInput is refused (IP.py:2565-2569)
        PE.doMetropolisHastings()                                      # the lock-step form (tests/test_gpu_mh_lockstep.py)
        assert PE.post_burn_in_samples.shape == (270, k)
    # a custom prior replaces the bounds check too (IP.py:2757-2759 returns before doInputParameterBoundsChecks)
    PEb = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), list(y), list(np.full(n, sig)), [0.0] * k, [2.0] * k, custom_logPrior=my_prior,
                  InputParameterPriorValues_lowerBounds=[-0.1] * k, InputParameterPriorValues_upperBounds=[0.1] * k)
    assert np.isfinite(PEb.getLogP([1.0, 1.0, 1.0]))


def test_simulation_uncertainties_as_a_function_of_the_parameters(env, golden):
    """model['responses_simulation_uncertainties'] = f(theta) (IP.py:563-572, 2883-2896, 2925-2927) as a vectorised torch callable: the
    covariance of the likelihood changes with every point (peuqse_loglik_simvar).  Against the unmodified reference's getLogP /
    getLogPrior / getLogLikelihood (tests/golden/simunc.npz), one logpdf per response and one per value; with a built-in model and with
    a torch simulator; then the batched drivers run on it."""
    PEUQSE, orc, _, torch = env
    g = golden("simunc")
    X = torch.tensor(X00, dtype=torch.float64, device='cuda')
    f = lambda t: (X[None, :] - t[:, :1]) ** 2 + t[:, 1:2]
    sd = lambda t: 0.03 * torch.abs(f(t)) + 50.0 * (1.0 + torch.abs(t[:, 1:2]) / 500.0)
    th = g['simunc_thetas']
    for tag, maxcov in (('mvn', 100), ('per_value', 2)):
        for model in (PEUQSE.QuadraticModel(X00), PEUQSE.TorchSimulator(f)):
            PE = make_pe(PEUQSE, model, Y00, S00, [200, 500], [100, 200], responses_simulation_uncertainties=sd,
                         settings={'responses_observed_max_covmat_size': maxcov, 'mcmc_length': 400, 'mcmc_burn_in': 40,
                                   'mcmc_relative_step_length': 0.2, 'mcmc_random_seed': 1, 'mcmc_nchains': 2, 'exportLog': False,
                                   'multistart_initialPointsDistributionType': 'grid', 'multistart_gridsearchSamplingInterval': [20, 40],
                                   'multistart_gridsearchSamplingRadii': [5, 5]})
            lp, pr, ll = parts(PE, th)
            want = g['simunc_' + tag]
            np.testing.assert_allclose(ll, want[:, 2], rtol=1e-12)
            np.testing.assert_allclose(lp, want[:, 0], rtol=1e-12)
            assert np.array_equal(pr, want[:, 1])
            assert abs(PE.getLogP(th[3]) - want[3, 0]) <= 1e-12 * abs(want[3, 0])
            l1, s1 = PE.getLogLikelihood(th[3])
            assert abs(l1 - want[3, 2]) <= 1e-12 * abs(want[3, 2]) and np.allclose(np.ravel(s1), (np.asarray(X00) - th[3, 0]) ** 2 + th[3, 1])
    # a NaN among the simulated values: zero likelihood (IP.py:2866-2882)
    bad = th[:4].copy()
    bad[1, 0] = np.nan
    assert parts(PE, bad)[2][1] == float('-inf')
    # the drivers: grid multistart (MAP = the best grid point by the reference's formula), lock-step Metropolis-Hastings, ensemble slice sampling
    PE.doMultiStart()
    grid = orc.grid_decode([200., 500.], [20, 40], [5, 5], np.arange(121))
    P = orc.build_problem([200, 500], [100, 200], Y00, S00, max_covmat_size=2)
    pt, _ = orc.sim_quadratic(X00)
    x = np.asarray(X00, dtype=float)
    sd_np = lambda t: 0.03 * np.abs((x - t[0]) ** 2 + t[1]) + 50.0 * (1.0 + np.abs(t[1]) / 500.0)
    ref = np.array([orc.log_likelihood_simvar_point(P, pt(tuple(t)), sd_np(t)) + orc.log_prior_point(P, t) for t in grid])
    assert PE.map_index == int(np.argmax(ref)) and abs(PE.map_logP - ref.max()) <= 1e-12 * abs(ref.max())
    res = PE.doMetropolisHastings()
    assert len(res) == 7 and np.all(np.isfinite(PE.post_burn_in_log_posteriors_un_normed_vec))
    lp2 = parts(PE, PE.post_burn_in_samples[:50])[0]
    np.testing.assert_allclose(lp2, np.ravel(PE.post_burn_in_log_posteriors_un_normed_vec)[:50], rtol=1e-12)


def test_reduced_parameter_space(env, golden):
    """model['reducedParameterSpace'] (IP.py:406-408, 519-561, 574-606): only the listed parameters vary, the rest stay at the initial
    guess.  getLogP / getLogPrior / getLogLikelihood against the unmodified reference (tests/golden/reduced.npz) for a built-in model
    (an inner full-space engine simulates) and a torch simulator, incl. the reference's handling of an unsorted index list; then the
    samplers run in the reduced space and UserInput.model is left as the user wrote it."""
    PEUQSE, orc, _, torch = env
    g = golden("reduced")
    X = torch.tensor(X00, dtype=torch.float64, device='cuda')
    f = lambda t: (X[None, :] - t[:, :1]) ** 2 + t[:, 1:2]
    for tag, idx in (('b', [1]), ('a', [0]), ('ba', [1, 0])):
        for model in (PEUQSE.QuadraticModel(X00), PEUQSE.TorchSimulator(f), f):
            PE = make_pe(PEUQSE, model, Y00, S00, [200, 500], [100, 200], InputParameterInitialGuess=[210.0, 480.0], reducedParameterSpace=idx,
                         settings={'mcmc_length': 600, 'mcmc_burn_in': 60, 'mcmc_relative_step_length': 0.3, 'mcmc_random_seed': 2,
                                   'mcmc_nwalkers': 8, 'exportLog': False})
            assert PE.k == len(idx) and PE.reducedParameterSpaceOn and PE.UserInput.model['reducedParameterSpace'] == idx
            assert len(PE.UserInput.model['InputParameterPriorValues']) == 2
            lp, pr, ll = parts(PE, g['reduced_%s_thetas' % tag])
            want = g['reduced_%s_vals' % tag]
            np.testing.assert_allclose(ll, want[:, 2], rtol=1e-12)
            np.testing.assert_allclose(lp, want[:, 0], rtol=1e-12)
            assert np.array_equal(pr, want[:, 1])
    assert list(PE.UserInput.parameterNamesList) == ['0', '1'] and np.array_equal(PE.UserInput.InputParameterInitialGuess, [210.0, 480.0])
    PE1 = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200], InputParameterInitialGuess=[210.0, 480.0],
                  reducedParameterSpace=[1], settings={'mcmc_length': 2000, 'mcmc_burn_in': 200, 'mcmc_relative_step_length': 0.3,
                                                       'mcmc_random_seed': 2, 'exportLog': False, 'mcmc_threshold_filter_samples': False})
    res = PE1.doMetropolisHastings()
    assert PE1.post_burn_in_samples.shape == (1800, 1) and len(res[0]) == 1
    # the posterior of b given a = 210: Gaussian prior N(500, 200^2) times three Gaussian observations of b + (x - 210)^2
    x = np.asarray(X00, dtype=float)
    w = 1.0 / np.asarray(S00, dtype=float) ** 2
    prec = w.sum() + 1 / 200.0 ** 2
    mean = (np.sum(w * (np.asarray(Y00) - (x - 210.0) ** 2)) + 500.0 / 200.0 ** 2) / prec
    assert abs(PE1.mu_AP_parameter_set[0] - mean) < 6 * prec ** -0.5 / np.sqrt(1800 / 20)
    PE1.UserInput.parameter_estimation_settings.update({'multistart_initialPointsDistributionType': 'grid',
                                                        'multistart_gridsearchSamplingInterval': [10.0], 'multistart_gridsearchSamplingRadii': [30]})
    PE1.doMultiStart()
    grid = 480.0 + 10.0 * np.arange(-30, 31)
    lpg = parts(PE1, grid[:, None])[0]
    assert PE1.map_parameter_set[0] == grid[int(np.argmax(lpg))]


def test_reduced_response_space(env, golden):
    """responses['reducedResponseSpace'] (IP.py:410-413, 470-512): the full simulation runs and only the listed response dimensions enter
    the likelihood.  Against the unmodified reference (tests/golden/reduced_responses.npz): three responses, keeping [0, 2] and keeping
    [1] (whose sigmas stay "as variances", the multi-response behaviour of returnShapedResponseCovMat); built-in model (an inner engine
    on the full response set simulates), torch simulator returning the concatenated block, and one returning a list per response."""
    PEUQSE, orc, _, torch = env
    g = golden("reduced_responses")
    Phi, y, sig, th = g['rr_Phi'], g['rr_y'], g['rr_sig'], g['rr_thetas']
    n = 6
    Phi_t = torch.from_numpy(Phi).cuda()
    obs = [list(y[r * n:(r + 1) * n]) for r in range(3)]
    unc = [list(sig[r * n:(r + 1) * n]) for r in range(3)]
    as_list = lambda t: [(t @ Phi_t.T)[:, r * n:(r + 1) * n] for r in range(3)]
    for tag, keep in (('02', [0, 2]), ('1', [1])):
        for model in (PEUQSE.LinearModel(Phi), PEUQSE.TorchSimulator(lambda t: t @ Phi_t.T), PEUQSE.TorchSimulator(as_list)):
            PE = make_pe(PEUQSE, model, obs, unc, [0.0, 0.0, 0.0], [2.0, 2.0, 3.0],
                         settings={'reducedResponseSpace': keep, 'mcmc_length': 500, 'mcmc_burn_in': 50, 'mcmc_relative_step_length': 0.01,
                                   'mcmc_random_seed': 2, 'exportLog': False})
            assert PE.UserInput.num_response_dimensions == len(keep) and PE.host_problem.n_total == n * len(keep)
            lp, pr, ll = parts(PE, th)
            want = g['rr_%s_vals' % tag]
            np.testing.assert_allclose(ll, want[:, 2], rtol=1e-12)
            np.testing.assert_allclose(lp, want[:, 0], rtol=1e-12)
            assert np.array_equal(pr, want[:, 1])
            sim = PE.simulate_batch(th[:2]).cpu().numpy()
            np.testing.assert_allclose(sim, np.hstack([(th[:2] @ Phi.T)[:, r * n:(r + 1) * n] for r in keep]), rtol=1e-12)
    res = PE.doMetropolisHastings()
    assert len(res) == 7 and PE.post_burn_in_samples.shape[1] == 3 and np.all(np.isfinite(PE.post_burn_in_log_posteriors_un_normed_vec))


def test_transient_kinetics_response_transforms(env, golden):
    """responses['data_overcategory'] = 'transient_kinetics' (transform_responses, IP.py:609-658): rate-type responses are integrated
    over their abscissa (explicit Euler from 0; an intermediate's rate twice) on the observed side once and on the simulated side per
    evaluation (peuqse_transform_transient, sequential per row = the reference's loop).  Against the unmodified reference
    (tests/golden/transient.npz): one product rate; an intermediate rate + a reactant concentration + a product rate on three unevenly
    spaced abscissas.  A built-in model goes through an inner engine and gives what the same function as a torch simulator gives."""
    PEUQSE, orc, _, torch = env
    g = golden("transient")
    t, th = g['tr_t'], g['tr_thetas']
    T = torch.from_numpy(t).cuda()

    def rates(p, resp):
        a, b = p[:, :1], p[:, 1:2]
        r = [a * torch.exp(-b * T[0][None, :]), 1.0 - 0.5 * a * T[1][None, :] / (1.0 + b * T[1][None, :]), a * b * torch.exp(-b * T[2][None, :])]
        return torch.cat([r[i] for i in resp], dim=1)
    for tag, resp, types, dtypes in (('P_r', [0], ['P'], ['r']), ('mixed', [0, 1, 2], ['I', 'R', 'P'], ['r', 'c', 'r'])):
        y, sig = g['tr_%s_y' % tag], g['tr_%s_sig' % tag]
        one = len(resp) == 1
        PE = make_pe(PEUQSE, PEUQSE.TorchSimulator(lambda p, resp=resp: rates(p, resp)), list(y[0]) if one else [list(v) for v in y],
                     list(sig[0]) if one else [list(v) for v in sig], [2.0, 1.5], [0.5, 0.4],
                     settings={'responses_abscissa': list(t[resp[0]]) if one else [list(t[r]) for r in resp],
                               'data_overcategory': 'transient_kinetics', 'response_types': types, 'response_data_types': dtypes,
                               'mcmc_length': 400, 'mcmc_burn_in': 40, 'mcmc_relative_step_length': 0.05, 'mcmc_random_seed': 1,
                               'exportLog': False})
        for a, b in zip(PE.UserInput.responses_observed_transformed, g['tr_%s_obs_transformed' % tag]):
            assert np.array_equal(a, b)
        for a, b in zip(PE.observed_responses_covmat_transformed, g['tr_%s_cov_transformed' % tag]):
            assert np.array_equal(a, b)
        lp, pr, ll = parts(PE, th)
        want = g['tr_%s_vals' % tag]
        np.testing.assert_allclose(ll, want[:, 2], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(lp, want[:, 0], rtol=1e-12, atol=1e-12)
        assert np.array_equal(pr, want[:, 1])
        # the transformed simulation is the oracle's sequential Euler sum (torch's exp differs from numpy's in the last ulp)
        sim = PE.simulate_batch(th[:3]).cpu().numpy()
        raw = rates(torch.from_numpy(th[:3]).cuda(), resp).cpu().numpy()
        n = t.shape[1]
        for i in range(3):
            st, _ = orc.transform_responses([t[r] for r in resp], types, dtypes, raw[i] if one else [raw[i][j * n:(j + 1) * n] for j in range(len(resp))])
            assert np.array_equal(sim[i], np.concatenate(st))
    res = PE.doMetropolisHastings()
    assert len(res) == 7 and np.all(np.isfinite(PE.post_burn_in_log_posteriors_un_normed_vec))
    # built-in model == the same function as a torch simulator
    rng = np.random.default_rng(2)
    n, k = 12, 3
    Phi = orc.cos_basis(n, k)
    tt = np.cumsum(0.1 + 0.1 * rng.random(n))
    yy = Phi @ np.array([1.0, 0.5, -0.3]) + 0.05 * rng.standard_normal(n)
    Phi_t = torch.from_numpy(Phi).cuda()
    pts = rng.standard_normal((40, k))
    vals = []
    for model in (PEUQSE.LinearModel(Phi), PEUQSE.TorchSimulator(lambda p: p @ Phi_t.T)):
        PEl = make_pe(PEUQSE, model, list(yy), list(np.full(n, 0.05)), [0.0] * k, [2.0] * k,
                      settings={'responses_abscissa': list(tt), 'data_overcategory': 'transient_kinetics', 'response_types': ['I'],
                                'response_data_types': ['r']})
        vals.append(parts(PEl, pts)[0])
    np.testing.assert_allclose(vals[0], vals[1], rtol=1e-12)
