"""Pins oracle/ against the reference: stored CSV rows of the reference repo and fixtures produced by
running the unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import peuqse_oracle as orc
from oracle import philox_np

X00 = [600, 1100, 1400]
Y00 = [360500, 580500, 1620500]
S00 = [200000, 300000, 200000]


def ex00(**kw):
    return orc.build_problem(kw.pop('mu', [200, 500]), kw.pop('unc', [100, 200]), Y00, S00, **kw)


def check_vals(P, sim_point, sim_batch, thetas, vals, exact=True, rtol=1e-13):
    """vals[:, 0:3] = reference getLogP, getLogPrior, getLogLikelihood."""
    lp_b, pr_b, ll_b = orc.log_p_batch(P, thetas, sim_batch)
    for i, th in enumerate(thetas):
        p = orc.log_p_point(P, th, sim_point)
        assert p == vals[i, 0] or (np.isnan(p) and np.isnan(vals[i, 0])), (i, p, vals[i, 0])
        assert orc.log_prior_point(P, th) == vals[i, 1]
    if exact:
        assert np.array_equal(lp_b, vals[:, 0])
        assert np.array_equal(pr_b, vals[:, 1])
        assert np.array_equal(ll_b, vals[:, 2])
    else:
        fin = np.isfinite(vals[:, 0])
        assert np.array_equal(np.isfinite(lp_b), fin)
        np.testing.assert_allclose(lp_b[fin], vals[fin, 0], rtol=rtol)
        np.testing.assert_allclose(ll_b, vals[:, 2], rtol=rtol)
        np.testing.assert_allclose(pr_b, vals[:, 1], rtol=rtol)


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert [int(x) for x in philox_np.philox4x32_10([0, 0, 0, 0], [0, 0])] == \
        [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert [int(x) for x in philox_np.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)] == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert [int(x) for x in philox_np.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                                                    [0xa4093822, 0x299f31d0])] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    ua, ub = philox_np.uniform_pair(7, 2, np.arange(10000), 0)
    assert ua.min() >= 0 and ua.max() < 1 and abs(ua.mean() - 0.5) < 0.02 and abs(ub.mean() - 0.5) < 0.02
    z, u = philox_np.mh_draws(0, np.arange(4000)[:, None], np.arange(1, 6)[None, :], 6)
    assert z.shape == (4000, 5, 6) and abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02


def test_example00_stored_csv_rows(golden):
    """The reference repo's own stored (logP, a, b) rows: Examples/Example00/logs_and_csvs/
    multistart_MAP_logP_and_parameters_values.csv (bit-exact)."""
    g = golden("example00")
    P = ex00()
    pt, bt = orc.sim_quadratic(X00)
    rows = g['ex00_csv_rows']
    for r in rows:
        assert orc.log_p_point(P, r[1:], pt) == r[0]
    assert np.array_equal(orc.log_p_batch(P, rows[:, 1:], bt)[0], rows[:, 0])
    # SURVEY 8c closed-form known answers
    assert orc.log_p_point(P, [200, 500], pt) == -42.81726459961094
    assert orc.log_p_point(P, [150, 400], pt) == -42.797773932944274
    assert orc.log_p_point(P, [255, 139153], pt) == -240350.85630424548


def test_example00_points(golden):
    g = golden("example00")
    pt, bt = orc.sim_quadratic(X00)
    check_vals(ex00(), pt, bt, g['ex00_thetas'], g['ex00_logP_prior_lik'])


def test_simulation_uncertainties_as_a_function_vs_reference(golden):
    """responses_simulation_uncertainties = f(theta) (tests/golden/make_golden_simunc.py ran the unmodified reference): the
    restatement reproduces the reference's likelihood bit for bit, one logpdf per response and one per value."""
    g = golden("simunc")
    pt, _ = orc.sim_quadratic(X00)
    x = np.asarray(X00, dtype=float)
    sd = lambda th: 0.03 * np.abs((x - th[0]) ** 2 + th[1]) + 50.0 * (1.0 + np.abs(th[1]) / 500.0)
    for tag, maxcov in (('mvn', 100), ('per_value', 2)):
        P = orc.build_problem([200, 500], [100, 200], Y00, S00, max_covmat_size=maxcov)
        rows = g['simunc_' + tag]
        for th, row in zip(g['simunc_thetas'], rows):
            ll = orc.log_likelihood_simvar_point(P, pt(tuple(th)), sd(th))
            assert ll == row[2] and orc.log_prior_point(P, th) == row[1]
            assert (ll + row[1] == row[0]) or (row[1] == float('-inf') and row[0] == float('-inf'))


def test_reduced_parameter_space_vs_reference(golden):
    """model['reducedParameterSpace'] (IP.py:519-561, 574-606; tests/golden/make_golden_reduced.py ran the unmodified reference): the
    problem built from the selected priors with a simulator that fills the other parameters from the initial guess is the reference's."""
    g = golden("reduced")
    pt, _ = orc.sim_quadratic(X00)
    guess = g['reduced_guess']
    for tag, idx in (('b', [1]), ('a', [0]), ('ba', [1, 0])):
        keep = sorted(idx)            # returnReducedIterable keeps the ORIGINAL order of the selected priors (IP.py:3559-3564) ...
        P = orc.build_problem(list(np.array([200., 500.])[keep]), list(np.array([100., 200.])[keep]), Y00, S00)

        def sim(th, idx=idx):         # ... while the simulation wrapper places value j at index idx[j] (IP.py:585-589)
            full = guess.copy()
            full[idx] = th
            return pt(tuple(full))
        for th, row in zip(g['reduced_%s_thetas' % tag], g['reduced_%s_vals' % tag]):
            assert orc.log_prior_point(P, th) == row[1]
            assert orc.log_likelihood_point(P, sim(th)) == row[2]
            assert orc.log_p_point(P, th, sim) == row[0]


def test_reduced_response_space_vs_reference(golden):
    """responses['reducedResponseSpace'] (IP.py:470-512; tests/golden/make_golden_reduced_responses.py ran the unmodified reference):
    the kept responses' blocks of the problem shaped on the FULL response set give the reference's likelihood bit for bit -- with one
    response left its sigmas are still used as variances, the multi-response behaviour of returnShapedResponseCovMat -- and the host
    problem of the package selects the same blocks."""
    import dataclasses
    from peuqse_b200.problem import build_host_problem
    g = golden("reduced_responses")
    Phi, y, sig, th = g['rr_Phi'], g['rr_y'], g['rr_sig'], g['rr_thetas']
    n = 6
    pt, _ = orc.sim_linear(Phi)
    obs = [list(y[r * n:(r + 1) * n]) for r in range(3)]
    unc = [list(sig[r * n:(r + 1) * n]) for r in range(3)]
    full = orc.build_problem([0.0, 0.0, 0.0], [2.0, 2.0, 3.0], obs, unc)
    for tag, keep in (('02', [0, 2]), ('1', [1])):
        P = dataclasses.replace(full, responses=[full.responses[r] for r in keep])
        want = g['rr_%s_vals' % tag]
        for t, row in zip(th, want):
            sim = pt(tuple(t))
            assert orc.log_likelihood_point(P, [sim[r * n:(r + 1) * n] for r in keep]) == row[2]
            assert orc.log_prior_point(P, t) == row[1]
        H = build_host_problem([0.0, 0.0, 0.0], [2.0, 2.0, 3.0], obs, unc, keep_responses=keep)
        assert len(H.responses_observed) == len(keep) and H.n_total == n * len(keep)
        for j, r in enumerate(keep):
            assert np.array_equal(H.responses_observed[j], y[r * n:(r + 1) * n]) and np.array_equal(H.responses_cov[j], sig[r * n:(r + 1) * n])


def test_transient_kinetics_transforms_vs_reference(golden):
    """responses['data_overcategory'] = 'transient_kinetics' (transform_responses, IP.py:609-658; tests/golden/make_golden_transient.py
    ran the unmodified reference): transformed observations, transformed covariances and the likelihood of transformed simulations,
    bit for bit; the package's host-side restatement (peuqse_b200/problem.py) gives the same arrays."""
    from peuqse_b200.problem import transform_transient
    g = golden("transient")
    t, th = g['tr_t'], g['tr_thetas']
    rates = lambda p: [p[0] * np.exp(-p[1] * t[0]), 1.0 - 0.5 * p[0] * t[1] / (1.0 + p[1] * t[1]), p[0] * p[1] * np.exp(-p[1] * t[2])]
    for tag, resp, types, dtypes in (('P_r', [0], ['P'], ['r']), ('mixed', [0, 1, 2], ['I', 'R', 'P'], ['r', 'c', 'r'])):
        y, sig = g['tr_%s_y' % tag], g['tr_%s_sig' % tag]
        ab = [t[r] for r in resp]
        one = len(resp) == 1
        yt, ut = orc.transform_responses(ab, types, dtypes, y[0] if one else list(y), sig[0] if one else list(sig))
        assert np.array_equal(np.array(yt), g['tr_%s_obs_transformed' % tag])
        yp, up = transform_transient(ab, types, dtypes, y[0] if one else list(y), sig[0] if one else list(sig))
        assert all(np.array_equal(a, b) for a, b in zip(yp, yt)) and all(np.array_equal(a, b) for a, b in zip(up, ut))
        P = orc.build_problem([2.0, 1.5], [0.5, 0.4], yt[0] if one else yt, ut[0] if one else ut)
        assert all(np.array_equal(np.asarray(b.cov), c) for b, c in zip(P.responses, g['tr_%s_cov_transformed' % tag]))
        for p_, row in zip(th, g['tr_%s_vals' % tag]):
            sims = [rates(p_)[r] for r in resp]
            st, _ = orc.transform_responses(ab, types, dtypes, sims[0] if one else sims)
            assert orc.log_likelihood_point(P, st[0] if one else st) == row[2] and orc.log_prior_point(P, p_) == row[1]


def test_grid_order(golden):
    g = golden("example00")
    for name in ("g3", "g3d"):
        c, h, r = g[f'grid_{name}_center'], g[f'grid_{name}_interval'], g[f'grid_{name}_radii']
        ref = g[f'grid_{name}_list']
        lst = np.array([np.array(x, dtype=float) for x in orc.combination_generator(c, h, r)])
        assert np.array_equal(lst, ref)
        assert orc.grid_size(r) == len(ref)
        assert np.array_equal(orc.grid_decode(c, h, r, np.arange(len(ref))), ref)


def test_example00_grid_multistart(golden):
    g = golden("example00")
    arr = g['ex00_grid_logP_theta']
    P = ex00()
    _, bt = orc.sim_quadratic(X00)
    th = orc.grid_decode([200., 500.], [5., 5.], [50, 50], np.arange(10201))
    assert np.array_equal(th, arr[:, 1:])
    logp = orc.log_p_batch(P, th, bt)[0]
    assert np.array_equal(logp, arr[:, 0])
    idx, best = orc.first_max(logp)
    assert (idx, best) == orc.first_max_np(logp)
    assert idx == 13 and best == float(g['ex00_grid_map_logP']) == -42.637391007597046
    assert np.array_equal(th[idx], g['ex00_grid_map_theta'])
    w, thr = orc.permutation_weights(logp, best, 2.0)
    assert np.array_equal(w, g['ex00_grid_multiplicity'])
    assert w.sum() == int(g['ex00_grid_nrows_expanded']) == 225166
    st = orc.weighted_statistics(logp, th, w)
    np.testing.assert_allclose(st['mu_AP'], g['ex00_grid_mu_AP'], rtol=1e-12)
    np.testing.assert_allclose(st['stdap'], g['ex00_grid_stdap'], rtol=1e-11)
    np.testing.assert_allclose(st['evidence'], float(g['ex00_grid_evidence']), rtol=1e-11)
    exp = orc.convert_permutations_to_samples(arr, best, 2.0)
    assert exp.shape == (225166, 3)
    np.testing.assert_allclose(exp[:, 1:].mean(axis=0), g['ex00_grid_mu_AP'], rtol=1e-13)


def test_example26_stored_rows(golden):
    """Examples/Example26 stored multistart rows; model = Example26/simulation_model_00.py (same quadratic,
    same data), prior from its runfile: mu [200,500], sigma [100,200]."""
    g = golden("example00")
    P = ex00()
    pt, _ = orc.sim_quadratic(X00)
    for key in ("ex26_perm_csv", "ex26_ms_csv"):
        for r in g[key]:
            np.testing.assert_allclose(orc.log_p_point(P, r[1:], pt), r[0], rtol=1e-12)


def test_priors_and_bounds(golden):
    g = golden("priors")
    pt, bt = orc.sim_quadratic(X00)
    P = ex00(mu=[200, 500], unc=[100, -1], upper_bounds=[None, 900], lower_bounds=[None, 100])
    check_vals(P, pt, bt, g['unif_thetas'], g['unif_vals'])
    assert orc.log_prior_point(P, [250, 500]) == -1.0439385332046727      # SURVEY 8c (z = 0.5)
    assert orc.log_prior_point(P, [200, 900]) > -np.inf and orc.log_prior_point(P, [200, np.nextafter(900, 1e9)]) == -np.inf
    P = ex00(mu=[200, 500], unc=[-1, -1], upper_bounds=[400, 900], lower_bounds=[0, 100])
    check_vals(P, pt, bt, g['unif2_thetas'], g['unif2_vals'])
    P = ex00(upper_bounds=[300, None], lower_bounds=[100, 0])
    check_vals(P, pt, bt, g['bnd_thetas'], g['bnd_vals'])
    P = ex00(mu=[1.0, 2.0], unc=[[1, .3], [.3, 4]])
    assert orc.log_prior_point(P, [1.5, 1.0]) == -2.467189810186834
    check_vals(P, pt, bt, g['covp_thetas'], g['covp_vals'], exact=False, rtol=1e-13)
    for tag, sc in [("off", "off"), ("mu", "mu"), ("fix", 3.0)]:
        check_vals(ex00(scaling_type=sc), pt, bt, g[f'scal_{tag}_thetas'], g[f'scal_{tag}_vals'],
                   exact=False, rtol=1e-13)
    P = ex00(simulation_uncertainties=[1e5, 5e4, 2e5])
    check_vals(P, pt, bt, g['simunc_thetas'], g['simunc_vals'])


def _check_vals_loose(P, sim_point, sim_batch, thetas, vals, rtol):
    lp_b, pr_b, ll_b = orc.log_p_batch(P, thetas, sim_batch)
    np.testing.assert_allclose(lp_b, vals[:, 0], rtol=rtol)
    np.testing.assert_allclose(pr_b, vals[:, 1], rtol=rtol)
    np.testing.assert_allclose(ll_b, vals[:, 2], rtol=rtol)
    for i in range(0, len(thetas), 7):
        np.testing.assert_allclose(orc.log_p_point(P, thetas[i], sim_point), vals[i, 0], rtol=rtol)


@pytest.mark.parametrize("n", [1024, 96])
def test_tpd_points(golden, n):
    g = golden("tpd")
    kBT, c1, c2 = orc.tpd2site_constants(n)
    pt, bt = orc.sim_tpd2site(kBT, c1, c2)
    P = orc.build_problem([41.5, 41.5, 13.0, 13.0, 0.1, 0.1], [20, 20, 13, 13, 0.1, 0.1],
                          g[f'tpd{n}_y'], np.full(n, float(g[f'tpd{n}_sigma'])))
    assert P.responses[0].per_value == (n > 100)
    th, vals = g[f'tpd{n}_thetas'], g[f'tpd{n}_vals']
    lp_b, pr_b, ll_b = orc.log_p_batch(P, th, bt)
    assert np.array_equal(pr_b, vals[:, 1])
    if n > 100:      # per-value path: same sequential order -> bit exact
        assert np.array_equal(ll_b, vals[:, 2]) and np.array_equal(lp_b, vals[:, 0])
        assert vals[0, 0] == -7571.424324564999      # SURVEY 8c known answer
    else:
        np.testing.assert_allclose(ll_b, vals[:, 2], rtol=1e-14)
    assert orc.log_p_point(P, th[0], pt) == vals[0, 0]


def test_multiresponse(golden):
    g = golden("multiresponse")
    for tag, n, full in [("mr_diag", 7, False), ("mr_full5", 5, True), ("mr_full50", 50, True), ("mr_diag40", 40, False)]:
        Phi, y, unc = g[f'{tag}_Phi'], g[f'{tag}_y'], g[f'{tag}_unc']
        pt, bt = orc.sim_linear(Phi)
        P = orc.build_problem([0., 0., 0.], [2., 2., 3.], [y[:n], y[n:]], [unc[0], unc[1]])

        def spt(th, pt=pt, n=n):
            s = pt(th)
            return [s[:n], s[n:]]

        def sbt(th, bt=bt, n=n):
            s = bt(th)
            return [s[:, :n], s[:, n:]]
        _check_vals_loose(P, spt, sbt, g[f'{tag}_thetas'], g[f'{tag}_vals'], rtol=1e-13 if not full else 1e-12)


def test_uniform_multistart(golden):
    g = golden("uniform_multistart")
    k, n = 10, 64
    Phi = orc.cos_basis(n, k)
    pt, bt = orc.sim_linear(Phi)
    P = orc.build_problem([0.0] * k, [1.0] * k, g['c5_y'], np.full(n, 0.1))
    u01 = philox_np.sweep_uniforms(0, 0, 400, k)
    assert np.array_equal(u01, g['c5_u01'])
    th = orc.uniform_points(u01, 0.866, P.std_prior, np.zeros(k))
    arr = g['c5_logP_theta']
    assert np.array_equal(th, arr[:, 1:])
    logp = orc.log_p_batch(P, th, bt)[0]
    np.testing.assert_allclose(logp, arr[:, 0], rtol=1e-13)
    idx, best = orc.first_max_np(logp)
    assert np.array_equal(th[idx], g['c5_map_theta'])
    w, _ = orc.permutation_weights(arr[:, 0], float(g['c5_map_logP']), 2.0)
    st = orc.weighted_statistics(arr[:, 0], th, w)
    np.testing.assert_allclose(st['mu_AP'], g['c5_mu_AP'], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(st['stdap'], g['c5_stdap'], rtol=1e-10, atol=1e-14)


def test_mh_step_for_step_example00(golden):
    g = golden("mh_example00")
    P = ex00()
    pt, bt = orc.sim_quadratic(X00)
    deltas, us = g['mh00_deltas'], g['mh00_us']
    z, u2 = philox_np.mh_draws(0, 0, np.arange(1, 1500), 2)
    assert np.array_equal(u2, us) and np.array_equal(z * (np.sqrt(0.1) * np.array([100., 200.])), deltas)
    res = orc.mh_chains_batch(P, np.array([[200., 500.]]), deltas[:, None, :], us[:, None], bt)
    burn = 100
    assert np.array_equal(res['samples'][burn:, 0], g['mh00_post_samples'])
    assert np.array_equal(res['samples'][:burn, 0], g['mh00_burn_samples'])
    assert np.array_equal(res['logP'][burn:, 0], g['mh00_post_logP'][:, 0])
    assert np.array_equal(res['log_priors'][burn:, 0], g['mh00_post_log_priors'][:, 0])
    assert np.array_equal(res['log_likelihoods'][burn:, 0], g['mh00_post_log_likelihoods'][:, 0])
    lit = orc.mh_chain_point(P, [200., 500.], deltas[:300], us[:300], pt)
    assert np.array_equal(lit['samples'], res['samples'][:301, 0])
    assert np.array_equal(lit['accepted'], res['accepted'][:301, 0])
    st = orc.post_burn_in_statistics(res['samples'][burn:, 0], res['logP'][burn:, 0])
    assert np.array_equal(st['mu_AP'], g['mh00_mu_AP']) and np.array_equal(st['stdap'], g['mh00_stdap'])
    assert st['map_logP'] == float(g['mh00_map_logP'])
    assert np.array_equal(st['map_parameter_set'], g['mh00_map_theta'])
    assert st['evidence'] == float(g['mh00_evidence'])
    acc = res['accepted'][1:, 0].mean()
    assert 0.1 < acc < 0.95


def test_mh_step_for_step_tpd(golden):
    g = golden("tpd")
    kBT, c1, c2 = orc.tpd2site_constants(1024)
    _, bt = orc.sim_tpd2site(kBT, c1, c2)
    P = orc.build_problem([41.5, 41.5, 13.0, 13.0, 0.1, 0.1], [20, 20, 13, 13, 0.1, 0.1],
                          g['tpd1024_y'], np.full(1024, float(g['tpd1024_sigma'])))
    truth = np.array([[41.5, 45, 13, 13, 0.1, 0.1]])
    res = orc.mh_chains_batch(P, truth, g['tpdmh_deltas'][:, None, :], g['tpdmh_us'][:, None], bt)
    assert np.array_equal(res['samples'][:, 0], g['tpdmh_samples'])
    assert np.array_equal(res['logP'][:, 0], g['tpdmh_logP'][:, 0])
    assert 0.05 < res['accepted'][1:, 0].mean() < 0.95


def test_ess_linear_gaussian_statistics():
    """ESS restatement (parity unpinned): recovers the closed-form posterior of a linear-Gaussian model."""
    rng = np.random.default_rng(3)
    k, n = 4, 30
    Phi = orc.cos_basis(n, k)
    truth = rng.standard_normal(k)
    y = Phi @ truth + 0.3 * rng.standard_normal(n)
    P = orc.build_problem([0.] * k, [2.] * k, y, np.full(n, 0.3))
    _, bt = orc.sim_linear(Phi)
    mean, cov = orc.linear_gaussian_posterior(Phi, Phi / 0.09, y, 4.0)
    start = 0.866 * 2.0 * (rng.random((16, k)) * 4 - 2)
    chain, lp, neval, mu = orc.ess_run(lambda X: orc.log_p_batch(P, X, bt)[0], start, 600, rng)
    s = chain[150:].reshape(-1, k)
    se = np.sqrt(np.diag(cov) / (len(s) / 20.0))
    assert np.all(np.abs(s.mean(axis=0) - mean) < 5 * se)
    d = np.sqrt(np.diag(cov))
    np.testing.assert_allclose(np.cov(s.T) / np.outer(d, d), cov / np.outer(d, d), atol=0.15)


def test_ess_philox_restatement_statistics_and_replay():
    """The counter-based form of the same algorithm (the one the device drivers are pinned against, tests/test_gpu_ess.py):
    recovers the closed-form posterior, is a pure function of (seed, start), and continues a run through first_iter."""
    rng = np.random.default_rng(5)
    k, n, sig, W = 3, 20, 0.3, 16
    Phi = orc.cos_basis(n, k)
    y = Phi @ rng.standard_normal(k) + sig * rng.standard_normal(n)
    P = orc.build_problem([0.] * k, [2.] * k, y, np.full(n, sig))
    _, bt = orc.sim_linear(Phi)
    f = lambda X: orc.log_p_batch(P, X, bt)[0]
    start = 0.5 * rng.standard_normal((W, k))
    chain, lp, info = orc.ess_run_philox(f, start, 500, 42)
    mean, cov = orc.linear_gaussian_posterior(Phi, Phi / sig ** 2, y, 4.0)
    s = chain[100:].reshape(-1, k)
    se = np.sqrt(np.diag(cov) / (len(s) / 20.0))
    assert np.all(np.abs(s.mean(axis=0) - mean) < 5 * se)
    np.testing.assert_allclose(np.diag(np.cov(s.T)) / np.diag(cov), 1.0, atol=0.15)
    assert np.array_equal(lp[-1], f(chain[-1]))                        # stored logP belongs to the stored positions
    assert info['neval'] > 2 * W * 500 and info['n_expansions'] > 0 and info['n_contractions'] > 0
    again, _, _ = orc.ess_run_philox(f, start, 40, 42)
    other, _, _ = orc.ess_run_philox(f, start, 40, 43)
    assert np.array_equal(again, chain[:40]) and not np.array_equal(other, chain[:40])
    a, _, _ = orc.ess_run_philox(f, start, 30, 42, tune=False, mu0=0.8)
    b, _, _ = orc.ess_run_philox(f, a[19], 10, 42, tune=False, mu0=0.8, first_iter=20)
    assert np.array_equal(b, a[20:])
    with pytest.raises(ValueError):
        orc.ess_run_philox(lambda X: np.full(len(X), -np.inf), start, 1, 0)


PBS_CASES = [('auto', True, 2.0, 'auto', 'pbs_logP'), ('auto1', True, 1.0, 'auto', 'pbs_logP'), ('MAP', True, 2.0, 'MAP', 'pbs_logP'),
             ('mu_AP', True, 2.0, 'mu_AP', 'pbs_logP'), ('off', False, 2.0, 'auto', 'pbs_logP'), ('odd', True, 2.0, 'auto', 'pbs_logP_odd')]


@pytest.mark.parametrize("case", PBS_CASES, ids=[c[0] for c in PBS_CASES])
def test_post_burn_in_statistics_with_filter(golden, case):
    """calculatePostBurnInStatistics of the unmodified reference (every filter benchmark, log-ratio info gain): the
    oracle restatement keeps exactly the same rows and reproduces the statistics bit for bit (same numpy calls)."""
    name, on, coef, bench, key = case
    g = golden("post_burn_in_statistics")
    P = ex00()
    sim_point, _ = orc.sim_quadratic(X00)
    with np.errstate(all='ignore'):
        st = orc.post_burn_in_statistics_filtered(g['pbs_samples'], g[key], g['pbs_log_priors'], benchmark=bench, coef=coef,
                                                  filter_samples=on, mu_ap_logp_fn=lambda th: orc.log_p_point(P, th, sim_point))
    rows = g[f'pbs_{name}_kept_rows']
    assert np.array_equal(st['samples'], g['pbs_samples'][rows])
    assert np.array_equal(st['logP'], g[key][rows], equal_nan=True)
    assert np.array_equal(st['log_priors'], g['pbs_log_priors'][rows])
    assert np.array_equal(st['mu_AP'], g[f'pbs_{name}_mu_AP']) and np.array_equal(st['stdap'], g[f'pbs_{name}_stdap'])
    assert float(st['map_logP']) == float(g[f'pbs_{name}_map_logP']) and st['map_index'] == int(g[f'pbs_{name}_map_index'])
    assert np.array_equal(st['map_parameter_set'], g[f'pbs_{name}_map_theta'])
    assert st['evidence'] == float(g[f'pbs_{name}_evidence'])
    assert st['info_gain_log_ratio'] == float(g[f'pbs_{name}_info_gain'])
    if name == 'odd':
        assert len(rows) == 3000          # a non-negative logP makes the threshold NaN: nothing is removed (IP.py:1894, 3650)


@pytest.mark.parametrize("fmt", ["xyz", "meshgrid"])
def test_doe_info_gain_matrices(golden, fmt):
    """doeGetInfoGainMatrix / doeParameterModulationPermutationsScanner of the unmodified reference (Example13's
    configuration, injected draws; tests/golden/make_golden_doe.py) vs the oracle restatement: bit for bit."""
    g = golden("doe")
    L, BURN, STEP, DH, DS, DHU, DSU, SIG = g['settings']
    build = lambda y: orc.build_problem([DH, DS], [DHU, DSU], y, [SIG] * 3)
    sim = lambda c: orc.sim_langmuir_log(c[0], c[1])
    conds = orc.doe_conditions([500, 0.5], [100, 0.1], [2, 2], fmt)
    for filt in (False, True):
        M, _ = orc.doe_info_gain_matrix(build, sim, conds, g['theta_true'], [DH, DS], g['deltas'], g['us'], int(BURN),
                                        dict(filter_samples=filt, benchmark='auto', coef=2.0))
        assert np.array_equal(M, g[f"{fmt}_{'filtered' if filt else 'unfiltered'}_info_gain_matrix"])
    mods = np.array(orc.combination_generator([DH, DS], np.array([1, 1]) * np.array([DHU, DSU]), [1, 1]), dtype=float)   # IP.py:1506-1509
    A = np.array([orc.doe_info_gain_matrix(build, sim, conds, m, [DH, DS], g['deltas'], g['us'], int(BURN))[0] for m in mods])
    assert np.array_equal(A, g[f'scanner_{fmt}_info_gains_matrices_array'])


def test_geweke_diagnostic_vs_reference(golden):
    """geweke_diagnostic and the Geweke half of calculateAndPlotConvergenceDiagnostics of the unmodified reference
    (tests/golden/make_golden_geweke.py; statsmodels' yule_walker supplied by the oracle's restatement): bit for bit."""
    import warnings
    g = golden("geweke")
    x = g['chains_a']
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        assert np.array_equal(orc.geweke_diagnostic(x[:, 0, 0]), g['single_chain_z'], equal_nan=True)
        assert np.array_equal(orc.geweke_diagnostic(x[:333, 1, 2]), g['single_chain_z_short'], equal_nan=True)
        r = orc.geweke_convergence(x)
    assert np.array_equal(r['window_indices'], g['a_window_indices'])
    assert np.array_equal(r['final_combined_parameter_values'], g['a_final_combined'], equal_nan=True)
    assert np.array_equal(r['final_combined_parameter_percent_outlier'], g['a_percent_outlier'])
    # the closed-form Yule-Walker restatement against a direct least-squares view of the same equations
    rng = np.random.default_rng(3)
    v = np.cumsum(rng.standard_normal(400)) * 0.1 + rng.standard_normal(400)
    rho, sigma = orc.yule_walker(v, 2)
    d = v - v.mean()
    r0, r1, r2 = (d * d).sum() / 400, (d[:-1] * d[1:]).sum() / 399, (d[:-2] * d[2:]).sum() / 398
    np.testing.assert_allclose(np.array([[r0, r1], [r1, r0]]) @ rho, [r1, r2], rtol=1e-12)
    np.testing.assert_allclose(sigma ** 2, r0 - r1 * rho[0] - r2 * rho[1], rtol=1e-12)


def test_autocorr_time_restatement_known_answers():
    """oracle.autocorr_time restates zeus.autocorr.AutoCorrTime (method 'mk'), a third-party routine that is neither vendored
    by the reference nor installed here: PARITY UNPINNED.  What can be checked without zeus: the FFT autocorrelation function
    equals the direct lag sums, white noise gives tau ~ 1, AR(1) chains give (1+rho)/(1-rho), the automated window is the
    first lag M >= c tau(M), and the growing windows are those of IP.py:3918."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal(300).cumsum()
    acf = orc.autocorr_func_1d(x, norm=False)
    xc = x - x.mean()
    direct = np.array([np.dot(xc[:len(x) - l], xc[l:]) for l in range(len(x))]) / (4 * 512)
    np.testing.assert_allclose(acf, direct, rtol=1e-9, atol=1e-12 * direct[0])
    assert orc.auto_window(np.array([1.0, 1.5, 0.3, 0.2]), 5.0) == 2          # 2 >= 5 * 0.3
    assert orc.auto_window(np.array([1.0, 1.0]), 5.0) == 0                    # never reached: numpy's argmin of all-True
    R, C = 20000, 6
    rho = np.array([0.0, 0.6, 0.9])
    e = rng.standard_normal((R, C, 3))
    y = np.empty_like(e)
    y[0] = e[0]
    for t in range(1, R):
        y[t] = rho * y[t - 1] + e[t]
    r = orc.autocorr_time_convergence(y)
    assert np.array_equal(r['window_indices'], np.exp(np.linspace(0, np.log(R), 21)).astype(int)[1:]) and r['taus'].shape == (20, 3)
    np.testing.assert_allclose(r['final_parameter_values'], (1 + rho) / (1 - rho), rtol=0.08)
    np.testing.assert_allclose(r['combined'], np.prod(r['taus'], axis=1))


def test_info_gain_cutoff_vs_reference(golden):
    """mcmc_info_gain_cutoff != 0 (IP.py:1805-1843): the truncated log-ratio information gain of the unmodified reference
    (tests/golden/make_golden_cutoff.py) on the sample arrays of the statistics fixture, bit for bit."""
    g, src = golden("info_gain_cutoff"), golden("post_burn_in_statistics")
    for key in g.files:
        filt, cutoff, info_gain, evidence, n_kept = g[key]
        st = orc.post_burn_in_statistics_filtered(src['pbs_samples'], src['pbs_logP'], src['pbs_log_priors'], 'auto', 2.0,
                                                  bool(filt), info_gain_cutoff=cutoff)
        assert len(st['samples']) == int(n_kept) and st['evidence'] == evidence
        assert np.array_equal(st['info_gain_log_ratio'], info_gain, equal_nan=True), key


def test_map_finding_restatement_vs_reference(golden):
    """mcmc_mode 'MAP_finding' (IP.py:2631-2635, 2688-2713): the oracle fed the fixture's draws reproduces the unmodified
    reference's chain, log-posteriors and step-coefficient history (the reference's own coefficient is recovered from the
    covariance it handed to np.random.multivariate_normal) bit for bit."""
    g = golden("mapfinding")
    P = orc.build_problem([200, 500], [100, 200], Y00, S00)
    _, bt = orc.sim_quadratic(X00)
    for scaled, deltas in ((False, g['base_deltas']), (True, g['scaled_deltas'])):
        out = orc.mh_chain_map_finding(P, [200.0, 500.0], deltas, g['mods'], g['us'], bt, scaled=scaled)
        assert np.array_equal(out['samples'], g['samples']) and np.array_equal(out['logP'], g['logP'])
        assert np.array_equal(out['dyn'], g['dyn']) and out['dyn_final'] == float(g['dyn_final'])
    step = float(g['step'])
    want = np.array([((100.0 ** 2 * g['dyn'][i + 1]) * g['mods'][i]) * step for i in range(len(g['mods']))])
    assert np.array_equal(g['ref_cov00'], want)
    assert len(np.unique(g['dyn'])) > 5
