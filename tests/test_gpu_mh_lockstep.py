"""GPU parity tests of the lock-step Metropolis-Hastings form (csrc/mh_steps.cu, peuqse_mhs_*): the sampler of
doMetropolisHastings (IP.py:2563-2754) for simulators that are torch callables, custom_logPrior / custom_logLikelihood and
full-covariance responses.  Pinned three ways: against the persistent kernel on a built-in model (identical trajectories),
against the numpy oracle replaying the exported draws, and through doMetropolisHastings() itself."""
import numpy as np
import pytest

from test_gpu_parity import make_pe, X00, Y00, S00, env  # noqa: F401

pytestmark = pytest.mark.gpu


def _mh_out(out):
    return {n: out[n].cpu().numpy() for n in ('samples', 'logp', 'logprior', 'loglik', 'deltas', 'us', 'stats')}


def test_lockstep_walks_the_persistent_kernels_trajectory(env):
    """Built-in model, both forms on the same (seed, chain) streams: identical rows, draws and statistics blocks -- unbiased mode with
    thinning and the accept-probability flattening, 'MAP_finding' with its step-size adaptation, and a continued run."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    PE = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200])
    start = np.array([200., 500.]) + np.random.default_rng(1).standard_normal((7, 2)) * [30, 60]
    chol = step_cholesky(PE.Q_covmat, 0.2)
    for kw in (dict(thin=1), dict(thin=3, modulate=4.0), dict(map_finding=True), dict(first_step=777, chain_offset=5, thin=2)):
        a = _mh_out(mh_launch(PE, start, 700, 60, 11, chol, export_draws=True, lockstep=False, **kw))
        b = _mh_out(mh_launch(PE, start, 700, 60, 11, chol, export_draws=True, lockstep=True, **kw))
        for n in ('deltas', 'us', 'samples', 'logp', 'logprior', 'loglik'):
            assert np.array_equal(a[n], b[n]), (kw, n)
        np.testing.assert_allclose(b['stats'], a['stats'], rtol=1e-12, atol=0, err_msg=str(kw))
        assert 0.05 < a['stats'][:, 0].mean() / 699 < 0.95
    # the remaining steps run as replays of ONE captured step (CUDA graph); the eager loop gives the same rows
    g = mh_launch(PE, start, 700, 60, 11, chol, lockstep=True)
    assert g['graph'] is True, g.get('graph_error')
    PE.UserInput.parameter_estimation_settings['mcmc_lockstep_graph'] = False
    e = mh_launch(PE, start, 700, 60, 11, chol, lockstep=True)
    assert e['graph'] is False and torch.equal(e['samples'], g['samples']) and torch.equal(e['stats'], g['stats'])
    PE.UserInput.parameter_estimation_settings['mcmc_lockstep_graph'] = 'auto'
    # a non-diagonal proposal covariance (full lower-triangular factor)
    Lf = np.linalg.cholesky(np.array([[400., 300.], [300., 1600.]]))
    a = _mh_out(mh_launch(PE, start, 300, 0, 3, Lf, export_draws=True, lockstep=False))
    b = _mh_out(mh_launch(PE, start, 300, 0, 3, Lf, export_draws=True, lockstep=True))
    assert np.array_equal(a['samples'], b['samples']) and np.array_equal(a['logp'], b['logp'])
    # no stored rows: the statistics block alone
    a = mh_launch(PE, start, 300, 40, 3, chol, store_samples=False, lockstep=False)['stats'].cpu().numpy()
    b = mh_launch(PE, start, 300, 40, 3, chol, store_samples=False, lockstep=True)['stats'].cpu().numpy()
    np.testing.assert_allclose(b, a, rtol=1e-12, atol=0)


def test_lockstep_torch_simulator_custom_functions_and_full_covariance_against_the_oracle(env):
    """What the persistent kernel cannot evaluate: the exported draws replayed through the oracle's literal loop give the same
    accept / reject decisions and the same chains (1e-12: the simulator is a torch matmul, the full covariance a dense contraction)."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200.mcmc import mh_launch, mh_needs_lockstep, step_cholesky
    rng = np.random.default_rng(8)
    k, n, sig = 4, 24, 0.2
    Phi = orc.cos_basis(n, k)
    truth = rng.standard_normal(k)
    y = Phi @ truth + sig * rng.standard_normal(n)
    Phi_t = torch.from_numpy(Phi).cuda()
    PhiT, yt = torch.from_numpy(Phi.T.copy()).cuda(), torch.from_numpy(y).cuda()
    LOG2PI = float(np.log(2 * np.pi))
    x = np.linspace(0, 1, n // 2)
    covs = [sig ** 2 * np.exp(-(x[:, None] - x[None, :]) ** 2 / (2 * ell ** 2)) + 0.05 ** 2 * np.eye(n // 2) for ell in (0.1, 0.3)]

    def my_prior(th):
        z = th / 2.0
        return -0.5 * (k * LOG2PI + (z * z).sum(dim=1))

    def my_lik(th):
        sim = th @ PhiT
        r = (yt - sim) / sig
        return -0.5 * (n * LOG2PI + n * float(np.log(sig ** 2)) + (r * r).sum(dim=1)), sim
    _, bt = orc.sim_linear(Phi)
    P_diag = orc.build_problem([0.0] * k, [2.0] * k, y, np.full(n, sig))
    P_full = orc.build_problem([0.0] * k, [2.0] * k, [y[:n // 2], y[n // 2:]], covs, max_covmat_size=600)
    bt_full = lambda t: [bt(t)[:, :n // 2], bt(t)[:, n // 2:]]
    diag_obs, full_obs = (list(y), list(np.full(n, sig))), ([list(y[:n // 2]), list(y[n // 2:])], covs)
    cases = [("torch simulator", PEUQSE.TorchSimulator(lambda t: t @ Phi_t.T), diag_obs, {}, P_diag, bt),
             ("custom prior", PEUQSE.LinearModel(Phi), diag_obs, dict(custom_logPrior=my_prior), P_diag, bt),
             ("custom likelihood", None, diag_obs, dict(custom_logLikelihood=my_lik), P_diag, bt),
             ("both custom", None, diag_obs, dict(custom_logPrior=my_prior, custom_logLikelihood=my_lik), P_diag, bt),
             ("full covariance, built-in model", PEUQSE.LinearModel(Phi), full_obs, {}, P_full, bt_full),
             ("full covariance, torch simulator", PEUQSE.TorchSimulator(lambda t: t @ Phi_t.T), full_obs, {}, P_full, bt_full)]
    start = truth + 0.05 * rng.standard_normal((5, k))
    for name, model, (obs, unc), kw, P, sim in cases:
        PE = make_pe(PEUQSE, model, obs, unc, [0.0] * k, [2.0] * k,
                     settings={'responses_observed_max_covmat_size': 600, 'mcmc_length': 400, 'mcmc_burn_in': 50, 'mcmc_nchains': 3,
                               'mcmc_relative_step_length': 0.002, 'mcmc_random_seed': 4, 'mcmc_threshold_filter_samples': False,
                               'mcmc_info_gain_returned': 'log_ratio', 'exportLog': False}, **kw)
        assert mh_needs_lockstep(PE), name
        raw = mh_launch(PE, start, 250, 30, 6, step_cholesky(PE.Q_covmat, 0.002), export_draws=True)
        assert raw['graph'] is True, (name, raw.get('graph_error'))           # every one of these is capturable
        out = _mh_out(raw)
        ref = orc.mh_chains_batch(P, start, out['deltas'], out['us'], sim)
        acc = np.any(out['samples'][1:] != out['samples'][:-1], axis=2)
        assert np.array_equal(acc, ref['accepted'][1:]), name
        assert 0.05 < acc.mean() < 0.95, name
        np.testing.assert_allclose(out['samples'], ref['samples'], rtol=1e-12, err_msg=name)
        np.testing.assert_allclose(out['logp'], ref['logP'], rtol=1e-11, atol=1e-12, err_msg=name)
        np.testing.assert_allclose(out['loglik'][1:], ref['log_likelihoods'][1:], rtol=1e-11, atol=1e-12, err_msg=name)
        np.testing.assert_allclose(out['logprior'][1:], ref['log_priors'][1:], rtol=1e-12, err_msg=name)
        # the reference-named call: three chains, stacked chain-major, statistics from the device kernels
        res = PE.doMetropolisHastings()
        assert len(res) == 7 and PE.post_burn_in_samples.shape == (3 * 350, k), name
        assert np.all(np.isfinite(PE.post_burn_in_log_posteriors_un_normed_vec)), name
        direct = mh_launch(PE, np.tile([0.0] * k, (3, 1)), 400, 50, 4, step_cholesky(PE.Q_covmat, 0.002))
        want = direct['samples'][50:].permute(1, 0, 2).reshape(-1, k).cpu().numpy()
        assert np.array_equal(PE.post_burn_in_samples, want), name
        assert np.allclose(PE.mu_AP_parameter_set, want.mean(axis=0), rtol=1e-10), name


def test_lockstep_multistart_of_chains(env):
    """doMultiStart(searchType='doMetropolisHastings') (IP.py:865-879, 975-990) through the lock-step form: the start points are the
    chains of one run; same per-start MAP rows and the same best start as through the persistent kernel."""
    PEUQSE, orc, _, torch = env
    settings = {'mcmc_length': 400, 'mcmc_burn_in': 50, 'mcmc_relative_step_length': 0.1, 'mcmc_random_seed': 3,
                'mcmc_info_gain_returned': 'log_ratio', 'multistart_searchType': 'doMetropolisHastings',
                'multistart_initialPointsDistributionType': 'grid', 'multistart_gridsearchSamplingInterval': [60, 120],
                'multistart_gridsearchSamplingRadii': [1, 1], 'multistart_keep_cumulative_post_burn_in_data': True,
                'responses_abscissa': X00}
    A = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200], settings=dict(settings))
    A.doMultiStart()
    B = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200], settings=dict(settings, mcmc_lockstep=True))
    B.doMultiStart()
    assert np.array_equal(A.permutations_MAP_logP_and_parameters_values, B.permutations_MAP_logP_and_parameters_values)
    assert np.array_equal(A.map_parameter_set, B.map_parameter_set) and A.map_logP == B.map_logP
    assert np.array_equal(A.post_burn_in_samples, B.post_burn_in_samples)


def test_lockstep_falls_back_to_the_eager_loop_when_the_simulator_cannot_be_captured(env):
    """A simulator that synchronises with the host (here: .item()) cannot be recorded into a CUDA graph: the driver notices, keeps the
    state and finishes the run with the eager loop -- same chain as a capturable simulator of the same function."""
    PEUQSE, orc, _, torch = env
    from peuqse_b200.mcmc import mh_launch, step_cholesky
    X = torch.tensor(X00, dtype=torch.float64, device='cuda')
    f = lambda t: (X[None, :] - t[:, :1]) ** 2 + t[:, 1:2]

    def f_sync(t):
        float(t.sum().item())                                          # a host round trip
        return f(t)
    runs = []
    for fn in (f, f_sync):
        PE = make_pe(PEUQSE, PEUQSE.TorchSimulator(fn), Y00, S00, [200, 500], [100, 200])
        runs.append(mh_launch(PE, np.tile([200., 500.], (3, 1)), 120, 10, 2, step_cholesky(PE.Q_covmat, 0.2)))
    assert runs[0]['graph'] is True and runs[1]['graph'] is False and 'graph_error' in runs[1]
    assert torch.equal(runs[0]['samples'], runs[1]['samples']) and torch.equal(runs[0]['logp'], runs[1]['logp'])
    # and the chain is the built-in model's (same function): same accept / reject decisions
    PEb = make_pe(PEUQSE, PEUQSE.QuadraticModel(X00), Y00, S00, [200, 500], [100, 200])
    b = mh_launch(PEb, np.tile([200., 500.], (3, 1)), 120, 10, 2, step_cholesky(PEb.Q_covmat, 0.2))
    np.testing.assert_allclose(runs[0]['samples'].cpu().numpy(), b['samples'].cpu().numpy(), rtol=1e-12)
