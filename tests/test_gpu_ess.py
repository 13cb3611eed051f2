"""Ensemble slice sampling pinned step for step (run with -m gpu).

zeus-mcmc, the engine behind the reference's doEnsembleSliceSampling (IP.py:2495-2518), is a third-party package that is
neither vendored nor installed, so parity with zeus itself stays unpinned.  What IS pinned here: the three implementations
of the published algorithm in this repository -- oracle.ess_run_philox (numpy), the per-walker kernels (csrc/ess.cu) and
the round-synchronous device driver (csrc/ess_rounds.cu, natively for built-in models and step by step from Python for
torch-callable simulators) -- consume the same Philox draws and must produce IDENTICAL walkers, log-posteriors, expansion /
contraction counts and tuned scale factors, iteration by iteration."""
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
    return PEUQSE, orc, torch


def make_pe(PEUQSE, model, observed, unc, mu, punc, **settings):
    UI = PEUQSE.UserInput.fresh()
    UI.responses['responses_observed'] = observed
    UI.responses['responses_observed_uncertainties'] = unc
    UI.model['InputParameterPriorValues'] = mu
    UI.model['InputParametersPriorValuesUncertainties'] = punc
    UI.model['simulateByInputParametersOnlyFunction'] = model
    for k, v in settings.items():
        (UI.responses if k in UI.responses else UI.parameter_estimation_settings)[k] = v
    return PEUQSE.parameter_estimation(UI)


def same_run(a, b, what):
    assert np.array_equal(a['chain'].cpu().numpy(), b['chain'].cpu().numpy()), what + ": walkers differ"
    assert np.array_equal(a['chain_logp'].cpu().numpy(), b['chain_logp'].cpu().numpy()), what + ": log-posteriors differ"
    assert np.array_equal(a['info'].cpu().numpy()[:5], b['info'].cpu().numpy()[:5]), what + ": counters / mu differ"


def same_as_oracle(run, ref, W, what, logp_rtol=0.0):
    chain, chain_lp, info = ref
    got = run['chain'].cpu().numpy()
    assert np.array_equal(got, chain), what + ": walkers differ from the oracle (first differing iteration %s)" % (
        np.nonzero((got != chain).any(axis=(1, 2)))[0][:1],)
    if logp_rtol == 0.0:
        assert np.array_equal(run['chain_logp'].cpu().numpy(), chain_lp), what + ": log-posteriors differ from the oracle"
    else:
        np.testing.assert_allclose(run['chain_logp'].cpu().numpy(), chain_lp, rtol=logp_rtol)
    ih = run['info'].cpu().numpy()
    assert int(ih[0]) == info['neval'] + W and int(ih[2]) == info['n_expansions'] and int(ih[3]) == info['n_contractions'], what
    assert ih[1] == info['mu'], what + ": tuned scale factor differs"


def test_ess_three_drivers_and_oracle_identical_small(env):
    """Linear model, diagonal responses (n = 30: the oracle's log-posterior is bit-equal to the kernels'), 32 walkers x 40
    iterations: per-walker kernels == round-synchronous driver == rounds issued from Python around a torch-callable
    simulator == numpy oracle, to the last bit."""
    PEUQSE, orc, torch = env
    from peuqse_b200.mcmc import ess_launch, ESS_DRIVER_WALKER, ESS_DRIVER_ROUNDS
    rng = np.random.default_rng(5)
    k, n, sig, W, iters, seed = 4, 30, 0.3, 32, 40, 1234567
    Phi = orc.cos_basis(n, k)
    y = Phi @ rng.standard_normal(k) + sig * rng.standard_normal(n)
    PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), list(y), list(np.full(n, sig)), [0.0] * k, [2.0] * k)
    start = 0.5 * rng.standard_normal((W, k))
    a = ess_launch(PE, start, iters, seed, driver=ESS_DRIVER_WALKER)
    b = ess_launch(PE, start, iters, seed, driver=ESS_DRIVER_ROUNDS)
    same_run(a, b, "per-walker kernels vs round-synchronous driver")
    # torch-callable simulator with the built-in model's operation order (parameter by parameter, multiply then add)
    PhiT = torch.from_numpy(Phi.T.copy()).cuda()

    def simulate(th):
        s = th[:, 0:1] * PhiT[0]
        for d in range(1, k):
            s = s + th[:, d:d + 1] * PhiT[d]
        return s
    PEx = make_pe(PEUQSE, PEUQSE.TorchSimulator(simulate), list(y), list(np.full(n, sig)), [0.0] * k, [2.0] * k)
    c = ess_launch(PEx, start, iters, seed)
    same_run(a, c, "built-in model vs torch-callable simulator")
    P = orc.build_problem([0.0] * k, [2.0] * k, y, np.full(n, sig))
    _, bt = orc.sim_linear(Phi)
    ref = orc.ess_run_philox(lambda th: orc.log_p_batch(P, th, bt)[0], start, iters, seed)
    same_as_oracle(a, ref, W, "per-walker kernels")
    same_as_oracle(b, ref, W, "round-synchronous driver")
    assert ref[2]['n_expansions'] > 0 and ref[2]['n_contractions'] > 0 and ref[2]['mu'] != 1.0
    # continuing a run: the draws are addressed by the iteration number (first_iter), so 25 + 15 iterations == 40
    full = ess_launch(PE, start, 40, seed, driver=ESS_DRIVER_ROUNDS, tune=False, mu0=0.7)
    p1 = ess_launch(PE, start, 25, seed, driver=ESS_DRIVER_ROUNDS, tune=False, mu0=0.7)
    p2 = ess_launch(PE, p1['X'].cpu().numpy(), 15, seed, driver=ESS_DRIVER_WALKER, tune=False, mu0=0.7, first_iter=25)
    assert np.array_equal(p2['chain'].cpu().numpy(), full['chain'].cpu().numpy()[25:])


def test_ess_odd_sizes_and_bounds(env):
    """Walker counts that do not fill a CTA or a warp, a uniform (bounded) parameter so that some candidate points have
    logP = -inf, W = 6 (three walkers per half: the smallest ensemble with distinct direction pairs)."""
    PEUQSE, orc, torch = env
    from peuqse_b200.mcmc import ess_launch, ESS_DRIVER_WALKER, ESS_DRIVER_ROUNDS
    rng = np.random.default_rng(9)
    k, n, sig = 3, 12, 0.5
    Phi = orc.cos_basis(n, k)
    y = Phi @ np.array([0.3, -0.2, 0.1]) + sig * rng.standard_normal(n)
    kw = dict(InputParameterPriorValues_lowerBounds=[-1.0, None, None], InputParameterPriorValues_upperBounds=[1.0, None, None])
    UI = PEUQSE.UserInput.fresh()
    UI.responses['responses_observed'] = list(y)
    UI.responses['responses_observed_uncertainties'] = list(np.full(n, sig))
    UI.model['InputParameterPriorValues'] = [0.0, 0.0, 0.0]
    UI.model['InputParametersPriorValuesUncertainties'] = [-1, 2.0, 2.0]
    UI.model['InputParameterPriorValues_lowerBounds'] = kw['InputParameterPriorValues_lowerBounds']
    UI.model['InputParameterPriorValues_upperBounds'] = kw['InputParameterPriorValues_upperBounds']
    UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.LinearModel(Phi)
    PE = PEUQSE.parameter_estimation(UI)
    P = orc.build_problem([0.0, 0.0, 0.0], [-1, 2.0, 2.0], y, np.full(n, sig), lower_bounds=[-1.0, None, None],
                          upper_bounds=[1.0, None, None])
    _, bt = orc.sim_linear(Phi)
    for W in (6, 70, 600):
        start = 0.4 * rng.standard_normal((W, k))
        start[:, 0] = np.clip(start[:, 0], -0.9, 0.9)
        a = ess_launch(PE, start, 12, 77, driver=ESS_DRIVER_WALKER)
        b = ess_launch(PE, start, 12, 77, driver=ESS_DRIVER_ROUNDS)
        same_run(a, b, "W = %d" % W)
        ref = orc.ess_run_philox(lambda th: orc.log_p_batch(P, th, bt)[0], start, 12, 77)
        same_as_oracle(b, ref, W, "W = %d" % W)
        assert np.all(np.abs(b['chain'].cpu().numpy()[:, :, 0]) <= 1.0)


def test_ess_warp_shape_drivers_identical(env):
    """n = 1024 (the C3 model): the per-walker kernel gives a warp to every walker, the round-synchronous driver a warp to
    every candidate row; both reduce the 1024 terms in the same order."""
    PEUQSE, orc, torch = env
    from peuqse_b200.mcmc import ess_launch, ESS_DRIVER_WALKER, ESS_DRIVER_ROUNDS
    n = 1024
    model = PEUQSE.Tpd2SiteModel.linear_ramp(n)
    truth = np.array([41.5, 45, 13, 13, 0.1, 0.1])
    kBT, c1, c2 = orc.tpd2site_constants(n)
    sim_truth = orc.sim_tpd2site(kBT, c1, c2)[0](truth)
    sig = 0.02 * np.max(sim_truth)
    y = sim_truth + sig * np.random.default_rng(0).standard_normal(n)
    PE = make_pe(PEUQSE, model, list(y), list(np.full(n, sig)), [41.5, 41.5, 13.0, 13.0, 0.1, 0.1], [20, 20, 13, 13, 0.1, 0.1])
    start = truth + 1e-3 * np.array([20, 20, 13, 13, 0.1, 0.1]) * np.random.default_rng(2).standard_normal((24, 6))
    a = ess_launch(PE, start, 10, 9, driver=ESS_DRIVER_WALKER)
    b = ess_launch(PE, start, 10, 9, driver=ESS_DRIVER_ROUNDS)
    same_run(a, b, "warp shape")


def _fullcov_case(orc, k, n, seed):
    rng = np.random.default_rng(seed)
    x = np.linspace(0, 1, n)
    Phi = orc.cos_basis(2 * n, k)
    Sig = 0.1 ** 2 * np.exp(-(x[:, None] - x[None, :]) ** 2 / (2 * 0.02 ** 2)) + 0.05 ** 2 * np.eye(n)
    truth = rng.standard_normal(k)
    y = Phi @ truth + 0.1 * rng.standard_normal(2 * n)
    return Phi, Sig, truth, y


def test_ess_full_covariance_rounds_vs_oracle_small(env):
    """Two responses with dense 24 x 24 covariances (the C4 path at a small size): the round-synchronous driver over the DMMA
    contraction against the oracle with scipy's hoisted whitening.  Positions depend on the log-posteriors only through
    comparisons, so the walkers are identical; the log-posteriors agree to 1e-12 relative (dense summation order)."""
    PEUQSE, orc, torch = env
    from peuqse_b200.mcmc import ess_launch
    k, n, W = 3, 24, 32
    Phi, Sig, truth, y = _fullcov_case(orc, k, n, 21)
    PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), [list(y[:n]), list(y[n:])], [Sig, Sig], [0.0] * k, [2.0] * k)
    P = orc.build_problem([0.0] * k, [2.0] * k, [y[:n], y[n:]], [Sig, Sig])
    _, lin = orc.sim_linear(Phi)
    bt = lambda th: [lin(th)[:, :n], lin(th)[:, n:]]
    start = truth + 0.3 * np.random.default_rng(3).standard_normal((W, k))
    run = ess_launch(PE, start, 30, 11)
    ref = orc.ess_run_philox(lambda th: orc.log_p_batch(P, th, bt)[0], start, 30, 11)
    same_as_oracle(run, ref, W, "full covariance", logp_rtol=1e-12)


def test_ess_c4_size_pinned(env):
    """Config C4 at its named size (2048 walkers, 12 parameters, two responses with 4096 x 4096 covariances), 2 ensemble
    iterations.  (i) The driver logic at this size: oracle.ess_run_philox replays the run with the DEVICE log-posterior as
    its evaluator -- every draw, mask, interval update, acceptance and the tuned mu must match bit for bit.  (ii) One
    iteration with the oracle's own CPU evaluator (scipy's whitening hoisted): identical walkers, log-posteriors to 1e-12."""
    PEUQSE, orc, torch = env
    from peuqse_b200.mcmc import ess_launch
    k, n, W = 12, 4096, 2048
    Phi, Sig, truth, y = _fullcov_case(orc, k, n, 1)
    PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), [list(y[:n]), list(y[n:])], [Sig, Sig], [0.0] * k, [2.0] * k,
                 responses_observed_max_covmat_size=4096)
    start = truth + 0.02 * np.random.default_rng(3).standard_normal((W, k))
    run = ess_launch(PE, start, 2, 7)
    dev_logp = lambda th: PE.getLogP_batch(np.ascontiguousarray(th)).cpu().numpy()
    ref = orc.ess_run_philox(dev_logp, start, 2, 7)
    same_as_oracle(run, ref, W, "C4 size, device evaluator")
    assert ref[2]['neval'] > 4 * W
    # (ii) CPU evaluator: y - sim whitened by U = u sqrt(1/s) of scipy's eigendecomposition (scipy _multivariate.py:208-236, 590)
    wh = orc.psd_whitening(Sig)
    c = wh.rank * orc.LOG2PI + wh.log_pdet
    Pp = orc.build_problem([0.0] * k, [2.0] * k, np.zeros(3), np.ones(3))          # the prior part only

    def cpu_logp(th):
        sim = th @ Phi.T
        ll = np.zeros(len(th))
        for r in range(2):
            t = (y[r * n:(r + 1) * n] - sim[:, r * n:(r + 1) * n]) @ wh.U
            ll = ll + (-0.5 * (c + np.sum(t * t, axis=1)))
        return ll + orc.log_prior_batch(Pp, th)
    run1 = ess_launch(PE, start, 1, 7)
    ref1 = orc.ess_run_philox(cpu_logp, start, 1, 7)
    same_as_oracle(run1, ref1, W, "C4 size, CPU evaluator", logp_rtol=1e-12)


def test_do_ensemble_slice_sampling_reproducible(env):
    """A seeded doEnsembleSliceSampling is reproducible (scipy's scrambled Sobol start points do not follow np.random.seed:
    they are seeded explicitly), through the per-walker kernels and through the round-synchronous driver."""
    PEUQSE, orc, torch = env
    rng = np.random.default_rng(5)
    k, n, sig = 4, 30, 0.3
    Phi = orc.cos_basis(n, k)
    y = Phi @ rng.standard_normal(k) + sig * rng.standard_normal(n)
    runs = []
    for _ in range(2):
        PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), list(y), list(np.full(n, sig)), [0.0] * k, [2.0] * k, mcmc_length=16 * 50,
                     mcmc_nwalkers=16, mcmc_burn_in=16 * 10, mcmc_random_seed=3, mcmc_threshold_filter_samples=False)
        PE.doEnsembleSliceSampling()
        runs.append(PE.post_burn_in_samples.copy())
    assert np.array_equal(runs[0], runs[1])


def test_multistart_of_ensemble_slice_runs(env):
    """doMultiStart(searchType='doEnsembleSliceSampling') (IP.py:832-836, 880-895, 933-947, 971-990): one ensemble per start point,
    walkers spread around it, 2k walkers by default; the start with the highest logP wins and its run is what a direct
    doEnsembleSliceSampling from that start (same seed) produces."""
    PEUQSE, orc, torch = env
    rng = np.random.default_rng(6)
    k, n, sig = 3, 30, 0.3
    Phi = orc.cos_basis(n, k)
    y = Phi @ rng.standard_normal(k) + sig * rng.standard_normal(n)
    settings = dict(mcmc_length=6 * 60, mcmc_burn_in=6 * 10, mcmc_random_seed=3, mcmc_threshold_filter_samples=False,
                    mcmc_info_gain_returned='log_ratio', multistart_searchType='doEnsembleSliceSampling',
                    multistart_initialPointsDistributionType='grid', multistart_gridsearchSamplingInterval=[0.5, 0.5, 0.5],
                    multistart_gridsearchSamplingRadii=[1, 0, 1], multistart_keep_cumulative_post_burn_in_data=True, exportLog=False)
    runs = []
    for _ in range(2):
        PE = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), list(y), list(np.full(n, sig)), [0.0] * k, [2.0] * k, **settings)
        best = PE.doMultiStart()
        runs.append((PE.permutations_MAP_logP_and_parameters_values.copy(), PE.post_burn_in_samples.copy()))
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    rows = runs[0][0]
    pts = PE.listOfPermutations
    assert rows.shape == (9, 1 + k) and pts.shape == (9, k) and PE.mcmc_nwalkers == 2 * k
    i = int(np.argmax(rows[:, 0]))
    assert PE.map_logP == rows[i, 0] and np.array_equal(PE.map_parameter_set, rows[i, 1:]) and len(best) == 7
    assert PE.post_burn_in_samples.shape == (6 * 50, k) and PE.cumulative_post_burn_in_samples.shape == (9 * 6 * 50, k)
    assert np.array_equal(PE.UserInput.InputParameterInitialGuess, [0.0] * k)
    assert len(np.unique(rows[:, 0])) == 9                               # every start ran its own ensemble
    one = make_pe(PEUQSE, PEUQSE.LinearModel(Phi), list(y), list(np.full(n, sig)), [0.0] * k, [2.0] * k,
                  **dict(settings, mcmc_random_seed=3 + 1000003 * i, mcmc_nwalkers=2 * k))
    one.UserInput.InputParameterInitialGuess = pts[i]
    one.doEnsembleSliceSampling()
    assert np.array_equal(one.post_burn_in_samples, PE.post_burn_in_samples)
    assert abs(float(np.ravel(one.map_logP)[0]) - PE.map_logP) == 0
