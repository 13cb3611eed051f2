"""CPU-only tests: host-side normalisation, the C-ABI library (load + exported symbols, no compute calls),
the UserInput mirror, and the multi-rank reduction logic over gloo with world_size 2."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

X00 = [600, 1100, 1400]
Y00 = [360500, 580500, 1620500]
S00 = [200000, 300000, 200000]


def test_library_loads_and_exports_every_declared_symbol():
    from peuqse_b200 import _lib
    L = _lib.load()
    hdr = open(os.path.join(REPO, "include", "peuqse_b200.h")).read()
    declared = set(re.findall(r"\b(peuqse_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"peuqse_problem_desc", "peuqse_sweep_spec", "peuqse_best", "peuqse_mh_spec", "peuqse_ess_spec"}
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/peuqse_b200.h but not exported"
    assert set(_lib.EXPORTS) == declared
    assert L.peuqse_version() >= 100


def test_ctypes_struct_layouts_match_header_sizes():
    from peuqse_b200 import _lib
    # sizes implied by the header (LP64): checked so that a field added on one side only is caught
    assert ctypes.sizeof(_lib.Best) == 16
    assert ctypes.sizeof(_lib.SweepSpec) == 8 + 16 * 8 * 3 + 8 + 16 * 8 + 8
    assert ctypes.sizeof(_lib.MHSpec) == 7 * 8 + 8 + 256 * 8 + 8
    assert ctypes.sizeof(_lib.ESSSpec) == 5 * 8 + 4 * 4
    assert ctypes.sizeof(_lib.ProblemDesc) == 16 + 32 + 8 + 24 + 8 + 8 + 24 + 32 + 8 + 8 + 8


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import peuqse_b200 as PEUQSE
    UI = PEUQSE.UserInput.fresh()
    UI.responses['responses_observed'] = Y00
    UI.responses['responses_observed_uncertainties'] = S00
    UI.model['InputParameterPriorValues'] = [200, 500]
    UI.model['InputParametersPriorValuesUncertainties'] = [100, 200]
    UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.QuadraticModel(X00)
    with pytest.raises(PEUQSE.PeuqseError):
        PEUQSE.parameter_estimation(UI)


def test_product_does_not_import_the_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "peuqse_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("peuqse_oracle", "oracle") or "import oracle" not in src
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_host_normalisation_matches_oracle_constants():
    """problem.py (product) and oracle.build_problem restate IP.py:146-350 independently: same constants."""
    from peuqse_b200 import _lib
    from peuqse_b200.problem import build_host_problem
    from oracle import peuqse_oracle as orc
    cases = [
        dict(mu=[200, 500], unc=[100, 200]),
        dict(mu=[200, 500], unc=[100, -1], ub=[None, 900], lb=[None, 100]),
        dict(mu=[1.0, 2.0], unc=[[1, .3], [.3, 4]]),
        dict(mu=[200, 500], unc=[100, 200], scaling="off"),
        dict(mu=[200, 500], unc=[100, 200], scaling=3.0),
        dict(mu=[200, 500], unc=[100, 200], simunc=[1e5, 5e4, 2e5]),
    ]
    for c in cases:
        H = build_host_problem(c['mu'], c['unc'], Y00, S00, lower_bounds=c.get('lb', ()), upper_bounds=c.get('ub', ()),
                               scaling_type=c.get('scaling', 'std'), simulation_uncertainties=c.get('simunc'))
        P = orc.build_problem(c['mu'], c['unc'], Y00, S00, lower_bounds=c.get('lb', ()), upper_bounds=c.get('ub', ()),
                              scaling_type=c.get('scaling', 'std'), simulation_uncertainties=c.get('simunc'))
        assert np.array_equal(H.mu_prior_scaled, P.mu_prior_scaled)
        assert np.array_equal(H.covmat_prior_scaled, P.covmat_prior_scaled)
        if P.prior_wh is not None:
            assert H.prior_c == P.prior_wh.rank * orc.LOG2PI + P.prior_wh.log_pdet
            if P.prior_wh.is_perm:
                assert H.prior_mode == _lib.PRIOR_DIAG and np.array_equal(H.prior_w, P.prior_wh.w)
                assert np.array_equal(H.prior_gidx, P.gauss_idx[P.prior_wh.perm])
            else:
                assert H.prior_mode == _lib.PRIOR_DENSE and np.array_equal(H.prior_U, P.prior_wh.U)
        blk = P.responses[0]
        assert H.resp_mode[0] == _lib.RESP_MVN_DIAG
        assert np.array_equal(H.w, blk.wh.w) and np.array_equal(H.y, blk.y[blk.wh.perm])
        assert H.resp_c[0] == blk.wh.rank * orc.LOG2PI + blk.wh.log_pdet
    # per-value path and full covariance
    n = 150
    y = np.linspace(1, 2, n)
    H = build_host_problem([0.0], [1.0], y, np.full(n, 0.1))
    P = orc.build_problem([0.0], [1.0], y, np.full(n, 0.1))
    assert H.resp_mode[0] == _lib.RESP_PER_VALUE
    assert np.array_equal(H.w, P.responses[0].w_val) and np.array_equal(H.cval, P.responses[0].c_val)
    rng = np.random.default_rng(0)
    A = rng.standard_normal((6, 6))
    cov = A @ A.T + np.eye(6)
    H = build_host_problem([0.0], [1.0], [y[:6], y[6:12]], [cov, cov])
    assert list(H.resp_mode) == [_lib.RESP_MVN_FULL] * 2 and H.resp_U[0].shape == (6, 6)
    with pytest.raises(NotImplementedError):
        build_host_problem([0.0], [1.0], y[:6], cov)        # single-response full covariance: dead code in the reference


def test_userinput_mirror_has_reference_keys():
    import peuqse_b200.UserInput as UI
    for d, keys in [('responses', ['responses_abscissa', 'responses_observed', 'responses_observed_uncertainties',
                                   'responses_observed_max_covmat_size']),
                    ('model', ['InputParameterPriorValues', 'InputParametersPriorValuesUncertainties', 'InputParameterInitialGuess',
                               'simulateByInputParametersOnlyFunction', 'InputParameterPriorValues_upperBounds',
                               'InputParameterPriorValues_lowerBounds', 'responses_simulation_uncertainties']),
                    ('parameter_estimation_settings',
                     ['scaling_uncertainties_type', 'mcmc_random_seed', 'mcmc_length', 'mcmc_burn_in', 'mcmc_relative_step_length',
                      'mcmc_modulate_accept_probability', 'mcmc_nwalkers', 'mcmc_maxiter', 'mcmc_threshold_filter_samples',
                      'multistart_searchType', 'multistart_numStartPoints', 'multistart_initialPointsDistributionType',
                      'multistart_relativeInitialDistributionSpread', 'multistart_gridsearchSamplingInterval',
                      'multistart_gridsearchSamplingRadii', 'multistart_permutationsToSamples',
                      'multistart_permutationsToSamples_threshold_filter_coefficient'])]:
        for k in keys:
            assert k in getattr(UI, d), (d, k)
    a, b = UI.fresh(), UI.fresh()
    a.model['InputParameterPriorValues'] = [1]
    assert b.model['InputParameterPriorValues'] == []
    assert UI.parameter_estimation_settings['multistart_relativeInitialDistributionSpread'] == 0.866
    assert UI.responses['responses_observed_max_covmat_size'] == 100


def test_merge_best_rule():
    from peuqse_b200.distributed import merge_best, shard_range
    assert merge_best([(-3.0, 10), (-1.0, 50), (-1.0, 20), (float('nan'), 1), (float('-inf'), 0), (-2.0, -1)]) == (-1.0, 20)
    assert merge_best([(float('-inf'), -1), (float('nan'), 3)]) == (float('-inf'), -1)
    cover = []
    for r in range(7):
        f, c = shard_range(100020001, r, 7)
        cover.append((f, c))
    assert cover[0][0] == 0 and sum(c for _, c in cover) == 100020001
    assert all(cover[i][0] + cover[i][1] == cover[i + 1][0] for i in range(6))
    # interleaved pieces (multistart_shard_pieces): every index exactly once, piece j of rank r is global piece r + j*size
    from peuqse_b200.sweep import interleaved_ranges
    for total, size, pieces in ((100020001, 8, 8), (10, 4, 5), (7, 3, 1)):
        allr = sorted(x for r in range(size) for x in interleaved_ranges(total, r, size, pieces) if x[1] > 0)
        assert allr[0][0] == 0 and sum(c for _, c in allr) == total
        assert all(allr[i][0] + allr[i][1] == allr[i + 1][0] for i in range(len(allr) - 1))
    a = interleaved_ranges(1000, 1, 4, 2)
    assert a == [(125, 125), (625, 125)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    sys.path.insert(0, REPO)
    from peuqse_b200 import distributed as D
    from oracle import peuqse_oracle as orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank evaluates its contiguous shard of the Example00 grid with the oracle (the GPU kernels' stand-in on CPU)
        c, h, r = [200., 500.], [5., 5.], [50, 50]
        N = orc.grid_size(r)
        first, count = D.shard_range(N)
        P = orc.build_problem([200, 500], [100, 200], Y00, S00)
        _, bt = orc.sim_quadratic(X00)
        th = orc.grid_decode(c, h, r, np.arange(first, first + count))
        lp = orc.log_p_batch(P, th, bt)[0]
        i, v = orc.first_max_np(lp)
        best_v, best_i = D.allgather_best(v, first + i if i >= 0 else -1)
        w, thr = orc.permutation_weights(lp, best_v)
        shift = orc.grid_decode(c, h, r, np.array([best_i]))[0]
        sums = np.concatenate([[w.sum()], (w[:, None] * (th - shift)).sum(axis=0)])
        sums = D.allreduce_np(sums, "sum")
        kept = w > 0
        mn = D.allreduce_np(th[kept].min(axis=0) if kept.any() else np.full(2, np.inf), "min")
        mx = D.allreduce_np(th[kept].max(axis=0) if kept.any() else np.full(2, -np.inf), "max")
        cnt = D.allreduce_int_np(np.array([int(kept.sum())]))
        rows = D.allgather_rows(torch.from_numpy(th[:3 + rank]))
        # ensemble slice sampling v2: a batch of walker positions evaluated shard-wise, gathered back in order (7 rows over 2
        # ranks: ragged shards of 3 and 4); a batch smaller than the per-rank minimum is evaluated whole on every rank
        calls = []
        fn = lambda r: (calls.append(r.shape[0]), orc_rows(r))[1]
        orc_rows = lambda r: torch.from_numpy(orc.log_p_batch(P, r.numpy(), bt)[0])
        common = orc.grid_decode(c, h, r, np.arange(0, 7))            # the same batch on every rank
        sh = D.sharded_rows(fn, torch.from_numpy(common.copy()))
        sh_small = D.sharded_rows(fn, torch.from_numpy(common.copy()), min_rows_per_rank=4)
        if rank == 0:
            np.savez(out, best_v=best_v, best_i=best_i, W=sums[0], mu=shift + sums[1:] / sums[0], mn=mn, mx=mx, cnt=cnt,
                     nrows=rows.shape[0], sharded=sh.numpy(), sharded_small=sh_small.numpy(), shard_calls=np.array(calls),
                     sharded_want=orc.log_p_batch(P, orc.grid_decode(c, h, r, np.arange(0, 7)), bt)[0])
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_reductions_match_single_process(tmp_path, golden):
    import torch.multiprocessing as mp
    out = str(tmp_path / "r.npz")
    port = _free_port()
    mp.spawn(_rank_main, args=(2, port, out), nprocs=2, join=True)
    res = np.load(out)
    g = golden("example00")
    assert int(res['best_i']) == 13 and float(res['best_v']) == float(g['ex00_grid_map_logP'])
    assert int(res['W']) == int(g['ex00_grid_nrows_expanded'])
    np.testing.assert_allclose(res['mu'], g['ex00_grid_mu_AP'], rtol=1e-12)
    assert int(res['cnt'][0]) == int((g['ex00_grid_multiplicity'] > 0).sum())
    assert int(res['nrows']) == 3 + 4
    assert np.array_equal(res['sharded'], res['sharded_want']) and np.array_equal(res['sharded_small'], res['sharded_want'])
    assert list(res['shard_calls']) == [3, 7]                     # rank 0: its 3-row shard, then the whole small batch


def _export_stub(golden):
    """An object carrying exactly the state the exporters read, filled from the reference run stored in the fixture."""
    import types
    from peuqse_b200 import UserInput as UIM
    g = golden("export_formats")
    UI = UIM.fresh()
    UI.responses['responses_abscissa'] = [600, 1100, 1400]
    UI.std_prior = np.array([100., 200.])
    UI.mu_prior = np.array([200., 500.])
    UI.model['exportResponses'] = False
    pe = types.SimpleNamespace(UserInput=UI)
    return g, UI, pe


def test_export_post_burn_in_statistics_matches_reference_files(golden, tmp_path):
    """exportPostBurnInStatistics (IP.py:1956-2001): text and CSV files byte for byte, pickles object for object, against
    the files the unmodified reference wrote from the same state (tests/golden/make_golden_export.py)."""
    import pickle
    from peuqse_b200 import export as EX
    g, UI, pe = _export_stub(golden)
    for name in ('post_burn_in_samples', 'post_burn_in_log_posteriors_un_normed_vec', 'during_burn_in_samples',
                 'during_burn_in_log_posteriors_un_normed_vec', 'map_logP', 'map_parameter_set', 'mu_AP_parameter_set',
                 'stdap_parameter_set', 'mcmc_last_point_sampled'):
        setattr(pe, name, g['mh_state__' + name])
    pe.map_index = int(g['mh_state__map_index'])
    pe.info_gain = float(g['mh_state__info_gain'])
    pe.evidence = np.float64(g['mh_state__evidence'])
    UI.InputParameterInitialGuess = g['mh_state__initial_guess']
    EX.exportPostBurnInStatistics(pe)
    keys = [k for k in g.files if k.startswith('mh__')]
    assert len(keys) == 10
    for key in keys:
        _, sub, fname = key.split('__', 2)
        got = open(os.path.join(sub, fname), 'rb').read()
        ref = bytes(g[key])
        if fname.endswith('.pkl'):
            a, b = pickle.loads(got), pickle.loads(ref)
            if isinstance(b, list):
                assert len(a) == len(b)
                for x, y in zip(a, b):
                    assert np.array_equal(np.asarray(x), np.asarray(y))
            else:
                assert np.array_equal(np.asarray(a), np.asarray(b))
        else:
            assert got == ref, fname
    assert sorted(os.listdir('pickles')) == sorted(k.split('__', 2)[2] for k in keys if '__pickles__' in k)
    # round trip through reload_samples' reader
    arr = np.genfromtxt(os.path.join('logs_and_csvs', 'mcmc_logP_and_parameter_samples.csv'), delimiter=",")
    assert np.array_equal(arr[:, 1:], pe.post_burn_in_samples) and np.array_equal(arr[:, 0:1], pe.post_burn_in_log_posteriors_un_normed_vec)
    assert np.array_equal(EX.unpickleAnObject('pickles/mcmc_last_point_sampled'), pe.mcmc_last_point_sampled)


def test_export_multistart_matches_reference_files(golden, tmp_path):
    import pickle
    import types
    from peuqse_b200 import export as EX
    g, UI, pe = _export_stub(golden)
    pe.sweep_result = types.SimpleNamespace(total=25)
    pe.map_logP = float(g['ms_state__map_logP'][0])
    pe.map_index = 1
    pe.map_parameter_set = g['ms_state__map_parameter_set']
    pe.listOfPermutations = g['ms_state__list']
    pe._perm_array = pe.permutations_MAP_logP_and_parameters_values = g['ms_state__perm_array']
    pe.getSimulatedResponses = lambda th: g['ms_state__map_sim']
    EX.exportMultistart(pe, np.array([200., 500.]), [pe.map_parameter_set, None, None], False)
    keys = [k for k in g.files if k.startswith('ms__')]
    assert len(keys) == 5
    for key in keys:
        _, sub, fname = key.split('__', 2)
        got = open(os.path.join(sub, fname), 'rb').read()
        if fname.endswith('.pkl'):
            assert np.array_equal(pickle.loads(got), pickle.loads(bytes(g[key])))
        else:
            assert got == bytes(g[key]), fname
    # with the weighted statistics of a permutationsToSamples run the two extra lines of IP.py:1050-1051 follow
    EX.exportMultistart(pe, np.array([200., 500.]), [pe.map_parameter_set, g['ms_state__mu_AP'], g['ms_state__stdap']], True)
    txt = open('logs_and_csvs/multistart_log_file.txt').read()
    assert txt.startswith(bytes(g['ms__logs_and_csvs__multistart_log_file.txt']).decode())
    assert "self.mu_AP_parameter_set : " + str(g['ms_state__mu_AP']) in txt and "self.stdap_parameter_set : " in txt


def test_lockstep_nelder_mead_equals_scipy():
    """peuqse_b200/optimize.py restates scipy's _minimize_neldermead for many start points at once: same optimum, same
    function value, same iteration and evaluation counts as a stand-alone scipy run of every start point, including runs cut
    short by maxiter and by the default evaluation budget."""
    import scipy.optimize
    from peuqse_b200.optimize import nelder_mead_batch
    rng = np.random.default_rng(0)

    def rosen(x):
        return 100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2 + (0 if len(x) < 3 else np.sum((x[2:] - 0.3) ** 4))
    for N, B, kw in ((2, 30, {}), (3, 12, {}), (2, 8, {'maxiter': 17}), (4, 6, {})):
        X0 = rng.normal(0, 2, (B, N))
        X0[0, 0] = 0.0                                   # a zero coordinate takes the absolute initial step
        r = nelder_mead_batch(lambda P: np.array([rosen(p) for p in P]), X0, **kw)
        for i in range(B):
            s = scipy.optimize.minimize(rosen, X0[i], method='Nelder-Mead', options=({'maxiter': kw['maxiter']} if kw else None))
            assert np.array_equal(s.x, r['x'][i]) and s.fun == r['fun'][i] and s.nit == r['nit'][i] and s.nfev == r['nfev'][i]


def test_reduced_space_wrappers_on_the_host():
    """The simulator wrappers of model['reducedParameterSpace'] (IP.py:574-606) and responses['reducedResponseSpace'] (IP.py:483-499) are
    plain torch code: reduced value j goes to full index idx[j], the other parameters come from the initial guess, priors / bounds / names
    are selected by membership in their original order; the response wrapper keeps the listed responses' columns."""
    import torch
    import peuqse_b200 as PEUQSE
    from peuqse_b200.inverse_problem import parameter_estimation
    UI = PEUQSE.UserInput.fresh()
    UI.model['InputParameterPriorValues'] = [1.0, 2.0, 3.0, 4.0]
    UI.model['InputParametersPriorValuesUncertainties'] = [0.1, 0.2, 0.3, 0.4]
    UI.model['InputParameterPriorValues_lowerBounds'] = [0.0, None, 0.0, 0.0]
    UI.model['InputParameterInitialGuess'] = [1.5, 2.5, 3.5, 4.5]
    UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.TorchSimulator(lambda th: th * 10.0)
    UI.model['simulationOutputProcessingFunction'] = lambda out: out + 1.0
    m = parameter_estimation._reduced_model(UI, [2, 0])
    assert m['InputParameterPriorValues'] == [1.0, 3.0] and list(m['InputParametersPriorValuesUncertainties']) == [0.1, 0.3]
    assert m['InputParameterPriorValues_lowerBounds'] == [0.0, 0.0] and m['InputParameterPriorValues_upperBounds'] == []
    assert m['InputParameterInitialGuess'] == [1.5, 3.5] and list(m['parameterNamesAndMathTypeExpressionsDict']) == ['0', '2']
    assert m['reducedParameterSpace'] == [] and m['simulationOutputProcessingFunction'] is None
    out = m['simulateByInputParametersOnlyFunction'](torch.tensor([[7.0, 8.0], [9.0, 6.0]], dtype=torch.float64))
    assert torch.equal(out, torch.tensor([[81.0, 26.0, 71.0, 46.0], [61.0, 26.0, 91.0, 46.0]], dtype=torch.float64))
    assert UI.model['reducedParameterSpace'] == [] and len(UI.model['InputParameterPriorValues']) == 4      # the user's dictionary is untouched
    cov = np.diag([0.1, 0.2, 0.3, 0.4]) ** 2
    cov[0, 2] = cov[2, 0] = 0.01
    UI.model['InputParametersPriorValuesUncertainties'] = cov
    m = parameter_estimation._reduced_model(UI, [0, 2])
    assert np.array_equal(m['InputParametersPriorValuesUncertainties'], cov[np.ix_([0, 2], [0, 2])])
    UI.responses['responses_observed'] = [[1.0, 2.0], [3.0, 4.0, 5.0], [6.0]]
    sub = parameter_estimation._response_subset_simulator(UI, PEUQSE.TorchSimulator(lambda th: torch.arange(12.0, dtype=torch.float64).reshape(2, 6)),
                                                          lambda o: o * 2.0, [2, 0])
    assert torch.equal(sub(torch.zeros((2, 1), dtype=torch.float64)), torch.tensor([[0.0, 2.0, 10.0], [12.0, 14.0, 22.0]], dtype=torch.float64))


def test_transient_transform_properties():
    """Size-independent properties of the 'transient_kinetics' transform (IP.py:609-658, 3501-3525): the Euler integral starts at 0, is
    linear in the response, equals the prefix sum of v[j] (t[j+1] - t[j]); two passes for an intermediate's rate; concentrations of
    reactants / products are untouched; the propagated uncertainty never decreases along the abscissa and starts at mean(sigma) / 10."""
    from peuqse_b200.problem import transform_transient, transient_passes
    assert [transient_passes(a, b) for a, b in (('I', 'c'), ('I', 'r'), ('R', 'c'), ('R', 'r'), ('P', 'c'), ('P', 'r'), ('O', 'o'), ('T', 'r'))] == \
        [1, 2, 0, 1, 0, 1, 0, 0]
    rng = np.random.default_rng(0)
    for n in (2, 3, 17, 400):
        t = np.cumsum(0.01 + rng.random(n))
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        sig = 0.1 + rng.random(n)
        one = lambda v, ty='P', dt='r': transform_transient([t], [ty], [dt], v, sig)
        tx, ux = one(x)
        assert tx[0][0] == 0.0 and np.allclose(tx[0][1:], np.cumsum(x[:-1] * np.diff(t)), rtol=1e-13, atol=1e-13)
        assert np.allclose(one(2.0 * x - 3.0 * y)[0][0], 2.0 * tx[0] - 3.0 * one(y)[0][0], rtol=1e-12, atol=1e-12)
        assert ux[0][0] == np.mean(sig) / 10 and np.all(np.diff(ux[0]) >= 0)
        twice = one(x, 'I', 'r')[0][0]
        assert np.allclose(twice, one(tx[0], 'P', 'r')[0][0], rtol=1e-13, atol=1e-13)
        same, usame = one(x, 'R', 'c')
        assert np.array_equal(same[0], x) and np.array_equal(usame[0], sig)
