"""Host drivers of the persistent MCMC kernels and the post-burn-in statistics.

doMetropolisHastings (PEUQSE/InverseProblem.py:2563-2754, "IP.py"), doEnsembleSliceSampling (IP.py:2362-2547,
engine = zeus) and calculatePostBurnInStatistics (IP.py:1861-1937).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import distributed as D


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def step_cholesky(Q_covmat, step):
    """Lower Cholesky factor of the proposal covariance Q_covmat*step (IP.py:2630: proposals are
    theta + N(Q_mu=0, Q_covmat*mcmc_relative_step_length))."""
    Q = np.asarray(Q_covmat, dtype=float) * step
    if np.count_nonzero(Q - np.diag(np.diagonal(Q))) == 0:
        return np.diag(np.sqrt(np.diagonal(Q)))
    return np.linalg.cholesky(Q)


def mh_launch(pe, theta0, n_steps, burn_in, seed, step_chol, first_step=0, chain_offset=0, modulate=0.0, thin=1,
              store_samples=True, store_parts=True, export_draws=False, map_finding=False, lockstep=None):
    """C chains starting at theta0[C,k]: one call of peuqse_mh_run (the persistent kernel), or -- for what cannot run inside
    it, see mh_needs_lockstep -- the lock-step form peuqse_mhs_* with the batched log-posterior between its two kernels per
    step.  Same draws, same arithmetic, same outputs.  Returns a dict of device tensors."""
    L = _lib.load()
    dev = _dev()
    if lockstep is None:
        lockstep = mh_needs_lockstep(pe)
    th0 = torch.as_tensor(np.asarray(theta0, dtype=float), dtype=torch.float64).reshape(-1, pe.k).to(dev).contiguous()
    Cn, k = th0.shape
    spec = _lib.MHSpec()
    spec.n_chains, spec.n_steps, spec.burn_in, spec.first_step = Cn, int(n_steps), int(burn_in), int(first_step)
    spec.chain_offset, spec.seed, spec.modulate, spec.k, spec.thin = int(chain_offset), int(seed) & (2 ** 64 - 1), float(modulate), k, int(thin)
    spec.map_finding = 1 if map_finding else 0
    flat = np.asarray(step_chol, dtype=float).reshape(-1)
    for i, v in enumerate(flat):
        spec.step_chol[i] = float(v)
    R = (int(n_steps) + thin - 1) // thin
    lay = _lib.mh_stat_layout(k)
    out = {}
    out['samples'] = torch.empty((R, Cn, k), dtype=torch.float64, device=dev) if store_samples else None
    out['logp'] = torch.empty((R, Cn), dtype=torch.float64, device=dev) if store_samples else None
    out['logprior'] = torch.empty((R, Cn), dtype=torch.float64, device=dev) if (store_samples and store_parts) else None
    out['loglik'] = torch.empty((R, Cn), dtype=torch.float64, device=dev) if (store_samples and store_parts) else None
    out['deltas'] = torch.empty((max(n_steps - 1, 0), Cn, k), dtype=torch.float64, device=dev) if export_draws else None
    # accept uniforms; with map_finding (accept uniform, modulation uniform + 0.5) per step
    us_shape = (max(n_steps - 1, 0), Cn, 2) if map_finding else (max(n_steps - 1, 0), Cn)
    out['us'] = torch.empty(us_shape, dtype=torch.float64, device=dev) if export_draws else None
    out['stats'] = torch.empty((Cn, lay['SIZE']), dtype=torch.float64, device=dev)
    out['theta0'] = th0
    if lockstep:
        return _mh_lockstep(pe, L, spec, th0, out)
    _lib.check(L.peuqse_mh_run(pe.device_problem.handle, C.byref(spec), _lib.ptr(th0), _lib.ptr(out['samples']),
                               _lib.ptr(out['logp']), _lib.ptr(out['logprior']), _lib.ptr(out['loglik']),
                               _lib.ptr(out['deltas']), _lib.ptr(out['us']), _lib.ptr(out['stats']), _lib.current_stream()))
    pe.gpu_launches += 1
    return out


def mh_needs_lockstep(pe):
    """What the persistent kernel cannot evaluate itself: a simulator that is a torch callable, custom_logPrior /
    custom_logLikelihood (IP.py:2757-2759, 2854-2857), a response with a full covariance matrix; or on request
    (parameter_estimation_settings['mcmc_lockstep'] = True)."""
    from .models import BuiltinModel
    return bool(not isinstance(pe.simulator, BuiltinModel) or getattr(pe, 'custom_logLikelihood', None) is not None
                or getattr(pe, 'custom_logPrior', None) is not None
                or np.any(np.asarray(pe.device_problem.H.resp_mode) == _lib.RESP_MVN_FULL)
                or pe.UserInput.parameter_estimation_settings.get('mcmc_lockstep', False))


def _mh_lockstep(pe, L, spec, th0, out):
    """doMetropolisHastings' loop (IP.py:2626-2675) with every chain's step taken together: k_mhs_propose, the batched
    log-prior / log-likelihood of the proposals (pe.getLogP_batch: built-in kernels, a torch simulator, the user's callables),
    k_mhs_accept.  Everything stays on the device and on one stream; nothing is read back inside the loop.  The step index lives
    in the device state (counter mode), so after one eager step the propose -> evaluate -> accept sequence is captured ONCE as a
    CUDA graph and replayed for the remaining steps (parameter_estimation_settings['mcmc_lockstep_graph'] = False keeps the eager
    loop; a simulator or callable that cannot be captured -- it synchronises, say -- falls back to it by itself)."""
    Cn, k = th0.shape
    dev = th0.device
    n = int(spec.n_steps)
    state = torch.empty((Cn, int(L.peuqse_mhs_state_doubles(k))), dtype=torch.float64, device=dev)
    prop = torch.empty((Cn, k), dtype=torch.float64, device=dev)
    rows = (_lib.ptr(out['samples']), _lib.ptr(out['logp']), _lib.ptr(out['logprior']), _lib.ptr(out['loglik']))

    def step():
        st = _lib.current_stream()
        _lib.check(L.peuqse_mhs_propose(C.byref(spec), -1, _lib.ptr(state), _lib.ptr(out['stats']), _lib.ptr(prop),
                                        _lib.ptr(out['deltas']), _lib.ptr(out['us']), st))
        _, lp, ll = pe.getLogP_batch(prop, return_parts=True)
        _lib.check(L.peuqse_mhs_accept(C.byref(spec), -1, _lib.ptr(prop), _lib.ptr(lp), _lib.ptr(ll), _lib.ptr(state),
                                       _lib.ptr(out['stats']), *rows, st))
        return lp, ll
    _, lp, ll = pe.getLogP_batch(th0, return_parts=True)
    _lib.check(L.peuqse_mhs_begin(C.byref(spec), _lib.ptr(th0), _lib.ptr(lp), _lib.ptr(ll), _lib.ptr(state), _lib.ptr(out['stats']), *rows,
                                  _lib.current_stream()))
    done = 1
    if n > 1:
        step()                                            # eager: allocator pools, library workspaces and the user's callables warm up
        done = 2
    use_graph = pe.UserInput.parameter_estimation_settings.get('mcmc_lockstep_graph', 'auto')
    out['graph'] = False
    if use_graph is not False and str(use_graph).lower() != 'false' and n - done >= 16:
        launches = pe.gpu_launches
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                keep = step()                             # noqa: F841  (the graph's pool owns these buffers)
            per_step = pe.gpu_launches - launches
            for _ in range(n - done):
                g.replay()
            pe.gpu_launches = launches + per_step * (n - done)
            done = n
            out['graph'] = True
        except Exception as e:                            # not capturable: the eager loop below takes over from the same state
            torch.cuda.synchronize()
            pe.gpu_launches = launches
            out['graph_error'] = repr(e)
    for _ in range(done, n):
        step()
    _lib.check(L.peuqse_mhs_finish(C.byref(spec), _lib.ptr(state), _lib.ptr(out['stats']), _lib.current_stream()))
    pe.gpu_launches += 2 * n                              # begin, finish, two kernels per step; getLogP_batch counts its own
    return out



class _RowsOnDevice:
    """Placeholder for an (N,1) array whose rows are still on the device (shape only)."""

    def __init__(self, n):
        self.shape = (int(n), 1)


def chain_moments(stats, theta0, k, n_rows):
    """Pooled mean / covariance / MAP over chains from the per-chain statistics blocks (host, tiny)."""
    lay = _lib.mh_stat_layout(k)
    st = stats.cpu().numpy()
    th0 = theta0.cpu().numpy()
    Cn = st.shape[0]
    S1 = st[:, lay['S1']:lay['S1'] + k]
    S2 = st[:, lay['S2']:lay['S2'] + k * k].reshape(Cn, k, k)
    ref = D.broadcast_np(th0[0])                                       # one origin for every rank's sums: rank 0's first chain
    sh = th0 - ref                                                     # shift of each chain's origin
    # sums about the common origin `ref`
    T1 = (S1 + n_rows * sh).sum(axis=0)
    T2 = (S2 + S1[:, :, None] * sh[:, None, :] + sh[:, :, None] * S1[:, None, :] + n_rows * sh[:, :, None] * sh[:, None, :]).sum(axis=0)
    N = float(n_rows) * Cn
    tot = np.array([N])
    T1, T2, tot = D.allreduce_np(T1), D.allreduce_np(T2.reshape(-1)).reshape(k, k), D.allreduce_np(tot)
    N = tot[0]
    m1 = T1 / N
    cov = T2 / N - np.outer(m1, m1)
    # per-chain columns of every rank, in global chain order (chains are sharded as contiguous blocks)
    cols = np.concatenate([st[:, [lay['MAPLOGP'], lay['NACC'], lay['SUMEXP']]], st[:, lay['MAPTHETA']:lay['MAPTHETA'] + k]], axis=1)
    if D.world()[1] > 1:
        cols = D.allgather_rows(torch.from_numpy(np.ascontiguousarray(cols))).cpu().numpy()
    i = int(np.argmax(cols[:, 0]))                                     # first chain attaining the highest logP
    return dict(mean=ref + m1, cov=cov, std=np.sqrt(np.maximum(np.diag(cov), 0)), n=N,
                map_logP=float(cols[i, 0]), map_theta=cols[i, 3:3 + k].copy(),
                map_chain=i, accept_rate=cols[:, 1], sumexp=float(cols[:, 2].sum()))


def run_metropolis_hastings(pe, calculatePostBurnInStatistics=True, continueSampling='auto', mcmc_exportLog='UserChoice'):
    UI = pe.UserInput
    pes = UI.parameter_estimation_settings
    if hasattr(UI, 'last_MCMC_run_type') and UI.last_MCMC_run_type != 'MH':                       # IP.py:2565-2569
        print("ERROR: Two different algorithms of MCMC runs were attempted with the same UserInput."); raise SystemExit
    if pes['mcmc_mode'] not in ('unbiased', 'MAP_finding'):
        raise ValueError("mcmc_mode must be 'unbiased' or 'MAP_finding' (IP.py:2629-2631)")
    map_finding = pes['mcmc_mode'] == 'MAP_finding'                                                  # IP.py:2631-2635, 2688-2713
    if continueSampling == 'auto':                                                                   # IP.py:2572-2578
        if ('mcmc_continueSampling' not in pes) or pes['mcmc_continueSampling'] == 'auto':
            continueSampling = hasattr(pe, 'mcmc_last_point_sampled')
        else:
            continueSampling = pes['mcmc_continueSampling']
    Lm = int(pes['mcmc_length'])
    burn = int(Lm * 0.1) if str(pes['mcmc_burn_in']).lower() == 'auto' else int(pes['mcmc_burn_in'])  # IP.py:2599-2600
    Cn = int(pes.get('mcmc_nchains', 1))
    first_step = 0
    if continueSampling:
        burn = 0                                                                                     # IP.py:2601-2602
        if hasattr(pe, 'mcmc_last_point_sampled'):
            pe.last_post_burn_in_log_posteriors_un_normed_vec = np.array(pe.post_burn_in_log_posteriors_un_normed_vec)
            pe.last_post_burn_in_samples = np.array(pe.post_burn_in_samples)
            first_step = getattr(pe, '_mh_steps_done', 0)
        else:
            # a fresh object continues the run an earlier process exported (IP.py:2583-2596): rows of
            # pickles/mcmc_logP_and_parameter_samples.pkl, start = its last row, initial guess from its pickle
            from .export import unpickleAnObject
            data = np.atleast_2d(np.asarray(unpickleAnObject(UI.directories['pickles'] + "mcmc_logP_and_parameter_samples"), dtype=float))
            pe.last_post_burn_in_log_posteriors_un_normed_vec = data[:, 0:1].copy()
            pe.last_post_burn_in_samples = data[:, 1:].copy()
            pe.mcmc_last_point_sampled = pe.last_post_burn_in_samples[-1]
            UI.InputParameterInitialGuess = unpickleAnObject(UI.directories['pickles'] + "mcmc_initial_point_parameters")
            pe.post_burn_in_samples, pe.post_burn_in_log_posteriors_un_normed_vec = pe.last_post_burn_in_samples, pe.last_post_burn_in_log_posteriors_un_normed_vec
            # the draws of the earlier process are not known here: a stream offset far beyond any exported run keeps them disjoint
            first_step = int(pes.get('mcmc_continue_first_step', 1 << 24))
        start = np.asarray(pe.mcmc_last_point_sampled, dtype=float).reshape(-1, pe.k)
        if start.shape[0] != Cn:
            start = np.tile(start[-1], (Cn, 1))
    else:
        start = np.tile(np.asarray(UI.InputParameterInitialGuess, dtype=float), (Cn, 1))              # IP.py:2610
    seed = pes.get('mcmc_random_seed')
    seed = int(np.random.SeedSequence().entropy % (2 ** 63)) if not isinstance(seed, int) else seed
    rank, size = D.world()
    if Cn < size:
        raise ValueError(f"mcmc_nchains = {Cn} is smaller than the number of ranks ({size}): every rank needs at least one chain")
    if not isinstance(pes.get('mcmc_random_seed'), int):
        seed = D.broadcast_int(seed)                                                                  # one Philox key for all ranks
    c0, cn = D.shard_range(Cn, rank, size)
    chol = step_cholesky(pe.Q_covmat, pes['mcmc_relative_step_length'])
    store = pes.get('mcmc_store_samples', 'auto')
    thin = int(pes.get('mcmc_thin', 1))
    if store == 'auto':
        store = (Lm // thin) * cn * (pe.k + 3) * 8 <= 64 * 2 ** 30
    out = mh_launch(pe, start[c0:c0 + cn], Lm, burn, seed, chol, first_step=first_step, chain_offset=c0,
                    modulate=float(pes['mcmc_modulate_accept_probability']), thin=thin, store_samples=bool(store),
                    map_finding=map_finding)
    pe._mh_steps_done = first_step + Lm
    pe._mh_out = out
    UI.last_MCMC_run_type = 'MH'                                                                     # IP.py:2729
    n_rows = Lm - burn
    mom = chain_moments(out['stats'], out['theta0'], pe.k, n_rows)
    pe.mcmc_acceptance_rate = float(np.mean(mom['accept_rate'])) / max(Lm - 1, 1)
    pe.mcmc_chain_moments = mom
    lay = _lib.mh_stat_layout(pe.k)
    last = D.allgather_rows(out['stats'][:, lay['LAST']:lay['LAST'] + pe.k].contiguous()).cpu().numpy()   # all Cn chains, global order
    pe.mcmc_last_point_sampled = last[0] if Cn == 1 else last                                        # IP.py:2725
    if not store:
        # statistics from the in-kernel accumulators only (no tail filter: it needs every row)
        pe.map_logP, pe.map_parameter_set = mom['map_logP'], mom['map_theta']
        pe.mu_AP_parameter_set, pe.stdap_parameter_set = mom['mean'], mom['std']
        pe.evidence, pe.info_gain = None, None
        pe.post_burn_in_samples = None
        return [pe.map_parameter_set, pe.mu_AP_parameter_set, pe.stdap_parameter_set, pe.evidence, pe.info_gain, None, None]
    b = burn // thin
    S = out['samples']                                                                                # [R, C, k]
    pe.discrete_chains_post_burn_in_samples = D.allgather_rows(S[b:].permute(1, 0, 2).contiguous()).permute(1, 0, 2).cpu().numpy()
    # chain-major stacking = consolidate_parallel_sampling_data('equal'): vstack of every chain's rows (IP.py:1282-1319)
    flat_dev = lambda t: D.allgather_rows(t.permute(1, 0, *range(2, t.ndim)).contiguous()).reshape(-1, *t.shape[2:])
    flat = lambda t: flat_dev(t).cpu().numpy()
    d_s, d_lp, d_pr = flat_dev(S[b:]), flat_dev(out['logp'][b:]), flat_dev(out['logprior'][b:])
    pe._dev_post = dict(samples=d_s, logp=d_lp, logprior=d_pr)              # device copies for the statistics kernels
    from . import stats as ST
    filtered = ST.filter_settings(pes)[0] != _lib.FILTER_NONE
    if calculatePostBurnInStatistics and filtered and not continueSampling:
        # the tail filter replaces these three arrays (IP.py:1919-1921): only the kept rows are copied to the host
        pe._host_arrays_pending = True
        pe.post_burn_in_samples = np.empty((d_s.shape[0], pe.k))[:0]
        pe.post_burn_in_log_posteriors_un_normed_vec = _RowsOnDevice(d_lp.shape[0])
        pe.post_burn_in_log_priors_vec = None
    else:
        pe.post_burn_in_samples = d_s.cpu().numpy()                                                   # IP.py:2718
        pe.post_burn_in_log_posteriors_un_normed_vec = d_lp.cpu().numpy().reshape(-1, 1)              # IP.py:2724
        pe.post_burn_in_log_priors_vec = d_pr.cpu().numpy().reshape(-1, 1)                            # IP.py:2727
    pe.post_burn_in_log_likelihoods_vec = flat(out['loglik'][b:]).reshape(-1, 1)                      # IP.py:2726
    if not continueSampling:
        pe.during_burn_in_samples = flat(S[:b])                                                       # IP.py:2716
        pe.during_burn_in_log_posteriors_un_normed_vec = flat(out['logp'][:b]).reshape(-1, 1)
    else:
        pe.post_burn_in_samples = np.vstack((pe.last_post_burn_in_samples, pe.post_burn_in_samples))  # IP.py:2731-2733
        pe.post_burn_in_log_posteriors_un_normed_vec = np.vstack((pe.last_post_burn_in_log_posteriors_un_normed_vec,
                                                                  pe.post_burn_in_log_posteriors_un_normed_vec))
        pe._dev_post = None                                                  # merged on the host: statistics upload them
    if calculatePostBurnInStatistics:
        post_burn_in_statistics(pe, calculate_post_burn_in_log_priors_vec=bool(continueSampling))
        if UI.user_requested_convergence_diagnostics:                                                 # IP.py:2741-2742
            pe._mh_burn_rows = b
            pe.getConvergenceDiagnostics(S[b:] if D.world()[1] == 1 else pe.discrete_chains_post_burn_in_samples)
        _export_if_requested(pe, mcmc_exportLog)                                                      # IP.py:2743-2746
        return [pe.map_parameter_set, pe.mu_AP_parameter_set, pe.stdap_parameter_set, pe.evidence, pe.info_gain,
                pe.post_burn_in_samples, pe.post_burn_in_log_posteriors_un_normed_vec]               # IP.py:2749
    lp = pe.post_burn_in_log_posteriors_un_normed_vec
    pe.map_logP = max(lp)                                                                             # IP.py:2751-2754
    pe.map_index = list(lp).index(pe.map_logP)
    pe.map_parameter_set = pe.post_burn_in_samples[pe.map_index]
    return pe.map_logP


def _export_if_requested(pe, mcmc_exportLog):
    if str(mcmc_exportLog) == 'UserChoice':
        mcmc_exportLog = bool(pe.UserInput.parameter_estimation_settings['mcmc_exportLog'])
    if mcmc_exportLog == True:                                                                        # noqa: E712
        from .export import exportPostBurnInStatistics
        exportPostBurnInStatistics(pe)


def post_burn_in_statistics(pe, calculate_post_burn_in_log_priors_vec=False):
    """calculatePostBurnInStatistics (IP.py:1861-1937) incl. the default tail filter (Q9: rows with
    log(-logP) > mean + 2 sigma removed) and calculateInfoGain (IP.py:1770-1847), computed by the streaming kernels of
    csrc/stats.cu on the device copies of the sample arrays (peuqse_b200/stats.py).  The sampler leaves those copies in
    pe._dev_post; arrays that only exist on the host (continued sampling, ESS, user-edited arrays) are uploaded."""
    from . import stats as ST
    pes = pe.UserInput.parameter_estimation_settings
    dev = getattr(pe, '_dev_post', None)
    pe._dev_post = None                       # consumed: a later call works from the (possibly edited) host arrays
    n_rows = int(np.shape(pe.post_burn_in_log_posteriors_un_normed_vec)[0])
    if dev is None or dev['samples'].shape[0] != n_rows:
        up = lambda a: torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=float)), dtype=torch.float64).to(_dev())
        dev = dict(samples=up(pe.post_burn_in_samples).reshape(n_rows, pe.k), logp=up(pe.post_burn_in_log_posteriors_un_normed_vec).reshape(-1),
                   logprior=None)
        lpv = getattr(pe, 'post_burn_in_log_priors_vec', None)
        if lpv is not None and not calculate_post_burn_in_log_priors_vec and np.shape(lpv)[0] == n_rows:
            dev['logprior'] = up(lpv).reshape(-1)
    if dev.get('logprior') is None or calculate_post_burn_in_log_priors_vec:                          # IP.py:1863-1869
        _, lp, _ = pe.getLogP_batch(dev['samples'], return_parts=True)
        dev['logprior'] = lp
        pe.post_burn_in_log_priors_vec = lp.cpu().numpy().reshape(-1, 1)
    mode, coef = ST.filter_settings(pes)
    want_kl = pes['mcmc_info_gain_returned'] == 'KL_divergence'
    R = ST.device_statistics(dev['samples'], dev['logp'], dev['logprior'], mode, coef,
                             prior_samples=pe.samples_of_prior if want_kl else None, kl_bins=100, logp_fn=pe.getLogP_batch)
    pe.gpu_launches += R.gpu_launches
    pe.device_statistics = R
    if not want_kl and pes['mcmc_info_gain_cutoff'] != 0:                                             # IP.py:1805-1843
        _, kl, kp = R.kept_rows()
        R.info_gain_log_ratio = ST.info_gain_log_ratio_cutoff(kl, kp, R.evidence, float(pes['mcmc_info_gain_cutoff']))
        pe.gpu_launches += 1
    if mode != _lib.FILTER_NONE or getattr(pe, '_host_arrays_pending', False):                        # IP.py:1919-1921
        ks, kl, kp = R.kept_rows()
        pe.gpu_launches += 1
        pe.post_burn_in_samples = ks.cpu().numpy()
        pe.post_burn_in_log_posteriors_un_normed_vec = kl.cpu().numpy()
        pe.post_burn_in_log_priors_vec = kp.cpu().numpy()
        pe._host_arrays_pending = False
    pe.mu_AP_parameter_set = R.mu_AP_parameter_set                                                    # IP.py:1923
    pe.stdap_parameter_set = R.stdap_parameter_set                                                    # IP.py:1924
    pe.map_logP = np.array([R.map_logP])                                                              # max() of (N,1) rows, IP.py:1925
    pe.map_index = R.map_index
    pe.map_parameter_set = R.map_parameter_set
    pe.evidence = R.evidence                                                                          # IP.py:1931
    pe.info_gain_log_ratio = R.info_gain_log_ratio
    if want_kl:
        pe.info_gain_KL_each_parameter = pe.info_gain_each_parameter = R.info_gain_KL_each_parameter
        pe.info_gain_KL = R.info_gain_KL
        pe.info_gain = R.info_gain_KL
    else:
        pe.info_gain = R.info_gain_log_ratio
    return [pe.map_parameter_set, pe.mu_AP_parameter_set, pe.stdap_parameter_set, pe.evidence, pe.info_gain,
            pe.post_burn_in_samples, pe.post_burn_in_log_posteriors_un_normed_vec]


# ----------------------------------------------------------------------------------------------
# ensemble slice sampling
# ----------------------------------------------------------------------------------------------
ESS_DRIVER_AUTO, ESS_DRIVER_WALKER, ESS_DRIVER_ROUNDS = 0, 1, 2


def _ess_spec(W, k, n_iters, seed, mu0, tune, max_rounds, first_iter, driver):
    spec = _lib.ESSSpec()
    spec.n_walkers, spec.n_iters, spec.first_iter, spec.seed = W, int(n_iters), int(first_iter), int(seed) & (2 ** 64 - 1)
    spec.mu0, spec.tune, spec.max_rounds, spec.k, spec.driver = float(mu0), int(bool(tune)), int(max_rounds), k, int(driver)
    return spec


def _ess_check(L, rc):
    """Library status -> the exceptions zeus raises (RuntimeError when a stepping-out / shrink loop runs away)."""
    if rc != 0:
        msg = L.peuqse_last_error().decode(errors="replace")
        if "maximum limit" in msg:
            raise RuntimeError(msg)
        raise _lib.PeuqseError(f"libpeuqse_b200 error {rc}: {msg}")


def ess_launch(pe, start, n_iters, seed, mu0=1.0, tune=True, max_rounds=10000, first_iter=0, store_chain=True,
               driver=ESS_DRIVER_AUTO, shard=True):
    """Ensemble slice sampling of `n_iters` ensemble iterations from walkers start[W,k] (zeus.EnsembleSampler.run_mcmc,
    IP.py:2495-2498).  Built-in model with diagonal responses: the per-walker CUDA kernels (csrc/ess.cu).  Full-covariance
    responses, torch-callable simulators, several ranks: the round-synchronous device driver (csrc/ess_rounds.cu), every
    round one batched log-posterior launch.  Both consume the same Philox draws: the walkers do not depend on the driver."""
    from .models import BuiltinModel
    L = _lib.load()
    dev = _dev()
    X = torch.as_tensor(np.asarray(start, dtype=float), dtype=torch.float64).reshape(-1, pe.k).to(dev).contiguous()
    W, k = X.shape
    custom = getattr(pe, 'custom_logLikelihood', None) is not None or getattr(pe, 'custom_logPrior', None) is not None
    builtin = isinstance(pe.simulator, BuiltinModel) and not custom      # user callables: the rounds are issued from Python
    full = int(pe.host_problem.resp_mode.max()) == _lib.RESP_MVN_FULL
    sharded = bool(shard) and D.world()[1] > 1 and (full or not builtin)
    if not builtin or sharded:
        return ess_rounds(pe, X, n_iters, seed, mu0=mu0, tune=tune, max_rounds=max_rounds, first_iter=first_iter,
                          store_chain=store_chain, shard=sharded)
    Z = pe.getLogP_batch(X).contiguous()
    if not bool(torch.isfinite(Z).all().item()):
        raise ValueError("Invalid walker initial positions! Initialise walkers from positions of finite log probability.")
    spec = _ess_spec(W, k, n_iters, seed, mu0, tune, max_rounds, first_iter, driver)
    chain = torch.empty((n_iters, W, k), dtype=torch.float64, device=dev) if store_chain else None
    chain_lp = torch.empty((n_iters, W), dtype=torch.float64, device=dev) if store_chain else None
    info = torch.zeros(8, dtype=torch.float64, device=dev)
    rounds = full or driver == ESS_DRIVER_ROUNDS
    work = None if rounds else torch.zeros(int(L.peuqse_ess_work_bytes(W, k)), dtype=torch.uint8, device=dev)
    _ess_check(L, L.peuqse_ess_run(pe.device_problem.handle, C.byref(spec), _lib.ptr(X), _lib.ptr(Z), _lib.ptr(chain),
                                   _lib.ptr(chain_lp), _lib.ptr(info), _lib.ptr(work), _lib.current_stream()))
    ih = info.cpu().numpy()
    if int(ih[4]) != 0:
        raise RuntimeError("Number of expansions / contractions exceeded maximum limit! Increase mcmc_maxiter.")
    # launches: per-walker driver 5 per iteration; rounds driver 2 per round + the batched log-posterior of each round
    # (a round evaluates at least one row, so the number of rounds is bounded by the evaluations; reported as a lower bound)
    pe.gpu_launches += 5 * int(n_iters) + 3
    info[0] += W                                                  # the start points were evaluated too
    return dict(X=X, Z=Z, chain=chain, chain_logp=chain_lp, info=info)


def ess_rounds(pe, X, n_iters, seed, mu0=1.0, tune=True, max_rounds=10000, first_iter=0, store_chain=True, shard=False):
    """The round-synchronous driver, rounds issued from here (peuqse_essr_*): for simulators that are torch callables (the
    block of candidate points goes through the callable, then peuqse_logp_from_sim) and for several ranks (SURVEY 8e, v2:
    every rank runs the same rounds on the same walkers -- same seed, same draws, same masks -- and evaluates its share of
    each block's rows; the log-posteriors are all-gathered with NCCL.  A row's value does not depend on the batch it is
    evaluated in, so the chain is the one a single GPU produces)."""
    L = _lib.load()
    dev = _dev()
    W, k = X.shape
    logp_rows = pe.getLogP_batch
    if shard:
        logp_rows = lambda th: D.sharded_rows(pe.getLogP_batch, th.contiguous(), min_rows_per_rank=8)
    Z = logp_rows(X).clone()
    if not bool(torch.isfinite(Z).all().item()):
        raise ValueError("Invalid walker initial positions! Initialise walkers from positions of finite log probability.")
    spec = _ess_spec(W, k, n_iters, seed, mu0, tune, max_rounds, first_iter, ESS_DRIVER_ROUNDS)
    h = C.c_void_p()
    _lib.check(L.peuqse_essr_create(C.byref(spec), C.byref(h)))
    chain = torch.empty((n_iters, W, k), dtype=torch.float64, device=dev) if store_chain else None
    chain_lp = torch.empty((n_iters, W), dtype=torch.float64, device=dev) if store_chain else None
    cand = torch.empty((W, k), dtype=torch.float64, device=dev)
    info = torch.zeros(8, dtype=torch.float64, device=dev)
    n = C.c_int64(0)
    st = _lib.current_stream()
    launches = 0
    try:
        for it in range(int(n_iters)):
            for half in (0, 1):
                _ess_check(L, L.peuqse_essr_begin_half(h, it, half, _lib.ptr(X), _lib.ptr(Z), _lib.ptr(cand), C.byref(n), st))
                launches += 4 - 2 * half
                while n.value > 0:
                    rows = logp_rows(cand[:n.value]).contiguous()
                    _ess_check(L, L.peuqse_essr_round(h, _lib.ptr(rows), _lib.ptr(cand), C.byref(n), st))
                    launches += 2
                _lib.check(L.peuqse_essr_end_half(h, _lib.ptr(X), _lib.ptr(Z), st))
            _lib.check(L.peuqse_essr_end_iteration(h, st))
            launches += 3
            if store_chain:
                chain[it] = X
                chain_lp[it] = Z
        _lib.check(L.peuqse_essr_info(h, _lib.ptr(info), st))
        torch.cuda.current_stream().synchronize()
    finally:
        torch.cuda.synchronize()
        L.peuqse_essr_destroy(h)
    pe.gpu_launches += launches
    info[0] += W
    return dict(X=X, Z=Z, chain=chain, chain_logp=chain_lp, info=info)


def run_ensemble_slice_sampling(pe, mcmc_nwalkers_direct_input=None, walkerInitialDistribution='UserChoice',
                                walkerInitialDistributionSpread='UserChoice', movesType='UserChoice',
                                calculatePostBurnInStatistics=True, continueSampling='auto', mcmc_exportLog='UserChoice'):
    UI = pe.UserInput
    pes = UI.parameter_estimation_settings
    if walkerInitialDistribution == 'UserChoice': walkerInitialDistribution = pes['mcmc_walkerInitialDistribution']
    if walkerInitialDistribution.lower() == 'auto': walkerInitialDistribution = 'sobol'              # IP.py:2369-2370
    if str(walkerInitialDistributionSpread) == 'UserChoice': walkerInitialDistributionSpread = pes['mcmc_walkerInitialDistributionSpread']
    if str(walkerInitialDistributionSpread).lower() == 'auto': walkerInitialDistributionSpread = 0.866   # IP.py:2373-2374
    if hasattr(UI, 'last_MCMC_run_type') and UI.last_MCMC_run_type != 'ESS':
        print("ERROR: Two different algorithms of MCMC runs were attempted with the same UserInput."); raise SystemExit
    if movesType == 'UserChoice': movesType = pes['mcmc_movesType']
    if movesType.lower() not in ('auto', 'differential'):
        raise NotImplementedError("only the differential move (zeus default, IP.py:2439-2440) is implemented")
    if continueSampling == 'auto':
        if ('mcmc_continueSampling' not in pes) or pes['mcmc_continueSampling'] == 'auto':
            continueSampling = hasattr(pe, 'mcmc_last_point_sampled')
        else:
            continueSampling = pes['mcmc_continueSampling']
    k = pe.k
    seed = pes.get('mcmc_random_seed')
    seeded = isinstance(seed, int)
    if seeded:
        np.random.seed(seed)                                                                         # IP.py:2413-2415
    else:
        # every rank must run the same ensemble (the rounds are shared out row-wise): rank 0 picks the seed
        seed = D.broadcast_int(int(np.random.SeedSequence().entropy % (2 ** 63)))
    if mcmc_nwalkers_direct_input is None:                                                           # IP.py:2416-2425
        nw = pes.get('mcmc_nwalkers', 'auto')
        nw = k * 4 if (isinstance(nw, str) and nw.lower() == 'auto') else int(nw)
    else:
        nw = int(mcmc_nwalkers_direct_input)
    if nw % 2 != 0:
        nw += 1                                                                                      # IP.py:2426-2428
    pe.mcmc_nwalkers = nw
    requested = int(pes['mcmc_length'])
    nsteps = max(1, int(requested / nw))                                                             # IP.py:2430-2432
    burn = int(requested * 0.1) if str(pes['mcmc_burn_in']).lower() == 'auto' else int(pes['mcmc_burn_in'])
    maxiter = int(pes.get('mcmc_maxiter', 1E6))
    first_iter = 0
    if continueSampling:
        burn = 0
        start = np.asarray(pe.mcmc_last_point_sampled, dtype=float)
        pe.last_post_burn_in_log_posteriors_un_normed_vec = np.array(pe.post_burn_in_log_posteriors_un_normed_vec)
        pe.last_post_burn_in_samples = np.array(pe.post_burn_in_samples)
        first_iter = getattr(pe, '_ess_iters_done', 0)
    out = None
    for trial in range(1000):                                                                        # IP.py:2496-2510
        if not continueSampling:
            # scipy's scrambled Sobol generator does not follow np.random.seed: seed it explicitly so that a seeded run is
            # reproducible, and take rank 0's points on every rank (quasi-random starts differ from process to process)
            start = pe.generateInitialPoints(initialPointsDistributionType=walkerInitialDistribution, numStartPoints=nw,
                                             relativeInitialDistributionSpread=walkerInitialDistributionSpread,
                                             seed=(seed + trial) if seeded else None)
            start = D.broadcast_np(start)
        try:
            out = ess_launch(pe, start, nsteps, seed + trial, max_rounds=min(maxiter, 2 ** 30), first_iter=first_iter)
            break
        except ValueError as e:
            if "finite" in str(e) and not continueSampling:
                print("One of the starting points has a non-finite probability. Picking new starting points.")
                continue
            raise
    pe._ess_iters_done = first_iter + nsteps
    pe._ess_out = out
    disc = int(burn / nw)                                                                            # IP.py:2512
    chain = out['chain'][disc:]
    chain_lp = out['chain_logp'][disc:]
    pe.discrete_chains_post_burn_in_samples = chain.cpu().numpy()                                    # get_chain(discard), IP.py:2514
    pe.post_burn_in_samples = chain.reshape(-1, k).cpu().numpy()                                     # samples.flatten(discard), IP.py:2513
    pe.post_burn_in_log_posteriors_un_normed_vec = chain_lp.reshape(-1, 1).cpu().numpy()             # IP.py:2517
    pe.mcmc_last_point_sampled = pe.discrete_chains_post_burn_in_samples[-1]                          # IP.py:2518
    info = out['info'].cpu().numpy()
    pe.ess_info = dict(n_logp_evaluations=int(info[0]), mu=float(info[1]), n_expansions=int(info[2]), n_contractions=int(info[3]))
    UI.last_MCMC_run_type = 'ESS'
    if continueSampling:
        pe.post_burn_in_samples = np.vstack((pe.last_post_burn_in_samples, pe.post_burn_in_samples))
        pe.post_burn_in_log_posteriors_un_normed_vec = np.vstack((pe.last_post_burn_in_log_posteriors_un_normed_vec,
                                                                  pe.post_burn_in_log_posteriors_un_normed_vec))
    if calculatePostBurnInStatistics:
        post_burn_in_statistics(pe, calculate_post_burn_in_log_priors_vec=True)                       # IP.py:2533
        _export_if_requested(pe, mcmc_exportLog)                                                      # IP.py:2536-2539
        return [pe.map_parameter_set, pe.mu_AP_parameter_set, pe.stdap_parameter_set, pe.evidence, pe.info_gain,
                pe.post_burn_in_samples, pe.post_burn_in_log_posteriors_un_normed_vec]
    lp = pe.post_burn_in_log_posteriors_un_normed_vec
    pe.map_logP = max(lp)
    pe.map_index = list(lp).index(pe.map_logP)
    pe.map_parameter_set = pe.post_burn_in_samples[pe.map_index]
    return pe.map_logP
