"""Bayesian design of experiments on the device: doeGetInfoGainMatrix and doeParameterModulationPermutationsScanner
(PEUQSE/InverseProblem.py:1363-1572, "IP.py") with populateResponsesWithSyntheticData (IP.py:1326-1354).

The reference loops `for modulation: for condition: populate independent variables -> synthetic data -> __init__ ->
doMetropolisHastings -> info_gain`.  Here every (modulation, condition) pair is one SCENARIO of a single batch
(csrc/doe.cu): the scenarios' model constants and synthetic observations live in HBM, one launch of the persistent MH
kernel walks all scenarios' chains, one launch of the per-scenario statistics kernel returns every information gain, and
the host reads back 8 + 3k doubles per scenario.  Scenarios are independent, so with several ranks each GPU takes a contiguous
range of them and the matrices are all-gathered (no data-path collective) -- the reference's 'parallel_conditions_exploration'
/ 'parallel_parameter_modulation' MPI modes with their pickle files (IP.py:1384-1397, 1530-1543) collapse into that.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import distributed as D
from . import stats as ST
from .mcmc import step_cholesky
from .models import BuiltinModel
from .sweep import make_spec, grid_size


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


class ScenarioBatch:
    """S scenarios of one problem: model constants under conditions[s] and observed values synthesised at theta_true[s]."""

    def __init__(self, pe, conditions, theta_true):
        if not isinstance(pe.simulator, BuiltinModel):
            raise NotImplementedError("design-of-experiments batches need a built-in CUDA model with independent variables "
                                      "(BuiltinModel.with_conditions)")
        self.pe = pe
        self._L = L = _lib.load()
        conditions = np.asarray(conditions, dtype=float)
        self.S = S = conditions.shape[0]
        order = pe.host_problem.src.astype(np.int64)
        rows = np.ascontiguousarray(np.stack([pe.simulator.with_conditions(c).constants(order) for c in conditions]))
        self.handle = C.c_void_p()
        _lib.check(L.peuqse_scenarios_create(pe.device_problem.handle, S, rows.ctypes.data_as(_lib.c_double_p), C.byref(self.handle)))
        th = np.asarray(theta_true, dtype=float)
        stride = 0 if th.ndim == 1 else pe.k
        self.theta_true = torch.as_tensor(np.ascontiguousarray(th.reshape(-1, pe.k)), dtype=torch.float64).to(_dev())
        if stride and self.theta_true.shape[0] != S:
            raise ValueError("theta_true must be one vector or one row per scenario")
        self.observed = torch.empty((S, pe.host_problem.n_total), dtype=torch.float64, device=_dev())   # natural order
        _lib.check(L.peuqse_scenarios_synthesize(pe.device_problem.handle, self.handle, _lib.ptr(self.theta_true), stride,
                                                 _lib.ptr(self.observed), _lib.current_stream()))
        pe.gpu_launches += 1

    def logp(self, thetas, return_parts=False):
        """thetas[S,B,k] -> logP[S,B] (and logPrior, logLikelihood) under each scenario's observations."""
        th = torch.as_tensor(np.asarray(thetas, dtype=float) if not torch.is_tensor(thetas) else thetas, dtype=torch.float64)
        th = th.to(_dev()).reshape(self.S, -1, self.pe.k).contiguous()
        B = th.shape[1]
        outs = [torch.empty((self.S, B), dtype=torch.float64, device=th.device) for _ in range(3 if return_parts else 1)]
        ptrs = [_lib.ptr(o) for o in outs] + [None] * (3 - len(outs))
        _lib.check(self._L.peuqse_scenarios_logp(self.pe.device_problem.handle, self.handle, _lib.ptr(th), B, *ptrs, _lib.current_stream()))
        self.pe.gpu_launches += 1
        return tuple(outs) if return_parts else outs[0]

    def run_mh(self, theta0, n_steps, burn_in, seed, step_chol, chains_per_scenario=1, modulate=0.0, export_draws=False,
               map_finding=False):
        """One launch of peuqse_mh_run_scenarios: S x chains_per_scenario chains.  Returns a dict of device tensors
        (samples[R, S*cps, k], logp / logprior / loglik [R, S*cps], stats)."""
        pe = self.pe
        cps = int(chains_per_scenario)
        Cn = self.S * cps
        th0 = torch.as_tensor(np.asarray(theta0, dtype=float), dtype=torch.float64).reshape(-1, pe.k)
        if th0.shape[0] == 1:
            th0 = th0.repeat(Cn, 1)
        th0 = th0.to(_dev()).contiguous()
        assert th0.shape[0] == Cn
        spec = _lib.MHSpec()
        spec.n_chains, spec.n_steps, spec.burn_in, spec.first_step = Cn, int(n_steps), int(burn_in), 0
        spec.chain_offset, spec.seed, spec.modulate, spec.k, spec.thin = 0, int(seed) & (2 ** 64 - 1), float(modulate), pe.k, 1
        spec.map_finding = 1 if map_finding else 0
        for i, v in enumerate(np.asarray(step_chol, dtype=float).reshape(-1)):
            spec.step_chol[i] = float(v)
        dev = _dev()
        R = int(n_steps)
        out = dict(samples=torch.empty((R, Cn, pe.k), dtype=torch.float64, device=dev),
                   logp=torch.empty((R, Cn), dtype=torch.float64, device=dev),
                   logprior=torch.empty((R, Cn), dtype=torch.float64, device=dev),
                   loglik=torch.empty((R, Cn), dtype=torch.float64, device=dev),
                   deltas=torch.empty((max(R - 1, 0), Cn, pe.k), dtype=torch.float64, device=dev) if export_draws else None,
                   us=torch.empty((max(R - 1, 0), Cn, 2) if map_finding else (max(R - 1, 0), Cn), dtype=torch.float64,
                                  device=dev) if export_draws else None,
                   stats=torch.empty((Cn, _lib.mh_stat_layout(pe.k)['SIZE']), dtype=torch.float64, device=dev))
        _lib.check(self._L.peuqse_mh_run_scenarios(pe.device_problem.handle, self.handle, C.byref(spec), cps, _lib.ptr(th0),
                                                   _lib.ptr(out['samples']), _lib.ptr(out['logp']), _lib.ptr(out['logprior']),
                                                   _lib.ptr(out['loglik']), _lib.ptr(out['deltas']), _lib.ptr(out['us']),
                                                   _lib.ptr(out['stats']), _lib.current_stream()))
        pe.gpu_launches += 1
        out['chains_per_scenario'] = cps
        return out

    def statistics(self, run, burn_in, filter_mode=_lib.FILTER_NONE, coef=2.0):
        """Per-scenario post-burn-in statistics + log-ratio information gain: host array [S, 8 + 3k] (_lib.scen_stats_layout)."""
        R = run['samples'].shape[0]
        out = torch.empty((self.S, _lib.scen_stats_layout(self.pe.k)['DOUBLES']), dtype=torch.float64, device=_dev())
        _lib.check(self._L.peuqse_scenarios_stats(self.pe.k, int(filter_mode), float(coef), self.S, run['chains_per_scenario'], R,
                                                  int(burn_in), _lib.ptr(run['samples']), _lib.ptr(run['logp']),
                                                  _lib.ptr(run['logprior']), _lib.ptr(out), _lib.current_stream()))
        self.pe.gpu_launches += 1
        self._d_stats = out
        return out.cpu().numpy()

    def cutoff_information_gain(self, run, burn_in, cutoff):
        """Log-ratio information gain of every scenario with mcmc_info_gain_cutoff != 0 (IP.py:1805-1843), after statistics():
        of a scenario's kept rows (logP >= its filter cut) those whose log-prior or log-posterior is NaN or below ln(cutoff)
        are dropped, the others contribute ln(exp(logP) / evidence) - logPrior (non-finite terms count 0); mean over the rows
        left, NaN if there are none.  One masked reduction over logp / logprior [R, S, chains] on the device; host array [S]."""
        L = _lib.scen_stats_layout(self.pe.k)
        cps = run['chains_per_scenario']
        lp = run['logp'][int(burn_in):].reshape(-1, self.S, cps)
        pr = run['logprior'][int(burn_in):].reshape(-1, self.S, cps)
        cut = self._d_stats[:, L['CUT']].reshape(1, self.S, 1)
        ev = self._d_stats[:, L['EVIDENCE']].reshape(1, self.S, 1)
        lc = float(np.log(cutoff))
        keep = ~(lp < cut) & ~(torch.isnan(lp) | torch.isnan(pr) | (lp < lc) | (pr < lc))
        ratios = torch.log(torch.exp(lp) / ev) - pr
        ratios = torch.where(keep & torch.isfinite(ratios), ratios, torch.zeros_like(ratios))
        n = keep.sum(dim=(0, 2)).to(torch.float64)
        self.pe.gpu_launches += 1
        return (ratios.sum(dim=(0, 2)) / n).cpu().numpy()            # 0 / 0 = NaN: numpy's mean of no rows

    def kl_information_gain(self, run, burn_in, prior_samples, bins=100):
        """KL-divergence information gain of every scenario (after statistics()): host array [S, 1 + k] = total, each parameter."""
        k = self.pe.k
        pri = torch.as_tensor(np.ascontiguousarray(np.asarray(prior_samples, dtype=float))).to(_dev()).reshape(-1, k).contiguous()
        out = torch.empty((self.S, 1 + k), dtype=torch.float64, device=_dev())
        _lib.check(self._L.peuqse_scenarios_kl(k, int(bins), self.S, run['chains_per_scenario'], run['samples'].shape[0], int(burn_in),
                                               _lib.ptr(run['samples']), _lib.ptr(run['logp']), _lib.ptr(pri), pri.shape[0],
                                               _lib.ptr(self._d_stats), _lib.ptr(out), _lib.current_stream()))
        self.pe.gpu_launches += 1
        return out.cpu().numpy()

    def close(self):
        if getattr(self, 'handle', None):
            self._L.peuqse_scenarios_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _grid_points(pe, center, interval, radii):
    """getGridPermutations (IP.py:755-770) through the device decode of the combination generator's order."""
    center = np.asarray(center, dtype=float)
    k = len(center)
    radii = np.ones(k, dtype='int') if len(radii) == 0 else np.array(radii, dtype='int')
    spec = make_spec("grid", k, center, interval=np.asarray(interval, dtype=float), radii=radii)
    return pe._points_from_spec(spec, 0, grid_size(radii))


def conditions_grid(pe, doe_settings):
    """The conditions doeGetInfoGainMatrix visits, in its loop order ('xyz': IP.py:1380-1382; 'meshgrid': IP.py:1434-1447)."""
    fmt = doe_settings['info_gains_matrices_array_format']
    c = doe_settings['independent_variable_grid_center']
    h = doe_settings['independent_variable_grid_interval_size']
    r = doe_settings['independent_variable_grid_num_intervals']
    if fmt == 'xyz':
        return _grid_points(pe, c, h, r)
    if fmt != 'meshgrid':
        raise ValueError("doe_settings['info_gains_matrices_array_format'] must be 'xyz' or 'meshgrid'")
    if len(c) != 2:
        print("CURRENTLY THE INFOGAIN MESHGRID OPTION IS ONLY SUPPORTED FOR TWO INDEPENDENT VARIABLES. Use "
              "doe_settings['independent_variable_grid_center'] = 'xyz' and run again.")                # IP.py:1427-1429
        raise SystemExit
    ax = []
    for i in range(2):
        hi = c[i] + h[i] * r[i]
        lo = c[i] - h[i] * r[i]
        ax.append(np.linspace(lo, hi, r[i] * 2 + 1))                                                   # IP.py:1436-1441
    pe.meshGrid_independentVariable1ValuesArray, pe.meshGrid_independentVariable2ValuesArray = ax[0], ax[1]
    return np.array([[v1, v2] for v2 in ax[1] for v1 in ax[0]], dtype=float)                           # IP.py:1448-1449


def run_info_gain_matrices(pe, modulations, searchType='doMetropolisHastings'):
    """All scenarios (modulations x conditions) in one batch.  Returns (matrices[M, Cn, nv+1], each[k][M, Cn, nv+1] or None)."""
    UI = pe.UserInput
    pes, doe = UI.parameter_estimation_settings, UI.doe_settings
    if searchType != 'doMetropolisHastings':
        raise NotImplementedError(f"design of experiments with searchType '{searchType}': only doMetropolisHastings is batched")
    pes['mcmc_continueSampling'] = False                                                               # IP.py:1367
    modulations = np.asarray(modulations, dtype=float).reshape(-1, pe.k)
    conds = conditions_grid(pe, doe)
    M, Cn, nv = modulations.shape[0], conds.shape[0], conds.shape[1]
    S = M * Cn
    rank, size = D.world()
    s0, sn = D.shard_range(S, rank, size)
    all_conds = np.tile(conds, (M, 1))
    all_true = np.repeat(modulations, Cn, axis=0)
    Lm = int(pes['mcmc_length'])
    burn = int(Lm * 0.1) if str(pes['mcmc_burn_in']).lower() == 'auto' else int(pes['mcmc_burn_in'])    # IP.py:2599-2600
    cps = int(pes.get('mcmc_nchains', 1))
    seed = pes.get('mcmc_random_seed')
    seed = int(np.random.SeedSequence().entropy % (2 ** 63)) if not isinstance(seed, int) else seed
    chol = step_cholesky(pe.Q_covmat, pes['mcmc_relative_step_length'])
    mode, coef = ST.filter_settings(pes)
    want_kl = pes['mcmc_info_gain_returned'] == 'KL_divergence'
    if mode == _lib.FILTER_MU_AP:
        raise NotImplementedError("design-of-experiments batches support the 'auto' and 'MAP' filter benchmarks")
    want_each = doe.get('info_gains_matrices_multiple_parameters', 'sum') == 'each'
    gains = np.zeros(sn)
    each = np.zeros((sn, pe.k)) if want_kl else None
    stats_rows = None
    if sn:
        batch = ScenarioBatch(pe, all_conds[s0:s0 + sn], all_true[s0:s0 + sn])
        run = batch.run_mh(np.asarray(UI.InputParameterInitialGuess, dtype=float), Lm, burn, seed, chol, chains_per_scenario=cps,
                           modulate=float(pes['mcmc_modulate_accept_probability']), map_finding=pes['mcmc_mode'] == 'MAP_finding')
        stats_rows = batch.statistics(run, burn, mode, coef)
        if not want_kl:
            gains = stats_rows[:, _lib.scen_stats_layout(pe.k)['INFO_GAIN_LOG_RATIO']].copy()
            if pes['mcmc_info_gain_cutoff'] != 0:                                                      # IP.py:1805-1843
                gains = batch.cutoff_information_gain(run, burn, float(pes['mcmc_info_gain_cutoff']))
        else:
            # the KL form bins prior and posterior samples on scenario-specific joint edges: one CTA per scenario (k_scen_kl)
            kl = batch.kl_information_gain(run, burn, pe.samples_of_prior, bins=100)
            gains, each = kl[:, 0].copy(), kl[:, 1:].copy()
        pe.responses_observed_last_scenario = batch.observed[-1].cpu().numpy()
        pe._doe_batch, pe._doe_run = batch, run
    if size > 1:
        full = np.zeros(S)
        full[s0:s0 + sn] = gains
        gains = D.allreduce_np(full)
        if each is not None:
            fe = np.zeros((S, pe.k))
            fe[s0:s0 + sn] = each
            each = D.allreduce_np(fe.reshape(-1)).reshape(S, pe.k)
    pe.doe_scenario_statistics = stats_rows
    mats = np.concatenate([all_conds, gains.reshape(-1, 1)], axis=1).reshape(M, Cn, nv + 1)            # IP.py:1411-1413
    each_mats = None
    if want_each and each is not None:
        each_mats = [np.concatenate([all_conds, each[:, d:d + 1]], axis=1).reshape(M, Cn, nv + 1) for d in range(pe.k)]
    return mats, each_mats


def doe_get_info_gain_matrix(pe, parameterPermutation, searchType='doMetropolisHastings'):
    """doeGetInfoGainMatrix (IP.py:1363-1498): one "true" parameter vector, every condition of the grid."""
    pe.middle_of_doe_flag = True
    mats, each = run_info_gain_matrices(pe, np.asarray(parameterPermutation, dtype=float).reshape(1, -1), searchType)
    pe.info_gain_matrix = mats[0]
    pe.info_gains_matrices_array_format = pe.UserInput.doe_settings['info_gains_matrices_array_format']
    if each is not None:
        pe.info_gain_matrices_each_parameter = [e[0] for e in each]
    pe.middle_of_doe_flag = False
    return np.array(pe.info_gain_matrix)


def doe_parameter_modulation_scanner(pe, searchType='doMetropolisHastings'):
    """doeParameterModulationPermutationsScanner (IP.py:1499-1572): a grid of "true" parameter vectors around the initial
    guess (spacing in units of the prior standard deviations), each with its own information-gain matrix."""
    UI = pe.UserInput
    doe = UI.doe_settings
    UI.user_requested_convergence_diagnostics = False
    center = np.asarray(UI.InputParameterInitialGuess, dtype=float)
    interval = np.asarray(doe['parameter_modulation_grid_interval_size'], dtype=float) * np.asarray(UI.std_prior, dtype=float)  # IP.py:1508
    mods = _grid_points(pe, center, interval, doe['parameter_modulation_grid_num_intervals'])
    pe.numParModulationPermutations = len(mods)
    pe.parModulationGridPermutations = mods
    pe.middle_of_doe_flag = True
    mats, each = run_info_gain_matrices(pe, mods, searchType)
    pe.middle_of_doe_flag = False
    pe.parModulationPermutationIndex = len(mods) - 1
    pe.info_gains_matrices_array_format = doe['info_gains_matrices_array_format']
    pe.info_gain_matrix = mats[-1]                                   # what the last doeGetInfoGainMatrix call leaves behind
    pe.info_gains_matrices_array = mats                                                                # IP.py:1550
    if each is not None:
        pe.info_gain_matrices_each_parameter = [e[-1] for e in each]
        pe.info_gains_matrices_arrays_one_for_each_parameter = np.array(each)                          # IP.py:1551-1555
    return pe.info_gains_matrices_array
