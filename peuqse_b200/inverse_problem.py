"""parameter_estimation: the reference's user-facing class (PEUQSE/InverseProblem.py:25, "IP.py") with the
log-posterior hot path and its three data-parallel drivers running as sm_100a CUDA kernels.

Same names, argument meanings and return conventions as the reference for the accelerated path:
getLogP / getLogPrior / getLogLikelihood (IP.py:1647, 2756, 2850), doMultiStart (1077),
doMetropolisHastings (2563), doEnsembleSliceSampling (2362), doeGetInfoGainMatrix (1363),
doeParameterModulationPermutationsScanner (1499).  Everything outside that path (plots, file export, optimisers, MPI)
is out of scope and raises NotImplementedError.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

from . import _lib
from . import distributed as D
from .models import BuiltinModel, TorchSimulator
from .problem import DeviceProblem, build_host_problem
from .sweep import Sweep, grid_size, make_spec


def _dev():
    if not torch.cuda.is_available():
        raise _lib.PeuqseError("peuqse_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class parameter_estimation:
    software_name = "peuqse_b200"

    def __init__(self, UserInput=None):
        self.UserInput = UI = UserInput
        reduced = list(UI.model.get('reducedParameterSpace', []))
        if len(reduced) and not getattr(UI, '_reduced_inner', False):
            # model['reducedParameterSpace'] (IP.py:406-408, 519-561, 574-606): the engine is built on the reduced problem; UserInput.model
            # itself stays the user's full-space dictionary, as in the reference
            full_model = UI.model
            UI.model = self._reduced_model(UI, reduced)
            try:
                self._init(UI)
            finally:
                UI.model = full_model
            self.reducedParameterSpaceOn = True                                                        # IP.py:52
            return
        self._init(UI)
        self.reducedParameterSpaceOn = False

    @staticmethod
    def _reduced_model(UI, idx):
        """The model dictionary of the reduced problem.  Priors, bounds and names keep the ORIGINAL order of the selected entries
        (returnReducedIterable tests membership, IP.py:3559-3564) while reduced value j is placed at full index idx[j] by the
        simulation wrapper (IP.py:585-589); the two differ only for an unsorted index list, and the reference's behaviour is kept.
        The other parameters stay at model['InputParameterInitialGuess'] (the prior values when no guess is given)."""
        import copy
        full = UI.model
        kfull = len(full['InputParameterPriorValues'])
        keep = [i for i in range(kfull) if i in idx]
        guess = full.get('InputParameterInitialGuess', [])
        if isinstance(guess, str):
            raise NotImplementedError("reducedParameterSpace with a pickled initial guess")
        guess = np.array(full['InputParameterPriorValues'] if guess is None or len(guess) == 0 else guess, dtype=float).reshape(-1)
        m = dict(full)
        take = lambda a: [a[i] for i in keep]
        m['reducedParameterSpace'] = []
        m['InputParameterPriorValues'] = take(list(full['InputParameterPriorValues']))
        unc = np.asarray(full['InputParametersPriorValuesUncertainties'], dtype=float)
        m['InputParametersPriorValuesUncertainties'] = unc[np.ix_(keep, keep)] if unc.ndim == 2 else list(unc[keep])
        for b in ('InputParameterPriorValues_lowerBounds', 'InputParameterPriorValues_upperBounds'):
            if len(full[b]):
                m[b] = take(list(full[b]))
        m['InputParameterInitialGuess'] = list(guess[keep])
        names = full['parameterNamesAndMathTypeExpressionsDict']
        if len(names) == 0:
            names = {str(i): 'ParInd_' + str(i) for i in range(kfull)}                                 # IP.py:68-71
        m['parameterNamesAndMathTypeExpressionsDict'] = {key: names[key] for j, key in enumerate(names) if j in idx}
        sim, post = full['simulateByInputParametersOnlyFunction'], full.get('simulationOutputProcessingFunction')
        m['simulationOutputProcessingFunction'] = None                       # applied inside the wrapper (IP.py:522-523, 600-603)
        cols = torch.as_tensor(list(idx), dtype=torch.long)
        base = torch.as_tensor(guess, dtype=torch.float64)
        if isinstance(sim, BuiltinModel):
            if post is not None:
                raise NotImplementedError("simulationOutputProcessingFunction cannot follow a built-in CUDA model; fold it into a TorchSimulator")
            # a built-in model evaluates full parameter vectors: an inner engine on the full space does the simulation
            import types
            inner_ui = types.SimpleNamespace(**{n: (copy.copy(v) if isinstance(v, dict) else v) for n, v in vars(UI).items()
                                                if not n.startswith('__')})
            inner_ui.model = dict(full, reducedParameterSpace=[])
            inner_ui._reduced_inner = True
            inner = parameter_estimation(inner_ui)
            fn = inner.simulate_batch
        else:
            fn = sim if not isinstance(sim, TorchSimulator) else sim.fn

        def subset_simulator(th):                                            # simulateWithSubsetOfParameters, IP.py:574-606
            fullth = base.to(th.device).repeat(th.shape[0], 1)
            fullth[:, cols.to(th.device)] = th
            out = fn(fullth)
            return out if post is None else post(out)
        m['simulateByInputParametersOnlyFunction'] = TorchSimulator(subset_simulator)
        for f in ('custom_logLikelihood', 'custom_logPrior', 'responses_simulation_uncertainties'):
            if callable(full.get(f)):
                raise NotImplementedError(f"reducedParameterSpace together with a callable model['{f}']")
        return m

    def _transient_simulator(self, UI, sim, post, H):
        """The simulated side of transform_responses (IP.py:2890): the simulation (and the user's output processing) runs, then
        peuqse_transform_transient integrates the rate-type responses in place.  A built-in model simulates through an inner engine
        without the transform."""
        import copy
        import types
        n = [int(len(o)) for o in H.responses_observed]
        resp_n = (C.c_int32 * len(n))(*n)
        passes = (C.c_int32 * len(n))(*self._transient_passes)
        L = _lib.load()
        dt = {}
        if isinstance(sim, BuiltinModel):
            inner_ui = types.SimpleNamespace(**{a: (copy.copy(v) if isinstance(v, dict) else v) for a, v in vars(UI).items()
                                                if not a.startswith('__')})
            inner_ui.responses['data_overcategory'] = ''
            fn = parameter_estimation(inner_ui).simulate_batch
        else:
            fn = sim.fn

        def transformed(th):
            out = fn(th)
            if post is not None:
                out = post(out)
            if isinstance(out, (list, tuple)):
                out = torch.cat([torch.as_tensor(o, dtype=torch.float64, device=th.device).reshape(th.shape[0], -1) for o in out], dim=1)
            out = torch.as_tensor(out, dtype=torch.float64, device=th.device).reshape(th.shape[0], -1).contiguous().clone()
            if out.shape[1] != sum(n):
                raise ValueError(f"the simulation returned {out.shape[1]} values per point, the observed responses hold {sum(n)}")
            if th.device not in dt:
                dt[th.device] = torch.as_tensor(self._transient_dt, device=th.device)
            _lib.check(L.peuqse_transform_transient(_lib.ptr(out), _lib.ptr(dt[th.device]), resp_n, passes, len(n), out.shape[0],
                                                    _lib.current_stream()))
            return out
        return TorchSimulator(transformed)

    @staticmethod
    def _response_subset_simulator(UI, sim, post, keep):
        """extractReducedResponsesOutputsWrapper (IP.py:483-499): the full simulation (and the user's output processing) runs, then only
        the response dimensions named in responses['reducedResponseSpace'] are kept (by membership, in their original order).  A
        built-in model simulates through an inner engine on the full response set."""
        import copy
        import types
        from .problem import nest_responses
        nfull = [len(o) for o in nest_responses(UI.responses['responses_observed'])]
        offs = np.concatenate(([0], np.cumsum(nfull)))
        cols = torch.as_tensor(np.concatenate([np.arange(offs[r], offs[r + 1]) for r in range(len(nfull)) if r in keep]), dtype=torch.long)
        if isinstance(sim, BuiltinModel):
            inner_ui = types.SimpleNamespace(**{n: (copy.copy(v) if isinstance(v, dict) else v) for n, v in vars(UI).items()
                                                if not n.startswith('__')})
            inner_ui.responses['reducedResponseSpace'] = []
            fn = parameter_estimation(inner_ui).simulate_batch
        else:
            fn = sim.fn

        def subset(th):
            out = fn(th)
            if post is not None:
                out = post(out)
            if isinstance(out, (list, tuple)):
                out = torch.cat([torch.as_tensor(o, dtype=torch.float64, device=th.device).reshape(th.shape[0], -1) for o in out], dim=1)
            out = torch.as_tensor(out, dtype=torch.float64, device=th.device).reshape(th.shape[0], -1)
            return out[:, cols.to(th.device)]
        return TorchSimulator(subset)

    def _init(self, UI):
        pes = UI.parameter_estimation_settings
        for key, default in (('mcmc_nchains', 1), ('mcmc_store_samples', 'auto'), ('mcmc_thin', 1),
                             ('multistart_store_logP', 'auto'), ('multistart_random_seed', None), ('multistart_shard_pieces', 'auto'),
                             ('multistart_histogram_bins', 10)):
            pes.setdefault(key, default)
        # custom_logPrior(theta) -> logPrior and custom_logLikelihood(theta) -> (logLikelihood, simulatedResponses) replace the
        # built-in prior (bounds check included) / simulation + likelihood (IP.py:2757-2759, 2854-2857).  Here they are VECTORISED
        # torch callables over a block of points: custom_logPrior(theta[B,k]) -> [B], custom_logLikelihood(theta[B,k]) ->
        # ([B], sim[B,n] or None), evaluated on the batched paths (getLogP*, multistart, ensemble slice sampling, optimisers).
        self.custom_logLikelihood = UI.model.get('custom_logLikelihood')
        self.custom_logPrior = UI.model.get('custom_logPrior')
        keep_resp = list(UI.responses.get('reducedResponseSpace', []))
        self._transient = UI.responses.get('data_overcategory', '') == 'transient_kinetics'      # other categories transform nothing (IP.py:664-676)
        self.parameterBoundsOn = bool(len(UI.model['InputParameterPriorValues_upperBounds']) +
                                      len(UI.model['InputParameterPriorValues_lowerBounds']))          # IP.py:49
        UI.InputParameterPriorValues_lowerBounds = UI.model['InputParameterPriorValues_lowerBounds']
        UI.InputParameterPriorValues_upperBounds = UI.model['InputParameterPriorValues_upperBounds']
        if len(UI.model['parameterNamesAndMathTypeExpressionsDict']) == 0:                             # IP.py:68-71
            for i in range(len(UI.model['InputParameterPriorValues'])):
                UI.model['parameterNamesAndMathTypeExpressionsDict'][str(i)] = 'ParInd_' + str(i)
        UI.parameterNamesList = list(UI.model['parameterNamesAndMathTypeExpressionsDict'])
        UI.user_requested_convergence_diagnostics = bool(pes.get('convergence_diagnostics', False))       # IP.py:2741

        obs_unc = UI.responses['responses_observed_uncertainties']
        observed = UI.responses['responses_observed']
        if isinstance(obs_unc, (list, tuple, np.ndarray)) and len(obs_unc) == 0:
            # blank list: 5 % of the value with an orthogonal 2 % of the maximum (IP.py:223-229)
            from .problem import nest_responses
            obs_n = nest_responses(observed)
            obs_unc = [((np.abs(o) * 0.05) ** 2 + (np.max(np.abs(o)) * 0.02) ** 2) ** 0.5 for o in obs_n]
            if len(obs_n) == 1:
                obs_unc = obs_unc[0]
        weighting = UI.responses.get('responses_observed_weighting', [])
        if len(weighting) > 0:
            # IP.py:238-256: the observed uncertainties are multiplied by weighting**(-0.5) (zeros become machine epsilon).
            # (The reference's own code path stops at np.float under numpy >= 1.24; this follows its formula.)
            from .problem import nest_responses
            wts = [np.array(w_, dtype=float) for w_ in nest_responses(weighting)]
            unc_n = nest_responses(obs_unc)
            if any(np.ndim(u_) != 1 for u_ in unc_n) or len(wts) != len(unc_n) or any(len(w_) != len(u_) for w_, u_ in zip(wts, unc_n)):
                raise NotImplementedError("responses_observed_weighting needs one weight per observed standard deviation "
                                          "(not covariance matrices)")
            for w_ in wts:
                w_[w_ == 0] = np.finfo(float).eps
            UI.responses_observed_weight_coefficients = [w_ ** (-0.5) for w_ in wts]
            obs_unc = [u_ * c_ for u_, c_ in zip(unc_n, UI.responses_observed_weight_coefficients)]
            if len(obs_unc) == 1:
                obs_unc = obs_unc[0]
        sim_unc = UI.model.get('responses_simulation_uncertainties')
        # a FUNCTION of the parameters (IP.py:563-572): here a VECTORISED torch callable f(theta[B,k]) -> standard deviations
        # [B, n_total] (or a list of [B, n_r] per response); its squares are added to the observed variances point by point
        # (IP.py:2925-2927) by peuqse_loglik_simvar -- installed below as the likelihood of every batched path
        self._sim_unc_fn = sim_unc if callable(sim_unc) else None
        if self._sim_unc_fn is not None:
            sim_unc = None
            UI.responses_simulation_uncertainties = self._sim_unc_fn                                   # IP.py:276
        raw_observed = observed
        if self._transient:
            # transform_responses on the observed side (IP.py:295, 609-658), once; the simulated side is peuqse_transform_transient
            from .problem import nest_responses, transform_transient, transient_passes
            if len(keep_resp) or sim_unc is not None or self._sim_unc_fn is not None or self.custom_logLikelihood is not None:
                raise NotImplementedError("data_overcategory 'transient_kinetics' together with reducedResponseSpace, simulation "
                                          "uncertainties or a custom likelihood")
            nresp = len(nest_responses(observed))
            types, dtypes = list(UI.responses['response_types']), list(UI.responses['response_data_types'])
            if len(types) != nresp or len(dtypes) != nresp:
                raise ValueError("responses['response_types'] and ['response_data_types'] need one entry per response")
            ab = nest_responses(UI.responses['responses_abscissa'])
            obs_t, unc_t = transform_transient(ab, types, dtypes, observed, obs_unc)
            self._transient_passes = [transient_passes(a, b) for a, b in zip(types, dtypes)]
            self._transient_dt = np.concatenate([np.append(np.diff(np.asarray(ab[0] if len(ab) == 1 else ab[r], dtype=float)), 0.0)
                                                 for r in range(nresp)])
            observed, obs_unc = (obs_t, unc_t) if nresp > 1 else (obs_t[0], unc_t[0])
        H = build_host_problem(UI.model['InputParameterPriorValues'], UI.model['InputParametersPriorValuesUncertainties'],
                               observed, obs_unc,
                               lower_bounds=UI.model['InputParameterPriorValues_lowerBounds'],
                               upper_bounds=UI.model['InputParameterPriorValues_upperBounds'],
                               scaling_type=pes['scaling_uncertainties_type'],
                               simulation_uncertainties=sim_unc,
                               max_covmat_size=UI.responses['responses_observed_max_covmat_size'],
                               covmat_whitening=UI.responses.get('responses_observed_covmat_whitening', 'eigh'),
                               keep_responses=keep_resp)
        self.host_problem = H
        # reference-named derived fields, written back into UserInput as the reference does (IP.py:146-350)
        UI.mu_prior, UI.std_prior, UI.var_prior, UI.covmat_prior = H.mu_prior, H.std_prior, H.var_prior, H.covmat_prior
        UI.mu_prior_scaled, UI.var_prior_scaled, UI.covmat_prior_scaled = H.mu_prior_scaled, H.var_prior_scaled, H.covmat_prior_scaled
        if H.scaling is not None:
            UI.scaling_uncertainties = H.scaling
        if H.uniform_idx is not None:
            UI.InputParametersPriorValuesUniformDistributionsIndices = H.uniform_idx
        UI.responses_observed = H.responses_observed if not self._transient else nest_responses(raw_observed)
        UI.responses_observed_transformed = H.responses_observed
        UI.num_response_dimensions = len(H.responses_observed)
        UI.num_data_points = int(sum(len(r) for r in H.responses_observed))
        self.observed_responses_covmat_transformed = H.responses_cov
        seed = pes.get('mcmc_random_seed')
        if type(seed) == type(1):                                                                     # IP.py:191-193
            np.random.seed(seed)
        self.samples_of_prior = np.random.multivariate_normal(H.mu_prior, H.covmat_prior, pes['mcmc_length'])  # IP.py:194
        self.Q_mu = H.mu_prior * 0                                                                    # IP.py:349
        self.Q_covmat = H.covmat_prior                                                                # IP.py:350
        guess = UI.model.get('InputParameterInitialGuess', [])
        if isinstance(guess, str):
            # a pickle file written by an earlier run, e.g. 'mcmc_last_point_sampled' (IP.py:355-380): looked for in the
            # working directory, then in the pickles directory; of a (walkers, parameters) array the first walker is taken
            from .export import unpickleAnObject
            name = guess.replace('.pkl', '')
            if os.path.exists(name + '.pkl'):
                loaded = unpickleAnObject(name)
            elif os.path.exists(UI.directories['pickles'] + name + '.pkl'):
                loaded = unpickleAnObject(UI.directories['pickles'] + name)
            else:
                print('The pickled object for initial guess points must exist in base directory or pickles directory. The pickled file should have extension of ".pkl"')
                sys.exit()
            loaded = np.array(loaded)
            if loaded.ndim not in (1, 2):
                print('The shape of the initial guess pickled array should be (num_walkers, num_parameters). The current pickled array does not have this shape and has:', loaded.shape)
                sys.exit()
            guess = loaded if loaded.ndim == 1 else loaded[0]
        if guess is None or len(guess) == 0:
            guess = np.array(UI.model['InputParameterPriorValues'], dtype='float')                    # IP.py:383-384
        UI.InputParameterInitialGuess = np.array(guess, dtype='float').reshape(-1)

        sim = UI.model['simulateByInputParametersOnlyFunction']
        if sim is None and self.custom_logLikelihood is not None:
            sim = TorchSimulator(None)         # never called: the custom likelihood does its own simulation
        if sim is None:
            raise ValueError("UserInput.model['simulateByInputParametersOnlyFunction'] must be provided")
        if UI.model.get('simulationOutputProcessingFunction') is not None and isinstance(sim, BuiltinModel):
            raise NotImplementedError("simulationOutputProcessingFunction cannot follow a built-in CUDA model; fold it into a TorchSimulator")
        if not isinstance(sim, (BuiltinModel, TorchSimulator)):
            sim = TorchSimulator(sim)      # a bare callable is taken as a vectorised torch simulator (no CPU path)
        self._post_fn = UI.model.get('simulationOutputProcessingFunction')     # None in a reduced space: applied inside the wrapper
        if self._transient and any(self._transient_passes):
            sim, self._post_fn = self._transient_simulator(UI, sim, self._post_fn, H), None
        if len(keep_resp):
            if self._sim_unc_fn is not None or self.custom_logLikelihood is not None:
                raise NotImplementedError("reducedResponseSpace together with a callable likelihood / simulation-uncertainty function")
            sim, self._post_fn = self._response_subset_simulator(UI, sim, self._post_fn, keep_resp), None
        self.simulator = sim
        _dev()
        self.device_problem = DeviceProblem(H, sim)
        if isinstance(sim, BuiltinModel):
            sim._owner = self
        self._L = _lib.load()
        self.k = H.k
        if self._sim_unc_fn is not None and self.custom_logLikelihood is None:
            self._setup_simulation_uncertainty_function(H)
        self.prepareResponsesForSplitLikelihood = False   # the reference's fall-back mode is not replicated
        self.gpu_launches = 0

    # ------------------------------------------------------------------------------------------
    # batched evaluation (device tensors in, device tensors out) -- the engine under everything else
    # ------------------------------------------------------------------------------------------
    def _as_theta(self, thetas):
        t = torch.as_tensor(np.asarray(thetas, dtype=float) if not torch.is_tensor(thetas) else thetas,
                            dtype=torch.float64)
        if t.ndim == 1:
            t = t.reshape(1, -1)
        if t.shape[1] != self.k:
            raise ValueError(f"expected {self.k} parameters per point, got {t.shape[1]}")
        return t.to(_dev()).contiguous()

    def simulate_batch(self, thetas):
        """sim[B, n_total] (natural order) on the device."""
        th = self._as_theta(thetas)
        B = th.shape[0]
        if isinstance(self.simulator, BuiltinModel):
            out = torch.empty((B, self.host_problem.n_total), dtype=torch.float64, device=th.device)
            _lib.check(self._L.peuqse_simulate_batch(self.device_problem.handle, _lib.ptr(th), B, _lib.ptr(out),
                                                     _lib.current_stream()))
            self.gpu_launches += 1
            return out
        out = self.simulator(th)
        post = self._post_fn
        if post is not None:
            out = post(out)
        out = torch.as_tensor(out, dtype=torch.float64, device=th.device)
        out = out.reshape(B, out.numel() // B if B else self.host_problem.n_total).contiguous()
        if out.shape[1] != self.host_problem.n_total:
            raise ValueError(f"simulator returned {out.shape[1]} values per point, expected {self.host_problem.n_total}")
        return out

    def getLogP_batch(self, thetas, return_parts=False):
        """logP[B] (and logPrior[B], logLikelihood[B]) as CUDA tensors for thetas[B,k]."""
        th = self._as_theta(thetas)
        B = th.shape[0]
        if B == 0:
            e = torch.empty(0, dtype=torch.float64, device=th.device)
            return (e, e.clone(), e.clone()) if return_parts else e
        if self.custom_logLikelihood is not None or self.custom_logPrior is not None:
            return self._getLogP_batch_custom(th, return_parts)
        logp = torch.empty(B, dtype=torch.float64, device=th.device)
        lp = torch.empty(B, dtype=torch.float64, device=th.device) if return_parts else None
        ll = torch.empty(B, dtype=torch.float64, device=th.device) if return_parts else None
        h = self.device_problem.handle
        if isinstance(self.simulator, BuiltinModel):
            _lib.check(self._L.peuqse_logp_batch(h, _lib.ptr(th), B, _lib.ptr(logp), _lib.ptr(lp), _lib.ptr(ll),
                                                 _lib.current_stream()))
        else:
            sim = self.simulate_batch(th)
            _lib.check(self._L.peuqse_logp_from_sim(h, _lib.ptr(th), _lib.ptr(sim), B, _lib.ptr(logp), _lib.ptr(lp),
                                                    _lib.ptr(ll), _lib.current_stream()))
        self.gpu_launches += 1
        return (logp, lp, ll) if return_parts else logp

    def _setup_simulation_uncertainty_function(self, H):
        var = []
        for r, c in enumerate(H.responses_cov):
            c = np.asarray(c, dtype=float)
            if c.ndim == 2:
                if np.count_nonzero(c - np.diag(np.diagonal(c))):
                    raise NotImplementedError("responses_simulation_uncertainties as a function of the parameters next to a full "
                                              "observed covariance matrix needs a factorisation per point; not part of the accelerated path")
                c = np.diagonal(c)
            var.append(c.reshape(-1))
        dev = _dev()
        self._simvar_y = torch.as_tensor(np.concatenate([np.asarray(y, dtype=float).reshape(-1) for y in H.responses_observed]), device=dev)
        self._simvar_var = torch.as_tensor(np.concatenate(var), device=dev)
        n = [int(len(v)) for v in var]
        maxcov = self.UserInput.responses['responses_observed_max_covmat_size']
        self._simvar_n = (C.c_int32 * len(n))(*n)
        self._simvar_pv = (C.c_int32 * len(n))(*[1 if m > maxcov else 0 for m in n])                  # IP.py:2952
        self._simvar_total = int(sum(n))
        self.custom_logLikelihood = self._loglik_with_simulation_uncertainties

    def _loglik_with_simulation_uncertainties(self, th):
        """(logLikelihood[B], sim[B, n_total]) with the covariance observed variances + f(theta)^2 (IP.py:2883-2896, 2925-2927)."""
        B = th.shape[0]
        sim = self.simulate_batch(th).contiguous()
        sd = self._sim_unc_fn(th)
        if isinstance(sd, (list, tuple)):
            sd = torch.cat([torch.as_tensor(x, dtype=torch.float64, device=th.device).reshape(B, -1) for x in sd], dim=1)
        sd = torch.as_tensor(sd, dtype=torch.float64, device=th.device).reshape(B, self._simvar_total).contiguous()
        if sim.shape != sd.shape:
            raise ValueError(f"responses_simulation_uncertainties returned {tuple(sd.shape)}, the simulation {tuple(sim.shape)}")
        ll = torch.empty(B, dtype=torch.float64, device=th.device)
        _lib.check(self._L.peuqse_loglik_simvar(_lib.ptr(sim), _lib.ptr(sd), _lib.ptr(self._simvar_y), _lib.ptr(self._simvar_var),
                                                self._simvar_n, self._simvar_pv, len(self._simvar_n), B, _lib.ptr(ll),
                                                _lib.current_stream()))
        return ll, sim

    def _getLogP_batch_custom(self, th, return_parts):
        """getLogP with a user-supplied prior and / or likelihood (vectorised torch callables, see __init__): the part the user
        did not supply still comes from the kernels; logP = logLikelihood + logPrior, -inf where the prior is -inf (IP.py:1647-1654)."""
        B = th.shape[0]
        h = self.device_problem.handle
        as_vec = lambda v: torch.as_tensor(v, dtype=torch.float64, device=th.device).reshape(B).contiguous()
        ll = lp = None
        if self.custom_logLikelihood is not None:
            out = self.custom_logLikelihood(th)
            ll_raw, sim = out if isinstance(out, (tuple, list)) else (out, None)
            ll = as_vec(ll_raw)
            self._last_custom_sim = sim
        if self.custom_logPrior is not None:
            lp = as_vec(self.custom_logPrior(th))
        if ll is None:                         # built-in likelihood (and the prior, if it is not the user's)
            ll = torch.empty(B, dtype=torch.float64, device=th.device)
            lpk = torch.empty(B, dtype=torch.float64, device=th.device) if lp is None else None
            if isinstance(self.simulator, BuiltinModel):
                _lib.check(self._L.peuqse_logp_batch(h, _lib.ptr(th), B, None, _lib.ptr(lpk), _lib.ptr(ll), _lib.current_stream()))
            else:
                sim = self.simulate_batch(th)
                _lib.check(self._L.peuqse_logp_from_sim(h, _lib.ptr(th), _lib.ptr(sim), B, None, _lib.ptr(lpk), _lib.ptr(ll),
                                                        _lib.current_stream()))
            lp = lpk if lp is None else lp
        elif lp is None:                       # the user's likelihood with the built-in prior
            lp = torch.empty(B, dtype=torch.float64, device=th.device)
            _lib.check(self._L.peuqse_logprior_batch(h, _lib.ptr(th), B, _lib.ptr(lp), _lib.current_stream()))
        self.gpu_launches += 1
        logp = torch.where(lp == float('-inf'), lp, ll + lp)
        return (logp, lp, ll) if return_parts else logp

    # ------------------------------------------------------------------------------------------
    # single-point API with the reference's signatures
    # ------------------------------------------------------------------------------------------
    def getLogP(self, proposal_sample):                                    # IP.py:1647
        return float(self.getLogP_batch(proposal_sample)[0].item())

    def getNegLogP(self, proposal_sample):                                 # IP.py:1656
        return -1 * self.getLogP(proposal_sample)

    def getLogPrior(self, discreteParameterVector):                        # IP.py:2756
        return float(self.getLogP_batch(discreteParameterVector, return_parts=True)[1][0].item())

    def getSimulatedResponses(self, discreteParameterVector):              # IP.py:2822
        if self.custom_logLikelihood is not None:                           # IP.py:2855-2857: the callable returns them
            out = self.custom_logLikelihood(self._as_theta(discreteParameterVector))
            sim = out[1] if isinstance(out, (tuple, list)) else None
            return None if sim is None else np.asarray(torch.as_tensor(sim).cpu()).reshape(-1)
        sim = self.simulate_batch(discreteParameterVector)[0].cpu().numpy()
        out, off = [], 0
        for r in self.host_problem.responses_observed:
            out.append(sim[off:off + len(r)])
            off += len(r)
        self.lastSimulatedResponses = out
        return out

    def getLogLikelihood(self, discreteParameterVector):                   # IP.py:2850 -> (logLikelihood, simulatedResponses)
        _, _, ll = self.getLogP_batch(discreteParameterVector, return_parts=True)
        ll = float(ll[0].item())
        if ll == float('-inf'):
            return ll, None                                                # IP.py:2865, 2882
        return ll, self.getSimulatedResponses(discreteParameterVector)

    # ------------------------------------------------------------------------------------------
    # multistart
    # ------------------------------------------------------------------------------------------
    def getGridPermutations(self, gridCenterVector, gridsearchSamplingInterval, gridsearchSamplingRadii):
        """Resolves the defaults like IP.py:755-770; returns (center, interval, radii, numPermutations)."""
        k = len(gridCenterVector)
        radii = np.ones(k, dtype='int') if len(gridsearchSamplingRadii) == 0 else np.array(gridsearchSamplingRadii, dtype='int')
        if len(gridsearchSamplingInterval) == 0:
            interval = np.asarray(self.UserInput.std_prior, dtype=float)    # IP.py:767
            if interval.ndim != 1:
                raise ValueError("the default grid interval needs 1-D prior uncertainties (SURVEY Q15)")
        else:
            interval = np.array(gridsearchSamplingInterval, dtype='float')
        return np.asarray(gridCenterVector, dtype=float), interval, radii, grid_size(radii)

    def generateInitialPoints(self, numStartPoints=0, initialPointsDistributionType='uniform',
                              relativeInitialDistributionSpread=1.0, numParameters=0, centerPoint=None,
                              gridsearchSamplingInterval=[], gridsearchSamplingRadii=[], seed=None):
        """IP.py:684-750.  'grid' and 'uniform' are generated on the device (index-addressed: grid decode /
        Philox); the quasi-random and Gaussian variants are drawn on the host exactly like the reference."""
        if str(centerPoint).lower() == 'none':
            centerPoint = np.array(self.UserInput.InputParameterInitialGuess) * 1.0
        centerPoint = np.asarray(centerPoint, dtype=float)
        kind = initialPointsDistributionType.lower()
        if kind not in ['grid', 'uniform', 'identical', 'gaussian', 'astroidal', 'sobol', 'shell']:
            kind = 'sobol'                                                                             # IP.py:693-695
        k = len(centerPoint) if numParameters == 0 else numParameters
        if kind == 'grid':
            c, h, r, n = self.getGridPermutations(centerPoint, gridsearchSamplingInterval, gridsearchSamplingRadii)
            spec = make_spec("grid", k, c, interval=h, radii=r)
            return self._points_from_spec(spec, 0, n)
        if numStartPoints == 0:
            numStartPoints = getattr(self, 'mcmc_nwalkers', 0) or 3 * k
        std = np.asarray(self.UserInput.std_prior, dtype=float)
        if kind == 'uniform':
            if seed is None:
                seed = self._sweep_seed()
            spec = make_spec("uniform", k, centerPoint, spread=relativeInitialDistributionSpread, std_prior=std, seed=seed)
            return self._points_from_spec(spec, 0, numStartPoints)
        if kind == 'identical':
            first = np.zeros((numStartPoints, k))
        elif kind == 'gaussian':
            first = np.random.randn(numStartPoints, k)                                                  # IP.py:710
        else:
            from warnings import catch_warnings, simplefilter
            from scipy.stats import qmc
            with catch_warnings():
                simplefilter("ignore")
                # IP.py:718-721, 741-744; `seed` (not in the reference, whose points differ from run to run) makes the
                # scrambling reproducible -- scipy's generator does not follow np.random.seed
                sob = qmc.Sobol(d=k, scramble=True, seed=seed).random(numStartPoints)
            if kind == 'sobol':
                first = -2 + 4 * sob                                                                    # IP.py:746
            else:
                first = -1 + 2 * sob
                neg = np.ones((numStartPoints, k), dtype=int)
                neg[first < 0] = -1
                first = np.abs(first)
                first = first ** k if kind == 'astroidal' else first ** (1 / k)                         # IP.py:729-732
                first = neg * first
                first *= 2
        return relativeInitialDistributionSpread * first * std + centerPoint                           # IP.py:749

    def _points_from_spec(self, spec, first, count):
        th = torch.empty((count, spec.k), dtype=torch.float64, device=_dev())
        _lib.check(self._L.peuqse_sweep_points(C.byref(spec), int(first), int(count), _lib.ptr(th), _lib.current_stream()))
        self.gpu_launches += 1
        return th.cpu().numpy()

    def _sweep_seed(self):
        pes = self.UserInput.parameter_estimation_settings
        s = pes.get('multistart_random_seed')
        if s is None:
            s = pes.get('mcmc_random_seed')
        return 0 if s is None else int(s)

    def doMultiStart(self, searchType='UserChoice', numStartPoints='UserChoice', relativeInitialDistributionSpread='UserChoice',
                     exportLog='UserChoice', initialPointsDistributionType='UserChoice', passThroughArgs='UserChoice',
                     calculatePostBurnInStatistics='UserChoice', keep_cumulative_post_burn_in_data='UserChoice',
                     walkerInitialDistribution='UserChoice', centerPoint=None, gridsearchSamplingInterval='UserChoice',
                     gridsearchSamplingRadii='UserChoice'):
        """IP.py:1077-1117 for searchType 'getLogP'.  Returns bestResultSoFar = [map_parameter_set, None x 6]."""
        pes = self.UserInput.parameter_estimation_settings
        if str(searchType) == 'UserChoice': searchType = pes['multistart_searchType']
        if str(initialPointsDistributionType) == 'UserChoice': initialPointsDistributionType = pes['multistart_initialPointsDistributionType']
        if str(numStartPoints) == 'UserChoice': numStartPoints = pes['multistart_numStartPoints']
        if str(relativeInitialDistributionSpread) == 'UserChoice': relativeInitialDistributionSpread = pes['multistart_relativeInitialDistributionSpread']
        if str(gridsearchSamplingInterval) == 'UserChoice': gridsearchSamplingInterval = pes['multistart_gridsearchSamplingInterval']
        if str(gridsearchSamplingRadii) == 'UserChoice': gridsearchSamplingRadii = pes['multistart_gridsearchSamplingRadii']
        if str(calculatePostBurnInStatistics) == 'UserChoice': calculatePostBurnInStatistics = pes['multistart_calculatePostBurnInStatistics']
        if numStartPoints == 0: numStartPoints = len(self.UserInput.InputParameterInitialGuess) * 3     # IP.py:1101-1102
        if relativeInitialDistributionSpread == 0: relativeInitialDistributionSpread = 0.866           # IP.py:1103-1104
        if searchType in ('doGetLogP', 'doSinglePoint'): searchType = 'getLogP'                        # IP.py:1105-1106
        if searchType not in ('getLogP', 'doOptimizeNegLogP', 'doOptimizeLogP', 'doMetropolisHastings', 'doEnsembleSliceSampling'):
            raise NotImplementedError(f"multistart searchType '{searchType}' (the emcee sampler or an SSR fit per start point) is outside the "
                                      "accelerated path; use 'getLogP', 'doMetropolisHastings', 'doEnsembleSliceSampling', 'doOptimizeNegLogP' "
                                      "or 'doOptimizeLogP'")
        if str(keep_cumulative_post_burn_in_data) == 'UserChoice':
            keep_cumulative_post_burn_in_data = pes.get('multistart_keep_cumulative_post_burn_in_data', False)
        if str(passThroughArgs) == 'UserChoice': passThroughArgs = pes.get('multistart_passThroughArgs', {})
        if str(centerPoint).lower() == 'none':
            centerPoint = np.array(self.UserInput.InputParameterInitialGuess) * 1.0
        centerPoint = np.asarray(centerPoint, dtype=float)
        kind = initialPointsDistributionType.lower()
        toSamples = bool(pes['multistart_permutationsToSamples']) and kind in ('grid', 'uniform', 'sobol', 'astroidal', 'shell')
        coef = pes['multistart_permutationsToSamples_threshold_filter_coefficient']
        coef = 2.0 if str(coef).lower() == 'auto' else float(coef)                                      # IP.py:1011-1014
        nbins = int(pes.get('multistart_histogram_bins', 10))
        k = self.k
        if searchType != 'getLogP':
            pts = self.generateInitialPoints(numStartPoints=numStartPoints, initialPointsDistributionType=kind,
                                             relativeInitialDistributionSpread=relativeInitialDistributionSpread, centerPoint=centerPoint,
                                             gridsearchSamplingInterval=gridsearchSamplingInterval, gridsearchSamplingRadii=gridsearchSamplingRadii)
            if searchType == 'doMetropolisHastings':
                return self._multistart_mh(pts, bool(calculatePostBurnInStatistics), bool(keep_cumulative_post_burn_in_data), centerPoint)
            if searchType == 'doEnsembleSliceSampling':
                return self._multistart_ess(pts, bool(calculatePostBurnInStatistics), bool(keep_cumulative_post_burn_in_data), centerPoint,
                                            walkerInitialDistribution)
            return self._multistart_optimize(searchType, pts, passThroughArgs, centerPoint)
        if kind == 'grid':
            c, h, r, n = self.getGridPermutations(centerPoint, gridsearchSamplingInterval, gridsearchSamplingRadii)
            spec = make_spec("grid", k, c, interval=h, radii=r)
        elif kind == 'uniform':
            n = int(numStartPoints)
            spec = make_spec("uniform", k, centerPoint, spread=relativeInitialDistributionSpread,
                             std_prior=np.asarray(self.UserInput.std_prior, dtype=float), seed=self._sweep_seed())
        else:
            pts = self.generateInitialPoints(numStartPoints=numStartPoints, initialPointsDistributionType=kind,
                                             relativeInitialDistributionSpread=relativeInitialDistributionSpread,
                                             centerPoint=centerPoint)
            return self.doListOfPermutationsSearch(pts, searchType='getLogP', permutationsToSamples=toSamples,
                                                   calculatePostBurnInStatistics=calculatePostBurnInStatistics, centerPoint=centerPoint)
        custom = self.custom_logLikelihood is not None or self.custom_logPrior is not None
        if custom or not isinstance(self.simulator, BuiltinModel) or self.host_problem.resp_mode.max() == _lib.RESP_MVN_FULL:
            # torch-callable simulators / full covariance / user-supplied prior or likelihood: decode points in chunks and
            # use the batched path
            return self._multistart_external(spec, n, toSamples and calculatePostBurnInStatistics, coef, nbins, centerPoint)
        # arrays left by an earlier search of another kind (list of permutations, optimiser or MH multistart, external simulator)
        # must not be mistaken for this sweep's: the properties below then read the sweep's own device buffers
        for stale in ('_perm_array', '_list_of_permutations'):
            if hasattr(self, stale):
                delattr(self, stale)
        store = pes.get('multistart_store_logP', 'auto')
        store = (n <= 2 ** 31) if store == 'auto' else bool(store)
        pieces = pes.get('multistart_shard_pieces', 'auto')            # multi-GPU: pieces per rank, dealt round-robin
        pieces = 1 if pieces == 'auto' else int(pieces)
        key = (bytes(spec), n, store, nbins, D.world(), pieces)
        if getattr(self, '_sweep_key', None) == key:
            sw = self._sweep                    # same sweep again: reuse the device buffers
            sw.launches = 0
        else:
            sw = Sweep(self.device_problem, spec, n, store_logp=store, nbins=nbins, pieces=pieces)
            self._sweep_key = key
        res = sw.run(coef=coef, statistics=bool(toSamples and calculatePostBurnInStatistics))
        self.gpu_launches += sw.launches
        self._sweep = sw
        return self._adopt_sweep_result(res, centerPoint)

    def _adopt_sweep_result(self, res, centerPoint):
        self.sweep_result = res
        self.highest_logP = res.map_logP                                                               # IP.py:935
        self.map_logP = res.map_logP                                                                   # IP.py:973
        self.map_parameter_set = res.map_parameter_set
        self.map_index = res.map_index
        self.UserInput.InputParameterInitialGuess = centerPoint                                        # IP.py:971
        if getattr(res, 'statistics', False):
            self.mu_AP_parameter_set = res.mu_AP_parameter_set
            self.stdap_parameter_set = res.stdap_parameter_set
            self.evidence = res.evidence
            self.info_gain = None      # the reference's log-ratio branch has no priors vector here
            self.marginal_histograms = (res.histogram_counts, res.histogram_edges)
        best = [self.map_parameter_set, None, None, None, None, None, None]                            # IP.py:862, 1074
        if bool(self.UserInput.parameter_estimation_settings.get('exportLog', True)) and res.map_index >= 0:
            from .export import exportMultistart                                                       # IP.py:1037-1072
            stats = bool(getattr(res, 'statistics', False))
            exportMultistart(self, centerPoint, [self.map_parameter_set, getattr(self, 'mu_AP_parameter_set', None) if stats else None,
                                                 getattr(self, 'stdap_parameter_set', None) if stats else None], stats,
                             initial_point=getattr(res, 'initial_point', None), optimizer=getattr(res, 'optimizer', False))
        return best

    # optimisers: scipy drives, every objective evaluation is one launch of the log-posterior kernel -------------------
    def doOptimizeNegLogP(self, simulationFunctionAdditionalArgs=(), method=None, optimizationAdditionalArgs={}, printOptimum=True,
                          verbose=True, maxiter=0):
        """IP.py:1670-1694: scipy.optimize.minimize(getNegLogP, InputParameterInitialGuess, method).  The optimiser is scipy's own
        code on the host, as in the reference; with a bit-exact getLogP it walks the reference's iterates."""
        import scipy.optimize
        initialGuess = self.UserInput.InputParameterInitialGuess
        options = {"maxiter": maxiter} if maxiter != 0 else ({"disp": True} if verbose else None)
        res = scipy.optimize.minimize(self.getNegLogP, initialGuess, method=method, options=options)
        self.map_parameter_set = res.x
        negLogP = res.fun
        self.map_logP = -1.0 * negLogP
        if printOptimum:
            print("Final results from doOptimizeNegLogP:", self.map_parameter_set, "final negLogP:", negLogP, "final logP:", self.map_logP)
        return [self.map_parameter_set, negLogP]

    def doOptimizeLogP(self, simulationFunctionAdditionalArgs=(), method=None, optimizationAdditionalArgs={}, printOptimum=True,
                       verbose=True, maxiter=0):
        """IP.py:1661-1666."""
        self.map_parameter_set, negLogP = self.doOptimizeNegLogP(simulationFunctionAdditionalArgs, method, optimizationAdditionalArgs,
                                                                 printOptimum, verbose, maxiter)
        self.map_logP = -1.0 * negLogP
        return [self.map_parameter_set, self.map_logP]

    find_MAP = doOptimizeLogP                                                                         # IP.py:1668

    def _multistart_optimize(self, searchType, pts, passThroughArgs, centerPoint):
        """The start-point loop of doListOfPermutationsSearch for an optimiser searchType (IP.py:844-956, 912-921)."""
        from .sweep import SweepResult
        pts = np.asarray(pts, dtype=float)
        self._list_of_permutations = pts
        highest, R, rows = float('-inf'), SweepResult(), []
        R.total, R.map_index, R.statistics, R.optimizer = len(pts), -1, False, True
        args = dict(passThroughArgs)
        batched = (str(args.get('method', '')).lower() == 'nelder-mead' and set(args) <= {'method', 'printOptimum', 'verbose', 'maxiter'}
                   and bool(self.UserInput.parameter_estimation_settings.get('multistart_batched_optimizer', True)) and len(pts) > 0)
        if batched:
            # every start point's simplex advances in lock step: one launch of the log-posterior kernel per Nelder-Mead stage
            from .optimize import nelder_mead_batch
            res = nelder_mead_batch(lambda q: (-1 * self.getLogP_batch(q)).cpu().numpy(), pts,
                                    maxiter=(args.get('maxiter', 0) or None))
            self.multistart_optimizer_info = res
            for i, p in enumerate(pts):
                logp = -1.0 * res['fun'][i]
                if logp > highest:                                                                    # IP.py:933
                    highest, R.map_index, R.map_parameter_set, R.initial_point = float(logp), i, np.copy(res['x'][i]), p
                rows.append(np.hstack((logp, res['x'][i])))                                           # IP.py:947
            pts_loop = []
        else:
            pts_loop = pts
        for i, p in enumerate(pts_loop):
            self.UserInput.InputParameterInitialGuess = p                                             # IP.py:858
            if searchType == 'doOptimizeLogP':
                x, logp = self.doOptimizeLogP(**passThroughArgs)
            else:
                x, neg = self.doOptimizeNegLogP(**passThroughArgs)
                logp = -1.0 * neg
            if logp > highest:                                                                        # IP.py:933
                highest, R.map_index, R.map_parameter_set, R.initial_point = float(logp), i, np.copy(x), p
            rows.append(np.hstack((logp, x)))                                                         # IP.py:947
        self._perm_array = np.array(rows)
        R.map_logP = highest
        if R.map_index < 0:
            R.map_parameter_set = np.full(self.k, np.nan)
        return self._adopt_sweep_result(R, centerPoint)

    def _multistart_mh(self, pts, calculatePostBurnInStatistics, keep_cumulative, centerPoint):
        """Multistart with searchType 'doMetropolisHastings' (IP.py:865-879, 975-990): the reference runs one whole chain per
        start point, one after the other; here the start points are the chains of ONE launch of the persistent MH kernel
        (chain c starts at point c and draws from Philox stream c), the per-start MAP comes from the kernel's own per-chain
        statistics, and the best start's rows go through calculatePostBurnInStatistics like the reference's bestResultSoFar."""
        from .mcmc import mh_launch, step_cholesky, post_burn_in_statistics, _export_if_requested
        from .sweep import SweepResult
        UI = self.UserInput
        pes = UI.parameter_estimation_settings
        pts = np.asarray(pts, dtype=float).reshape(-1, self.k)
        self._list_of_permutations = pts
        Lm = int(pes['mcmc_length'])
        burn = int(Lm * 0.1) if str(pes['mcmc_burn_in']).lower() == 'auto' else int(pes['mcmc_burn_in'])
        seed = pes.get('mcmc_random_seed')
        seed = int(np.random.SeedSequence().entropy % (2 ** 63)) if not isinstance(seed, int) else seed
        rank, size = D.world()
        c0, cn = D.shard_range(len(pts), rank, size)
        from .mcmc import step_cholesky as _chol
        out = mh_launch(self, pts[c0:c0 + cn], Lm, burn, seed, _chol(self.Q_covmat, pes['mcmc_relative_step_length']), chain_offset=c0,
                        modulate=float(pes['mcmc_modulate_accept_probability']), map_finding=pes['mcmc_mode'] == 'MAP_finding')
        lay = _lib.mh_stat_layout(self.k)
        st = out['stats']
        rows = torch.cat([st[:, lay['MAPLOGP']:lay['MAPLOGP'] + 1], st[:, lay['MAPTHETA']:lay['MAPTHETA'] + self.k]], dim=1)
        rows = D.allgather_rows(rows.contiguous()).cpu().numpy()                                   # [starts, 1 + k]: IP.py:947
        self._perm_array = rows
        lp = np.where(np.isnan(rows[:, 0]), -np.inf, rows[:, 0])
        R = SweepResult()
        R.total, R.statistics, R.optimizer = len(pts), False, False
        if not np.any(lp > -np.inf):
            R.map_logP, R.map_index, R.map_parameter_set = float('-inf'), -1, np.full(self.k, np.nan)
            return self._adopt_sweep_result(R, centerPoint)
        i = int(np.argmax(lp))                                                                     # first strict maximum, IP.py:933
        own = c0 <= i < c0 + cn
        j = i - c0 if own else 0
        pick = lambda t, lo, hi: D.allgather_rows((t[lo:hi, j] if own else t[0:0, 0]).contiguous())
        d_s, d_lp, d_pr = pick(out['samples'], burn, Lm), pick(out['logp'], burn, Lm), pick(out['logprior'], burn, Lm)
        self.discrete_chains_post_burn_in_samples = d_s.cpu().numpy()[:, None, :]
        self.post_burn_in_samples = d_s.cpu().numpy()
        self.post_burn_in_log_posteriors_un_normed_vec = d_lp.cpu().numpy().reshape(-1, 1)
        self.post_burn_in_log_priors_vec = d_pr.cpu().numpy().reshape(-1, 1)
        self.during_burn_in_samples = pick(out['samples'], 0, burn).cpu().numpy()
        self.during_burn_in_log_posteriors_un_normed_vec = pick(out['logp'], 0, burn).cpu().numpy().reshape(-1, 1)
        self.mcmc_last_point_sampled = self.post_burn_in_samples[-1]
        UI.last_MCMC_run_type = 'MH'
        if keep_cumulative:                                                                        # IP.py:869-879: every start, stacked
            flat = lambda t: D.allgather_rows(t[burn:].permute(1, 0, *range(2, t.ndim)).contiguous()).reshape(-1, *t.shape[2:]).cpu().numpy()
            self.cumulative_post_burn_in_samples = flat(out['samples'])
            self.cumulative_post_burn_in_log_posteriors_un_normed_vec = flat(out['logp']).reshape(-1, 1)
            self.cumulative_post_burn_in_log_priors_vec = flat(out['logprior']).reshape(-1, 1)
        R.map_logP, R.map_index, R.map_parameter_set, R.initial_point = float(rows[i, 0]), i, rows[i, 1:].copy(), pts[i]
        self.highest_logP = R.map_logP
        UI.InputParameterInitialGuess = pts[i]            # what the best start's own run recorded (mcmc_log_file.txt, IP.py:1971)
        best = [R.map_parameter_set, None, None, None, None, self.post_burn_in_samples, self.post_burn_in_log_posteriors_un_normed_vec]
        if calculatePostBurnInStatistics:
            from . import stats as ST
            self._dev_post = dict(samples=d_s, logp=d_lp, logprior=d_pr)
            post_burn_in_statistics(self)                   # the best start's own doMetropolisHastings (IP.py:866) ...
            if ST.filter_settings(pes)[0] != _lib.FILTER_NONE:
                post_burn_in_statistics(self)               # ... and once more on bestResultSoFar's filtered rows (IP.py:979-982)
            _export_if_requested(self, 'UserChoice')
            best = [self.map_parameter_set, self.mu_AP_parameter_set, self.stdap_parameter_set, self.evidence, self.info_gain,
                    self.post_burn_in_samples, self.post_burn_in_log_posteriors_un_normed_vec]
            R.statistics, R.mcmc = True, True
        self.sweep_result = R
        self.map_logP, self.map_index_of_start, self.map_parameter_set = R.map_logP, i, R.map_parameter_set
        UI.InputParameterInitialGuess = centerPoint                                                # IP.py:971
        if bool(pes.get('exportLog', True)):
            from .export import exportMultistart
            saved, self.map_index = getattr(self, 'map_index', None), i
            exportMultistart(self, centerPoint, best, R.statistics, initial_point=pts[i], mcmc=True)
            self.map_index = saved if saved is not None else i
        return best

    def _multistart_ess(self, pts, calculatePostBurnInStatistics, keep_cumulative, centerPoint, walkerInitialDistribution):
        """Multistart with searchType 'doEnsembleSliceSampling' (IP.py:832-836, 880-895, 933-947, 971-990): one ensemble per start
        point, one after the other as in the reference -- an ensemble is a batch already, its rounds are what fills the device.  The
        walkers of a run are spread around its start point (InputParameterInitialGuess = the start point, IP.py:858); 2k walkers when
        mcmc_nwalkers is 'auto' (IP.py:833-836); the start whose run holds the highest logP wins (strict '>', IP.py:933) and its rows
        go through calculatePostBurnInStatistics once more (IP.py:979-982).  With an integer mcmc_random_seed start i draws from
        seed + 1000003 i, so the runs are reproducible and distinct."""
        from .mcmc import _export_if_requested
        from .sweep import SweepResult
        UI = self.UserInput
        pes = UI.parameter_estimation_settings
        pts = np.asarray(pts, dtype=float).reshape(-1, self.k)
        self._list_of_permutations = pts
        nw = pes.get('mcmc_nwalkers', 'auto')
        nw = 2 * self.k if str(nw).lower() == 'auto' else int(nw)
        seed0 = pes.get('mcmc_random_seed')
        rows, best, best_i, best_state = [], None, -1, None
        self.highest_logP = float('-inf')                                                              # IP.py:824
        cum = []
        kept = ('post_burn_in_samples', 'post_burn_in_log_posteriors_un_normed_vec', 'post_burn_in_log_priors_vec',
                'discrete_chains_post_burn_in_samples', 'mcmc_last_point_sampled', 'ess_info')
        try:
            for i, pt in enumerate(pts):
                UI.InputParameterInitialGuess = pt                                                     # IP.py:858
                if isinstance(seed0, int):
                    pes['mcmc_random_seed'] = seed0 + 1000003 * i
                self.map_logP = float('-inf')                                                          # IP.py:881
                res = self.doEnsembleSliceSampling(mcmc_nwalkers_direct_input=nw, calculatePostBurnInStatistics=calculatePostBurnInStatistics,
                                                   walkerInitialDistribution=walkerInitialDistribution, continueSampling=False,
                                                   mcmc_exportLog=False)
                lp_i = float(np.ravel(self.map_logP)[0])
                rows.append(np.hstack((lp_i, np.asarray(self.map_parameter_set, dtype=float))))        # IP.py:947
                if keep_cumulative:                                                                    # IP.py:884-895
                    cum.append((self.post_burn_in_samples, self.discrete_chains_post_burn_in_samples,
                                getattr(self, 'post_burn_in_log_priors_vec', None), self.post_burn_in_log_posteriors_un_normed_vec))
                if lp_i > self.highest_logP:                                                           # IP.py:933-941
                    self.highest_logP, best, best_i = lp_i, res, i
                    best_map = np.array(self.map_parameter_set, dtype=float)
                    best_state = {n: getattr(self, n) for n in kept if hasattr(self, n)}
        finally:
            pes['mcmc_random_seed'] = seed0
            UI.InputParameterInitialGuess = centerPoint                                                # IP.py:971
        self._perm_array = np.vstack(rows)
        R = SweepResult()
        R.total, R.statistics, R.optimizer = len(pts), False, False
        if best is None:
            R.map_logP, R.map_index, R.map_parameter_set = float('-inf'), -1, np.full(self.k, np.nan)
            return self._adopt_sweep_result(R, centerPoint)
        for n, v in best_state.items():
            setattr(self, n, v)
        if keep_cumulative:
            self.cumulative_post_burn_in_samples = np.vstack([c[0] for c in cum])
            self.cumulative_discrete_chains_post_burn_in_samples = np.vstack([c[1] for c in cum])
            if all(c[2] is not None for c in cum):
                self.cumulative_post_burn_in_log_priors_vec = np.vstack([c[2] for c in cum])
            self.cumulative_post_burn_in_log_posteriors_un_normed_vec = np.vstack([c[3] for c in cum])
        self.map_logP, self.map_parameter_set = self.highest_logP, best_map                            # IP.py:973-974
        R.map_logP, R.map_index, R.map_parameter_set, R.initial_point = self.highest_logP, best_i, best_map, pts[best_i]
        out = [best_map, None, None, None, None, self.post_burn_in_samples, self.post_burn_in_log_posteriors_un_normed_vec]
        if calculatePostBurnInStatistics:
            self._dev_post = None
            self.calculatePostBurnInStatistics(calculate_post_burn_in_log_priors_vec=True)             # IP.py:979-982
            UI.InputParameterInitialGuess = pts[best_i]       # what the best start's own run recorded (mcmc_log_file.txt, IP.py:1971)
            _export_if_requested(self, 'UserChoice')
            UI.InputParameterInitialGuess = centerPoint
            out = [self.map_parameter_set, self.mu_AP_parameter_set, self.stdap_parameter_set, self.evidence, self.info_gain,
                   self.post_burn_in_samples, self.post_burn_in_log_posteriors_un_normed_vec]
            R.statistics, R.mcmc = True, True
        self.sweep_result = R
        self.map_index_of_start = best_i
        if bool(pes.get('exportLog', True)):
            from .export import exportMultistart
            saved, self.map_index = getattr(self, 'map_index', None), best_i
            exportMultistart(self, centerPoint, out, R.statistics, initial_point=pts[best_i], mcmc=True)
            self.map_index = saved if saved is not None else best_i
        return out

    def doGridSearch(self, searchType='getLogP', gridSamplingAbsoluteIntervalSize=[], gridSamplingNumOfIntervals=[], **kw):
        """Deprecated alias kept by the reference (IP.py:1120-1128)."""
        return self.doMultiStart(searchType=searchType, initialPointsDistributionType='grid',
                                 gridsearchSamplingInterval=gridSamplingAbsoluteIntervalSize,
                                 gridsearchSamplingRadii=gridSamplingNumOfIntervals)

    # lazily materialised big arrays of the reference's object state --------------------------------
    @property
    def permutations_MAP_logP_and_parameters_values(self):
        """[N, 1+k] = hstack((logP, theta)) of every point, in list order (IP.py:947, 1004).  Materialised on
        demand: 8*(1+k)*N bytes cross PCIe."""
        if hasattr(self, '_perm_array'):
            return self._perm_array
        sw = self._sweep
        if D.world()[1] != 1:
            raise NotImplementedError("gather of all sweep points across ranks is not provided; use rank-local ranges")
        if sw.logp is None:
            raise _lib.PeuqseError("logP was not stored for this sweep (multistart_store_logP)")
        th = sw.points(sw.first, sw.count)
        return torch.cat([sw.logp.reshape(-1, 1), th], dim=1).cpu().numpy()

    @property
    def listOfPermutations(self):
        if hasattr(self, '_list_of_permutations'):
            return self._list_of_permutations
        sw = self._sweep
        return sw.points(0, sw.total).cpu().numpy()

    def doListOfPermutationsSearch(self, listOfPermutations, numPermutations=None, searchType='getLogP', exportLog=True,
                                   walkerInitialDistribution='UserChoice', passThroughArgs={}, calculatePostBurnInStatistics=True,
                                   keep_cumulative_post_burn_in_data=False, centerPoint=None, permutationsToSamples=False):
        """IP.py:772-1074 for an explicit list of points and searchType 'getLogP': one batched evaluation."""
        if searchType in ('doGetLogP', 'doSinglePoint'): searchType = 'getLogP'
        if searchType != 'getLogP':
            raise NotImplementedError("only searchType 'getLogP' is accelerated")
        pes = self.UserInput.parameter_estimation_settings
        pts = np.asarray([np.asarray(p, dtype=float) for p in listOfPermutations], dtype=float)
        self._list_of_permutations = pts
        if str(centerPoint).lower() == 'none':
            centerPoint = np.array(self.UserInput.InputParameterInitialGuess) * 1.0
        logp = self.getLogP_batch(pts).cpu().numpy()
        self._perm_array = np.hstack([logp.reshape(-1, 1), pts])
        res = self._reduce_list(logp, pts, permutationsToSamples and calculatePostBurnInStatistics,
                                pes['multistart_permutationsToSamples_threshold_filter_coefficient'],
                                int(pes.get('multistart_histogram_bins', 10)))
        return self._adopt_sweep_result(res, centerPoint)

    def _reduce_list(self, logp, pts, statistics, coef, nbins):
        """Small explicit lists: the reductions on the host in numpy (same formulas as the kernels)."""
        from .sweep import SweepResult
        coef = 2.0 if str(coef).lower() == 'auto' else float(coef)
        R = SweepResult()
        lp = np.where(np.isnan(logp), -np.inf, logp)
        R.total = len(lp)
        if len(lp) == 0 or not np.any(lp > -np.inf):
            R.map_logP, R.map_index, R.map_parameter_set, R.statistics = float('-inf'), -1, np.full(self.k, np.nan), False
            return R
        i = int(np.argmax(lp))                                           # first maximum == strict '>' loop (IP.py:933)
        R.map_logP, R.map_index, R.map_parameter_set = float(lp[i]), i, pts[i].copy()
        R.statistics = bool(statistics)
        if not statistics:
            return R
        thr = R.map_logP + np.log(10 ** (-1 * coef))
        with np.errstate(over='ignore', invalid='ignore'):
            w = np.where(lp > thr, np.trunc(np.exp(lp - thr)), 0.0)
        keep = w > 0
        W = w.sum()
        R.threshold, R.sum_weights, R.n_retained = float(thr), W, int(keep.sum())
        R.mu_AP_parameter_set = (w[:, None] * pts).sum(axis=0) / W
        dlt = pts - R.mu_AP_parameter_set
        R.posterior_covariance = (w[:, None, None] * dlt[:, :, None] * dlt[:, None, :]).sum(axis=0) / W
        R.stdap_parameter_set = np.sqrt(np.diag(R.posterior_covariance))
        R.evidence = ((w * np.exp(lp)).sum() / W) / np.sqrt((w[:, None] * pts ** 2).sum())
        R.retained_min, R.retained_max = pts[keep].min(axis=0), pts[keep].max(axis=0)
        hs, es = [], []
        for d in range(self.k):
            h, e = np.histogram(pts[keep, d], bins=nbins, weights=w[keep])
            hs.append(h.astype(np.int64)); es.append(e)
        R.histogram_counts, R.histogram_edges = np.array(hs), np.array(es)
        return R

    def _multistart_external(self, spec, n, statistics, coef, nbins, centerPoint, chunk=1 << 22):
        """Sweep with a torch-callable simulator or full-covariance responses: points are decoded on the device
        in chunks, simulated by the callable, evaluated by logp_from_sim; logP stays in HBM for passes 2-3."""
        if D.world()[1] != 1:
            raise NotImplementedError("external-simulator sweeps are single-GPU for now")
        sw = Sweep(self.device_problem, spec, n, store_logp=True)
        for first in range(0, n, chunk):
            cnt = min(chunk, n - first)
            th = sw.points(first, cnt)
            sw.logp[first:first + cnt] = self.getLogP_batch(th)
        lp = torch.where(torch.isnan(sw.logp), torch.full_like(sw.logp, float('-inf')), sw.logp)
        from .sweep import SweepResult
        R = SweepResult()
        R.total = n
        mx = float(lp.max().item()) if n else float('-inf')
        if not (mx > float('-inf')):
            R.map_logP, R.map_index, R.map_parameter_set, R.statistics = float('-inf'), -1, np.full(self.k, np.nan), False
            self._sweep = sw
            return self._adopt_sweep_result(R, centerPoint)
        idx = int(torch.nonzero(lp == mx)[0].item())                      # first maximum
        pts_logp = lp.cpu().numpy()
        pts = sw.points(0, n).cpu().numpy()
        self._sweep = sw
        self._list_of_permutations = pts
        self._perm_array = np.hstack([sw.logp.cpu().numpy().reshape(-1, 1), pts])
        assert idx == int(np.argmax(pts_logp))
        res = self._reduce_list(sw.logp.cpu().numpy(), pts, statistics, coef, nbins)
        return self._adopt_sweep_result(res, centerPoint)

    def expandedPostBurnInSamples(self, max_rows=50_000_000):
        """post_burn_in_samples of a permutationsToSamples run: every retained point repeated
        int(exp(logP - threshold)) times in list order (IP.py:3480-3491).  Built on demand."""
        res = self.sweep_result
        arr = self.permutations_MAP_logP_and_parameters_values
        with np.errstate(over='ignore', invalid='ignore'):
            w = np.where(arr[:, 0] > res.threshold, np.trunc(np.exp(arr[:, 0] - res.threshold)), 0).astype(np.int64)
        if w.sum() > max_rows:
            raise MemoryError(f"{int(w.sum())} expanded rows requested; use the weighted statistics instead")
        exp = np.repeat(arr, w, axis=0)
        self.post_burn_in_samples = exp[:, 1:]
        self.post_burn_in_log_posteriors_un_normed_vec = exp[:, 0:1]
        return self.post_burn_in_samples

    # ------------------------------------------------------------------------------------------
    # Metropolis-Hastings
    # ------------------------------------------------------------------------------------------
    def doMetropolisHastings(self, calculatePostBurnInStatistics=True, mcmc_exportLog='UserChoice', continueSampling='auto'):
        from .mcmc import run_metropolis_hastings
        return run_metropolis_hastings(self, calculatePostBurnInStatistics, continueSampling, mcmc_exportLog)

    def doEnsembleSliceSampling(self, mcmc_nwalkers_direct_input=None, walkerInitialDistribution='UserChoice',
                                walkerInitialDistributionSpread='UserChoice', movesType='UserChoice',
                                calculatePostBurnInStatistics=True, mcmc_exportLog='UserChoice', continueSampling='auto'):
        from .mcmc import run_ensemble_slice_sampling
        return run_ensemble_slice_sampling(self, mcmc_nwalkers_direct_input, walkerInitialDistribution,
                                           walkerInitialDistributionSpread, movesType, calculatePostBurnInStatistics,
                                           continueSampling, mcmc_exportLog)

    def calculatePostBurnInStatistics(self, calculate_post_burn_in_log_priors_vec=False):
        from .mcmc import post_burn_in_statistics
        return post_burn_in_statistics(self, calculate_post_burn_in_log_priors_vec)

    # ------------------------------------------------------------------------------------------
    # design of experiments: every (modulation, condition) pair is one scenario of a device batch
    # ------------------------------------------------------------------------------------------
    def doeGetInfoGainMatrix(self, parameterPermutation, searchType='doMetropolisHastings'):     # IP.py:1363
        from .doe import doe_get_info_gain_matrix
        return doe_get_info_gain_matrix(self, parameterPermutation, searchType)

    def doeParameterModulationPermutationsScanner(self, searchType='doMetropolisHastings'):       # IP.py:1499
        from .doe import doe_parameter_modulation_scanner
        return doe_parameter_modulation_scanner(self, searchType)

    # ------------------------------------------------------------------------------------------
    # convergence diagnostics: Geweke z-scores of every chain, window and parameter on the device
    # ------------------------------------------------------------------------------------------
    def getConvergenceDiagnostics(self, discrete_chains_post_burn_in_samples=[], showFigure=None):   # IP.py:2121
        from .diagnostics import convergence_diagnostics
        return convergence_diagnostics(self, discrete_chains_post_burn_in_samples)

    # ------------------------------------------------------------------------------------------
    # on-disk formats
    # ------------------------------------------------------------------------------------------
    def exportPostBurnInStatistics(self):                                                          # IP.py:1956
        from .export import exportPostBurnInStatistics
        return exportPostBurnInStatistics(self)

    def exportResponses(self, choiceOrResponsesData='', filenameString='', consolidated='auto'):  # IP.py:2044
        from .export import exportResponses
        return exportResponses(self, choiceOrResponsesData, filenameString, consolidated)

    def reload_samples(self, sampling_type='', filepath=''):                                       # IP.py:441
        from .export import reload_samples
        return reload_samples(self, sampling_type, filepath)

    # out of scope ------------------------------------------------------------------------------
    def _out_of_scope(self, *a, **k):
        raise NotImplementedError("outside the accelerated hot path (see DESIGN.md, 'Out of scope')")

    createAllPlots = makeHistogramsForEachParameter = makeSamplingScatterMatrixPlot = createSimulatedResponsesPlots = _out_of_scope
    doOptimizeSSR = doEnsembleModifiedMHSampling = _out_of_scope
    save_to_dill = load_from_dill = createInfoGainPlots = createInfoGainModulationPlots = _out_of_scope
