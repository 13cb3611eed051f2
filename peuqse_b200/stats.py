"""Post-burn-in statistics on device-resident samples (csrc/stats.cu): calculatePostBurnInStatistics
(PEUQSE/InverseProblem.py:1861-1937, "IP.py") with its tail filter (arrayThresholdFilter, IP.py:3635-3664) and
calculateInfoGain (IP.py:1770-1847).  The sampler kernels leave samples / logP / logPrior in HBM; these passes read them
there, so the host only ever sees the ~1 KB packed result (and, if it asks for them, the kept rows)."""
import ctypes as C

import numpy as np
import torch

from . import _lib

_FILTERS = {'auto': _lib.FILTER_AUTO, 'MAP': _lib.FILTER_MAP, 'mu_AP': _lib.FILTER_MU_AP}


class DeviceStatistics:
    """Result of device_statistics(): attributes named like the reference's (map_parameter_set, mu_AP_parameter_set, ...)."""

    def kept_rows(self):
        """(samples[n_kept,k], logP[n_kept,1], logPrior[n_kept,1] or None) as CUDA tensors, in the reference's row order."""
        return self._compact()


def filter_settings(pes):
    """(filter mode, coefficient) from parameter_estimation_settings (IP.py:1871-1876, 1901-1918)."""
    if not bool(pes['mcmc_threshold_filter_samples']):
        return _lib.FILTER_NONE, 0.0
    coef = pes['mcmc_threshold_filter_coefficient']
    coef = 2.0 if (isinstance(coef, str) and coef.lower() == 'auto') else float(coef)
    bench = pes['mcmc_threshold_filter_benchmark']
    if bench not in _FILTERS:
        # the reference's if/elif chain leaves truncatedMergedArray undefined for any other string (IP.py:1901-1918)
        raise ValueError(f"mcmc_threshold_filter_benchmark must be 'auto', 'MAP' or 'mu_AP', got {bench!r}")
    return _FILTERS[bench], coef


def device_statistics(samples, logp, logprior=None, filter_mode=_lib.FILTER_NONE, coef=2.0, prior_samples=None, kl_bins=100,
                      logp_fn=None):
    """samples[N,k], logp[N], logprior[N] (optional): contiguous float64 CUDA tensors.  prior_samples[Np,k] (optional, for
    the KL-divergence information gain).  logp_fn(theta[1,k] cuda) -> logP[1] cuda is needed for the 'mu_AP' benchmark only."""
    L = _lib.load()
    assert samples.is_cuda and samples.dtype == torch.float64 and samples.ndim == 2
    samples = samples.contiguous()
    N, k = samples.shape
    logp = logp.reshape(-1).contiguous()
    assert logp.numel() == N and logp.dtype == torch.float64
    if logprior is not None:
        logprior = logprior.reshape(-1).contiguous()
        assert logprior.numel() == N
    dev = samples.device
    spec = _lib.StatsSpec()
    spec.k, spec.filter, spec.coef, spec.kl_bins = k, int(filter_mode), float(coef), int(kl_bins if prior_samples is not None else 0)
    out = torch.empty(_lib.STATS['DOUBLES'], dtype=torch.float64, device=dev)
    work = torch.empty(int(L.peuqse_stats_work_bytes()), dtype=torch.uint8, device=dev)
    st = _lib.current_stream()
    _lib.check(L.peuqse_stats_begin(C.byref(spec), _lib.ptr(samples), _lib.ptr(logp), N, _lib.ptr(out), _lib.ptr(work), st))
    mu_lp = None
    if filter_mode == _lib.FILTER_MU_AP:
        if logp_fn is None:
            raise ValueError("the 'mu_AP' filter benchmark needs logp_fn")
        o = _lib.STATS['MEAN_ALL']
        mu_lp = logp_fn(out[o:o + k].reshape(1, k)).reshape(1).contiguous()          # getLogP(mean of all samples), IP.py:1900
    pri = None
    if prior_samples is not None:
        pri = torch.as_tensor(np.asarray(prior_samples, dtype=float) if not torch.is_tensor(prior_samples) else prior_samples,
                              dtype=torch.float64).to(dev).contiguous()
        assert pri.ndim == 2 and pri.shape[1] == k
    _lib.check(L.peuqse_stats_finish(C.byref(spec), _lib.ptr(samples), _lib.ptr(logp), _lib.ptr(logprior), N, _lib.ptr(mu_lp),
                                     _lib.ptr(pri), 0 if pri is None else pri.shape[0], _lib.ptr(out), _lib.ptr(work), st))
    h = out.cpu().numpy()                                                              # the one device->host read
    S = _lib.STATS
    R = DeviceStatistics()
    as_i64 = lambda x: int(np.array([x]).view(np.int64)[0])
    R.n_total, R.n_kept = int(h[S['N_TOTAL']]), int(h[S['N_KEPT']])
    R.cut, R.filtering_threshold = float(h[S['CUT']]), float(h[S['THRESHOLD']])
    R.map_logP = float(h[S['MAP_LOGP']])
    R.map_index, R.map_index_unfiltered = as_i64(h[S['MAP_INDEX']]), as_i64(h[S['MAP_INDEX_ALL']])
    R.map_parameter_set = h[S['MAP_THETA']:S['MAP_THETA'] + k].copy()
    R.mu_AP_parameter_set = h[S['MU_AP']:S['MU_AP'] + k].copy()
    R.stdap_parameter_set = h[S['STDAP']:S['STDAP'] + k].copy()
    R.evidence = float(h[S['EVIDENCE']])
    R.info_gain_log_ratio = float(h[S['INFO_GAIN_LOG_RATIO']]) if logprior is not None else None
    R.info_gain_KL = float(h[S['INFO_GAIN_KL']]) if pri is not None else None
    R.info_gain_KL_each_parameter = list(h[S['KL_EACH']:S['KL_EACH'] + k]) if pri is not None else None
    R.retained_min, R.retained_max = h[S['MIN']:S['MIN'] + k].copy(), h[S['MAX']:S['MAX'] + k].copy()
    R.gpu_launches = 4 + (1 if filter_mode == _lib.FILTER_AUTO else 0) + (2 if pri is not None else 0)

    def compact():
        o_s = torch.empty((R.n_kept, k), dtype=torch.float64, device=dev)
        o_l = torch.empty(R.n_kept, dtype=torch.float64, device=dev)
        o_p = torch.empty(R.n_kept, dtype=torch.float64, device=dev) if logprior is not None else None
        if R.n_kept:
            _lib.check(L.peuqse_stats_compact(C.byref(spec), _lib.ptr(samples), _lib.ptr(logp), _lib.ptr(logprior), N, _lib.ptr(out),
                                              _lib.ptr(o_s), _lib.ptr(o_l), _lib.ptr(o_p), _lib.ptr(work), _lib.current_stream()))
        return o_s, o_l.reshape(-1, 1), (o_p.reshape(-1, 1) if o_p is not None else None)
    R._compact = compact
    return R


def info_gain_log_ratio_cutoff(logp, logprior, evidence, cutoff):
    """The truncated log-ratio information gain of calculateInfoGain for mcmc_info_gain_cutoff != 0 (IP.py:1805-1843) on the
    device vectors of the kept rows: rows whose log-prior or log-posterior is NaN or below ln(cutoff) are dropped, the others
    contribute ln(exp(logP) / evidence) - logPrior with non-finite terms counted as 0; no rows left -> NaN (numpy's mean of
    an empty array)."""
    lp, pr = logp.reshape(-1), logprior.reshape(-1)
    lc = float(np.log(cutoff))
    keep = ~(torch.isnan(lp) | torch.isnan(pr) | (lp < lc) | (pr < lc))
    lp, pr = lp[keep], pr[keep]
    if lp.numel() == 0:
        return float('nan')
    ratios = torch.log(torch.exp(lp) / evidence) - pr
    ratios = torch.where(torch.isfinite(ratios), ratios, torch.zeros_like(ratios))
    return float(ratios.mean().item())
