"""Generates tests/golden/info_gain_cutoff.npz: calculatePostBurnInStatistics of the UNMODIFIED reference with
parameter_estimation_settings['mcmc_info_gain_cutoff'] != 0 (the truncated log-ratio information gain, IP.py:1805-1843) on the
sample arrays of post_burn_in_statistics.npz.  Run in the build container only:  python tests/golden/make_golden_cutoff.py
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as G  # noqa: E402

src = np.load(os.path.join(G.HERE, "post_burn_in_statistics.npz"))
th, logp, logprior = src['pbs_samples'], src['pbs_logP'], src['pbs_log_priors']
UI = G.ui_example00()
UI.parameter_estimation_settings['mcmc_info_gain_returned'] = 'log_ratio'
PE = G.make_pe(UI)
out = {}
# cutoffs in probability units: exp(-46.05) keeps part of the rows, 1e-5 removes every row (mean of nothing = NaN),
# 1e-300 keeps them all; with and without the tail filter in front
for name, filt, cutoff in (('off_1e-20', False, 1e-20), ('auto_1e-20', True, 1e-20), ('off_1e-5', False, 1e-5),
                           ('off_1e-300', False, 1e-300), ('auto_2e-20', True, 2e-20)):
    pes = PE.UserInput.parameter_estimation_settings
    pes['mcmc_threshold_filter_samples'] = filt
    pes['mcmc_threshold_filter_coefficient'] = 'auto'
    pes['mcmc_threshold_filter_benchmark'] = 'auto'
    pes['mcmc_info_gain_cutoff'] = cutoff
    PE.post_burn_in_samples = th.copy()
    PE.post_burn_in_log_posteriors_un_normed_vec = logp.copy()
    PE.post_burn_in_log_priors_vec = logprior.copy()
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = PE.calculatePostBurnInStatistics()
    out['case_' + name] = np.array([float(filt), cutoff, float(res[4]), float(res[3]), float(len(res[5]))])
path = os.path.join(G.HERE, "info_gain_cutoff.npz")
np.savez_compressed(path, **out)
for k, v in out.items():
    print(k, v)
