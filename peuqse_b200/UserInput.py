"""Configuration namespace with the same dictionaries and keys as the reference's PEUQSE/UserInput.py, so
run-files written for PEUQSE keep working:  ``import peuqse_b200.UserInput as UserInput``.

Only keys that reach the accelerated path are interpreted; the others are accepted and ignored.
Unlike the reference (where UserInput is a process-wide module that parameter_estimation.__init__ writes
derived fields into), `fresh()` returns an independent copy so that several problems can coexist.
"""
import copy
import types

directories = {'graphs': "./graphs/", 'logs_and_csvs': "./logs_and_csvs/", 'pickles': "./pickles/"}

responses = {
    'responses_abscissa': [],
    'responses_observed': [],
    'responses_observed_uncertainties': [],
    'responses_observed_max_covmat_size': 100,
    'responses_observed_covmat_whitening': 'eigh',   # peuqse_b200 only: 'eigh' = scipy's _PSD (the reference's arithmetic); 'cholesky' = triangular whitening, half the flops
    'responses_observed_weighting': [],
    'reducedResponseSpace': [],
    'independent_variables_values': [],
    'independent_variables_names': [],
    'num_responses': 'auto',
    'data_overcategory': '',
    'response_types': [],
    'response_data_types': [],
}

model = {
    'InputParameterPriorValues': [],
    'InputParametersPriorValuesUncertainties': [],
    'InputParameterInitialGuess': [],
    'parameterNamesAndMathTypeExpressionsDict': {},
    'populateIndependentVariablesFunction': None,
    'simulateByInputParametersOnlyFunction': None,
    'simulationOutputProcessingFunction': None,
    'reducedParameterSpace': [],
    'responses_simulation_uncertainties': None,
    'custom_logLikelihood': None,
    'custom_logPrior': None,
    'InputParameterPriorValues_upperBounds': [],
    'InputParameterPriorValues_lowerBounds': [],
    'simulatedResponses_upperBounds': [],
    'simulatedResponses_lowerBounds': [],
    'PosteriorParameterBounds': {},
    'exportResponses': False,
}

parameter_estimation_settings = {
    'verbose': False,
    'exportLog': True,
    'exportAllSimulatedOutputs': False,
    'checkPointFrequency': None,
    'scaling_uncertainties_type': "std",
    'undo_scaling_uncertainties_type': False,
    'convergence_diagnostics': False,
    'mcmc_exportLog': True,            # as the reference (UserInput.py:62): sampler results are written to logs_and_csvs/ and pickles/
    'mcmc_random_seed': None,
    'mcmc_mode': 'unbiased',
    'mcmc_length': 10000,
    'mcmc_burn_in': 'auto',
    'mcmc_relative_step_length': 0.1,
    'mcmc_modulate_accept_probability': 0,
    'mcmc_info_gain_cutoff': 0,
    'mcmc_info_gain_returned': 'KL_divergence',
    'mcmc_threshold_filter_samples': True,
    'mcmc_threshold_filter_coefficient': 'auto',
    'mcmc_threshold_filter_benchmark': 'auto',
    'mcmc_nwalkers': 'auto',
    'mcmc_maxiter': 1E6,
    'mcmc_walkerInitialDistribution': 'auto',
    'mcmc_walkerInitialDistributionSpread': 'auto',
    'mcmc_checkPointFrequency': None,
    'mcmc_parallel_sampling': False,
    'mcmc_continueSampling': False,
    'mcmc_store_samplingObject': False,
    'mcmc_movesType': 'auto',
    'multistart_searchType': 'getLogP',
    'multistart_checkPointFrequency': None,
    'multistart_parallel_sampling': False,
    'multistart_centerPoint': None,
    'multistart_numStartPoints': 0,
    'multistart_initialPointsDistributionType': 'sobol',
    'multistart_relativeInitialDistributionSpread': 0.866,
    'multistart_gridsearchSamplingInterval': [],
    'multistart_gridsearchSamplingRadii': [],
    'multistart_permutationsToSamples': True,
    'multistart_permutationsToSamples_threshold_filter_samples': True,
    'multistart_permutationsToSamples_threshold_filter_coefficient': 'auto',
    'multistart_continueSampling': 'auto',
    'multistart_passThroughArgs': {},
    'multistart_calculatePostBurnInStatistics': True,
    'multistart_keep_cumulative_post_burn_in_data': False,
    'multistart_exportLog': False,
    # ---- additions of this engine (absent keys in a reference UserInput mean the defaults below)
    'multistart_batched_optimizer': True,   # Nelder-Mead multistarts advance all start points in lock step (peuqse_b200/optimize.py)
    'export_max_rows': 1_000_000,      # sample / point arrays with more rows than this are not written to file
    'mcmc_nchains': 1,                 # independent MH chains advanced by one kernel (reference: one per MPI rank)
    'mcmc_store_samples': 'auto',      # 'auto': keep every row while it fits comfortably in device memory
    'mcmc_thin': 1,
    'multistart_store_logP': 'auto',   # materialise logP of every sweep point in HBM ('auto': when <= 2^31 points)
    'multistart_shard_pieces': 'auto', # multi-GPU: index-range pieces per rank, dealt round-robin ('auto': 1 = one contiguous range; more balances pass 2-3 of centre-heavy grids at the cost of extra launches)
    'multistart_random_seed': None,    # Philox seed of the uniform sweep (None -> mcmc_random_seed or 0)
    'multistart_histogram_bins': 10,
}

plotting_ouput_settings = {'setMatPlotLibAgg': 'auto'}
histogram_plot_settings = {}
simulated_response_plot_settings = {}
scatter_matrix_plots_settings = {}
scatter_heatmap_plots_settings = {}
contour_plot_settings = {}
doe_settings = {'parallel_conditions_exploration': False, 'parallel_parameter_modulation': False}

_DICTS = ('directories', 'responses', 'model', 'parameter_estimation_settings', 'plotting_ouput_settings',
          'histogram_plot_settings', 'simulated_response_plot_settings', 'scatter_matrix_plots_settings',
          'scatter_heatmap_plots_settings', 'contour_plot_settings', 'doe_settings')
_DEFAULTS = {name: copy.deepcopy(globals()[name]) for name in _DICTS}


def fresh():
    """An independent namespace holding the default dictionaries."""
    ns = types.SimpleNamespace()
    for name in _DICTS:
        setattr(ns, name, copy.deepcopy(_DEFAULTS[name]))
    return ns


def reset():
    """Restore this module's dictionaries to their defaults (what importlib.reload does for the reference)."""
    g = globals()
    for name in _DICTS:
        g[name].clear()
        g[name].update(copy.deepcopy(_DEFAULTS[name]))
    for name in list(g):
        if name not in _KEEP:
            del g[name]


_KEEP = set(globals()) | {'_KEEP'}
