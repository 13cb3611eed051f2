"""On-disk formats of the accelerated path: the files the reference writes after a sampler or a
multistart and reads back for continued sampling / re-analysis, byte-compatible so that downstream scripts, `createAllPlots`
of the reference and `reload_samples` keep working on results produced here.

  exportPostBurnInStatistics       PEUQSE/InverseProblem.py ("IP.py"):1956-2001
  exportMultistart                 IP.py:1037-1072 (tail of doListOfPermutationsSearch)
  exportResponses                  IP.py:2044-2119
  reload_samples                   IP.py:441-480
  pickleAnObject / unpickleAnObject IP.py:3685-3701

Everything here is host-side file I/O on arrays the kernels already returned; nothing is computed differently.  One addition:
arrays with more than parameter_estimation_settings['export_max_rows'] rows (default 1e6) are not written as CSV / pickle
(a notice is printed and the statistics pickle holds None in their place) -- the reference cannot reach such sizes.
"""
import os
import pickle
import sys

import numpy as np


def _dirs(UI):
    d = UI.directories
    for key in ('logs_and_csvs', 'pickles'):
        if d[key] not in ('', './'):
            os.makedirs(d[key], exist_ok=True)                                       # IP.py:44-46 creates them in __init__
    return d['logs_and_csvs'], d['pickles']


def pickleAnObject(objectToPickle, base_file_name, file_name_prefix='', file_name_suffix='', file_name_extension='.pkl'):
    try:                                                                             # IP.py:3685-3693
        base_file_name = base_file_name.replace(file_name_extension, "")
        with open(file_name_prefix + base_file_name + file_name_suffix + file_name_extension, 'wb') as picklefile:
            pickle.dump(objectToPickle, picklefile)
    except Exception as theError:
        print("pickleAnObject was unable to perform pickling. The error was:", theError)


def unpickleAnObject(base_file_name, file_name_prefix='', file_name_suffix='', file_name_extension='.pkl'):
    base_file_name = base_file_name.replace(file_name_extension, "")                 # IP.py:3695-3701
    with open(file_name_prefix + base_file_name + file_name_suffix + file_name_extension, 'rb') as picklefile:
        return pickle.load(picklefile)


def _max_rows(pe):
    return int(pe.UserInput.parameter_estimation_settings.get('export_max_rows', 1_000_000))


def _rank_suffix():
    """Per-rank file names of the reference's MPI mode (getParallelProcessingPrefixAndSuffix, IP.py:1939-1954): results are
    already consolidated over the ranks here, so only rank 0 writes and no suffix is used."""
    from . import distributed as D
    return D.world()[0]


def exportPostBurnInStatistics(pe):
    if _rank_suffix() != 0:
        return
    UI = pe.UserInput
    pes = UI.parameter_estimation_settings
    logs, pkls = _dirs(UI)
    logP = np.asarray(pe.post_burn_in_log_posteriors_un_normed_vec)
    samples = np.asarray(pe.post_burn_in_samples)
    small = samples.shape[0] <= _max_rows(pe)
    if small:
        mcmc_samples_array = np.hstack((logP, samples))                              # IP.py:1959
        np.savetxt(logs + 'mcmc_logP_and_parameter_samples.csv', mcmc_samples_array, delimiter=",")
        pickleAnObject(mcmc_samples_array, pkls + 'mcmc_logP_and_parameter_samples')
    else:
        print(f"Notification: {samples.shape[0]} post-burn-in rows exceed export_max_rows; the sample arrays are not written to file.")
    with open(logs + 'mcmc_log_file.txt', 'w') as out_file:                          # IP.py:1970-1987
        out_file.write("self.initial_point_parameters:" + str(UI.InputParameterInitialGuess) + "\n")
        out_file.write("MAP_logP:" + str(pe.map_logP) + "\n")
        out_file.write("self.map_parameter_set:" + str(pe.map_parameter_set) + "\n")
        out_file.write("self.map_index:" + str(pe.map_index) + "\n")
        out_file.write("self.mu_AP_parameter_set:" + str(pe.mu_AP_parameter_set) + "\n")
        out_file.write("self.stdap_parameter_set:" + str(pe.stdap_parameter_set) + "\n")
        out_file.write("self.info_gain:" + str(pe.info_gain) + "\n")
        out_file.write("evidence:" + str(pe.evidence) + "\n")
        cov = np.cov(samples.T) if small else getattr(pe, 'mcmc_chain_moments', {}).get('cov')
        out_file.write("posterior_cov_matrix:" + "\n" + str(cov) + "\n")
        relativeDifferenceArray = abs((np.asarray(pe.map_parameter_set) - np.asarray(pe.mu_AP_parameter_set)) / np.asarray(UI.std_prior))
        if np.array(relativeDifferenceArray > 0.10).any() == True:                   # noqa: E712
            out_file.write("Convergence Status: Warning - The difference between the MAP parameter set and mu_AP parameter set is greater than 10% of the prior standard deviations in at least one parameter. This means that either your posterior distribution is asymmetric or that it is not yet converged. This may mean that you need to increase your mcmc_length, increase or decrease your mcmc_relative_step_length, or change what is used for the model response.  There is no general method for knowing the right  value for mcmc_relative_step_length since it depends on the sharpness and smoothness of the response. See for example https://www.sciencedirect.com/science/article/pii/S0039602816300632")
        else:
            out_file.write("Convergence Status: The difference between the MAP parameter set and mu_AP parameter set is less than 10% of the prior standard deviations for each parameter. There is no general method for knowing when the correct solution has been obtained, but this small difference is one sign of a converged solution.")
    postBurnInStatistics = [pe.map_parameter_set, pe.mu_AP_parameter_set, pe.stdap_parameter_set, pe.evidence, pe.info_gain,
                            samples if small else None, logP if small else None]     # IP.py:1988
    if hasattr(pe, 'during_burn_in_samples') and np.shape(pe.during_burn_in_samples)[0] <= _max_rows(pe):
        burn = np.hstack((pe.during_burn_in_log_posteriors_un_normed_vec, pe.during_burn_in_samples))
        pickleAnObject(burn, pkls + 'mcmc_burn_in_logP_and_parameter_samples')
        np.savetxt(logs + 'mcmc_burn_in_logP_and_parameter_samples.csv', burn, delimiter=",")
    pickleAnObject(postBurnInStatistics, pkls + 'mcmc_post_burn_in_statistics')
    pickleAnObject(pe.map_logP, pkls + 'mcmc_map_logP')
    pickleAnObject(pe.map_parameter_set, pkls + 'mcmc_map_parameter_set')
    if pes.get('mcmc_continueSampling') == True:                                     # noqa: E712  IP.py:1995-1998
        pickleAnObject(UI.InputParameterInitialGuess, pkls + 'mcmc_continued_initial_point_parameters')
    else:
        pickleAnObject(UI.InputParameterInitialGuess, pkls + 'mcmc_initial_point_parameters')
    if hasattr(pe, 'mcmc_last_point_sampled'):
        pickleAnObject(pe.mcmc_last_point_sampled, pkls + 'mcmc_last_point_sampled')
    if UI.model.get('exportResponses') in ('auto', True):
        exportResponses(pe)


def exportMultistart(pe, centerPoint, bestResultSoFar, permutationsToSamples, initial_point=None, optimizer=False, mcmc=False):
    """The file output at the end of doListOfPermutationsSearch for searchType 'getLogP' (IP.py:1037-1072)."""
    if _rank_suffix() != 0:
        return
    UI = pe.UserInput
    logs, pkls = _dirs(UI)
    res = pe.sweep_result
    with open(logs + "multistart_log_file.txt", 'w') as out_file:
        exportResponses(pe, "map")
        start = bestResultSoFar[0] if initial_point is None else initial_point
        start = tuple(np.float64(v) for v in start)                                  # the reference keeps the list's tuple
        best = np.asarray(bestResultSoFar[0]) if (optimizer or mcmc) else start      # optimisers and samplers return arrays
        out_file.write("centerPoint: " + str(centerPoint) + "\n")
        out_file.write("highest_MAP_logP: " + str(np.array([pe.map_logP])) + "\n")
        out_file.write("highest_MAP_logP_parameter_set: " + str(best) + "\n")
        out_file.write("highest_MAP_initial_point_index: " + str(pe.map_index) + "\n")
        out_file.write("highest_MAP_initial_point_parameters: " + str(start) + "\n")
        if permutationsToSamples == True:                                            # noqa: E712
            caveat = ' (for the above initial point) ' if mcmc else ''                # IP.py:1046-1049
            out_file.write("self.mu_AP_parameter_set : " + caveat + str(bestResultSoFar[1]) + "\n")
            out_file.write("self.stdap_parameter_set : " + caveat + str(bestResultSoFar[2]) + "\n")
        if optimizer:                                                                # IP.py:1052-1053
            out_file.write("\n WARNING: It appears this run used a doOptimize with multi-start. In this case, the MAP_logP and map_parameter_set are the optimum from the run.  However, the mu_AP_parameter_set and stdap_parameter_set are not meaningful, since this was not an even weighted exploration of the posterior. \n")
    from . import distributed as D
    sw = getattr(pe, '_sweep', None)
    have_arrays = hasattr(pe, '_perm_array') or (sw is not None and sw.logp is not None and D.world()[1] == 1)
    if not have_arrays:
        print("Notification: logP of the sweep was not kept (multistart_store_logP) or is sharded over ranks; the point arrays are not written to file.")
    elif res.total <= _max_rows(pe):
        np.savetxt(logs + 'multistart_initial_points_parameters_values.csv', pe.listOfPermutations, delimiter=",")
        arr = pe.permutations_MAP_logP_and_parameters_values
        np.savetxt(logs + 'multistart_MAP_logP_and_parameters_values.csv', arr, delimiter=",")
        pickleAnObject(arr, pkls + 'multistart_MAP_logP_and_parameters_values')
    else:
        print(f"Notification: {res.total} multistart points exceed export_max_rows; the point arrays are not written to file.")


def exportResponses(pe, choiceOrResponsesData='', filenameString='', consolidated='auto'):
    """IP.py:2044-2119: simulated responses at the last / map / muap / initial / prior point next to the abscissa."""
    UI = pe.UserInput
    logs, _ = _dirs(UI)
    if isinstance(choiceOrResponsesData, str):
        choice = choiceOrResponsesData.lower()
        choice = {'initial': 'initial_guess', 'last': 'last_called'}.get(choice, choice)
    else:
        choice, responsesData = '', choiceOrResponsesData
    if isinstance(choiceOrResponsesData, str) and choice == '':
        for c in ("last", "map", "muap", "initial_guess", "prior"):
            exportResponses(pe, c)
        return None
    if choice == "last_called":
        if not hasattr(pe, "lastSimulatedResponses"):
            return None
        responsesData = pe.lastSimulatedResponses
    if choice == "initial_guess":
        responsesData = pe.getSimulatedResponses(UI.InputParameterInitialGuess)
    if choice == "prior":
        responsesData = pe.getSimulatedResponses(UI.mu_prior)
    if choice == "map":
        if not hasattr(pe, "map_parameter_set"):
            return None
        responsesData = pe.getSimulatedResponses(pe.map_parameter_set)
    if choice == "muap":
        if not hasattr(pe, "map_parameter_set") or getattr(pe, "mu_AP_parameter_set", None) is None:
            return None
        responsesData = pe.getSimulatedResponses(pe.mu_AP_parameter_set)
    if choice == "observed":
        responsesData = UI.responses_observed
    abscissa = _abscissa(UI)
    if consolidated == 'auto':
        consolidated = len(abscissa) == 1
    if consolidated:
        headerString = 'ResponseAbscissa,' + ",".join("Response" + str(i) for i in range(len(responsesData)))
        np.savetxt(logs + "ExportedResponsesData__" + filenameString + choice + "__ResponsesConsolidated" + ".csv",
                   np.vstack((abscissa, responsesData)).transpose(), delimiter=",", encoding=None, header=headerString, comments='')
    else:
        for responseIndex, responseData in enumerate(responsesData):
            headerString = 'ResponseAbscissa,' + "Response" + str(responseIndex)
            np.savetxt(logs + "ExportedResponsesData__" + filenameString + choice + "__Response" + str(responseIndex) + ".csv",
                       np.vstack((abscissa[responseIndex], responseData)).transpose(), delimiter=",", encoding=None,
                       header=headerString, comments='')


def _abscissa(UI):
    """UserInput.responses_abscissa as the reference's __init__ shapes it (IP.py:204-213): given, or linspace(0, n, n)."""
    given = UI.responses.get('responses_abscissa', [])
    if len(given) == 0:
        return [np.linspace(0, len(r), len(r)) for r in UI.responses_observed]
    a = np.asarray(given, dtype=object if isinstance(given[0], (list, tuple, np.ndarray)) and len({len(g) for g in given}) > 1 else float)
    return [np.asarray(a, dtype=float)] if a.ndim == 1 else [np.asarray(r, dtype=float) for r in a]


def reload_samples(pe, sampling_type='', filepath=''):
    """IP.py:441-480: rebuild post_burn_in_* from the CSV files and recompute the statistics."""
    if sampling_type not in ('multistart', 'mcmc'):
        print("ERROR: reload_samples requires specifying either 'multistart' or 'mcmc' as the first argument"); sys.exit()
    UI = pe.UserInput
    pes = UI.parameter_estimation_settings
    if filepath == '':
        filepath = UI.directories['logs_and_csvs']
    if sampling_type == 'multistart':
        arr = np.genfromtxt(os.path.join(filepath, "multistart_MAP_logP_and_parameters_values.csv"), delimiter=",")
        arr = np.atleast_2d(arr)
        pe._perm_array = arr
        coef = pes['multistart_permutationsToSamples_threshold_filter_coefficient']
        coef = 2.0 if str(coef).lower() == 'auto' else float(coef)
        with np.errstate(invalid='ignore'):
            best = np.nanmax(arr[:, 0])
        thr = best + np.log(10 ** (-1 * coef))                                       # convertPermutationsToSamples, IP.py:3463-3491
        with np.errstate(over='ignore', invalid='ignore'):
            w = np.where(arr[:, 0] > thr, np.trunc(np.exp(arr[:, 0] - thr)), 0).astype(np.int64)
        expanded = np.repeat(arr, w, axis=0)
        pe.post_burn_in_samples = expanded[:, 1:]
        pe.post_burn_in_log_posteriors_un_normed_vec = expanded[:, 0:1]
        pe.post_burn_in_log_priors_vec = None
        pes['mcmc_threshold_filter_samples'] = False                                 # IP.py:458
    else:
        arr = np.atleast_2d(np.genfromtxt(os.path.join(filepath, "mcmc_logP_and_parameter_samples.csv"), delimiter=","))
        pe.post_burn_in_samples = arr[:, 1:]
        pe.post_burn_in_log_posteriors_un_normed_vec = arr[:, 0:1]
        pe.post_burn_in_log_priors_vec = None
        try:
            pe.during_burn_in_samples = np.atleast_2d(np.genfromtxt(os.path.join(filepath, "mcmc_burn_in_logP_and_parameter_samples.csv"),
                                                                    delimiter=","))[:, 1:]
        except Exception:
            pass
    pe._dev_post = None
    return pe.calculatePostBurnInStatistics(calculate_post_burn_in_log_priors_vec=(sampling_type == 'mcmc'))
