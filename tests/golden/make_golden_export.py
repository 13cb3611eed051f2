"""Generates tests/golden/export_formats.npz: the files the UNMODIFIED reference writes after doMetropolisHastings
(exportPostBurnInStatistics, PEUQSE/InverseProblem.py:1956-2001) and after a grid doMultiStart (IP.py:1037-1072), as raw
bytes, together with the object state they were written from.  tests/test_host_cpu.py feeds that state to
peuqse_b200/export.py and compares the files byte for byte (text / CSV) and object for object (pickles).

Run in the build container only:  python tests/golden/make_golden_export.py
"""
import contextlib
import io
import os
import pickle
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as G  # noqa: E402


def grab(out, tag):
    for sub in ('logs_and_csvs', 'pickles'):
        for f in sorted(os.listdir(sub)):
            path = os.path.join(sub, f)
            out[f'{tag}__{sub}__{f}'] = np.frombuffer(open(path, 'rb').read(), dtype=np.uint8)
            os.remove(path)


def main():
    out = {}
    # ---- Metropolis-Hastings
    UI = G.ui_example00()
    pes = UI.parameter_estimation_settings
    pes.update({'mcmc_length': 60, 'mcmc_burn_in': 10, 'mcmc_relative_step_length': 0.1, 'mcmc_random_seed': 0,
                'mcmc_threshold_filter_samples': True, 'mcmc_info_gain_returned': 'log_ratio'})
    UI.model['exportResponses'] = False
    PE = G.make_pe(UI)
    with contextlib.redirect_stdout(io.StringIO()):
        PE.doMetropolisHastings()
    grab(out, 'mh')
    for name in ('post_burn_in_samples', 'post_burn_in_log_posteriors_un_normed_vec', 'during_burn_in_samples',
                 'during_burn_in_log_posteriors_un_normed_vec', 'map_logP', 'map_parameter_set', 'mu_AP_parameter_set',
                 'stdap_parameter_set', 'mcmc_last_point_sampled'):
        out['mh_state__' + name] = np.array(getattr(PE, name))
    out['mh_state__map_index'] = np.array(PE.map_index)
    out['mh_state__info_gain'] = np.array(PE.info_gain)
    out['mh_state__evidence'] = np.array(PE.evidence)
    out['mh_state__initial_guess'] = np.array(UI.InputParameterInitialGuess)
    # ---- grid multistart with permutationsToSamples
    UI = G.ui_example00()
    pes = UI.parameter_estimation_settings
    pes.update({'multistart_initialPointsDistributionType': 'grid', 'multistart_gridsearchSamplingInterval': [50, 100],
                'multistart_gridsearchSamplingRadii': [2, 2], 'multistart_searchType': 'getLogP',
                'multistart_permutationsToSamples': True})
    UI.model['exportResponses'] = False
    PE = G.make_pe(UI)
    with contextlib.redirect_stdout(io.StringIO()):
        best = PE.doMultiStart()
    grab(out, 'ms')
    out['ms_state__perm_array'] = np.array(PE.permutations_MAP_logP_and_parameters_values)
    out['ms_state__list'] = np.array(PE.listOfPermutations)
    out['ms_state__map_logP'] = np.array(PE.map_logP)
    out['ms_state__map_parameter_set'] = np.array(best[0])
    out['ms_state__mu_AP'] = np.array(PE.mu_AP_parameter_set)
    out['ms_state__stdap'] = np.array(PE.stdap_parameter_set)
    out['ms_state__map_sim'] = np.array(PE.lastSimulatedResponses)
    path = os.path.join(G.HERE, "export_formats.npz")
    np.savez_compressed(path, **out)
    print("export_formats ->", path)
    for k, v in out.items():
        print("  ", k, v.shape)


if __name__ == "__main__":
    main()
    os.chdir(G.REPO)
