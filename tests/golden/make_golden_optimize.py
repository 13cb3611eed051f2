"""Generates tests/golden/optimize.npz: doOptimizeNegLogP / doOptimizeLogP (PEUQSE/InverseProblem.py:1661-1694) and a
multistart with searchType 'doOptimizeNegLogP' (IP.py:912-921) of the UNMODIFIED reference on Example00.
Run in the build container only:  python tests/golden/make_golden_optimize.py"""
import contextlib
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as G  # noqa: E402


def main():
    out = {}
    for tag, method in (('nm', 'Nelder-Mead'), ('bfgs', None), ('powell', 'Powell')):
        PE = G.make_pe(G.ui_example00())
        with contextlib.redirect_stdout(io.StringIO()):
            x, f = PE.doOptimizeNegLogP(method=method, printOptimum=False, verbose=False)
        out[f'{tag}_x'], out[f'{tag}_negLogP'] = np.array(x), np.array(f)
    PE = G.make_pe(G.ui_example00())
    with contextlib.redirect_stdout(io.StringIO()):
        x, f = PE.doOptimizeLogP(method='Nelder-Mead', printOptimum=False, verbose=False, maxiter=30)
    out['nm30_x'], out['nm30_logP'] = np.array(x), np.array(f)
    UI = G.ui_example00()
    pes = UI.parameter_estimation_settings
    pes.update({'multistart_initialPointsDistributionType': 'grid', 'multistart_gridsearchSamplingInterval': [50, 100],
                'multistart_gridsearchSamplingRadii': [1, 1], 'multistart_searchType': 'doOptimizeNegLogP',
                'multistart_passThroughArgs': {'method': 'Nelder-Mead', 'printOptimum': False, 'verbose': False}})
    UI.model['exportResponses'] = False
    PE = G.make_pe(UI)
    with contextlib.redirect_stdout(io.StringIO()):
        best = PE.doMultiStart()
    out['ms_best_x'] = np.array(best[0])
    out['ms_map_logP'] = np.array(PE.map_logP)
    out['ms_perm_array'] = np.array(PE.permutations_MAP_logP_and_parameters_values)
    out['ms_list'] = np.array(PE.listOfPermutations)
    out['ms_log'] = np.frombuffer(open('logs_and_csvs/multistart_log_file.txt', 'rb').read(), dtype=np.uint8)
    path = os.path.join(G.HERE, "optimize.npz")
    np.savez_compressed(path, **out)
    print("optimize ->", path)
    for k, v in out.items():
        print("  ", k, v.shape, v if v.size < 8 else '')
    print(bytes(out['ms_log']).decode())


if __name__ == "__main__":
    main()
    os.chdir(G.REPO)
