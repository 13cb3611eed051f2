"""Where the host time of PE.doMultiStart() on the C2 grid goes: cProfile over 300 calls after the graph is captured (the device
step is ~0.30 ms, the call ~0.35 ms).  usage: python profiles/c2_e2e_profile.py"""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import peuqse_b200 as PEUQSE

radii, total = bench.grid_for(1)
UI = PEUQSE.UserInput.fresh()
UI.responses['responses_abscissa'] = bench.X00
UI.responses['responses_observed'] = bench.Y00
UI.responses['responses_observed_uncertainties'] = bench.S00
UI.model['InputParameterPriorValues'] = bench.MU
UI.model['InputParametersPriorValuesUncertainties'] = bench.SD
UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.QuadraticModel(bench.X00)
pes = UI.parameter_estimation_settings
pes['multistart_initialPointsDistributionType'] = 'grid'
pes['multistart_gridsearchSamplingInterval'] = bench.INTERVAL
pes['multistart_gridsearchSamplingRadii'] = radii
pes['multistart_searchType'] = 'getLogP'
pes['exportLog'] = False
PE = PEUQSE.parameter_estimation(UI)
for _ in range(5):
    PE.doMultiStart()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    PE.doMultiStart()
torch.cuda.synchronize()
print("ms per call", (time.perf_counter() - t0) / 300 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    PE.doMultiStart()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
