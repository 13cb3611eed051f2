"""Lock-step Metropolis-Hastings (csrc/mh_steps.cu): seconds per doMetropolisHastings() call with a torch-callable simulator,
graph-replayed against the eager loop, and the persistent kernel on the same function for scale.  python profiles/mh_lockstep_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import peuqse_b200 as PEUQSE

X00, Y00, S00 = [600, 1100, 1400], [360500, 580500, 1620500], [200000, 300000, 200000]


def make(model, **pes):
    UI = PEUQSE.UserInput.fresh()
    UI.responses['responses_observed'] = Y00
    UI.responses['responses_observed_uncertainties'] = S00
    UI.model['InputParameterPriorValues'] = [200, 500]
    UI.model['InputParametersPriorValuesUncertainties'] = [100, 200]
    UI.model['simulateByInputParametersOnlyFunction'] = model
    UI.parameter_estimation_settings.update({'mcmc_length': 10000, 'mcmc_relative_step_length': 0.1, 'mcmc_random_seed': 0,
                                             'mcmc_exportLog': False, 'exportLog': False})
    UI.parameter_estimation_settings.update(pes)
    return PEUQSE.parameter_estimation(UI)


def timed(PE, reps=3):
    ts = []
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        PE.doMetropolisHastings()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts[1:])


X = torch.tensor(X00, dtype=torch.float64, device='cuda')
f = lambda t: (X[None, :] - t[:, :1]) ** 2 + t[:, 1:2]
for chains in (1, 1024):
    for name, PE in (("persistent kernel, built-in model", make(PEUQSE.QuadraticModel(X00), mcmc_nchains=chains)),
                     ("lock-step, torch simulator, graph replay", make(PEUQSE.TorchSimulator(f), mcmc_nchains=chains)),
                     ("lock-step, torch simulator, eager loop", make(PEUQSE.TorchSimulator(f), mcmc_nchains=chains, mcmc_lockstep_graph=False))):
        s = timed(PE)
        print(f"{chains:5d} chains x 10000 steps  {name:45s} {s * 1e3:9.1f} ms  {chains * 9999 / s:.3e} chain-steps/s", flush=True)
