"""How long do the tiny kernels of the C2 step take back to back (warm) vs inside the step?"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE
from peuqse_b200.sweep import Sweep, make_spec

def timeit(fn, n=50, warm=3):
    for _ in range(warm): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

X00, Y00, S00 = [600, 1100, 1400], [360500, 580500, 1620500], [200000, 300000, 200000]
UI = PEUQSE.UserInput.fresh()
UI.responses['responses_observed'] = Y00; UI.responses['responses_observed_uncertainties'] = S00
UI.model['InputParameterPriorValues'] = [200, 500]; UI.model['InputParametersPriorValuesUncertainties'] = [100, 200]
UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.QuadraticModel(X00)
UI.parameter_estimation_settings['mcmc_length'] = 10
PE = PEUQSE.parameter_estimation(UI)
spec = make_spec("grid", 2, [200., 500.], interval=[0.08, 0.16], radii=[5000, 5000])
sw = Sweep(PE.device_problem, spec, 100_020_001)
lr = float(np.log(1e-2))
sw.step(lr); torch.cuda.synchronize()
out = {}
out["step_us"] = timeit(lambda: sw.step(lr))
out["pass1_us"] = timeit(sw.pass1)
out["pass2_us"] = timeit(lambda: sw.pass2(lr))
out["pass3_us(edges+bin)"] = timeit(lambda: sw.pass3(lr))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    sw.step(lr)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        sw.step(lr)
torch.cuda.synchronize()
out["step_graph_us"] = timeit(g.replay)
print(json.dumps(out, indent=1))
