"""Single-GPU timing of ONE rank's shard of the N-GPU bench grid, without collectives: separates what the shard's
shape costs from what the exchanges and rank skew cost in the multi-GPU bench."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import peuqse_b200 as PEUQSE
from peuqse_b200.sweep import Sweep, make_spec
import bench

def ev_time(fn, n=100, warm=5):
    for _ in range(warm): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

UI = PEUQSE.UserInput.fresh()
UI.responses['responses_observed'] = bench.Y00
UI.responses['responses_observed_uncertainties'] = bench.S00
UI.model['InputParameterPriorValues'] = bench.MU
UI.model['InputParametersPriorValuesUncertainties'] = bench.SD
UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.QuadraticModel(bench.X00)
PE = PEUQSE.parameter_estimation(UI)
lr = float(np.log(1e-2))
out = {}
for n in (1, 2, 8):
    radii, total = bench.grid_for(n)
    spec = make_spec("grid", 2, [float(bench.MU[0]), float(bench.MU[1])], interval=bench.INTERVAL, radii=radii)
    for r in sorted({0, n // 2, n - 1}):
        sw = Sweep(PE.device_problem, spec, total, store_logp=True, nbins=10, shard=(r, n))
        sw.world = 1
        out[f"n{n}_rank{r}_step_ms"] = ev_time(lambda: sw.step(lr))
        out[f"n{n}_rank{r}_pass1_ms"] = ev_time(sw.pass1)
        del sw
        torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
