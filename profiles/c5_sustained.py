"""C5 sweep (1e10 points, no-store) four times back to back with the SM clock and power sampled: is the first run's rate sustained?
usage: python profiles/c5_sustained.py [points]"""
import json, os, subprocess, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE
from peuqse_b200.sweep import Sweep, make_spec
from oracle import peuqse_oracle as orc
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000_000
k, n = 10, 64
Phi = orc.cos_basis(n, k)
truth = np.random.default_rng(5).standard_normal(k) * 0.5
y = Phi @ truth + 0.1 * np.random.default_rng(6).standard_normal(n)
UI = PEUQSE.UserInput.fresh()
UI.responses['responses_observed'] = list(y)
UI.responses['responses_observed_uncertainties'] = list(np.full(n, 0.1))
UI.model['InputParameterPriorValues'] = [0.0] * k
UI.model['InputParametersPriorValuesUncertainties'] = [1.0] * k
UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.LinearModel(Phi)
PE = PEUQSE.parameter_estimation(UI)
spec = make_spec("uniform", k, np.zeros(k), spread=0.866, std_prior=np.ones(k), seed=0)
sw = Sweep(PE.device_problem, spec, N, store_logp=False)
lr = float(np.log(1e-2))
Sweep(PE.device_problem, spec, 1 << 24, store_logp=False).step(lr, True)
torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                      "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
time.sleep(0.5)
out = []
for r in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); sw.step(lr, True); b.record(); torch.cuda.synchronize()
    out.append(N / (a.elapsed_time(b) * 1e-3))
p.terminate()
lines = p.communicate()[0].strip().splitlines()
print(json.dumps({"evals_per_s": out}))
print("\n".join(lines[::3]))
