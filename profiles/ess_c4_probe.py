"""Where a C4 ensemble-slice iteration spends its time: the batched log-posterior at several batch sizes, then two
iterations of the round-synchronous driver under torch.profiler (top kernels / host calls).  usage: python profiles/ess_c4_probe.py"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE
from peuqse_b200.mcmc import ess_launch
from oracle import peuqse_oracle as orc
k, n, W = 12, 4096, 2048
x = np.linspace(0, 1, n)
Phi = orc.cos_basis(2 * n, k)
Sig = 0.1 ** 2 * np.exp(-(x[:, None] - x[None, :]) ** 2 / (2 * 0.02 ** 2)) + 0.05 ** 2 * np.eye(n)
truth = np.random.default_rng(1).standard_normal(k)
y = Phi @ truth + 0.1 * np.random.default_rng(2).standard_normal(2 * n)
UI = PEUQSE.UserInput.fresh()
UI.responses['responses_observed'] = [list(y[:n]), list(y[n:])]
UI.responses['responses_observed_uncertainties'] = [Sig, Sig]
UI.model['InputParameterPriorValues'] = [0.0] * k
UI.model['InputParametersPriorValuesUncertainties'] = [2.0] * k
UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.LinearModel(Phi)
UI.responses['responses_observed_max_covmat_size'] = 4096
PE = PEUQSE.parameter_estimation(UI)
out = {}
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for B in (2048, 1536, 1024, 512, 256, 128, 64, 16, 2):
    th = torch.from_numpy(truth + 0.05 * np.random.default_rng(3).standard_normal((B, k))).cuda()
    for _ in range(2):
        PE.getLogP_batch(th)
    torch.cuda.synchronize()
    a.record()
    for _ in range(5):
        PE.getLogP_batch(th)
    b.record(); torch.cuda.synchronize()
    out["logp_ms_B%d" % B] = a.elapsed_time(b) / 5
print(json.dumps(out))
start = 0.866 * 2.0 * np.random.default_rng(3).standard_normal((W, k)) * 0.25
ess_launch(PE, start, 1, 7, store_chain=False)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    ess_launch(PE, start, 2, 7, store_chain=False)
    torch.cuda.synchronize()
print("wall_s_2_iterations_under_profiler", time.perf_counter() - t0)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=10, max_name_column_width=60))
