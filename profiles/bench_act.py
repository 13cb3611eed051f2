"""Integrated autocorrelation time on the device (peuqse_b200.diagnostics.autocorr_times: csrc/diag.cu k_act_concat + cuFFT) on
C3-shaped chains: R rows x C chains x 6 parameters, the reference's 20 growing windows.  The layout kernel is timed alone
with CUDA events (HBM-bound: 8 B read + 8 B written per sample).  CPU: the oracle's numpy restatement of
zeus.autocorr.AutoCorrTime on a bounded sample (the last window of a C/64-chain subset, one parameter).
usage: python profiles/bench_act.py [R=20000] [C=4096]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peuqse_b200.diagnostics import autocorr_times
from peuqse_b200 import _lib

R = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
k = 6
gen = torch.Generator(device="cuda").manual_seed(3)
x = torch.randn((R, C, k), dtype=torch.float64, device="cuda", generator=gen)
rho = torch.tensor([0.0, 0.5, 0.8, 0.9, 0.95, 0.3], dtype=torch.float64, device="cuda")
# AR(1) along the rows by a log-depth scan: x_t = rho x_{t-1} + e_t
s = 1
while s < R:
    x[s:] += (rho ** s) * x[:-s].clone()
    s *= 2
taus, windows, launches = autocorr_times(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record()
taus, windows, launches = autocorr_times(x)
b.record(); torch.cuda.synchronize()
wall = time.perf_counter() - t0
ms = a.elapsed_time(b)
# the layout kernel alone, last window
L = _lib.load()
n2 = 2
while n2 < 2 * R * C:
    n2 <<= 1
mean = x.mean(dim=(0, 1)).contiguous()
series = torch.empty((k, n2), dtype=torch.float64, device="cuda")
for _ in range(2):
    _lib.check(L.peuqse_act_concat(_lib.ptr(x), R, C, k, _lib.ptr(mean), _lib.ptr(series), n2, _lib.current_stream()))
a.record()
for _ in range(5):
    _lib.check(L.peuqse_act_concat(_lib.ptr(x), R, C, k, _lib.ptr(mean), _lib.ptr(series), n2, _lib.current_stream()))
b.record(); torch.cuda.synchronize()
ms_concat = a.elapsed_time(b) / 5
del series
out = dict(rows=R, chains=C, k=k, windows=[int(w) for w in windows], ms_per_call=ms, wall_ms=wall * 1e3, launches_own=launches,
           taus_last_window=[float(v) for v in taus[-1].cpu()], expected_ar1=[float(v) for v in ((1 + rho) / (1 - rho)).cpu()],
           concat_kernel=dict(ms=ms_concat, algorithmic_GB=2 * R * C * k * 8 / 1e9, padding_GB=k * (n2 - R * C) * 8 / 1e9,
                              algorithmic_GBps=2 * R * C * k * 8 / (ms_concat * 1e-3) / 1e9))
from oracle import peuqse_oracle as orc
Cs = max(1, C // 64)
xs = x[:, :Cs, 3:4].cpu().numpy()
t0 = time.perf_counter()
t_cpu = orc.autocorr_time(xs)
dt = time.perf_counter() - t0
t_gpu = autocorr_times(x[:, :Cs, 3:4].contiguous(), windows=[R])[0].cpu().numpy()[0]
out['cpu'] = dict(kind='port', cores=1, sample='last window only, %d of %d chains, one parameter' % (Cs, C), seconds=dt,
                  samples_per_s=R * Cs / dt, tau=float(t_cpu[0]), tau_device_same_input=float(t_gpu[0]))
out['gpu_samples_per_s_all_windows'] = float(sum(int(w) for w in windows)) * C * k / (ms * 1e-3)
print(json.dumps(out))
