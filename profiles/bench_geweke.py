"""Throughput of the device Geweke diagnostic (csrc/diag.cu) on C3-shaped chains: R rows x C chains x 6 parameters.
usage: python profiles/bench_geweke.py [R=20000] [C=4096]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peuqse_b200.diagnostics import geweke_z_scores

R = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
k = 6
gen = torch.Generator(device="cuda").manual_seed(3)
x = torch.randn((R, C, k), dtype=torch.float64, device="cuda", generator=gen).cumsum(dim=0) * 0.05 + torch.randn((R, C, k), dtype=torch.float64, device="cuda", generator=gen)
for _ in range(2):
    z = geweke_z_scores(x)[0]
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record()
for _ in range(3):
    z = geweke_z_scores(x)[0]
b.record(); torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 3
ms = a.elapsed_time(b) / 3
out = dict(rows=R, chains=C, k=k, ms_per_call=ms, wall_ms=wall * 1e3, z_scores=int(z.numel()), sample_GB=R * C * k * 8 / 1e9,
           algorithmic_GBps=R * C * k * 8 / (ms * 1e-3) / 1e9)
from oracle import peuqse_oracle as orc
import warnings
xc = x[:, 0, 0].cpu().numpy()
t0 = time.perf_counter()
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    for n in np.linspace(0, R, 21).astype(int)[1:]:
        orc.geweke_diagnostic(xc[:n])
dt = time.perf_counter() - t0
out['cpu'] = dict(kind='port', cores=1, sample='one chain, one parameter, 20 windows', s_per_chain_parameter=dt, chain_parameters_per_s=1 / dt)
out['gpu_chain_parameters_per_s'] = C * k / (ms * 1e-3)
print(json.dumps(out))
