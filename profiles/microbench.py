"""Microbenchmarks that give the roofline denominators / diagnose the C2 kernels (run on the B200 box)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE
from peuqse_b200 import _lib
from peuqse_b200.sweep import Sweep, make_spec


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


out = {}
N = 100_020_001
x = torch.empty(N, dtype=torch.float64, device="cuda")
y = torch.empty(N, dtype=torch.float64, device="cuda")
ms = timeit(lambda: x.fill_(1.0))
out["fill_800MB_write_only_GBs"] = 8 * N / ms / 1e6
ms = timeit(lambda: y.copy_(x))
out["copy_800MB_read_plus_write_GBs"] = 16 * N / ms / 1e6
ms = timeit(lambda: x.sum())
out["sum_800MB_read_only_GBs"] = 8 * N / ms / 1e6
L = _lib.load()
scratch = torch.zeros(8, dtype=torch.float64, device="cuda")
for mode, name in ((0, "fp64_dfma_instr_per_s"), (1, "fp64_dmul_dadd_instr_per_s"), (3, "fp64_dmma_m8n8k4_flop_per_s_8acc"),
                   (4, "fp64_dmma_m8n8k4_flop_per_s_4acc")):
    r = C.c_double()
    _lib.check(L.peuqse_bench_fp64(mode, C.byref(r), _lib.ptr(scratch), _lib.current_stream()))
    out[name] = r.value

X00, Y00, S00 = [600, 1100, 1400], [360500, 580500, 1620500], [200000, 300000, 200000]
UI = PEUQSE.UserInput.fresh()
UI.responses['responses_observed'] = Y00
UI.responses['responses_observed_uncertainties'] = S00
UI.model['InputParameterPriorValues'] = [200, 500]
UI.model['InputParametersPriorValuesUncertainties'] = [100, 200]
UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.QuadraticModel(X00)
UI.parameter_estimation_settings['mcmc_length'] = 10
PE = PEUQSE.parameter_estimation(UI)
spec = make_spec("grid", 2, [200., 500.], interval=[0.08, 0.16], radii=[5000, 5000])
for store in (True, False):
    sw = Sweep(PE.device_problem, spec, N, store_logp=store)
    out[f"pass1_ms_store_{store}"] = timeit(sw.pass1)
    del sw
print(json.dumps(out, indent=1))
