"""fp64 DFMA and DMMA (mma.sync.m8n8k4.f64) ceilings of the device, measured by the library's register-only probes
(peuqse_bench_fp64 modes 0, 3, 4; modes 5-8: 16 accumulators per warp from 4 x 4 operand fragments, the register
pattern of csrc/fullcov.cu, with 8 / 16 / 32 / 64 warps per SM): the denominators of the fp64 / DMMA rooflines.  usage: python profiles/dmma_peak.py"""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peuqse_b200 import _lib
L = _lib.load()
scratch = torch.zeros(64, dtype=torch.float64, device="cuda")
out = {}
for mode, name in ((0, "fp64_dfma_instr_per_s"), (3, "dmma_m8n8k4_flop_per_s_8_accumulators"), (4, "dmma_m8n8k4_flop_per_s_4_accumulators"),
                   (5, "dmma_tile_pattern_1_cta_per_sm"), (6, "dmma_tile_pattern_2_cta_per_sm"), (7, "dmma_tile_pattern_4_cta_per_sm"),
                   (8, "dmma_tile_pattern_8_cta_per_sm")):
    best = 0.0
    for _ in range(3):
        r = C.c_double()
        _lib.check(L.peuqse_bench_fp64(mode, C.byref(r), _lib.ptr(scratch), _lib.current_stream()))
        best = max(best, r.value)
    out[name] = best
out["fp64_dfma_flop_per_s"] = 2 * out["fp64_dfma_instr_per_s"]
print(json.dumps(out))
