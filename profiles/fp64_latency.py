"""Dependent-DFMA probe: one warp per scheduler, 1..16 independent chains per thread (run on the B200 box)."""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peuqse_b200 import _lib
L = _lib.load()
scratch = torch.zeros(8, dtype=torch.float64, device="cuda")
sms = torch.cuda.get_device_properties(0).multi_processor_count
out = {}
for lg in range(5):
    r = C.c_double()
    _lib.check(L.peuqse_bench_fp64(10 + lg, C.byref(r), C.c_void_p(scratch.data_ptr()), None))
    out[f"chains_{1 << lg}_warp_dfma_per_s_per_scheduler"] = r.value / 32 / (sms * 4)
sc2 = torch.cat([torch.full((12,), 1e-3, dtype=torch.float64), torch.full((4,), 0.25, dtype=torch.float64), torch.zeros(8, dtype=torch.float64)]).cuda()
r = C.c_double()
_lib.check(L.peuqse_bench_fp64(2, C.byref(r), C.c_void_p(sc2.data_ptr()), None))
out["mode2_horner_three_register_operands_instr_per_s"] = r.value
for m in (0, 1):
    r = C.c_double()
    _lib.check(L.peuqse_bench_fp64(m, C.byref(r), C.c_void_p(scratch.data_ptr()), None))
    out[f"mode{m}_full_occupancy_instr_per_s"] = r.value
for m, name in ((20, "4_int"), (21, "8_int"), (22, "16_int"), (23, "2_lds"), (24, "4_lds"), (25, "5_int_2_lds")):
    r = C.c_double()
    _lib.check(L.peuqse_bench_fp64(m, C.byref(r), C.c_void_p(scratch.data_ptr()), None))
    out[f"dfma_per_s_with_{name}_per_8_dfma"] = r.value
print(json.dumps(out, indent=1))
