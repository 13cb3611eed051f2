"""The batched log-posterior of the C4 problem (two 4096 x 4096 dense-covariance responses, k = 12) at batch sizes that hit every
kernel shape of csrc/fullcov.cu: 2048 / 1024 rows (k_qf_tile<false>), 200 (k_qf_tile<true> + k_qf_finish_tc), 64 / 16 / 2
(k_qf_skinny<8 / 2 / 1>).  Prints the CUDA-event time per call of peuqse_logp_batch (20 calls back to back, no host work in
between) and is the command of the ncu captures.  usage: python profiles/fullcov_once.py [cholesky] [reps]"""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE
from peuqse_b200 import _lib
from oracle import peuqse_oracle as orc
wh = 'cholesky' if 'cholesky' in sys.argv[1:] else 'eigh'
reps = next((int(a) for a in sys.argv[1:] if a.isdigit()), 20)
k, n = 12, 4096
x = np.linspace(0, 1, n)
Phi = orc.cos_basis(2 * n, k)
Sig = 0.1 ** 2 * np.exp(-(x[:, None] - x[None, :]) ** 2 / (2 * 0.02 ** 2)) + 0.05 ** 2 * np.eye(n)
truth = np.random.default_rng(1).standard_normal(k)
y = Phi @ truth + 0.1 * np.random.default_rng(2).standard_normal(2 * n)
UI = PEUQSE.UserInput.fresh()
UI.responses['responses_observed'] = [list(y[:n]), list(y[n:])]
UI.responses['responses_observed_uncertainties'] = [Sig, Sig]
UI.responses['responses_observed_covmat_whitening'] = wh
UI.model['InputParameterPriorValues'] = [0.0] * k
UI.model['InputParametersPriorValuesUncertainties'] = [2.0] * k
UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.LinearModel(Phi)
UI.responses['responses_observed_max_covmat_size'] = 4096
PE = PEUQSE.parameter_estimation(UI)
L = _lib.load()
out = {"whitening": wh}
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flops = 2 * (2.0 * n * n if wh == 'eigh' else 1.0 * n * (n + 128))     # executed per row, both responses
for B in [int(v) for v in os.environ.get('FULLCOV_BS', '2048 1024 200 64 16 2').split()]:   # FULLCOV_BS="1024 2": the ncu captures
    th = torch.from_numpy(truth + 0.05 * np.random.default_rng(3).standard_normal((B, k))).cuda()
    lp = torch.empty(B, dtype=torch.float64, device="cuda")
    st = _lib.current_stream()
    call = lambda: _lib.check(L.peuqse_logp_batch(PE.device_problem.handle, _lib.ptr(th), B, _lib.ptr(lp), None, None, st))
    call(); call()
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        call()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    out["B%d" % B] = dict(ms=round(ms, 4), tflops=round(B * flops / (ms * 1e-3) / 1e12, 2), U_read_TBps=round(2 * n * n * 8 * (1 if wh == 'eigh' else 0.52) / (ms * 1e-3) / 1e12, 2))
print(json.dumps(out))
