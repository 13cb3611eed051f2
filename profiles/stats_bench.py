"""Throughput of the device post-burn-in statistics (csrc/stats.cu) on a C3-shaped sample set: N rows x 6 parameters."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from peuqse_b200 import stats as ST, _lib

N, k = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000, 6
gen = torch.Generator(device="cuda").manual_seed(5)
th = torch.randn((N, k), dtype=torch.float64, device="cuda", generator=gen)
logp = -0.5 * (th * th).sum(dim=1) - 5.0
logpr = -0.02 * (th * th).sum(dim=1)
prior = torch.as_tensor(np.random.default_rng(0).standard_normal((100000, k)) * 3).cuda()   # samples_of_prior, uploaded once
out = {}
for name, mode, pri in (("auto_logratio", _lib.FILTER_AUTO, None), ("auto_kl", _lib.FILTER_AUTO, prior), ("none", _lib.FILTER_NONE, None)):
    for _ in range(2):
        R = ST.device_statistics(th, logp, logpr, mode, 2.0, prior_samples=pri)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(3):
        R = ST.device_statistics(th, logp, logpr, mode, 2.0, prior_samples=pri)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    # algorithmic bytes: begin reads logP once or twice (8 B/row each), finish reads samples + logP + logPrior twice
    nbytes = N * (8 * (2 if mode == _lib.FILTER_AUTO else 1) + 2 * 8 * (k + 2))
    out[name] = dict(ms=ms, rows_per_s=N / (ms * 1e-3), algorithmic_GBps=nbytes / (ms * 1e-3) / 1e9, n_kept=R.n_kept)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
R = ST.device_statistics(th, logp, logpr, _lib.FILTER_AUTO, 2.0)
torch.cuda.synchronize(); a.record()
ks, kl, kp = R.kept_rows()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
out["compact"] = dict(ms=ms, algorithmic_GBps=(N * 8 + 2 * R.n_kept * 8 * (k + 2)) / (ms * 1e-3) / 1e9)
print(json.dumps(out, indent=1))
