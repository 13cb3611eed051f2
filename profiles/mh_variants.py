import os, sys, json
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench, peuqse_b200 as PEUQSE
from peuqse_b200.mcmc import mh_launch, step_cholesky
steps = 2000
PE, y, sig = bench.c3_problem(PEUQSE, steps, 4096)
th0 = np.tile(np.asarray(bench.C3["truth"]), (4096, 1))
chol = step_cholesky(PE.Q_covmat, bench.C3["step"])
run = lambda: mh_launch(PE, th0, steps, 200, 0, chol, store_samples=False)
run(); run()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(3): out = run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
st = out['stats'].cpu().numpy()
print(os.environ.get("PEUQSE_MH_VARIANT"), "chain-steps/s %.4e" % (4096 * (steps - 1) / (ms * 1e-3)), "checksum", repr(float(st.sum())))
