"""C3 Metropolis-Hastings kernel: chain-steps/s over the number of chains for the launcher's shapes (PEUQSE_MH_SHAPE, see launch_mh
in csrc/mh.cu; 0 = the launcher's own choice).  usage: python profiles/c3_shapes_sweep.py [steps]"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE
from peuqse_b200.mcmc import mh_launch, step_cholesky
from oracle import peuqse_oracle as orc

n = 1024
kBT, c1, c2 = orc.tpd2site_constants(n)
truth = np.array([41.5, 45, 13, 13, 0.1, 0.1])
sim_t = orc.sim_tpd2site(kBT, c1, c2)[0](truth)
sig = 0.02 * np.max(sim_t)
y = sim_t + sig * np.random.default_rng(0).standard_normal(n)
UI = PEUQSE.UserInput.fresh()
UI.responses['responses_observed_max_covmat_size'] = 5000
UI.responses['responses_observed'] = list(y); UI.responses['responses_observed_uncertainties'] = list(np.full(n, sig))
UI.model['InputParameterPriorValues'] = [41.5, 41.5, 13.0, 13.0, 0.1, 0.1]
UI.model['InputParametersPriorValuesUncertainties'] = [20, 20, 13, 13, 0.1, 0.1]
UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.Tpd2SiteModel.linear_ramp(n)
PE = PEUQSE.parameter_estimation(UI)
Lm = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
chol = step_cholesky(PE.Q_covmat, 1e-5)
out = {}
for C in (256, 1024, 1776, 2048, 3552, 4096, 8192, 16384):
    th0 = np.tile(truth, (C, 1))
    for shape in os.environ.get('SHAPES', '0 4 5 7 8').split():
        os.environ['PEUQSE_MH_SHAPE'] = shape
        mh_launch(PE, th0, 50, 20, 0, chol, store_samples=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        st = mh_launch(PE, th0, Lm, 20, 0, chol, store_samples=False)
        e1.record(); torch.cuda.synchronize()
        out[f"{C}/{shape}"] = C * (Lm - 1) / (e0.elapsed_time(e1) * 1e-3)
        print(C, shape, f"{out[f'{C}/{shape}']:.4g}", float(st['stats'][:, 0].mean()) / (Lm - 1), flush=True)
print(json.dumps(out))
