"""C4: doEnsembleSliceSampling with 2048 walkers on a 12-parameter linear model with two 4096-point responses of dense
covariance (the DMMA likelihood path), round-synchronous driver (peuqse_b200/csrc/ess_rounds.cu).  Under torchrun every
round's batch of walkers is shared out over the ranks and the log-posteriors are all-gathered; the final
walker positions must be those of the single-GPU run (a checksum is printed).
usage: [torchrun --nproc-per-node N] python profiles/bench_ess_c4.py [iterations=10]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE
from peuqse_b200.mcmc import ess_launch
from oracle import peuqse_oracle as orc

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
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
start = 0.866 * 2.0 * np.random.default_rng(3).standard_normal((W, k)) * 0.25
ess_launch(PE, start, 2, 7, store_chain=False)                       # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
out = ess_launch(PE, start, iters, 7, store_chain=False)
b.record(); torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
info = out['info'].cpu().numpy()
if rank == 0:
    s = float(ms.item()) * 1e-3
    print(json.dumps(dict(config="C4: 2048 walkers, k=12, two 4096x4096 dense-covariance responses", n_gpus=world, iterations=iters,
                          seconds=s, walker_updates_per_s=W * iters / s, logp_evals=int(info[0]), logp_evals_per_s=info[0] / s,
                          mu=float(info[1]), expansions=int(info[2]), contractions=int(info[3]),
                          checksum_X=float(out['X'].double().sum().item()), checksum_Z=float(out['Z'].sum().item()))))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
