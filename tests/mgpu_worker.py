"""Run under torchrun with N >= 2 GPUs: the sharded multistart / MH paths must reproduce the single-GPU results.
(Launched by tests/test_multi_gpu.py; not collected by pytest itself.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "example00.npz"))
X00, Y00, S00 = [600, 1100, 1400], [360500, 580500, 1620500], [200000, 300000, 200000]


def make(settings):
    UI = PEUQSE.UserInput.fresh()
    UI.responses['responses_observed'] = Y00
    UI.responses['responses_observed_uncertainties'] = S00
    UI.model['InputParameterPriorValues'] = [200, 500]
    UI.model['InputParametersPriorValuesUncertainties'] = [100, 200]
    UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.QuadraticModel(X00)
    UI.parameter_estimation_settings.update(settings)
    return PEUQSE.parameter_estimation(UI)


# 1. reference grid (generic kernels: short last axis), sharded over the ranks
PE = make({'multistart_initialPointsDistributionType': 'grid', 'multistart_gridsearchSamplingInterval': [5, 5],
           'multistart_gridsearchSamplingRadii': [50, 50]})
PE.doMultiStart()
assert PE.map_index == 13 and PE.map_logP == float(g['ex00_grid_map_logP'])
assert PE.sweep_result.sum_weights == int(g['ex00_grid_nrows_expanded'])
np.testing.assert_allclose(PE.mu_AP_parameter_set, g['ex00_grid_mu_AP'], rtol=1e-12)
np.testing.assert_allclose(PE.stdap_parameter_set, g['ex00_grid_stdap'], rtol=1e-10)
# 2. long last axis (specialised + TMA kernels), compared with a one-rank run of the same sweep on every rank
settings = {'multistart_initialPointsDistributionType': 'grid', 'multistart_gridsearchSamplingInterval': [2.5, 0.4],
            'multistart_gridsearchSamplingRadii': [37, 2900]}
PE = make(settings)
PE.doMultiStart()
from peuqse_b200.sweep import Sweep, make_spec
spec = make_spec("grid", 2, [200., 500.], interval=[2.5, 0.4], radii=[37, 2900])
N = 75 * 5801
one = Sweep(PE.device_problem, spec, N, shard=(0, 1))
one.world = 1
R1 = one.unpack(one_res := (one.step(float(np.log(1e-2))), one.res.cpu().numpy())[1], 2.0)
assert PE.map_index == R1.map_index and PE.map_logP == R1.map_logP
assert PE.sweep_result.sum_weights == R1.sum_weights and PE.sweep_result.n_retained == R1.n_retained
np.testing.assert_allclose(PE.mu_AP_parameter_set, R1.mu_AP_parameter_set, rtol=1e-13)
np.testing.assert_allclose(PE.stdap_parameter_set, R1.stdap_parameter_set, rtol=1e-11)
assert np.array_equal(PE.marginal_histograms[0], R1.histogram_counts)
# 2b. the same sweep with each rank's share dealt as 3 interleaved pieces (multistart_shard_pieces)
try:
    multi = Sweep(PE.device_problem, spec, N, pieces=3)
    assert len(multi.ranges) == 3
    Rm = multi.run(2.0)
    assert (Rm.map_index, Rm.map_logP) == (R1.map_index, R1.map_logP)
    assert Rm.sum_weights == R1.sum_weights and Rm.n_retained == R1.n_retained
    assert np.array_equal(Rm.histogram_counts, R1.histogram_counts)
    np.testing.assert_allclose(Rm.mu_AP_parameter_set, R1.mu_AP_parameter_set, rtol=1e-12)
    np.testing.assert_allclose(Rm.stdap_parameter_set, R1.stdap_parameter_set, rtol=1e-10)
except Exception:
    import traceback
    traceback.print_exc()
    sys.stdout.flush(); sys.stderr.flush()
    raise
# 3. MH chains sharded by global chain index: same pooled moments as one rank running all chains
PE = make({'mcmc_length': 400, 'mcmc_burn_in': 100, 'mcmc_nchains': 64, 'mcmc_random_seed': 5, 'mcmc_relative_step_length': 0.5,
           'mcmc_threshold_filter_samples': False, 'mcmc_info_gain_returned': 'log_ratio'})
PE.doMetropolisHastings()
from peuqse_b200.mcmc import mh_launch, step_cholesky, chain_moments
all_ = mh_launch(PE, np.tile([200., 500.], (64, 1)), 400, 100, 5, step_cholesky(PE.Q_covmat, 0.5))
S = all_['samples'][100:].permute(1, 0, 2).reshape(-1, 2).cpu().numpy()
assert PE.post_burn_in_samples.shape == S.shape and np.array_equal(PE.post_burn_in_samples, S)
# 3b. the same chains with the samples not stored (in-kernel moments): MAP, last points and acceptance must be those of ALL
#     chains on every rank (not of the rank's own block), the pooled moments those of the stored run
from peuqse_b200 import _lib
PEn = make({'mcmc_length': 400, 'mcmc_burn_in': 100, 'mcmc_nchains': 64, 'mcmc_random_seed': 5, 'mcmc_relative_step_length': 0.5,
            'mcmc_store_samples': False})
outn = PEn.doMetropolisHastings()
lay = _lib.mh_stat_layout(2)
st = all_['stats'].cpu().numpy()
ibest = int(np.argmax(st[:, lay['MAPLOGP']]))
assert PEn.map_logP == st[ibest, lay['MAPLOGP']] and np.array_equal(PEn.map_parameter_set, st[ibest, lay['MAPTHETA']:lay['MAPTHETA'] + 2])
assert np.array_equal(PEn.mcmc_last_point_sampled, st[:, lay['LAST']:lay['LAST'] + 2])
np.testing.assert_allclose(outn[1], S.mean(axis=0), rtol=1e-12)
np.testing.assert_allclose(outn[2], S.std(axis=0), rtol=1e-10)
try:
    make({'mcmc_length': 50, 'mcmc_nchains': 1, 'mcmc_random_seed': 5}).doMetropolisHastings()
    raise AssertionError("one chain on several ranks must be refused")
except ValueError:
    pass
# 4. design of experiments: the 9 x 25 scenarios of the reference's Example13 unit test sharded over the ranks
gd = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "doe.npz"))
Ld, Bd, STEP, DH, DS, DHU, DSU, SIG = gd['settings']
UI = PEUQSE.UserInput.fresh()
UI.responses['responses_observed'] = [0.0, 0.0, 0.0]
UI.responses['responses_observed_uncertainties'] = [SIG] * 3
UI.model['InputParameterPriorValues'] = [DH, DS]
UI.model['InputParametersPriorValuesUncertainties'] = [DHU, DSU]
UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.LangmuirLogModel(400, 0.1)
UI.parameter_estimation_settings.update({'mcmc_random_seed': 0, 'mcmc_burn_in': int(Bd), 'mcmc_length': int(Ld),
                                         'mcmc_relative_step_length': float(STEP), 'mcmc_info_gain_returned': 'log_ratio',
                                         'mcmc_threshold_filter_samples': False})
UI.doe_settings.update({'info_gains_matrices_array_format': 'xyz', 'independent_variable_grid_center': [500, 0.5],
                        'independent_variable_grid_interval_size': [100, 0.1], 'independent_variable_grid_num_intervals': [2, 2],
                        'parameter_modulation_grid_interval_size': [1, 1], 'parameter_modulation_grid_num_intervals': [1, 1]})
PE = PEUQSE.parameter_estimation(UI)
A = PE.doeParameterModulationPermutationsScanner()
ref = gd['scanner_xyz_info_gains_matrices_array']
assert A.shape == ref.shape and np.array_equal(A[:, :, :2], ref[:, :, :2])
np.testing.assert_allclose(A[:, :, 2], ref[:, :, 2], rtol=1e-9)
# 5. ensemble slice sampling v2: the round-synchronous driver with every round's batch shared out over the ranks
#    and all-gathered must produce the chain of the unsharded driver (full-covariance responses: the C4 path at a small size)
from oracle import peuqse_oracle as orc
from peuqse_b200.mcmc import ess_launch
rng = np.random.default_rng(21)
kk, nn = 3, 24
Phi = orc.cos_basis(2 * nn, kk)
xx = np.linspace(0, 1, nn)
Sig = 0.1 ** 2 * np.exp(-(xx[:, None] - xx[None, :]) ** 2 / (2 * 0.05 ** 2)) + 0.05 ** 2 * np.eye(nn)
Lc = np.linalg.cholesky(Sig)
yy = Phi @ rng.standard_normal(kk) + np.concatenate([Lc @ rng.standard_normal(nn), Lc @ rng.standard_normal(nn)])
UI = PEUQSE.UserInput.fresh()
UI.responses['responses_observed'] = [list(yy[:nn]), list(yy[nn:])]
UI.responses['responses_observed_uncertainties'] = [Sig, Sig]
UI.model['InputParameterPriorValues'] = [0.0] * kk
UI.model['InputParametersPriorValuesUncertainties'] = [2.0] * kk
UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.LinearModel(Phi)
PE = PEUQSE.parameter_estimation(UI)
start = 0.5 * np.random.default_rng(3).standard_normal((64, kk))
a = ess_launch(PE, start, 25, 11, shard=True)
b = ess_launch(PE, start, 25, 11, shard=False)
assert torch.equal(a['chain'], b['chain']) and torch.equal(a['chain_logp'], b['chain_logp']) and torch.equal(a['info'], b['info'])
# 6. doEnsembleSliceSampling itself, unseeded (every rank must end up with rank 0's seed and start walkers) and seeded
for seed in (None, 5):
    UI.parameter_estimation_settings.update(mcmc_length=64 * 30, mcmc_nwalkers=64, mcmc_burn_in=64 * 5, mcmc_random_seed=seed,
                                            mcmc_threshold_filter_samples=False, mcmc_exportLog=False)
    PE2 = PEUQSE.parameter_estimation(UI)
    PE2.doEnsembleSliceSampling()
    mine = torch.from_numpy(PE2.post_burn_in_samples).cuda()
    both = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    assert all(torch.equal(both[0], b_) for b_ in both), "ranks sampled different ensembles"
    if hasattr(UI, 'last_MCMC_run_type'):
        del UI.last_MCMC_run_type
dist.barrier()
if rank == 0:
    print("MGPU_OK", world)
dist.destroy_process_group()
