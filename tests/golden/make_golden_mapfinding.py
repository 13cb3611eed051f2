"""tests/golden/mapfinding.npz: doMetropolisHastings with mcmc_mode 'MAP_finding' run by the UNMODIFIED reference
(/root/reference) on Example00, fed injected draws, next to the oracle's restatement (oracle.mh_chain_map_finding).

The reference draws, per step, np.random.uniform() (modulation), np.random.multivariate_normal(0, Q*dyn*mod*step) and
np.random.uniform() (accept) (IP.py:2633-2635, 2650).  The fixture first runs the ORACLE on Philox draws (modulations, unbiased
increments L z scaled by sqrt(dyn*mod), accept uniforms), then feeds the reference those uniforms and the oracle's scaled
increments; the covariance argument of every injected multivariate_normal call is recorded, so that the reference's own
step-coefficient history dyn_i = cov_i[0,0] / (Q[0,0] * mod_i * step) is pinned as well (it comes from the reference's polyfit).
Run in the build container only:  python tests/golden/make_golden_mapfinding.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG                      # noqa: E402  (import shim, fresh_userinput, make_pe; chdirs to a scratch directory)
from oracle import peuqse_oracle as orc       # noqa: E402
from oracle import philox_np                  # noqa: E402

L, SEED, STEP = 1200, 3, 0.3
X00, Y00, S00 = MG.X00, MG.Y00, MG.S00
z, u = philox_np.mh_draws(SEED, 0, np.arange(1, L), 2)                       # normals z[L-1,2], accept uniforms
_, um = philox_np.uniform_pair(SEED, philox_np.STREAM_MH, 0, np.uint64(8) * np.arange(1, L, dtype=np.uint64) + np.uint64(7))
mods = um + 0.5
base = z * np.sqrt(STEP * np.array([100.0, 200.0]) ** 2)                     # L z for the diagonal Q = diag(sigma_prior^2)
P = orc.build_problem([200, 500], [100, 200], Y00, S00)
_, bt = orc.sim_quadratic(X00)
ora = orc.mh_chain_map_finding(P, [200.0, 500.0], base, mods, u, bt, scaled=False)
scaled = np.sqrt(ora['dyn'][1:] * mods)[:, None] * base

UI = MG.ui_example00()
UI.parameter_estimation_settings.update({'mcmc_length': L, 'mcmc_burn_in': 100, 'mcmc_relative_step_length': STEP, 'mcmc_mode': 'MAP_finding',
                                         'mcmc_random_seed': None, 'mcmc_threshold_filter_samples': False,
                                         'mcmc_info_gain_returned': 'log_ratio', 'mcmc_exportLog': False, 'exportLog': False})
PE = MG.make_pe(UI)
uq = []
for i in range(L - 1):
    uq += [um[i], u[i]]
uq = iter(uq)
dq = iter(scaled)
covs = []
real_mvn, real_uni = np.random.multivariate_normal, np.random.uniform


def fake_mvn(mean, cov, *a, **k):
    covs.append(float(np.asarray(cov)[0, 0]))
    return next(dq)


np.random.multivariate_normal = fake_mvn
np.random.uniform = lambda *a, **k: next(uq)
import contextlib, io                        # noqa: E402
with contextlib.redirect_stdout(io.StringIO()):
    PE.doMetropolisHastings()
np.random.multivariate_normal, np.random.uniform = real_mvn, real_uni
ref_samples = np.vstack([PE.during_burn_in_samples, PE.post_burn_in_samples])
ref_logp = np.vstack([PE.during_burn_in_log_posteriors_un_normed_vec, PE.post_burn_in_log_posteriors_un_normed_vec])[:, 0]
assert np.array_equal(ref_samples, ora['samples']), "oracle restatement differs from the reference"
assert np.array_equal(ref_logp, ora['logP'])
# the reference's own dyn: cov[0,0] = ((Q00 * dyn) * mod) * step evaluated left to right (IP.py:2635)
want = np.array([((100.0 ** 2 * ora['dyn'][i + 1]) * mods[i]) * STEP for i in range(L - 1)])
assert np.array_equal(np.array(covs), want), "step-coefficient history differs from the reference"
assert len(set(np.round(ora['dyn'], 12))) > 3, "the coefficient never adapted: uninformative fixture"
np.savez_compressed(os.path.join(HERE, "mapfinding.npz"), seed=SEED, step=STEP, L=L, mods=mods, us=u, base_deltas=base,
                    scaled_deltas=scaled, samples=ref_samples, logP=ref_logp, dyn=ora['dyn'], dyn_final=ora['dyn_final'],
                    ref_cov00=np.array(covs))
print("mapfinding.npz: reference == oracle over", L, "steps; dyn history", sorted(set(np.round(ora['dyn'], 6)))[:8], "final", ora['dyn_final'])
