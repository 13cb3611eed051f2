"""Throughput of the other BASELINE.json configurations (C1 replicas, C3, C5) on one GPU -- not the driver's bench line."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import peuqse_b200 as PEUQSE
from peuqse_b200.mcmc import mh_launch, step_cholesky
from peuqse_b200.sweep import Sweep, make_spec

def ev_time(fn, n=3, warm=1):
    for _ in range(warm): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def make(model, obs, unc, mu, sd, **pes):
    UI = PEUQSE.UserInput.fresh()
    UI.responses['responses_observed_max_covmat_size'] = 5000
    UI.responses['responses_observed'] = obs; UI.responses['responses_observed_uncertainties'] = unc
    UI.model['InputParameterPriorValues'] = mu; UI.model['InputParametersPriorValuesUncertainties'] = sd
    UI.model['simulateByInputParametersOnlyFunction'] = model
    UI.parameter_estimation_settings['mcmc_length'] = 10
    UI.parameter_estimation_settings.update(pes)
    return PEUQSE.parameter_estimation(UI)

out = {}
which = sys.argv[1:] or ["c1", "c3", "c5", "c4"]
if "c1" in which:
    PE = make(PEUQSE.QuadraticModel([600, 1100, 1400]), [360500, 580500, 1620500], [200000, 300000, 200000], [200, 500], [100, 200])
    C, Lm = 65536, 2000
    th0 = np.tile([200., 500.], (C, 1))
    ms = ev_time(lambda: mh_launch(PE, th0, Lm, 200, 0, step_cholesky(PE.Q_covmat, 0.1), store_samples=False))
    out["c1_replicas_chain_steps_per_s"] = C * (Lm - 1) / (ms * 1e-3)
    ms1 = ev_time(lambda: mh_launch(PE, th0[:1], 10000, 1000, 0, step_cholesky(PE.Q_covmat, 0.1), store_samples=True), n=3)
    out["c1_single_chain_10k_steps_ms"] = ms1
if "c3" in which:
    n = 1024
    model = PEUQSE.Tpd2SiteModel.linear_ramp(n)
    from oracle import peuqse_oracle as orc
    kBT, c1, c2 = orc.tpd2site_constants(n)
    truth = np.array([41.5, 45, 13, 13, 0.1, 0.1])
    sim_t = orc.sim_tpd2site(kBT, c1, c2)[0](truth)
    sig = 0.02 * np.max(sim_t)
    y = sim_t + sig * np.random.default_rng(0).standard_normal(n)
    sd = [20, 20, 13, 13, 0.1, 0.1]
    PE = make(model, list(y), list(np.full(n, sig)), [41.5, 41.5, 13.0, 13.0, 0.1, 0.1], sd)
    C, Lm = 4096, 1000
    th0 = np.tile(truth, (C, 1))
    ms = ev_time(lambda: mh_launch(PE, th0, Lm, 100, 0, step_cholesky(PE.Q_covmat, 1e-5), store_samples=False), n=2)
    out["c3_chain_steps_per_s"] = C * (Lm - 1) / (ms * 1e-3)
    out["c3_ms_for_4096x1000"] = ms
    st = mh_launch(PE, th0, Lm, 100, 0, step_cholesky(PE.Q_covmat, 1e-5), store_samples=False)['stats'].cpu().numpy()
    out["c3_accept_rate"] = float(st[:, 0].mean() / (Lm - 1))
    B = 1 << 18
    th = torch.from_numpy(truth + 0.01 * np.random.default_rng(1).standard_normal((B, 6))).cuda()
    ms = ev_time(lambda: PE.getLogP_batch(th))
    out["c3_logp_batch_evals_per_s"] = B / (ms * 1e-3)
if "c5" in which:
    k, n = 10, 64
    from oracle import peuqse_oracle as orc
    Phi = orc.cos_basis(n, k)
    truth = np.random.default_rng(5).standard_normal(k) * 0.5
    y = Phi @ truth + 0.1 * np.random.default_rng(6).standard_normal(n)
    PE = make(PEUQSE.LinearModel(Phi), list(y), list(np.full(n, 0.1)), [0.0] * k, [1.0] * k)
    N = 1 << 27
    spec = make_spec("uniform", k, np.zeros(k), spread=0.866, std_prior=np.ones(k), seed=0)
    sw = Sweep(PE.device_problem, spec, N, store_logp=False)
    ms = ev_time(sw.pass1, n=2)
    out["c5_pass1_evals_per_s"] = N / (ms * 1e-3)
    out["c5_pass1_ms_for_2^27"] = ms
    lr = float(np.log(10 ** -2.0))
    ms = ev_time(lambda: sw.step(lr, True), n=2)
    out["c5_whole_sweep_ms_for_2^27"] = ms
    out["c5_whole_sweep_evals_per_s"] = N / (ms * 1e-3)
    R = sw.unpack(sw.res.cpu().numpy(), 2.0)
    out["c5_n_retained"] = R.n_retained
if "c4" in which:
    from oracle import peuqse_oracle as orc
    k, n = 12, 4096
    x = np.linspace(0, 1, n)
    Phi = orc.cos_basis(2 * n, k)
    Sig = 0.1 ** 2 * np.exp(-(x[:, None] - x[None, :]) ** 2 / (2 * 0.02 ** 2)) + 0.05 ** 2 * np.eye(n)
    truth = np.random.default_rng(1).standard_normal(k)
    y = Phi @ truth + 0.1 * np.random.default_rng(2).standard_normal(2 * n)
    t0 = time.time()
    PE = make(PEUQSE.LinearModel(Phi), [list(y[:n]), list(y[n:])], [Sig, Sig], [0.0] * k, [2.0] * k)
    PE.UserInput  # noqa
    out["c4_host_setup_s"] = time.time() - t0
    B = 1024
    th = torch.from_numpy(truth + 0.05 * np.random.default_rng(3).standard_normal((B, k))).cuda()
    ms = ev_time(lambda: PE.getLogP_batch(th), n=5)
    out["c4_logp_batch_ms_B1024_two_4096_responses"] = ms
    out["c4_evals_per_s"] = B / (ms * 1e-3)
    out["c4_dmma_tflops_executed"] = 2 * (2.0 * B * n * n) / (ms * 1e-3) / 1e12
print(json.dumps(out, indent=1))
