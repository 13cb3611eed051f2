"""TEST INFRASTRUCTURE ONLY -- CPU (numpy/scipy) restatement of the PEUQSE log-posterior hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product (peuqse_b200/) never does; it fails loudly without its CUDA library.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function here against
  * the reference's own stored outputs (Examples/Example00 and Example26 multistart CSV rows),
  * fixtures produced by importing and running the UNMODIFIED reference in the build container
    (tests/golden/make_golden.py, outputs tests/golden/*.npz).
Exception: ensemble slice sampling (`ess_*`) restates the published algorithm of zeus-mcmc
(Karamanis & Beutler, arXiv:2002.06212; zeus-mcmc >= 2.4.0 per the reference's setup.py:36, cited
at InverseProblem.py:2356-2359).  zeus is not vendored in the reference and not installed here, so
for ESS "parity unpinned" -- it is checked statistically against closed-form posteriors only.

All file:line citations are to /root/reference/PEUQSE/InverseProblem.py ("IP.py") unless stated.
Two flavours of every evaluation are kept:
  *_point   -- one theta at a time through scipy.stats.multivariate_normal.logpdf, exactly as the
               reference does it (eigh per call).  This is the timed CPU baseline ("port").
  *_batch   -- vectorised numpy with the eigendecomposition hoisted out (done once), same
               floating-point operation order per point, for parity checks at larger sizes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np
import scipy.linalg
from scipy.stats import multivariate_normal

LOG2PI = float(np.log(2 * np.pi))  # scipy/stats/_multivariate.py:42


# --------------------------------------------------------------------------------------------
# Host-side normalisation (IP.py:146-350): turns user input into the constants of the hot path
# --------------------------------------------------------------------------------------------
@dataclass
class Whitening:
    """scipy _PSD of one covariance (scipy/stats/_multivariate.py:208-236), computed once."""
    rank: int
    log_pdet: float
    U: np.ndarray                     # (n, n) = u * sqrt(1/s)
    is_perm: bool = False             # U has exactly one non-zero per column (diagonal cov)
    perm: Optional[np.ndarray] = None  # row of the non-zero in column j
    w: Optional[np.ndarray] = None     # |U[perm[j], j]|


def psd_whitening(cov) -> Whitening:
    """What multivariate_normal.logpdf builds from `cov` on every call (allow_singular=False)."""
    cov = np.asarray(cov, dtype=float)
    if cov.ndim == 0:
        cov = cov.reshape(1, 1)
    elif cov.ndim == 1:
        cov = np.diag(cov)          # _multivariate.py:533
    s, u = scipy.linalg.eigh(cov, lower=True, check_finite=True)
    eps = 1e6 * np.finfo('d').eps * np.max(np.abs(s))   # _eigvalsh_to_eps
    if np.min(s) < -eps:
        raise ValueError("The input matrix must be symmetric positive semidefinite.")
    d = s[s > eps]
    if len(d) < len(s):
        raise np.linalg.LinAlgError("singular covariance (allow_singular=False)")
    s_pinv = np.array([0 if abs(x) <= eps else 1 / x for x in s], dtype=float)  # _pinv_1d
    U = np.multiply(u, np.sqrt(s_pinv))
    wh = Whitening(rank=len(d), log_pdet=float(np.sum(np.log(d))), U=U)
    nz = U != 0
    if np.all(nz.sum(axis=0) == 1):
        wh.is_perm = True
        wh.perm = np.argmax(nz, axis=0).astype(np.int64)
        wh.w = np.abs(U[wh.perm, np.arange(U.shape[1])])
    return wh


@dataclass
class ResponseBlock:
    """One response dimension of the likelihood (IP.py:2949-2979)."""
    y: np.ndarray                 # observed (transformed) values, (n,)
    cov: np.ndarray               # variances (n,) or full covariance (n,n), sim-uncertainty already added
    per_value: bool               # True: n > responses_observed_max_covmat_size (IP.py:2951)
    wh: Optional[Whitening] = None          # mvn path
    w_val: Optional[np.ndarray] = None      # per-value path: sqrt(1/V_j)
    c_val: Optional[np.ndarray] = None      # per-value path: 1*LOG2PI + ln V_j


@dataclass
class OracleProblem:
    k: int
    mu_prior: np.ndarray
    std_prior: np.ndarray         # 1-D sigma (IP.py:173) or elementwise sqrt of covmat (IP.py:179)
    covmat_prior: np.ndarray
    scaling: Optional[np.ndarray]  # None == 'off'
    mu_prior_scaled: np.ndarray
    var_prior_scaled: np.ndarray
    covmat_prior_scaled: np.ndarray
    uniform_idx: Optional[List[int]]      # None: attribute absent (no -1 uncertainties)
    lb: Optional[np.ndarray]              # object arrays with None entries, or None when list empty
    ub: Optional[np.ndarray]
    responses: List[ResponseBlock] = field(default_factory=list)
    prior_wh: Optional[Whitening] = None  # of the (possibly truncated) scaled prior covariance
    gauss_idx: Optional[np.ndarray] = None
    max_covmat_size: int = 100
    Q_covmat: Optional[np.ndarray] = None


def build_problem(prior_values, prior_uncertainties, observed, observed_uncertainties,
                  lower_bounds=(), upper_bounds=(), scaling_type="std",
                  simulation_uncertainties=None, max_covmat_size=100) -> OracleProblem:
    """Restates parameter_estimation.__init__ for the hot-path constants.

    prior: IP.py:146-179; scaling: IP.py:311-339; observed covmat: IP.py:343-346 + 3591-3606;
    simulation uncertainties added as variances: IP.py:2895, 2931; Q: IP.py:349-350.
    `observed`, `observed_uncertainties`: 1-D (single response) or list of 1-D / 2-D per response.
    """
    mu = np.array(prior_values, dtype=float)
    unc = np.array(prior_uncertainties, dtype=float)
    k = len(mu)
    lb = np.array(list(lower_bounds), dtype=object) if len(lower_bounds) else None
    ub = np.array(list(upper_bounds), dtype=object) if len(upper_bounds) else None
    uniform_idx = None
    if -1.0 in unc:                                           # IP.py:149
        uniform_idx = []
        for i, u in enumerate(unc):
            if u == -1.0:
                uniform_idx.append(i)
                if lb is None or ub is None:
                    raise SystemExit("uniform prior needs both bounds (IP.py:164-165)")
                unc[i] = (ub[i] - lb[i]) / (12 ** 0.5)        # IP.py:166
                mu[i] = (ub[i] + lb[i]) / 2                   # IP.py:169
    if unc.ndim == 1:                                         # IP.py:172-175
        std = np.array(unc, dtype=float)
        var = np.power(unc, 2)
        covmat = np.diagflat(var)
    else:                                                     # IP.py:176-179
        covmat = np.array(unc, dtype=float)
        var = np.diagonal(covmat)
        std = np.power(covmat, 0.5)
    if scaling_type == "off":                                 # IP.py:311-314
        scaling = None
        mu_s, var_s, cov_s = mu * 1.0, var * 1.0, covmat * 1.0
    else:
        if scaling_type == "std":
            scaling = std
        elif scaling_type == "mu":
            scaling = mu
        else:
            scaling = (mu / mu) * float(scaling_type)         # IP.py:321-322
        if np.ndim(scaling) == 2:                             # IP.py:327-328
            scaling = np.diagonal(scaling) ** 0.5
        mu_s = np.array(mu / scaling)                         # IP.py:333
        var_s = np.array(var / (scaling * scaling))           # IP.py:334
        cov_s = covmat * 1.0
        for i, sv in enumerate(scaling):                      # IP.py:336-339
            cov_s[i, :] = cov_s[i, :] / sv
            cov_s[:, i] = cov_s[:, i] / sv

    P = OracleProblem(k=k, mu_prior=mu, std_prior=std, covmat_prior=covmat, scaling=scaling,
                      mu_prior_scaled=mu_s, var_prior_scaled=var_s, covmat_prior_scaled=cov_s,
                      uniform_idx=uniform_idx, lb=lb, ub=ub, max_covmat_size=max_covmat_size,
                      Q_covmat=covmat)
    # prior whitening, hoisted (IP.py:2772-2785)
    if uniform_idx is None:
        P.gauss_idx = np.arange(k)
        P.prior_wh = psd_whitening(cov_s)
    else:
        P.gauss_idx = np.delete(np.arange(k), uniform_idx)
        if len(P.gauss_idx):
            P.prior_wh = psd_whitening(np.diagflat(np.delete(var_s, uniform_idx)))  # IP.py:2778-2781

    # responses: make 2-D nested
    obs = _nest(observed)
    ounc = _nest_unc(observed_uncertainties, len(obs))
    sunc = None if simulation_uncertainties is None else _nest_unc(simulation_uncertainties, len(obs))
    single = len(obs) == 1
    for r, y in enumerate(obs):
        cov = _shaped_response_cov(ounc, r)
        if sunc is not None:
            cov = cov + _shaped_response_cov(sunc, r)                     # IP.py:2931
        if single and cov.ndim == 2:
            raise NotImplementedError("single-response full covariance is dead code in the reference "
                                      "(SURVEY Q2); use >= 2 responses")
        per_value = len(y) > max_covmat_size                              # IP.py:2951
        blk = ResponseBlock(y=np.asarray(y, dtype=float), cov=cov, per_value=per_value)
        if per_value:
            if cov.ndim != 1:
                raise NotImplementedError("per-value path needs variances (IP.py:2971)")
            wl = [psd_whitening(v) for v in cov] if len(cov) <= 4096 else None
            # scalar logpdf: U = sqrt(1/V), log_pdet = ln V, rank 1
            blk.w_val = np.sqrt(1.0 / cov)
            blk.c_val = 1 * LOG2PI + np.log(cov)
            if wl is not None:   # guard: hoisted constants equal what scipy builds per value
                assert np.array_equal(blk.w_val, np.array([abs(w.U[0, 0]) for w in wl]))
                assert np.array_equal(blk.c_val, np.array([w.rank * LOG2PI + w.log_pdet for w in wl]))
        else:
            blk.wh = psd_whitening(cov)
        P.responses.append(blk)
    return P


def _shaped_response_cov(unc_list, r):
    """returnShapedResponseCovMat, IP.py:3591-3606.

    One response: a 1-D list of sigmas arrives as shape (1, n) and is squared into variances (3595-3596);
    anything else is taken as a covariance matrix.
    Two or more responses (reference quirk, kept for parity; verified against the fixtures): entry r is
    squared only if np.shape(entry)[0] == 1, and the array is re-created on every loop iteration (3601),
    so only the LAST response can ever be squared and only when it holds a single value.  A vector of n > 1
    sigmas is therefore used AS IF it were variances.
    """
    u = np.asarray(unc_list[r], dtype=float)
    nresp = len(unc_list)
    if nresp == 1:
        return np.square(u) if u.ndim == 1 else u
    if r == nresp - 1 and u.shape[0] == 1:
        return np.square(u)
    return u


def _nest(x):
    a = np.asarray(x, dtype=object) if _is_ragged(x) else np.asarray(x, dtype=float)
    if a.dtype != object and a.ndim == 1:
        return [np.asarray(a, dtype=float)]
    return [np.asarray(r, dtype=float) for r in x]


def _nest_unc(x, nresp):
    if nresp == 1:
        a = np.asarray(x, dtype=float)
        if a.ndim == 1:
            return [a]
        if a.ndim == 2 and a.shape[0] == 1:
            return [a[0]]
        return [a]            # full covariance for the single response
    return [np.asarray(r, dtype=float) for r in x]


def _is_ragged(x):
    try:
        np.asarray(x, dtype=float)
        return False
    except (ValueError, TypeError):
        return True


# --------------------------------------------------------------------------------------------
# Faithful per-point path (what the reference executes; the timed CPU baseline)
# --------------------------------------------------------------------------------------------
def bounds_check(values, bounds, kind):
    """boundsCheck, IP.py:3608-3633 (inclusive; None = unbounded)."""
    values = np.array(values).flatten()
    bounds = np.array(bounds).flatten()
    v = values[bounds != None].flatten()       # noqa: E711  (numpy elementwise, as in the reference)
    b = bounds[bounds != None].flatten()       # noqa: E711
    if kind == 'upper':
        if False in (v <= b):
            return False
    if kind == 'lower':
        if False in (v >= b):
            return False
    return True


def log_prior_point(P: OracleProblem, theta) -> float:
    """getLogPrior, IP.py:2756-2794 (custom_logPrior and undo_scaling omitted)."""
    if P.ub is not None and not bounds_check(theta, P.ub, 'upper'):     # IP.py:2797-2800
        return float('-inf')
    if P.lb is not None and not bounds_check(theta, P.lb, 'lower'):     # IP.py:2801-2804
        return float('-inf')
    if P.scaling is None:
        ts = np.array(theta) * 1.0                                      # IP.py:2764
    else:
        ts = np.array(theta) / P.scaling                                # IP.py:2767
    if P.uniform_idx is None:
        return multivariate_normal.logpdf(x=ts, mean=P.mu_prior_scaled, cov=P.covmat_prior_scaled)  # IP.py:2773
    ts_t = np.delete(ts, P.uniform_idx)
    mu_t = np.delete(P.mu_prior_scaled, P.uniform_idx)
    cov_t = np.diagflat(np.delete(P.var_prior_scaled, P.uniform_idx))
    if len(cov_t) == 0:
        return 0                                                        # IP.py:2782-2783
    return multivariate_normal.logpdf(x=ts_t, mean=mu_t, cov=cov_t)     # IP.py:2785


def log_likelihood_point(P: OracleProblem, sim) -> float:
    """getLogLikelihood + getLogLikelihood_byResponses, IP.py:2850-2980, given simulated responses.

    `sim`: None / 1-D (single response) / list of 1-D per response.
    """
    if sim is None:
        return float('-inf')                                            # IP.py:2863-2865
    sims = _nest(sim)
    if any(np.isnan(s).any() for s in sims):
        return float('-inf')                                            # IP.py:2880-2882
    total = 0                                                           # IP.py:2948
    for r, blk in enumerate(P.responses):
        if not blk.per_value:
            ll = multivariate_normal.logpdf(mean=sims[r], x=blk.y, cov=blk.cov)     # IP.py:2960
        else:
            ll = 0                                                      # IP.py:2968
            for j in range(len(sims[r])):                               # IP.py:2969-2978
                cur = multivariate_normal.logpdf(mean=sims[r][j], x=blk.y[j], cov=blk.cov[j])
                if float(cur) == float('-inf'):
                    cur = -1E100
                ll = cur + ll
        total = total + ll                                              # IP.py:2979
    return total


def log_likelihood_simvar_point(P: OracleProblem, sim, sim_sd) -> float:
    """getLogLikelihood when model['responses_simulation_uncertainties'] is a function of the parameters (IP.py:563-572,
    2883-2896): `sim_sd` (standard deviations, shaped like `sim`) is squared by returnShapedResponseCovMat and ADDED to the
    observed variances (IP.py:2925-2927) before the same per-response densities as log_likelihood_point.  Responses with 1-D
    (variance) uncertainties."""
    if sim is None:
        return float('-inf')
    sims, sds = _nest(sim), _nest(sim_sd)
    if any(np.isnan(s).any() for s in sims):
        return float('-inf')
    total = 0
    for r, blk in enumerate(P.responses):
        cov = np.asarray(blk.cov, dtype=float) + np.square(np.asarray(sds[r], dtype=float))      # IP.py:2927
        if not blk.per_value:
            ll = multivariate_normal.logpdf(mean=sims[r], x=blk.y, cov=cov)
        else:
            ll = 0
            for j in range(len(sims[r])):
                cur = multivariate_normal.logpdf(mean=sims[r][j], x=blk.y[j], cov=cov[j])
                if float(cur) == float('-inf'):
                    cur = -1E100
                ll = cur + ll
        total = total + ll
    return total


def little_euler(t, dydt):
    """littleEulerGivenArray with y_initial = 0 (IP.py:3501-3513): explicit Euler integral of dydt over t, sequential."""
    y = np.zeros(len(t))
    for j in range(len(t) - 1):
        y[j + 1] = y[j] + dydt[j] * (t[j + 1] - t[j])
    return y


def little_euler_uncertainty(unc, t, initial):
    """littleEulerUncertaintyPropagation (IP.py:3516-3525)."""
    unc = np.asarray(unc, dtype=float)
    out = unc * 0.0
    out[0] = initial
    for j in range(len(unc) - 1):
        out[j + 1] = ((out[j]) ** 2 + (unc[j] * (t[j + 1] - t[j])) ** 2) ** 0.5
    if initial == 0:
        out[0] = out[1]
    return out


def transient_passes(response_type, data_type):
    """How often transform_responses integrates a response of data_overcategory 'transient_kinetics' (IP.py:620-657): an intermediate
    given as a concentration once and as a rate twice, a reactant or product given as a rate once, everything else not at all."""
    if response_type == 'I':
        return {'c': 1, 'r': 2}.get(data_type, 0)
    if response_type in ('R', 'P'):
        return 1 if data_type == 'r' else 0
    return 0


def transform_responses(abscissa, types, data_types, responses, uncertainties=None):
    """transform_responses for 'transient_kinetics' (IP.py:609-658).  abscissa: one per response, or a single shared one."""
    resp = [np.array(r, dtype=float) for r in _nest(responses)]
    ab = _nest(abscissa)
    unc = None if uncertainties is None else [np.array(u, dtype=float) for u in _nest_unc(uncertainties, len(resp))]
    for r in range(len(resp)):
        t = ab[0] if len(ab) == 1 else ab[r]
        base = None if unc is None else np.mean(unc[r]) / 10                     # of the UNtransformed uncertainties, both passes
        for _ in range(transient_passes(types[r], data_types[r])):
            resp[r] = little_euler(t, resp[r])
            if unc is not None:
                unc[r] = little_euler_uncertainty(unc[r], t, base)
    return resp, unc


def log_p_point(P: OracleProblem, theta, simulate: Callable) -> float:
    """getLogP, IP.py:1647-1654."""
    lp = log_prior_point(P, theta)
    if lp == float('-inf'):
        return float('-inf')
    ll = log_likelihood_point(P, simulate(tuple(np.ndarray.copy(np.array(theta)))))   # IP.py:2851
    return ll + lp


# --------------------------------------------------------------------------------------------
# Vectorised path: same per-point operation order, eigendecompositions hoisted
# --------------------------------------------------------------------------------------------
def _np_rowsum(a):
    """np.sum(a, axis=-1) on a C-contiguous (B, n) array == numpy pairwise order per row, which is
    what scipy's `np.sum(np.square(...), axis=-1)` does for its (1, n) input (_multivariate.py:590)."""
    return np.sum(np.ascontiguousarray(a), axis=-1)


def _maha(dev, wh: Whitening):
    if wh.is_perm:
        q = np.square(dev[:, wh.perm] * wh.w)       # (dev @ U)_j = dev[perm_j] * (+-w_j): exact
    else:
        q = np.square(dev @ wh.U)                   # BLAS order: tolerance, not bit parity
    return _np_rowsum(q)


def log_prior_batch(P: OracleProblem, thetas) -> np.ndarray:
    th = np.atleast_2d(np.asarray(thetas, dtype=float))
    B = th.shape[0]
    ok = np.ones(B, dtype=bool)
    if P.ub is not None:
        for i, b in enumerate(P.ub):
            if b is not None:
                ok &= th[:, i] <= float(b)
    if P.lb is not None:
        for i, b in enumerate(P.lb):
            if b is not None:
                ok &= th[:, i] >= float(b)
    ts = th * 1.0 if P.scaling is None else th / P.scaling
    if P.prior_wh is None:
        lp = np.zeros(B)
    else:
        dev = ts[:, P.gauss_idx] - P.mu_prior_scaled[P.gauss_idx]
        wh = P.prior_wh
        lp = -0.5 * (wh.rank * LOG2PI + wh.log_pdet + _maha(dev, wh))     # _multivariate.py:591
    return np.where(ok, lp, -np.inf)


def log_likelihood_batch(P: OracleProblem, sims: Sequence[np.ndarray]) -> np.ndarray:
    """sims: list over responses of (B, n_r) arrays."""
    B = sims[0].shape[0]
    bad = np.zeros(B, dtype=bool)
    total = np.zeros(B)
    for r, blk in enumerate(P.responses):
        s = np.asarray(sims[r], dtype=float)
        bad |= np.isnan(s).any(axis=1)
        dev = blk.y - s                                                   # x - mean, IP.py:2960
        with np.errstate(invalid='ignore', over='ignore'):
            if not blk.per_value:
                wh = blk.wh
                ll = -0.5 * (wh.rank * LOG2PI + wh.log_pdet + _maha(dev, wh))
            else:
                terms = -0.5 * (blk.c_val + np.square(dev * blk.w_val))   # scalar logpdf per value
                terms = np.where(terms == -np.inf, -1E100, terms)         # IP.py:2975-2977
                ll = np.zeros(B)
                for j in range(terms.shape[1]):                           # sequential, IP.py:2978
                    ll = terms[:, j] + ll
            total = total + ll
    return np.where(bad, -np.inf, total)


def log_p_batch(P: OracleProblem, thetas, simulate_batch: Callable):
    """-> (logP, logPrior, logLik); logLik is computed for every row (as MH does) while
    logP short-circuits to -inf for out-of-bounds rows (IP.py:1649)."""
    th = np.atleast_2d(np.asarray(thetas, dtype=float))
    lp = log_prior_batch(P, th)
    with np.errstate(all='ignore'):
        sims = simulate_batch(th)
    if isinstance(sims, np.ndarray):
        sims = [sims]
    ll = log_likelihood_batch(P, sims)
    logp = np.where(lp == -np.inf, -np.inf, ll + lp)
    return logp, lp, ll


# --------------------------------------------------------------------------------------------
# Built-in simulation models (the "user plugin" side, L0 in SURVEY section 1)
# --------------------------------------------------------------------------------------------
def sim_quadratic(x):
    """Examples/Example00/simulation_model_00.py:6-10, 19-24: y = (x-a)**2 + b."""
    x = np.asarray(x)

    def point(theta):
        return (x - theta[0]) ** 2 + theta[1]

    def batch(th):
        return (x[None, :] - th[:, 0:1]) ** 2 + th[:, 1:2]
    return point, batch


KB_KJ = 1.380649e-26 * 6.0221409e+23   # Examples/Example01/tprmodel.py:17


def tpd2site_constants(n=1024):
    """prescribed coverages, linear temperature ramp."""
    t = np.linspace(0, 400, n)
    T = 150 + 0.5 * t
    cov1 = 0.5 * np.exp(-t / 150)
    cov2 = 0.5 * np.exp(-t / 300)
    return KB_KJ * T, cov1, cov2


def sim_tpd2site(kBT, cov1, cov2):
    """Rate law of Examples/Example01/tprmodel.py:20 with prescribed coverages, summed over the
    two sites: sim_j = sum_i cov_ij*exp(-(Ea_i - kBT_j*logA_i - gamma_i*cov_ij)/kBT_j).
    theta = (Ea1, Ea2, logA1, logA2, gamma1, gamma2)."""
    def batch(th):
        th = np.atleast_2d(th)
        Ea1, Ea2, lA1, lA2, g1, g2 = (th[:, i:i + 1] for i in range(6))
        r1 = cov1 * np.exp(-(Ea1 - kBT * lA1 - g1 * cov1) / kBT)
        r2 = cov2 * np.exp(-(Ea2 - kBT * lA2 - g2 * cov2) / kBT)
        return r1 + r2

    def point(theta):
        return batch(np.asarray(theta, dtype=float)[None, :])[0]
    return point, batch


def sim_linear(Phi):
    """sim = Phi @ theta, accumulated parameter by parameter in index order (the kernels' order)."""
    Phi = np.asarray(Phi, dtype=float)

    def batch(th):
        th = np.atleast_2d(th)
        acc = Phi[None, :, 0] * th[:, 0:1]
        for d in range(1, Phi.shape[1]):
            acc = acc + Phi[None, :, d] * th[:, d:d + 1]
        return acc

    def point(theta):
        return batch(np.asarray(theta, dtype=float)[None, :])[0]
    return point, batch


KB_EV = 8.61733035E-5   # eV/K, Examples/Example13doeFunctionExample/simulation_functions_Langmuir_CO_H2O_four_parameters.py:31


def sim_langmuir_log(T, pA, offsets=(0.0, -25.0, 25.0)):
    """Langmuir_replacement_three_temperatures_log of the reference's design-of-experiments example
    (Examples/Example13doeFunctionExample/simulation_functions_Langmuir_CO_H2O_four_parameters.py:23-52):
    for each temperature T_j = T + offsets[j] (the reference's order: middle, low, high)
        delta_G = dH - T_j*dS;  K = exp(-delta_G/(kB*T_j));  theta_A = (K*pA)/(1 + K*pA);  sim_j = log10(theta_A).
    theta = (dH, dS).  T and pA are the independent variables the DOE loop modulates."""
    Tj = np.array([T + o for o in offsets], dtype=float)            # T, T-25, T+25 (T - 25 == T + (-25) in IEEE arithmetic)

    def batch(th):
        th = np.atleast_2d(np.asarray(th, dtype=float))
        dG = th[:, 0:1] - Tj[None, :] * th[:, 1:2]
        K = np.exp(-dG / (KB_EV * Tj[None, :]))
        return np.log10((K * pA) / (1 + K * pA))

    def point(theta):
        return batch(np.asarray(theta, dtype=float)[None, :])[0]
    return point, batch


def cos_basis(n, k):
    x = np.linspace(0, 1, n)
    return np.cos(np.pi * np.arange(k)[None, :] * x[:, None])


# --------------------------------------------------------------------------------------------
# Multistart: grid order, uniform points, best-of, permutations -> samples, statistics
# --------------------------------------------------------------------------------------------
def combination_generator(center, interval, radii):
    """CombinationGeneratorModule.py:40-81 ("Addition" spread), literal list construction."""
    combos = [np.asarray(center, dtype=float)]
    for p in range(len(center)):
        for base_i in range(len(combos)):
            base = combos[base_i]
            if radii[p] > 0:
                for i in range(1, radii[p] + 1):
                    lo = list(base)
                    lo[p] = base[p] - (interval[p] * i)      # :59
                    combos.append(tuple(lo))
                    hi = list(base)
                    hi[p] = base[p] + (interval[p] * i)      # :71
                    combos.append(tuple(hi))
    return combos


def grid_size(radii) -> int:
    n = 1
    for r in radii:
        n *= 2 * int(r) + 1
    return n


def grid_decode(center, interval, radii, idx) -> np.ndarray:
    """Closed-form element `idx` of combination_generator's list."""
    center = np.asarray(center, dtype=float)
    interval = np.asarray(interval, dtype=float)
    g = np.asarray(idx, dtype=np.int64).copy()
    k = len(center)
    out = np.empty(g.shape + (k,))
    M = [1]
    for d in range(k):
        M.append(M[-1] * (2 * int(radii[d]) + 1))
    for p in range(k - 1, -1, -1):
        r2 = 2 * int(radii[p])
        inner = g < M[p]
        gp = np.where(inner, 0, g - M[p])
        if r2 > 0:
            base = gp // r2
            j = gp % r2
        else:
            base, j = gp, np.zeros_like(gp)
        i = j // 2 + 1
        step = interval[p] * i.astype(float)                 # (BaseFactor * i) first, then +/-
        val = np.where(j % 2 == 0, center[p] - step, center[p] + step)
        out[..., p] = np.where(inner, center[p], val)
        g = np.where(inner, g, base)
    return out


def uniform_points(u01, spread, std_prior, center):
    """generateInitialPoints('uniform'), IP.py:706 + 749, from given U[0,1) draws.
    np.random.uniform(-2,2) == -2 + 4*u (numpy legacy: low + (high-low)*random_sample)."""
    first = -2.0 + 4.0 * np.asarray(u01)
    return spread * first * np.asarray(std_prior) + np.asarray(center)


def first_max(logp):
    """Best-of loop, IP.py:824, 933-938: strict '>' from -inf; NaN never selected."""
    best, idx = float('-inf'), -1
    for i, v in enumerate(logp):
        if v > best:
            best, idx = v, i
    return idx, best


def first_max_np(logp):
    lp = np.where(np.isnan(logp), -np.inf, logp)
    i = int(np.argmax(lp))
    return (i, float(lp[i])) if lp[i] > -np.inf else (-1, float('-inf'))


def permutation_weights(logp, max_logp, coef=2.0):
    """convertPermutationsToSamples, IP.py:3463-3491: copies = int(exp(logP - thr)) for logP > thr."""
    thr = max_logp + np.log(10 ** (-1 * coef))               # IP.py:1016, 3477
    with np.errstate(over='ignore', invalid='ignore'):
        reps = np.exp(np.asarray(logp) - thr)
    w = np.where(np.asarray(logp) > thr, np.floor(reps), 0.0)
    return w.astype(np.int64), thr


def convert_permutations_to_samples(logp_and_params, max_logp, coef=2.0):
    """The literal expansion (small inputs only)."""
    w, _ = permutation_weights(logp_and_params[:, 0], max_logp, coef)
    return np.repeat(logp_and_params, w, axis=0)


def weighted_statistics(logp, thetas, w, nbins=10):
    """calculatePostBurnInStatistics on the expanded array (IP.py:1923-1931) expressed with weights;
    histograms as plotting_functions.py:135 (ax.hist default: 10 equal bins over [min,max])."""
    w = np.asarray(w, dtype=float)
    keep = w > 0
    W = w.sum()
    mu = (w[:, None] * thetas).sum(axis=0) / W
    std = np.sqrt((w[:, None] * (thetas - mu) ** 2).sum(axis=0) / W)
    evidence = ((w * np.exp(logp)).sum() / W) / np.sqrt((w[:, None] * thetas ** 2).sum())
    hists, edges = [], []
    for d in range(thetas.shape[1]):
        lo, hi = thetas[keep, d].min(), thetas[keep, d].max()
        h, e = np.histogram(thetas[keep, d], bins=nbins, range=(lo, hi), weights=w[keep])
        hists.append(h)
        edges.append(e)
    return dict(W=W, mu_AP=mu, stdap=std, evidence=evidence, hist=np.array(hists), edges=np.array(edges))


# --------------------------------------------------------------------------------------------
# Metropolis-Hastings with injected draws (IP.py:2606-2675, 2715-2727)
# --------------------------------------------------------------------------------------------
def mh_chain_point(P: OracleProblem, theta0, deltas, us, simulate: Callable, modulate=0.0):
    """Literal restatement: re-evaluates the current point each step, one point at a
    time through scipy.  deltas[(L-1),k]: proposal increments; us[L-1]: accept uniforms."""
    L = len(us) + 1
    k = len(theta0)
    samples = np.zeros((L, k))
    logp = np.zeros((L, 1))
    lps = np.zeros((L, 1))
    lls = np.zeros((L, 1))
    acc = np.zeros(L, dtype=bool)
    samples[0] = theta0
    logp[0] = log_p_point(P, samples[0], simulate)                     # IP.py:2625
    for i in range(1, L):
        prop = samples[i - 1] + deltas[i - 1]                          # IP.py:2630
        lp_p = log_prior_point(P, prop)
        ll_p = log_likelihood_point(P, simulate(tuple(prop)))
        lp_c = log_prior_point(P, samples[i - 1])
        ll_c = log_likelihood_point(P, simulate(tuple(samples[i - 1])))
        with np.errstate(invalid='ignore'):
            la = (ll_p + lp_p) - (ll_c + lp_c)                         # IP.py:2641
        if modulate != 0:
            la = (1 / float(modulate)) * la                            # IP.py:2649
        with np.errstate(divide='ignore'):
            lu = np.log(us[i - 1])                                     # IP.py:2651
        if la > lu:                                                    # IP.py:2652
            samples[i], logp[i], lls[i], lps[i], acc[i] = prop, ll_p + lp_p, ll_p, lp_p, True
        else:
            samples[i], logp[i], lls[i], lps[i] = samples[i - 1], ll_c + lp_c, ll_c, lp_c
    return dict(samples=samples, logP=logp, log_priors=lps, log_likelihoods=lls, accepted=acc)


def mh_chains_batch(P: OracleProblem, theta0, deltas, us, simulate_batch: Callable, modulate=0.0):
    """C chains advanced together with cached current values (bit-identical to the literal loop
    because re-evaluating the unchanged current point reproduces the same doubles).
    theta0[C,k], deltas[L-1,C,k], us[L-1,C]."""
    theta = np.array(theta0, dtype=float)
    C, k = theta.shape
    L = us.shape[0] + 1
    samples = np.zeros((L, C, k))
    logp = np.zeros((L, C))
    lps = np.zeros((L, C))
    lls = np.zeros((L, C))
    acc = np.zeros((L, C), dtype=bool)
    cur_p, cur_lp, cur_ll = log_p_batch(P, theta, simulate_batch)
    samples[0], logp[0] = theta, cur_p
    cur_sum = cur_ll + cur_lp                       # MH uses ll+lp even when lp == -inf (Q5)
    for i in range(1, L):
        prop = theta + deltas[i - 1]
        _, lp_p, ll_p = log_p_batch(P, prop, simulate_batch)
        with np.errstate(invalid='ignore', divide='ignore'):
            la = (ll_p + lp_p) - cur_sum
            if modulate != 0:
                la = (1 / float(modulate)) * la
            a = la > np.log(us[i - 1])
        theta = np.where(a[:, None], prop, theta)
        cur_lp = np.where(a, lp_p, cur_lp)
        cur_ll = np.where(a, ll_p, cur_ll)
        cur_sum = np.where(a, ll_p + lp_p, cur_sum)
        samples[i], logp[i], lps[i], lls[i], acc[i] = theta, cur_sum, cur_lp, cur_ll, a
    return dict(samples=samples, logP=logp, log_priors=lps, log_likelihoods=lls, accepted=acc)


def mh_chain_map_finding(P: OracleProblem, theta0, deltas, mods, us, simulate_batch: Callable, modulate=0.0, scaled=True):
    """doMetropolisHastings with mcmc_mode 'MAP_finding' (IP.py:2631-2635, 2688-2713) for ONE chain fed its draws:
    mods[i] = np.random.uniform() + 0.5 of step i+1, us[i] its accept uniform, deltas[i] the proposal increment -- already
    multiplied by sqrt(dyn * mod) (scaled=True: what np.random.multivariate_normal(0, Q*dyn*mod*step) returned / what the
    kernel exports) or the unbiased increment L z, which is then scaled here.  Every 100 steps the slope of a straight-line
    fit (np.polynomial.polynomial.polyfit, as the reference) of the previous hundred proposals' log-posteriors against
    their mods adapts dyn: x 1.05 while < 10 if positive, x 0.95 while > 0.1 if negative.  The history arrays start as
    zeros, so the first window holds the entry (0, 0) of index 0 (IP.py:2608, 2616).
    Returns samples[L,k], logP[L], dyn[L] (the coefficient used by each step; dyn[0] unused) and the final coefficient."""
    th = np.array(theta0, dtype=float).reshape(1, -1)
    L = len(us) + 1
    k = th.shape[1]
    samples = np.zeros((L, k))
    logp = np.zeros(L)
    drawn = np.zeros(L)
    hist = np.zeros(L)
    dyn_used = np.ones(L)
    cur_p, cur_lp, cur_ll = log_p_batch(P, th, simulate_batch)
    samples[0], logp[0] = th[0], cur_p[0]
    cur_sum = float(cur_ll[0] + cur_lp[0])
    dyn = 1
    for i in range(1, L):
        m = mods[i - 1]
        hist[i] = m
        dyn_used[i] = dyn
        d = np.asarray(deltas[i - 1], dtype=float)
        if not scaled:
            d = np.sqrt(dyn * m) * d
        prop = th + d
        _, lp_p, ll_p = log_p_batch(P, prop, simulate_batch)
        sum_p = float(ll_p[0] + lp_p[0])
        with np.errstate(invalid='ignore', divide='ignore'):
            la = sum_p - cur_sum
            if modulate != 0:
                la = (1 / float(modulate)) * la
            acc = la > np.log(us[i - 1])
        drawn[i] = sum_p
        if acc:
            th, cur_sum = prop, sum_p
        samples[i], logp[i] = th[0], cur_sum
        if i % 100 == 0:                                             # IP.py:2688-2713
            y, x = drawn[i - 100:i], hist[i - 100:i]
            ok = np.isfinite(y)
            fit = np.polynomial.polynomial.polyfit(x[ok], y[ok], 1)
            if fit[1] > 0:
                if dyn < 10:
                    dyn = dyn * 1.05
            if fit[1] < 0:
                if dyn > 0.1:
                    dyn = dyn * 0.95
    return dict(samples=samples, logP=logp, dyn=dyn_used, dyn_final=dyn, drawn=drawn)


def post_burn_in_statistics(samples, logp):
    """IP.py:1923-1931 without the tail filter."""
    mu = np.mean(samples, axis=0)
    sd = np.std(samples, axis=0)
    lp = np.asarray(logp).reshape(-1)
    m = max(lp)
    idx = list(lp).index(m)
    evidence = np.mean(np.exp(lp)) / np.linalg.norm(samples)
    return dict(mu_AP=mu, stdap=sd, map_logP=m, map_index=idx, map_parameter_set=samples[idx],
                evidence=evidence)


def array_threshold_filter(merged, key, threshold, remove='below'):
    """arrayThresholdFilter (IP.py:3635-3664): rows whose key is beyond the threshold become NaN rows and are dropped
    (comparisons with a NaN key or a NaN threshold are False, so such rows stay)."""
    key = np.asarray(key).reshape(-1)
    with np.errstate(invalid='ignore'):
        bad = key > threshold if remove == 'above' else key < threshold
    marked = merged * 1.0
    marked[bad] = np.nan
    return marked[~np.all(np.isnan(marked), axis=1)]


def info_gain_kl(prior_samples, post_samples, bins=100):
    """calculateInfoGain, KL form (IP.py:1772-1791).  plt.hist([a, b], bins=100, density=True) bins both data sets on
    the common edges np.histogram_bin_edges(concatenate([a, b]), 100) and normalises each separately
    (matplotlib Axes.hist; matplotlib itself is not installed here, so this restates its documented behaviour:
    parity unpinned at that one call).  Returns (total, per-parameter list)."""
    each = []
    total = 0
    for d in range(post_samples.shape[1]):
        a, b = prior_samples[:, d].flatten(), post_samples[:, d].flatten()
        edges = np.histogram_bin_edges(np.concatenate([a, b]), bins)
        d0, _ = np.histogram(a, edges, density=True)
        d1, _ = np.histogram(b, edges, density=True)
        with np.errstate(divide='ignore', invalid='ignore'):
            cur = d1 * np.log(d1 / d0)                                    # IP.py:1785
        cur = np.sum(cur[np.isfinite(cur)])                               # IP.py:1786-1787
        each.append(cur)
        total = total + cur
    return total, each


def info_gain_log_ratio_cutoff(logp, logprior, evidence, cutoff):
    """The truncated log-ratio information gain of calculateInfoGain for mcmc_info_gain_cutoff != 0 (IP.py:1805-1843): rows
    whose log-prior or log-posterior is NaN or below ln(cutoff) are removed, the rest contribute
    ln(exp(logP) / evidence) - logPrior (non-finite terms count as 0); the mean of no rows is NaN."""
    logp = np.asarray(logp, dtype=float).reshape(-1)
    logprior = np.asarray(logprior, dtype=float).reshape(-1)
    stacked = np.vstack((logprior, logp))                                  # IP.py:1809
    with np.errstate(invalid='ignore', divide='ignore'):
        remove = np.isnan(stacked).any(axis=0) | (stacked < np.log(cutoff)).any(axis=0)      # IP.py:1818-1826
        kept = stacked[:, ~remove]
        post = np.log(np.exp(kept[1]) / evidence)                          # IP.py:1834
        ratios = post - kept[0]
        ratios[np.isinf(ratios)] = 0
        ratios = np.nan_to_num(ratios)
        return np.mean(ratios) if ratios.size else np.nan


def post_burn_in_statistics_filtered(samples, logp, logprior, benchmark='auto', coef=2.0, filter_samples=True,
                                     mu_ap_logp_fn=None, prior_samples=None, info_gain_cutoff=0):
    """calculatePostBurnInStatistics (IP.py:1861-1937) including the tail filter and calculateInfoGain's log-ratio form
    (IP.py:1793-1803).  logp, logprior: (N,1).  Returns a dict with the filtered arrays and the statistics."""
    samples = np.asarray(samples, dtype=float)
    logp = np.asarray(logp, dtype=float).reshape(-1, 1)
    logprior = np.asarray(logprior, dtype=float).reshape(-1, 1)
    thr = None
    if filter_samples:
        merged = np.hstack((logp, logprior, samples))                    # IP.py:1887
        with np.errstate(invalid='ignore', divide='ignore'):
            log_neg = np.log(-1 * logp)                                   # IP.py:1894
        if benchmark == 'mu_AP':
            thr = mu_ap_logp_fn(np.mean(samples, axis=0)) - coef          # IP.py:1899-1904
            merged = array_threshold_filter(merged, logp, thr, 'below')
        elif benchmark == 'MAP':
            thr = np.max(logp) - coef                                     # IP.py:1897, 1910
            merged = array_threshold_filter(merged, logp, thr, 'below')
        else:
            thr = np.mean(log_neg) + coef * np.std(log_neg)               # IP.py:1895-1896, 1915
            merged = array_threshold_filter(merged, log_neg, thr, 'above')
        logp = np.atleast_2d(merged[:, 0]).transpose()
        logprior = np.atleast_2d(merged[:, 1]).transpose()
        samples = merged[:, 2:]
    st = post_burn_in_statistics(samples, logp)
    post = logp - np.log(st['evidence'])                                  # IP.py:1799
    with np.errstate(invalid='ignore'):
        ratios = post - logprior
    ratios[np.isinf(ratios)] = 0
    ratios = np.nan_to_num(ratios)
    st.update(samples=samples, logP=logp, log_priors=logprior, threshold=thr, info_gain_log_ratio=np.mean(ratios))
    if info_gain_cutoff != 0:
        st['info_gain_log_ratio'] = info_gain_log_ratio_cutoff(logp, logprior, st['evidence'], info_gain_cutoff)
    if prior_samples is not None:
        st['info_gain_KL'], st['info_gain_KL_each_parameter'] = info_gain_kl(np.asarray(prior_samples, dtype=float), samples)
    return st


# --------------------------------------------------------------------------------------------
# Design of experiments: information gain over a grid of conditions (IP.py:1326-1354, 1363-1572)
# --------------------------------------------------------------------------------------------
def doe_conditions(center, interval, num_intervals, array_format='xyz'):
    """The conditions visited by doeGetInfoGainMatrix, in its loop order: 'xyz' walks getGridPermutations
    (IP.py:1382, 755-770: the combination generator's order); 'meshgrid' walks np.linspace axes with the SECOND
    variable in the outer loop (IP.py:1426-1438)."""
    center = np.asarray(center, dtype=float)
    if array_format == 'xyz':
        return np.array(combination_generator(center, np.asarray(interval, dtype=float), np.asarray(num_intervals, dtype=int)), dtype=float)
    lo = [center[i] - interval[i] * num_intervals[i] for i in range(2)]
    hi = [center[i] + interval[i] * num_intervals[i] for i in range(2)]
    ax = [np.linspace(lo[i], hi[i], num_intervals[i] * 2 + 1) for i in range(2)]
    return np.array([[v1, v2] for v2 in ax[1] for v1 in ax[0]], dtype=float)


def doe_info_gain_matrix(build_problem_for, simulate_for, conditions, theta_true, theta0, deltas, us, burn_in,
                         filter_kwargs=None, prior_samples=None, info_gain_returned='log_ratio'):
    """doeGetInfoGainMatrix (IP.py:1363-1472) for one "true" parameter vector: for every condition c the observed
    responses are replaced by the simulation at theta_true (populateResponsesWithSyntheticData, IP.py:1326-1354; the
    observed covariance keeps the values of the first __init__ because the synthetic uncertainties are None, IP.py:342-346),
    one Metropolis-Hastings chain is run and its post-burn-in information gain is appended to the condition values.
    build_problem_for(y) -> OracleProblem with observed responses y; simulate_for(c) -> (point, batch) simulators under
    condition c; deltas[L-1,k] / us[L-1]: the draws, the SAME for every condition (the reference reseeds numpy in every
    __init__, IP.py:191-193).  Returns (matrix[n_cond, n_var+1], list of per-condition statistics dicts)."""
    rows, stats = [], []
    fk = dict(filter_samples=False) if filter_kwargs is None else dict(filter_kwargs)
    for c in np.asarray(conditions, dtype=float):
        point, batch = simulate_for(c)
        y = point(np.asarray(theta_true, dtype=float))                                    # IP.py:1332, 1351
        P = build_problem_for(y)
        run = mh_chains_batch(P, np.asarray(theta0, dtype=float)[None, :], deltas[:, None, :], us[:, None], batch)
        b = int(burn_in)
        st = post_burn_in_statistics_filtered(run['samples'][b:, 0, :], run['logP'][b:, 0], run['log_priors'][b:, 0],
                                              prior_samples=prior_samples, **fk)
        ig = st['info_gain_KL'] if info_gain_returned == 'KL_divergence' else st['info_gain_log_ratio']
        rows.append(np.hstack((c, ig)))                                                   # IP.py:1412-1413
        stats.append(st)
    return np.array(rows), stats


# --------------------------------------------------------------------------------------------
# Convergence diagnostics: Geweke z-scores (IP.py:3960-3975, 4012-4062)
# --------------------------------------------------------------------------------------------
def yule_walker(x, order=2):
    """statsmodels.regression.linear_model.yule_walker(x, order, method='adjusted', demean=True) -- a third-party
    dependency (statsmodels, unpinned in the reference's setup.py extras, imported at IP.py:4044) that is not installed
    here: restated from its published algorithm (parity unpinned at this one call): demean, r_0 = sum x^2 / n,
    r_k = sum x_t x_{t+k} / (n - k), solve the order x order Toeplitz system, sigma^2 = r_0 - r_{1..}.rho."""
    x = np.array(x, dtype=np.float64)
    x -= x.mean()
    n = x.shape[0]
    r = np.zeros(order + 1, np.float64)
    r[0] = (x ** 2).sum() / n
    for k in range(1, order + 1):
        r[k] = (x[0:-k] * x[k:]).sum() / (n - k)
    R = scipy.linalg.toeplitz(r[:-1])
    rho = np.linalg.solve(R, r[1:])
    sigmasq = r[0] - (r[1:] * rho).sum()
    sigma = np.sqrt(sigmasq) if sigmasq > 0 else np.nan
    return rho, sigma


def geweke_slices(n, initial_window=0.1, comparison_window=0.5, intervals=20):
    """The (start, first slice, last slice) index triples of geweke_diagnostic for a chain of n rows (IP.py:4030-4052)."""
    last_index = n - 1
    last_start_index = (1 - comparison_window) * last_index
    out = []
    for start in np.linspace(0, last_start_index, num=intervals, endpoint=True, dtype=int):
        out.append((int(start), (int(start), int(start) + int(initial_window * (last_index - start))),
                    (int(last_index - comparison_window * (last_index - start)), n)))
    return out


def geweke_diagnostic(x, initial_window=0.1, comparison_window=0.5, intervals=20):
    """geweke_diagnostic (IP.py:4012-4062): [(start, z)] with z = (mean_first - mean_last) / sqrt(sv_first/n_first + sv_last/n_last),
    sv = sigma^2 / (1 - sum(beta))^2 of an order-2 Yule-Walker fit (the spectral density at frequency zero)."""
    x = np.asarray(x, dtype=float)
    rows = []
    for start, (a1, b1), (a2, b2) in geweke_slices(len(x), initial_window, comparison_window, intervals):
        first, last = x[a1:b1], x[a2:b2]

        def spec_var(v):
            beta, sigma = yule_walker(v, 2)
            return sigma ** 2 / (1. - np.sum(beta)) ** 2
        z = first.mean() - last.mean()
        z /= np.sqrt(spec_var(first) / len(first) + spec_var(last) / len(last))
        rows.append([start, z])
    return np.array(rows)


def geweke_convergence(chains):
    """The Geweke half of calculateAndPlotConvergenceDiagnostics (IP.py:3960-4005) for chains[R, C, k]: z[k, C, 20, 20] over
    the 20 growing windows, and the reference's combined outputs (window indices, final combined values, percent outliers)."""
    R, C, k = chains.shape
    windows = np.linspace(0, R, 21).astype(int)[1:]
    z = np.empty((k, C, len(windows), 20))
    for d in range(k):
        for c in range(C):
            for w, n in enumerate(windows):
                z[d, c, w] = geweke_diagnostic(chains[:n, c, d])[:, 1]
    comb = np.mean(np.mean(np.abs(z), axis=0), axis=0)                    # over parameters, then chains
    final = comb[:, -1]
    outlier = np.count_nonzero(comb > 1, axis=1) / comb.shape[1]
    return dict(z=z, window_indices=windows, final_combined_parameter_values=final, final_combined_parameter_percent_outlier=outlier)


# --------------------------------------------------------------------------------------------
# Convergence diagnostics: integrated autocorrelation time (IP.py:3914-3948) -- PARITY UNPINNED
# --------------------------------------------------------------------------------------------
def autocorr_func_1d(x, norm=True):
    """zeus.autocorr._autocorr_func_1d (zeus-mcmc, pinned '>=2.4.0' only in the reference's setup.py extras, imported at
    IP.py:3916; neither vendored by the reference nor installed here, so this is a restatement of the published routine --
    the same one as in emcee's autocorrelation tutorial -- and parity is UNPINNED): autocorrelation function of a 1-D series
    by FFT, zero-padded to twice the next power of two."""
    x = np.atleast_1d(x)
    if len(x.shape) != 1:
        raise ValueError("invalid dimensions for 1D autocorrelation function")
    n = 1
    while n < len(x):
        n <<= 1
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[:len(x)].real
    acf /= 4 * n
    if norm:
        acf /= acf[0]
    return acf


def auto_window(taus, c):
    """Sokal's automated window: the first lag M with M >= c * tau(M).  np.argmin of a boolean vector is the first False --
    and 0 when every entry is True (a chain too short for its window then reports tau = taus[0] = 1)."""
    m = np.arange(len(taus)) < c * taus
    if np.any(m):
        return np.argmin(m)
    return len(taus) - 1


def autocorr_time(samples, c=5.0):
    """zeus.autocorr.AutoCorrTime(samples[nsteps, nwalkers, ndim], c=5.0, method='mk') -> tau[ndim] (PARITY UNPINNED, see
    autocorr_func_1d): per parameter the walkers' chains are laid end to end (walker-major), the normalised autocorrelation
    function of that one series is truncated to the first nsteps lags, tau(M) = 2 * cumsum - 1 at the automated window."""
    samples = np.asarray(samples, dtype=float)
    nsteps, _, ndim = samples.shape
    taus = np.empty(ndim)
    for d in range(ndim):
        y = samples[:, :, d].T                                             # (walkers, steps)
        f = autocorr_func_1d(y.reshape((-1), order='C'))[:nsteps]
        t = 2.0 * np.cumsum(f) - 1.0
        taus[d] = t[auto_window(t, c)]
    return taus


def autocorr_time_convergence(chains):
    """The autocorrelation half of calculateAndPlotConvergenceDiagnostics (IP.py:3918-3948) for chains[R, C, k]: taus[20, k]
    over windows growing on a log scale, the final row, and the product over parameters per window (IP.py:3936-3944)."""
    with np.errstate(all='ignore'):
        windows = np.exp(np.linspace(0, np.log(chains.shape[0]), 21)).astype(int)[1:]          # IP.py:3918
        taus = np.empty((len(windows), chains.shape[2]))
        for i, n in enumerate(windows):
            taus[i] = autocorr_time(chains[:n, :, :])                                          # IP.py:3930
    return dict(window_indices=windows, taus=taus, final_parameter_values=taus[-1, :], combined=np.prod(taus, axis=1))


# --------------------------------------------------------------------------------------------
# Ensemble slice sampling -- published algorithm (parity unpinned, see module header)
# --------------------------------------------------------------------------------------------
def ess_run(logp_batch: Callable, start, nsteps, rng: np.random.Generator, mu0=1.0, maxiter=10000,
            tune=True, tolerance=0.05, patience=5, maxsteps=10000):
    """Karamanis & Beutler 2020 (arXiv:2002.06212) Algorithms 2-4 with the differential move, as
    implemented by zeus.EnsembleSampler.run_mcmc which the reference calls at IP.py:2495-2498.
    logp_batch(X[B,k]) -> logp[B].  Returns chain[nsteps,W,k], logp[nsteps,W], neval."""
    X = np.array(start, dtype=float)
    W, k = X.shape
    Z = logp_batch(X)
    if not np.all(np.isfinite(Z)):
        raise ValueError("Invalid walker initial positions! Initialise walkers from positions of finite log probability.")
    chain = np.empty((nsteps, W, k))
    chain_lp = np.empty((nsteps, W))
    mu = mu0
    neval = W
    tuning = tune
    n_tuned = 0
    half = W // 2
    for it in range(nsteps):
        perm = rng.permutation(W)
        sets = [perm[:half], perm[half:]]
        nexp = ncon = 0
        for s in range(2):
            active, inactive = sets[s], sets[1 - s]
            na = len(active)
            # differential move: direction from two distinct complementary walkers
            pairs = np.array([rng.choice(len(inactive), 2, replace=False) for _ in range(na)])
            directions = 2.0 * mu * (X[inactive[pairs[:, 0]]] - X[inactive[pairs[:, 1]]])
            Z0 = Z[active] - rng.exponential(size=na)
            Lb = -rng.uniform(size=na)
            Rb = Lb + 1.0
            Jl = np.floor(maxsteps * rng.uniform(size=na)).astype(int)
            Kr = (maxsteps - 1) - Jl
            X0 = X[active]
            # stepping out, left
            mask = np.ones(na, dtype=bool)
            cnt = 0
            while mask.any():
                mask = mask & (Jl >= 1)
                if not mask.any():
                    break
                zl = logp_batch(X0[mask] + Lb[mask, None] * directions[mask])
                neval += int(mask.sum())
                out = zl < Z0[mask]
                idx = np.where(mask)[0]
                Lb[idx[~out]] -= 1.0
                Jl[idx[~out]] -= 1
                nexp += int((~out).sum())
                mask[idx[out]] = False
                cnt += 1
                if cnt > maxiter:
                    raise RuntimeError("Number of expansions exceeded maximum limit! maxiter")
            mask = np.ones(na, dtype=bool)
            cnt = 0
            while mask.any():
                mask = mask & (Kr >= 1)
                if not mask.any():
                    break
                zr = logp_batch(X0[mask] + Rb[mask, None] * directions[mask])
                neval += int(mask.sum())
                out = zr < Z0[mask]
                idx = np.where(mask)[0]
                Rb[idx[~out]] += 1.0
                Kr[idx[~out]] -= 1
                nexp += int((~out).sum())
                mask[idx[out]] = False
                cnt += 1
                if cnt > maxiter:
                    raise RuntimeError("Number of expansions exceeded maximum limit! maxiter")
            # shrinking
            mask = np.ones(na, dtype=bool)
            Wd = np.zeros(na)
            Xn = X0.copy()
            Zn = Z[active].copy()
            cnt = 0
            while mask.any():
                idx = np.where(mask)[0]
                Wd[idx] = Lb[idx] + rng.uniform(size=len(idx)) * (Rb[idx] - Lb[idx])
                cand = X0[idx] + Wd[idx, None] * directions[idx]
                zc = logp_batch(cand)
                neval += len(idx)
                ok = zc >= Z0[idx]
                Xn[idx[ok]] = cand[ok]
                Zn[idx[ok]] = zc[ok]
                mask[idx[ok]] = False
                rej = idx[~ok]
                neg = Wd[rej] < 0
                Lb[rej[neg]] = Wd[rej[neg]]
                Rb[rej[~neg]] = Wd[rej[~neg]]
                ncon += len(rej)
                cnt += 1
                if cnt > maxiter:
                    raise RuntimeError("Number of contractions exceeded maximum limit! maxiter")
            X[active] = Xn
            Z[active] = Zn
        if tuning:
            nexp_eff = max(1, nexp)
            mu = mu * 2.0 * nexp_eff / (nexp_eff + ncon)
            if abs(nexp_eff / (nexp_eff + ncon) - 0.5) < tolerance:
                n_tuned += 1
            else:
                n_tuned = 0
            if n_tuned >= patience:
                tuning = False
        chain[it] = X
        chain_lp[it] = Z
    return chain, chain_lp, neval, mu


def ess_run_philox(logp_batch: Callable, start, nsteps, seed, mu0=1.0, maxiter=10000, tune=True, tolerance=0.05,
                   patience=5, maxsteps=10000, first_iter=0):
    """The same algorithm as ess_run (zeus.EnsembleSampler.run_mcmc with the differential move, called by the reference
    at IP.py:2495-2498; Karamanis & Beutler, arXiv:2002.06212), drawing from the counter-based Philox generator in the
    layout of peuqse_b200/csrc/ess_common.cuh, walker by walker, so that a device run can be replayed step for step:
        counter index = ((first_iter + it) * 2 + half) << 32 | walker, stream STREAM_ESS
        block 0xffffffff (half 0)  words (0,1): shuffle key -- the walkers are ordered by key (ties by index)
        block 0   u_a -> first complementary walker floor(u_a * H), u_b -> second floor(u_b * (H-1)), skipping the first
        block 1   u_a -> Exp(1) = -log(1 - u_a) of the slice level, u_b -> left end L = -u_b, R = L + 1
        block 2   u_a -> J = floor(maxsteps * u_a), K = maxsteps - 1 - J
        block 3 + r // 2   shrink round r: u_a for even r, u_b for odd r
    Comparisons follow zeus: expand / accept iff Z0 < Z (strict).  Returns chain[nsteps,W,k], logp[nsteps,W], info dict."""
    from . import philox_np as px
    X = np.array(start, dtype=float)
    W, k = X.shape
    H = W // 2
    Z = np.asarray(logp_batch(X), dtype=float).copy()
    if not np.all(np.isfinite(Z)):
        raise ValueError("Invalid walker initial positions! Initialise walkers from positions of finite log probability.")
    chain = np.empty((nsteps, W, k))
    chain_lp = np.empty((nsteps, W))
    mu, tuning, ncount = float(mu0), bool(tune), 0
    neval, tot_exp, tot_con = 0, 0, 0
    walkers = np.arange(W, dtype=np.uint64)
    for it in range(nsteps):
        base = np.uint64((first_iter + it) * 2)
        r = px.raw_block(seed, px.STREAM_ESS, (base << np.uint64(32)) | walkers, 0xffffffff)
        keys = (r[:, 0].astype(np.uint64) << np.uint64(32)) | r[:, 1].astype(np.uint64)
        perm = np.lexsort((np.arange(W), keys))                       # rank by key, ties by walker index
        nexp = ncon = 0
        for half in range(2):
            act, comp = perm[half * H:(half + 1) * H], perm[(1 - half) * H:(2 - half) * H]
            idx = ((base + np.uint64(half)) << np.uint64(32)) | act.astype(np.uint64)
            ua, ub = px.uniform_pair(seed, px.STREAM_ESS, idx, 0)
            i1 = np.minimum((ua * float(H)).astype(np.int64), H - 1)
            i2 = np.minimum((ub * float(H - 1)).astype(np.int64), H - 2)
            i2 = i2 + (i2 >= i1)
            dirs = (2.0 * mu) * (X[comp[i1]] - X[comp[i2]])
            ua, ub = px.uniform_pair(seed, px.STREAM_ESS, idx, 1)
            Z0 = Z[act] - (-np.log(1.0 - ua))
            Lb = -ub
            Rb = Lb + 1.0
            ua, _ = px.uniform_pair(seed, px.STREAM_ESS, idx, 2)
            J = np.floor(float(maxsteps) * ua)
            K = (float(maxsteps) - 1.0) - J
            X0 = X[act].copy()
            for bound, budget, step in ((Lb, J, -1.0), (Rb, K, 1.0)):   # stepping out: left, then right
                mask = budget >= 1.0
                rounds = 0
                while mask.any():
                    ii = np.where(mask)[0]
                    z = np.asarray(logp_batch(X0[ii] + bound[ii, None] * dirs[ii]), dtype=float)
                    neval += len(ii)
                    grow = Z0[ii] < z
                    g = ii[grow]
                    bound[g] += step
                    budget[g] -= 1.0
                    nexp += len(g)
                    mask[ii[~grow]] = False
                    mask &= budget >= 1.0
                    if grow.any() and rounds > maxiter:
                        raise RuntimeError("Number of expansions exceeded maximum limit! Increase mcmc_maxiter.")
                    rounds += 1
            mask = np.ones(H, dtype=bool)
            Xn, Zn = X0.copy(), Z[act].copy()
            r_sh = 0
            while mask.any():                                         # shrinking
                ii = np.where(mask)[0]
                ua, ub = px.uniform_pair(seed, px.STREAM_ESS, idx[ii], 3 + (r_sh >> 1))
                u = ub if (r_sh & 1) else ua
                wd = Lb[ii] + u * (Rb[ii] - Lb[ii])
                cand = X0[ii] + wd[:, None] * dirs[ii]
                z = np.asarray(logp_batch(cand), dtype=float)
                neval += len(ii)
                ok = Z0[ii] < z
                Xn[ii[ok]], Zn[ii[ok]] = cand[ok], z[ok]
                mask[ii[ok]] = False
                rej, wr = ii[~ok], wd[~ok]
                Lb[rej[wr < 0]] = wr[wr < 0]
                Rb[rej[wr > 0]] = wr[wr > 0]
                ncon += len(rej)
                if len(rej) and r_sh > maxiter:
                    raise RuntimeError("Number of contractions exceeded maximum limit! Increase mcmc_maxiter.")
                r_sh += 1
            X[act], Z[act] = Xn, Zn
        tot_exp += nexp
        tot_con += ncon
        if tuning:
            ne = max(1, nexp)
            frac = ne / (ne + ncon)
            mu *= 2.0 * frac
            if abs(frac - 0.5) < tolerance:
                ncount += 1
            if ncount > patience:
                tuning = False
        chain[it], chain_lp[it] = X, Z
    return chain, chain_lp, dict(neval=neval, mu=mu, n_expansions=tot_exp, n_contractions=tot_con)


def linear_gaussian_posterior(Phi, Sigma_inv_Phi, y, prior_var):
    """Closed form for y ~ N(Phi theta, Sigma), theta ~ N(0, prior_var I): (mean, cov)."""
    A = Phi.T @ Sigma_inv_Phi + np.eye(Phi.shape[1]) / prior_var
    cov = np.linalg.inv(A)
    return cov @ (Sigma_inv_Phi.T @ y), cov
