"""Host-side normalisation: UserInput -> constants of the hot path -> device handle.

Restates what the reference's parameter_estimation.__init__ derives (PEUQSE/InverseProblem.py, "IP.py",
lines 146-350) with the reference's own formulas so that the constants are bit-identical, and hoists the
eigendecomposition scipy repeats inside every multivariate_normal.logpdf call
(scipy/stats/_multivariate.py:208-236, _PSD) to this one-time set-up.  Runs once per problem on the host.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import scipy.linalg

from . import _lib
from .models import BuiltinModel

LOG2PI = float(np.log(2 * np.pi))     # scipy/stats/_multivariate.py:42


@dataclass
class PSD:
    rank: int
    log_pdet: float
    U: np.ndarray
    is_perm: bool
    perm: Optional[np.ndarray] = None
    w: Optional[np.ndarray] = None


def psd(cov) -> PSD:
    """scipy's _PSD(cov, allow_singular=False): eigh(lower) -> U = u*sqrt(1/s), log_pdet = sum(log s)."""
    cov = np.asarray(cov, dtype=float)
    if cov.ndim == 0:
        cov = cov.reshape(1, 1)
    elif cov.ndim == 1:
        cov = np.diag(cov)
    s, u = scipy.linalg.eigh(cov, lower=True, check_finite=True)
    eps = 1e6 * np.finfo('d').eps * np.max(np.abs(s))
    if np.min(s) < -eps:
        raise ValueError("The input matrix must be symmetric positive semidefinite.")
    if np.any(s <= eps):
        raise np.linalg.LinAlgError("When `allow_singular is False`, the input matrix must be symmetric positive definite.")
    U = np.multiply(u, np.sqrt(1.0 / s))
    out = PSD(rank=len(s), log_pdet=float(np.sum(np.log(s))), U=U, is_perm=False)
    nz = U != 0
    if np.all(nz.sum(axis=0) == 1):
        out.is_perm = True
        out.perm = np.argmax(nz, axis=0).astype(np.int64)
        out.w = np.abs(U[out.perm, np.arange(U.shape[1])])
    return out


def psd_cholesky(cov) -> PSD:
    """Triangular whitening of a dense covariance (responses['responses_observed_covmat_whitening'] = 'cholesky'): with
    cov = L L^T, U = (L^-1)^T is upper triangular, ||dev @ U||^2 = dev^T cov^-1 dev and log det = 2 sum log L_ii.  The
    contraction kernel skips the zero half of U (n^2 instead of scipy's 2 n^2 flops per point); the value
    agrees with scipy's eigen-whitening to ~cond(cov) * eps.  Falls back to scipy's form when cov is not positive definite
    enough for a Cholesky factorisation (so the error conventions of _PSD are kept)."""
    cov = np.asarray(cov, dtype=float)
    try:
        Lc = scipy.linalg.cholesky(cov, lower=True, check_finite=True)
    except np.linalg.LinAlgError:
        return psd(cov)
    n = cov.shape[0]
    Linv = scipy.linalg.solve_triangular(Lc, np.eye(n), lower=True)
    U = np.ascontiguousarray(np.triu(Linv.T))
    return PSD(rank=n, log_pdet=float(2.0 * np.sum(np.log(np.diagonal(Lc)))), U=U, is_perm=False)


def nest_responses(x):
    """makeAtLeast_2dNested for responses: 1-D -> [x]; list of sequences -> list of 1-D arrays."""
    try:
        a = np.asarray(x, dtype=float)
        if a.ndim == 1:
            return [a]
        if a.ndim == 2:
            return [np.array(r, dtype=float) for r in a]
    except (ValueError, TypeError):
        pass
    return [np.asarray(r, dtype=float) for r in x]


def nest_uncertainties(x, nresp):
    a = None
    try:
        a = np.asarray(x, dtype=float)
    except (ValueError, TypeError):
        pass
    if nresp == 1:
        if a is None:
            raise ValueError("single-response uncertainties must be a regular array")
        if a.ndim == 1:
            return [a]
        if a.ndim == 2 and a.shape[0] == 1:
            return [a[0]]
        return [a]
    if a is not None and a.ndim >= 2 and a.shape[0] == nresp:
        return [np.array(r, dtype=float) for r in a]
    return [np.asarray(r, dtype=float) for r in x]


def shaped_response_cov(unc_list, r):
    """returnShapedResponseCovMat (IP.py:3591-3606), including its multi-response behaviour: with two or
    more responses an entry is squared only when it is the last response and holds a single value, so a
    vector of n > 1 sigmas is used as if it held variances (kept for parity with the reference)."""
    u = np.asarray(unc_list[r], dtype=float)
    if len(unc_list) == 1:
        return np.square(u) if u.ndim == 1 else u
    if r == len(unc_list) - 1 and u.shape[0] == 1:
        return np.square(u)
    return u


def transient_passes(response_type, data_type):
    """How often transform_responses integrates a response of data_overcategory 'transient_kinetics' (IP.py:620-657): an intermediate
    given as a concentration once and as a rate twice, a reactant or product given as a rate once, everything else not at all."""
    if response_type == 'I':
        return {'c': 1, 'r': 2}.get(data_type, 0)
    if response_type in ('R', 'P'):
        return 1 if data_type == 'r' else 0
    return 0


def transform_transient(abscissa, types, data_types, responses, uncertainties):
    """The observed side of transform_responses (IP.py:609-658), done once at construction: every pass replaces a response by its
    explicit-Euler integral over the abscissa starting at 0 (littleEulerGivenArray, IP.py:3501-3513) and propagates its standard
    deviations alongside (littleEulerUncertaintyPropagation, IP.py:3516-3525; the first point takes a tenth of the mean of the
    UNtransformed uncertainties).  The simulated side is the kernel behind peuqse_transform_transient."""
    resp = [np.array(r, dtype=float) for r in nest_responses(responses)]
    ab = nest_responses(abscissa)
    unc = [np.array(u, dtype=float) for u in nest_uncertainties(uncertainties, len(resp))]
    for r in range(len(resp)):
        if unc[r].ndim != 1:
            raise NotImplementedError("data_overcategory 'transient_kinetics' needs standard deviations per value, not a covariance matrix")
        t = ab[0] if len(ab) == 1 else ab[r]
        base = np.mean(unc[r]) / 10
        for _ in range(transient_passes(types[r], data_types[r])):
            y, u = np.zeros(len(t)), unc[r] * 0.0
            u[0] = base
            for j in range(len(t) - 1):
                dt = t[j + 1] - t[j]
                y[j + 1] = y[j] + resp[r][j] * dt
                u[j + 1] = ((u[j]) ** 2 + (unc[r][j] * dt) ** 2) ** 0.5
            if base == 0:
                u[0] = u[1]
            resp[r], unc[r] = y, u
    return resp, unc


@dataclass
class HostProblem:
    """Everything the kernels need, in host memory, plus the reference-named derived quantities."""
    k: int
    mu_prior: np.ndarray
    std_prior: np.ndarray
    var_prior: np.ndarray
    covmat_prior: np.ndarray
    scaling: Optional[np.ndarray]
    mu_prior_scaled: np.ndarray
    var_prior_scaled: np.ndarray
    covmat_prior_scaled: np.ndarray
    uniform_idx: Optional[List[int]]
    lb: Optional[list]
    ub: Optional[list]
    responses_observed: List[np.ndarray]
    responses_cov: List[np.ndarray]
    max_covmat_size: int
    # derived for the device
    prior_mode: int = _lib.PRIOR_NONE
    prior_gidx: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    prior_w: np.ndarray = field(default_factory=lambda: np.zeros(0))
    prior_U: np.ndarray = field(default_factory=lambda: np.zeros(0))
    prior_c: float = 0.0
    resp_n: np.ndarray = None
    resp_mode: np.ndarray = None
    resp_c: np.ndarray = None
    y: np.ndarray = None
    w: np.ndarray = None
    cval: np.ndarray = None
    src: np.ndarray = None
    resp_U: list = None

    @property
    def n_total(self):
        return int(self.resp_n.sum())


def build_host_problem(prior_values, prior_uncertainties, observed, observed_uncertainties, lower_bounds=(),
                       upper_bounds=(), scaling_type="std", simulation_uncertainties=None,
                       max_covmat_size=100, covmat_whitening='eigh', keep_responses=()) -> HostProblem:
    mu = np.array(prior_values, dtype='float')                             # IP.py:146
    unc = np.array(prior_uncertainties, dtype='float')                     # IP.py:148
    k = len(mu)
    lb = list(lower_bounds) if len(lower_bounds) else None
    ub = list(upper_bounds) if len(upper_bounds) else None
    uniform_idx = None
    if -1.0 in unc:                                                        # IP.py:149-169
        if unc.ndim != 1:
            raise ValueError("uniform priors (-1) need a 1-D uncertainties array (IP.py:151-152)")
        if lb is None or ub is None:
            raise SystemExit("ERROR: an uncertainty of -1 (uniform prior) needs both upper and lower bounds (IP.py:164-165)")
        uniform_idx = []
        for i, u in enumerate(unc):
            if u == -1.0:
                uniform_idx.append(i)
                unc[i] = (ub[i] - lb[i]) / (12 ** 0.5)                     # IP.py:166
                mu[i] = (ub[i] + lb[i]) / 2                                # IP.py:169
    if unc.ndim == 1:                                                      # IP.py:172-175
        std = np.array(unc, dtype='float')
        var = np.power(unc, 2)
        covmat = np.diagflat(var)
    else:                                                                  # IP.py:176-179
        covmat = np.array(unc, dtype='float')
        var = np.diagonal(covmat)
        std = np.power(covmat, 0.5)
    if scaling_type == "off":                                              # IP.py:311-314
        scaling = None
        mu_s, var_s, cov_s = mu * 1.0, var * 1.0, covmat * 1.0
    else:
        if scaling_type == "std":
            scaling = std                                                  # IP.py:317
        elif scaling_type == "mu":
            scaling = mu                                                   # IP.py:319
        else:
            scaling = (mu / mu) * float(scaling_type)                      # IP.py:321-322
        if np.ndim(scaling) == 2:
            scaling = (np.diagonal(scaling)) ** 0.5                        # IP.py:328
        mu_s = np.array(mu / scaling)                                      # IP.py:333
        var_s = np.array(var / (scaling * scaling))                        # IP.py:334
        cov_s = covmat * 1.0
        for i, sv in enumerate(scaling):                                   # IP.py:336-339
            cov_s[i, :] = cov_s[i, :] / sv
            cov_s[:, i] = cov_s[:, i] / sv

    obs = nest_responses(observed)
    ounc = nest_uncertainties(observed_uncertainties, len(obs))
    sunc = None if simulation_uncertainties is None else nest_uncertainties(simulation_uncertainties, len(obs))
    covs = []
    for r in range(len(obs)):
        cov = shaped_response_cov(ounc, r)
        if sunc is not None:
            cov = cov + shaped_response_cov(sunc, r)                       # IP.py:2931
        if len(obs) == 1 and np.ndim(cov) == 2:
            raise NotImplementedError("a single response with a full covariance matrix never reaches scipy in the "
                                      "reference (IP.py:2938-2945 raises and falls through); split it into >= 2 responses")
        covs.append(np.asarray(cov, dtype=float))
    if len(keep_responses):
        # responses['reducedResponseSpace'] (reduceResponseSpace, IP.py:470-512): the observed values and covariances are reduced by
        # MEMBERSHIP after they were shaped on the full response set -- a vector of sigmas of a multi-response problem stays "as
        # variances" (shaped_response_cov) even when one response is left
        sel = [r for r in range(len(obs)) if r in keep_responses]
        obs, covs = [obs[r] for r in sel], [covs[r] for r in sel]

    H = HostProblem(k=k, mu_prior=mu, std_prior=std, var_prior=var, covmat_prior=covmat, scaling=scaling,
                    mu_prior_scaled=mu_s, var_prior_scaled=var_s, covmat_prior_scaled=cov_s,
                    uniform_idx=uniform_idx, lb=lb, ub=ub, responses_observed=obs, responses_cov=covs,
                    max_covmat_size=int(max_covmat_size))

    # ---- prior whitening (what logpdf would rebuild on every call at IP.py:2773 / 2785)
    if uniform_idx is None:
        gauss = np.arange(k)
        pw = psd(cov_s)
    else:
        gauss = np.delete(np.arange(k), uniform_idx)
        pw = psd(np.diagflat(np.delete(var_s, uniform_idx))) if len(gauss) else None   # IP.py:2778-2783
    if pw is None:
        H.prior_mode = _lib.PRIOR_NONE
    elif pw.is_perm:
        H.prior_mode = _lib.PRIOR_DIAG
        H.prior_gidx = gauss[pw.perm].astype(np.int32)
        H.prior_w = np.ascontiguousarray(pw.w)
        H.prior_c = pw.rank * LOG2PI + pw.log_pdet
    else:
        H.prior_mode = _lib.PRIOR_DENSE
        H.prior_gidx = gauss.astype(np.int32)
        H.prior_U = np.ascontiguousarray(pw.U)
        H.prior_c = pw.rank * LOG2PI + pw.log_pdet

    # ---- responses (IP.py:2949-2979)
    n_list, modes, cs, ys, ws, cvals, srcs, Us = [], [], [], [], [], [], [], []
    off = 0
    for r, yv in enumerate(obs):
        n = len(yv)
        cov = covs[r]
        ident = np.arange(n)
        if n > H.max_covmat_size:                                          # per-value path, IP.py:2951, 2969-2978
            if cov.ndim != 1:
                raise NotImplementedError("responses longer than responses_observed_max_covmat_size need variances, "
                                          "not a covariance matrix (IP.py:2971)")
            modes.append(_lib.RESP_PER_VALUE)
            cs.append(0.0)
            ys.append(yv)
            ws.append(np.sqrt(1.0 / cov))                                  # scalar _PSD: U = sqrt(1/V)
            cvals.append(1 * LOG2PI + np.log(cov))                         # rank*LOG2PI + log_pdet
            srcs.append(off + ident)
            Us.append(None)
        else:
            ps = psd(cov) if (covmat_whitening == 'eigh' or np.ndim(cov) != 2) else psd_cholesky(cov)
            cs.append(ps.rank * LOG2PI + ps.log_pdet)
            cvals.append(np.zeros(n))
            if ps.is_perm:
                modes.append(_lib.RESP_MVN_DIAG)
                ys.append(yv[ps.perm])
                ws.append(ps.w)
                srcs.append(off + ps.perm)
                Us.append(None)
            else:
                modes.append(_lib.RESP_MVN_FULL)
                ys.append(yv)
                ws.append(np.zeros(n))
                srcs.append(off + ident)
                Us.append(np.ascontiguousarray(ps.U))
        n_list.append(n)
        off += n
    H.resp_n = np.array(n_list, dtype=np.int32)
    H.resp_mode = np.array(modes, dtype=np.int32)
    H.resp_c = np.array(cs, dtype=float)
    H.y = np.ascontiguousarray(np.concatenate(ys))
    H.w = np.ascontiguousarray(np.concatenate(ws))
    H.cval = np.ascontiguousarray(np.concatenate(cvals))
    H.src = np.ascontiguousarray(np.concatenate(srcs).astype(np.int32))
    H.resp_U = Us
    return H


def _dp(a):
    return a.ctypes.data_as(_lib.c_double_p)


def _ip(a):
    return a.ctypes.data_as(_lib.c_int32_p)


class DeviceProblem:
    """Owns a peuqse_problem handle (immutable device copies of the constants)."""

    def __init__(self, H: HostProblem, model):
        self.H = H
        self.model = model
        L = _lib.load()
        k = H.k
        d = _lib.ProblemDesc()
        keep = []

        def hold(a, dtype=float):
            a = np.ascontiguousarray(np.asarray(a, dtype=dtype))
            keep.append(a)
            return a
        d.k = k
        d.scaling_on = 0 if H.scaling is None else 1
        d.has_lb = 0 if H.lb is None else 1
        d.has_ub = 0 if H.ub is None else 1
        lb = hold([-np.inf if (H.lb is None or v is None) else float(v) for v in (H.lb or [None] * k)])
        ub = hold([np.inf if (H.ub is None or v is None) else float(v) for v in (H.ub or [None] * k)])
        scale = hold(np.ones(k) if H.scaling is None else H.scaling)
        mus = hold(H.mu_prior_scaled)
        d.lb, d.ub, d.scale, d.mu_scaled = _dp(lb), _dp(ub), _dp(scale), _dp(mus)
        d.prior_mode = H.prior_mode
        d.kp = len(H.prior_gidx)
        gidx = hold(H.prior_gidx, np.int32)
        pw = hold(H.prior_w if len(H.prior_w) else np.zeros(max(1, d.kp)))
        pU = hold(H.prior_U if H.prior_U.size else np.zeros(1))
        d.prior_gidx, d.prior_w, d.prior_U = _ip(gidx), _dp(pw), _dp(pU)
        d.prior_c = float(H.prior_c)
        d.n_resp = len(H.resp_n)
        d.n_total = H.n_total
        rn, rm, rc = hold(H.resp_n, np.int32), hold(H.resp_mode, np.int32), hold(H.resp_c)
        d.resp_n, d.resp_mode, d.resp_c = _ip(rn), _ip(rm), _dp(rc)
        y, w, cv, src = hold(H.y), hold(H.w), hold(H.cval), hold(H.src, np.int32)
        d.y, d.w, d.cval, d.src = _dp(y), _dp(w), _dp(cv), _ip(src)
        Uarr = (_lib.c_double_p * d.n_resp)()
        for r, U in enumerate(H.resp_U):
            if U is not None:
                Uh = hold(U)
                Uarr[r] = _dp(Uh)
            else:
                Uarr[r] = None
        keep.append(Uarr)
        d.resp_U = C.cast(Uarr, C.POINTER(_lib.c_double_p))
        if isinstance(model, BuiltinModel):
            if model.n_total() != H.n_total:
                raise ValueError(f"model produces {model.n_total()} values but {H.n_total} were observed")
            if model.n_params != k:
                raise ValueError(f"model has {model.n_params} parameters but the prior has {k}")
            mc = hold(model.constants(H.src.astype(np.int64)))
            d.model_id = model.model_id
        else:
            mc = hold(np.zeros(0))
            d.model_id = _lib.MODEL_EXTERNAL
        d.n_model_const = len(mc)
        d.model_const = _dp(mc) if len(mc) else None
        handle = C.c_void_p()
        _lib.check(L.peuqse_problem_create(C.byref(d), C.byref(handle)))
        self.handle = handle
        self._L = L

    def close(self):
        if getattr(self, "handle", None):
            self._L.peuqse_problem_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
