"""Lock-step Nelder-Mead over many start points ("batched doOptimizeNegLogP multistart").

The reference's multistart with searchType 'doOptimizeNegLogP' / 'doOptimizeLogP' (PEUQSE/InverseProblem.py:912-921) runs
scipy.optimize.minimize(getNegLogP, start, method=...) once per start point, one objective evaluation (= one simulation) at a
time.  For method 'Nelder-Mead' the simplex logic is a short deterministic state machine, so all start points can advance
together: every stage of an iteration (reflection; expansion / outside / inside contraction; shrink) becomes ONE batched
launch of the log-posterior kernel over the start points that need that stage.

The arithmetic below restates scipy.optimize._optimize._minimize_neldermead (scipy 1.18, the version the reference runs with
here; non-adaptive coefficients, default initial simplex, xatol = fatol = 1e-4, maxiter = maxfev = 200 N, no bounds) operation
for operation, so each start point ends on exactly the point a stand-alone scipy run reaches (checked in tests/test_gpu_parity.py
against scipy itself and against the unmodified reference's multistart).
"""
import numpy as np


def nelder_mead_batch(f_batch, X0, maxiter=None, maxfev=None, xatol=1e-4, fatol=1e-4):
    """f_batch(points[M, N]) -> values[M] (numpy).  X0[B, N].  Returns dict(x[B, N], fun[B], nit[B], nfev[B], launches)."""
    X0 = np.array(X0, dtype=np.float64)
    B, N = X0.shape
    rho, chi, psi, sigma = 1, 2, 0.5, 0.5
    nonzdelt, zdelt = 0.05, 0.00025
    sim = np.empty((B, N + 1, N))
    sim[:, 0] = X0
    for k in range(N):
        y = np.array(X0, copy=True)
        nz = y[:, k] != 0
        y[nz, k] = (1 + nonzdelt) * y[nz, k]
        y[~nz, k] = zdelt
        sim[:, k + 1] = y
    maxfun = maxfev
    if maxiter is None and maxfun is None:
        maxiter, maxfun = N * 200, N * 200
    elif maxiter is None:
        maxiter = N * 200 if maxfun == np.inf else np.inf
    elif maxfun is None:
        maxfun = N * 200 if maxiter == np.inf else np.inf
    launches = [0]

    def f(points):
        launches[0] += 1
        return np.asarray(f_batch(np.ascontiguousarray(points)), dtype=float).reshape(-1)

    def sort(rows):
        ind = np.argsort(fsim[rows], axis=1)
        fsim[rows] = np.take_along_axis(fsim[rows], ind, axis=1)
        sim[rows] = np.take_along_axis(sim[rows], ind[:, :, None], axis=1)

    fsim = np.full((B, N + 1), np.inf)
    fcalls = np.zeros(B, dtype=np.int64)
    for k in range(N + 1):                               # the first N + 1 evaluations, stopped by maxfun like scipy's wrapper
        ok = fcalls < maxfun
        if ok.any():
            fsim[ok, k] = f(sim[ok, k])
            fcalls[ok] += 1
    sort(np.arange(B))
    iterations = np.ones(B, dtype=np.int64)
    active = np.ones(B, dtype=bool)
    while True:
        active &= (fcalls < maxfun) & (iterations < maxiter)
        a = np.nonzero(active)[0]
        if len(a):
            conv = (np.max(np.abs(sim[a, 1:] - sim[a, :1]).reshape(len(a), -1), axis=1) <= xatol) & \
                   (np.max(np.abs(fsim[a, :1] - fsim[a, 1:]), axis=1) <= fatol)
            active[a[conv]] = False
            a = a[~conv]
        if not len(a):
            break
        xbar = np.add.reduce(sim[a, :-1], 1) / N
        xr = (1 + rho) * xbar - rho * sim[a, -1]
        fxr = f(xr)
        fcalls[a] += 1
        aborted = np.zeros(len(a), dtype=bool)           # an evaluation refused by maxfun ends the iteration without counting it
        better = fxr < fsim[a, 0]
        accept_r = ~better & (fxr < fsim[a, -2])
        outside = ~better & ~accept_r & (fxr < fsim[a, -1])
        inside = ~better & ~accept_r & ~outside
        second = better | outside | inside
        refused = second & (fcalls[a] >= maxfun)
        aborted |= refused
        second &= ~refused
        pts = np.empty((len(a), N))
        pts[better] = (1 + rho * chi) * xbar[better] - rho * chi * sim[a[better], -1]
        pts[outside] = (1 + psi * rho) * xbar[outside] - psi * rho * sim[a[outside], -1]
        pts[inside] = (1 - psi) * xbar[inside] + psi * sim[a[inside], -1]
        f2 = np.full(len(a), np.nan)
        if second.any():
            f2[second] = f(pts[second])
            fcalls[a[second]] += 1
        doshrink = np.zeros(len(a), dtype=bool)
        # expansion
        m = better & second
        take_e = m & (f2 < fxr)
        sim[a[take_e], -1], fsim[a[take_e], -1] = pts[take_e], f2[take_e]
        take_r = (m & ~take_e) | accept_r
        sim[a[take_r], -1], fsim[a[take_r], -1] = xr[take_r], fxr[take_r]
        # outside contraction
        m = outside & second
        ok = m & (f2 <= fxr)
        sim[a[ok], -1], fsim[a[ok], -1] = pts[ok], f2[ok]
        doshrink |= m & ~ok
        # inside contraction
        m = inside & second
        ok = m & (f2 < fsim[a, -1])
        sim[a[ok], -1], fsim[a[ok], -1] = pts[ok], f2[ok]
        doshrink |= m & ~ok
        if doshrink.any():
            s = a[doshrink]
            live = np.ones(len(s), dtype=bool)
            for j in range(1, N + 1):
                sl = s[live]
                if not len(sl):
                    break
                sim[sl, j] = sim[sl, 0] + sigma * (sim[sl, j] - sim[sl, 0])
                can = fcalls[sl] < maxfun
                if can.any():
                    fsim[sl[can], j] = f(sim[sl[can], j])
                    fcalls[sl[can]] += 1
                stop = sl[~can]
                if len(stop):
                    aborted[np.isin(a, stop)] = True
                    live[np.isin(s, stop)] = False
        iterations[a[~aborted]] += 1
        sort(a)
    return dict(x=sim[:, 0].copy(), fun=np.min(fsim, axis=1), nit=iterations, nfev=fcalls, launches=launches[0])
