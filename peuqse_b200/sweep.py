"""Host driver of the multistart sweeps: three kernel passes + two tiny collectives.

pass 1  logP of every point of this rank's index range (optionally kept in HBM) + first strict maximum
        -> all-gather (max, index)            [doListOfPermutationsSearch loop, IP.py:844-956]
pass 2  weights int(exp(logP - thr)), shifted moments, min/max of the retained points
        -> all-reduce                         [convertPermutationsToSamples + calculatePostBurnInStatistics,
                                               IP.py:3463-3491, 1923-1931]
pass 3  weighted marginal histograms          [plotting_functions.py:135]
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import distributed as D


def grid_size(radii):
    n = 1
    for r in radii:
        n *= 2 * int(r) + 1
    return n


def make_spec(kind, k, center, interval=None, radii=None, spread=0.0, std_prior=None, seed=0):
    s = _lib.SweepSpec()
    s.kind = 0 if kind == "grid" else 1
    s.k = k
    for d in range(k):
        s.center[d] = float(center[d])
        s.interval[d] = float(interval[d]) if interval is not None else 0.0
        s.radii[d] = int(radii[d]) if radii is not None else 0
        s.std_prior[d] = float(std_prior[d]) if std_prior is not None else 0.0
    s.spread = float(spread)
    s.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return s


class SweepResult:
    pass


class Sweep:
    def __init__(self, dp, spec, total, store_logp=True):
        self.dp = dp
        self.spec = spec
        self.k = spec.k
        self.total = int(total)
        self.L = _lib.load()
        self.first, self.count = D.shard_range(self.total)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.work = torch.empty(int(self.L.peuqse_sweep_work_bytes()), dtype=torch.uint8, device=dev)
        self.best = torch.zeros(2, dtype=torch.float64, device=dev)      # peuqse_best {double, int64}
        self.logp = torch.empty(self.count, dtype=torch.float64, device=dev) if store_logp else None
        self.acc = torch.zeros(_lib.acc_layout(self.k)["SIZE"], dtype=torch.float64, device=dev)
        self.launches = 0

    # -- individual passes (device-resident in/out; used directly by bench.py for the kernel-only timing)
    def pass1(self):
        _lib.check(self.L.peuqse_sweep_pass1(self.dp.handle, C.byref(self.spec), self.first, self.count,
                                             _lib.ptr(self.logp), _lib.ptr(self.best), _lib.ptr(self.work),
                                             _lib.current_stream()))
        self.launches += 1

    def read_best(self):
        b = self.best.cpu()
        return float(b[0].item()), int(b.view(torch.int64)[1].item())

    def pass2(self, thr, keep=None):
        lay = _lib.acc_layout(self.k)
        self.acc.zero_()
        self.acc[lay["MIN"]:lay["MIN"] + self.k] = float("inf")
        self.acc[lay["MAX"]:lay["MAX"] + self.k] = float("-inf")
        kidx = klogp = kcount = None
        cap = 0
        if keep is not None:
            kidx, klogp, kcount = keep
            cap = kidx.numel()
            kcount.zero_()
        _lib.check(self.L.peuqse_sweep_pass2(self.dp.handle, C.byref(self.spec), self.first, self.count,
                                             _lib.ptr(self.logp), float(thr), _lib.ptr(self.acc), _lib.ptr(kidx),
                                             _lib.ptr(klogp), cap, _lib.ptr(kcount), _lib.ptr(self.work),
                                             _lib.current_stream()))
        self.launches += 1

    def pass3(self, thr, edges_dev, hist_dev, nbins):
        hist_dev.zero_()
        _lib.check(self.L.peuqse_sweep_hist(self.dp.handle, C.byref(self.spec), self.first, self.count,
                                            _lib.ptr(self.logp), float(thr), _lib.ptr(edges_dev), int(nbins),
                                            _lib.ptr(hist_dev), _lib.current_stream()))
        self.launches += 1

    def points(self, first, count):
        th = torch.empty((count, self.k), dtype=torch.float64, device=self.dev)
        if count:
            _lib.check(self.L.peuqse_sweep_points(C.byref(self.spec), int(first), int(count), _lib.ptr(th),
                                                  _lib.current_stream()))
        return th

    # -- the whole doMultiStart('getLogP') computation
    def run(self, coef=2.0, nbins=10, statistics=True):
        k = self.k
        R = SweepResult()
        self.pass1()
        lv, li = self.read_best()                                        # D2H: 16 bytes
        R.map_logP, R.map_index = D.allgather_best(lv, li)
        R.total = self.total
        if R.map_index < 0:
            R.map_parameter_set = np.full(k, np.nan)
            R.statistics = False
            return R
        R.map_parameter_set = self.points(R.map_index, 1).cpu().numpy()[0]
        R.statistics = bool(statistics)
        if not statistics:
            return R
        thr = R.map_logP + np.log(10 ** (-1 * coef))                     # IP.py:1016, 3477
        R.threshold = float(thr)
        self.pass2(thr)
        lay = _lib.acc_layout(k)
        acc = self.acc.cpu().numpy()                                      # D2H: (3+3k+k^2)*8 bytes
        sums = D.allreduce_np(acc[:lay["MIN"]], "sum")
        mn = D.allreduce_np(acc[lay["MIN"]:lay["MIN"] + k], "min")
        mx = D.allreduce_np(acc[lay["MAX"]:lay["MAX"] + k], "max")
        W = sums[lay["W"]]
        c = np.array([self.spec.center[d] for d in range(k)])
        S1 = sums[lay["S1"]:lay["S1"] + k]
        S2 = sums[lay["S2"]:lay["S2"] + k * k].reshape(k, k)
        R.sum_weights = W
        R.n_retained = int(round(sums[lay["NKEPT"]]))
        m1 = S1 / W
        R.mu_AP_parameter_set = c + m1                                   # np.mean of the expanded rows, IP.py:1923
        cov = S2 / W - np.outer(m1, m1)
        R.posterior_covariance = cov
        R.stdap_parameter_set = np.sqrt(np.maximum(np.diag(cov), 0.0))   # np.std (ddof 0), IP.py:1924
        sumsq = np.diag(S2) + 2.0 * c * S1 + c * c * W                   # sum w*theta_d^2
        R.evidence = (sums[lay["WEXP"]] / W) / np.sqrt(np.sum(sumsq))    # IP.py:1931
        R.retained_min, R.retained_max = mn, mx
        # histograms: np.histogram(bins=nbins) of every retained column (plotting_functions.py:135)
        edges = np.stack([np.linspace(mn[d], mx[d], nbins + 1) if mx[d] > mn[d]
                          else np.linspace(mn[d] - 0.5, mx[d] + 0.5, nbins + 1) for d in range(k)])
        ed = torch.from_numpy(edges).to(self.dev)
        hist = torch.zeros((k, nbins), dtype=torch.int64, device=self.dev)
        self.pass3(thr, ed, hist, nbins)
        R.histogram_counts = D.allreduce_int_np(hist.cpu().numpy())
        R.histogram_edges = edges
        return R
